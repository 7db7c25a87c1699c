"""ctypes binding of the CUDA library behind the reference's `chambolle_pock` seam.

The product path has NO CPU fallback: if `libproxsdp_b200.so` is missing or no CUDA
device is present, every entry point raises.  (The CPU oracle lives under oracle/ and
is test infrastructure only.)
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64

import numpy as np

from ._abi import OptionsPOD, ProblemPOD, ResultPOD, bind_solve, call_solve
from .options import Options
from .structs import AffineSets, ConicSets, Result

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libproxsdp_b200.so")
_lib = None

_pd = POINTER(c_double)
_pi = POINTER(c_int64)


class ExtensionMissing(RuntimeError):
    pass


def lib():
    """Load the CUDA extension; fail loudly when it is absent or its ABI does not match."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ExtensionMissing(
            f"{LIB_PATH} not found: build it with `python -m proxsdp_b200.build` "
            "(there is no CPU fallback for the product path)")
    L = ctypes.CDLL(LIB_PATH)
    L.proxsdp_b200_last_error.restype = c_char_p
    L.proxsdp_b200_version.restype = c_char_p
    for nm in ("problem", "options", "result"):
        getattr(L, f"proxsdp_b200_sizeof_{nm}").restype = c_int64
    if (L.proxsdp_b200_sizeof_problem() != ctypes.sizeof(ProblemPOD)
            or L.proxsdp_b200_sizeof_options() != ctypes.sizeof(OptionsPOD)
            or L.proxsdp_b200_sizeof_result() != ctypes.sizeof(ResultPOD)):
        raise ExtensionMissing("ABI mismatch between proxsdp_b200/_abi.py and libproxsdp_b200.so")
    L.proxsdp_b200_psd_project.argtypes = [c_int64, _pi, _pd, _pi, POINTER(OptionsPOD), c_int64, c_int64, _pd,
                                           _pi, _pd, _pi, _pi, c_int64, _pd]
    L.proxsdp_b200_psd_project.restype = c_int
    L.proxsdp_b200_soc_project.argtypes = [c_int64, _pi, _pd]
    L.proxsdp_b200_soc_project.restype = c_int
    L.proxsdp_b200_lanczos.argtypes = [c_int64, _pd, _pd, c_int64, c_int64, c_int64, c_double, _pd, _pd,
                                       _pi, _pi, _pi, _pi, c_int64, _pd]
    L.proxsdp_b200_lanczos.restype = c_int
    L.proxsdp_b200_eigh.argtypes = [c_int64, _pd, _pd, _pd]
    L.proxsdp_b200_eigh.restype = c_int
    L.proxsdp_b200_device_count.restype = c_int
    L.proxsdp_b200_host_alloc.argtypes = [c_int64]
    L.proxsdp_b200_host_alloc.restype = ctypes.c_void_p
    L.proxsdp_b200_host_free.argtypes = [ctypes.c_void_p]
    L.proxsdp_b200_host_free.restype = c_int
    _lib = L
    return L


def pinned_empty(count: int, dtype=np.float64) -> np.ndarray:
    """Uninitialised array in page-locked host memory from the library's cache (`proxsdp_b200_host_alloc`); the
    block goes back to the cache when the last view of the array is collected."""
    import weakref
    L = lib()
    dt = np.dtype(dtype)
    nbytes = max(int(count) * dt.itemsize, 1)
    ptr = L.proxsdp_b200_host_alloc(nbytes)
    if not ptr:
        raise RuntimeError(f"proxsdp_b200_host_alloc failed: {L.proxsdp_b200_last_error().decode()}")
    raw = (ctypes.c_char * nbytes).from_address(ptr)
    weakref.finalize(raw, L.proxsdp_b200_host_free, ctypes.c_void_p(ptr))
    return np.frombuffer(raw, dtype=dt, count=int(count))


def pinned_copy(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    out = pinned_empty(a.size, a.dtype)
    out[...] = a.ravel()
    return out.reshape(a.shape)


def pin_problem(aff: AffineSets, con: ConicSets):
    """The same problem with every array of the `chambolle_pock` arguments in page-locked memory and the matrices as
    SparseMatrixCSC{Float64,Int64} (what a Julia caller holds): the C ABI then reads the caller's buffers in place."""
    from .structs import SDPSet, SOCSet, SparseMatrixCSC

    def mat(M):
        M = M if isinstance(M, SparseMatrixCSC) else SparseMatrixCSC.from_scipy(M)
        return SparseMatrixCSC(M.m, M.n, pinned_copy(M.colptr), pinned_copy(M.rowval), pinned_copy(M.nzval))

    def vec(v, dtype=np.float64):
        return pinned_copy(np.asarray(v, dtype=dtype))

    aff2 = AffineSets(aff.n, aff.p, aff.m, aff.extra, mat(aff.A), mat(aff.G), vec(aff.b), vec(aff.h), vec(aff.c))
    con2 = ConicSets([SDPSet(vec(s.vec_i, np.int64), s.tri_len, s.sq_side) for s in con.sdpcone],
                     [SOCSet(vec(s.idx, np.int64), s.len) for s in con.socone])
    return aff2, con2


def _check(rc: int):
    if rc != 0:
        raise RuntimeError(f"proxsdp_b200 error {rc}: {lib().proxsdp_b200_last_error().decode()}")


def device_count() -> int:
    return int(lib().proxsdp_b200_device_count())


def _dp(a):
    return a.ctypes.data_as(_pd)


def _ip(a):
    return a.ctypes.data_as(_pi)


def chambolle_pock(aff: AffineSets, con: ConicSets, opt: Options, eig_resid=None) -> Result:
    """Drop-in for reference src/pdhg.jl:1 `chambolle_pock(affine_sets, conic_sets, opt)::Result`."""
    L = lib()
    fn = bind_solve(L, "proxsdp_b200_solve")
    return call_solve(fn, aff, con, opt, eig_resid, err_fn=L.proxsdp_b200_last_error, empty=pinned_empty)


class Solve:
    """`chambolle_pock` in three calls (create / iterate / finish) for callers that own the loop —
    one `iterate` step is one pass of the reference's `for k in 1:2*opt.max_iter_local` body
    (src/pdhg.jl:145-484).  Device state stays resident in HBM between calls."""

    COUNT_NAMES = ("iterations", "launches", "lanczos_matvecs", "lanczos_calls", "lanczos_timed_calls",
                   "full_eig_calls", "linesearch_trials", "sum_target_rank")
    TIME_NAMES = ("psd_proj_ms", "lanczos_ms", "rest_ms", "l2_flush_ms")

    def __init__(self, aff: AffineSets, con: ConicSets, opt: Options, eig_resid=None):
        from ._abi import MarshalledProblem
        L = lib()
        self._L = L
        self._mp = MarshalledProblem(aff, con, eig_resid)
        self._opt = opt
        self._h = ctypes.c_void_p()
        opod = opt.to_pod()
        L.proxsdp_b200_create.argtypes = [POINTER(ProblemPOD), POINTER(OptionsPOD), POINTER(ctypes.c_void_p)]
        L.proxsdp_b200_iterate.argtypes = [ctypes.c_void_p, c_int64, c_int64, _pi, _pi, _pd]
        L.proxsdp_b200_counters.argtypes = [ctypes.c_void_p, _pi, _pd]
        L.proxsdp_b200_finish.argtypes = [ctypes.c_void_p, POINTER(ResultPOD)]
        L.proxsdp_b200_destroy.argtypes = [ctypes.c_void_p]
        _check(L.proxsdp_b200_create(ctypes.byref(self._mp.pod), ctypes.byref(opod), ctypes.byref(self._h)))

    def iterate(self, max_steps: int = -1, flush_l2: bool = False):
        """Returns (steps_done, finished, device_ms)."""
        done, fin, ms = c_int64(0), c_int64(0), c_double(0.0)
        _check(self._L.proxsdp_b200_iterate(self._h, max_steps, int(flush_l2), ctypes.byref(done), ctypes.byref(fin),
                                            ctypes.byref(ms)))
        return done.value, bool(fin.value), ms.value

    def counters(self) -> dict:
        c = np.zeros(8, dtype=np.int64)
        t = np.zeros(4)
        _check(self._L.proxsdp_b200_counters(self._h, _ip(c), _dp(t)))
        out = dict(zip(self.COUNT_NAMES, (int(v) for v in c)))
        out.update(dict(zip(self.TIME_NAMES, (float(v) for v in t))))
        return out

    def finish(self) -> Result:
        from ._abi import MarshalledResult
        mr = MarshalledResult(self._mp.n, self._mp.p, self._mp.m, self._mp.n_sdp, int(self._opt.trace_cap))
        _check(self._L.proxsdp_b200_finish(self._h, ctypes.byref(mr.pod)))
        return mr.to_result()

    def close(self):
        if self._h:
            self._L.proxsdp_b200_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def psd_project(sides, x, target_rank, opt: Options, iter: int = 1, mode: int = 0, resid=None, repeat: int = 1):
    """One `psd_projection!` (reference src/prox_operators.jl:33-66) on concatenated svec blocks.
    Returns (x_projected, current_rank, min_eig, converged, numops, ms_per_call)."""
    sides = np.ascontiguousarray(sides, dtype=np.int64)
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    tr = np.ascontiguousarray(target_rank, dtype=np.int64)
    k = len(sides)
    cur = np.zeros(max(k, 1), dtype=np.int64)
    mineig = np.zeros(max(k, 1))
    conv = np.zeros(max(k, 1), dtype=np.int64)
    nops = c_int64(0)
    ms = c_double(0.0)
    opod = opt.to_pod()
    if resid is not None:
        resid = np.ascontiguousarray(resid, dtype=np.float64)
        rp = _dp(resid)
    else:
        rp = ctypes.cast(None, _pd)
    _check(lib().proxsdp_b200_psd_project(k, _ip(sides), _dp(x), _ip(tr), ctypes.byref(opod), iter, mode, rp,
                                          _ip(cur), _dp(mineig), _ip(conv), ctypes.byref(nops), repeat,
                                          ctypes.byref(ms)))
    return x, cur[:k], mineig[:k], conv[:k], nops.value, ms.value


def dual_step(A, G, n, p, m, opt: Options, **state):
    """`linesearch!` / `dual_step!` (reference src/pdhg.jl:532-609) on explicit working-space state: A (p x n), G (m x n)
    and y, Mx, Mx_old, Mty, b, h, primal_step, primal_step_old, theta, beta, dual_step as keywords.
    Returns (y_new, Mty_new, scalars, trials)."""
    from ._abi import call_dual_step
    L = lib()
    return call_dual_step(L.proxsdp_b200_dual_step, A, G, n, p, m, opt, err_fn=L.proxsdp_b200_last_error, **state)


def residuals(n, p, m, opt: Options, **state):
    """`compute_residual!` + `compute_gap!` (reference src/residuals.jl:2-71) on explicit working-space state."""
    from ._abi import call_residuals
    L = lib()
    return call_residuals(L.proxsdp_b200_residuals, n, p, m, opt, err_fn=L.proxsdp_b200_last_error, **state)


def soc_project(lens, x):
    """`soc_projection!` (reference src/prox_operators.jl:138-158)."""
    lens = np.ascontiguousarray(lens, dtype=np.int64)
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    _check(lib().proxsdp_b200_soc_project(len(lens), _ip(lens), _dp(x)))
    return x


def lanczos(A, x0, howmany: int, krylovdim: int, maxiter: int = 100, tol: float = 1e-12, repeat: int = 1):
    """KrylovKit-style eigsolve(:LR) on the device (reference src/eigsolver.jl:802-812)."""
    A = np.asfortranarray(A, dtype=np.float64)
    n = A.shape[0]
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    vals = np.zeros(krylovdim)
    vecs = np.zeros((n, krylovdim), order="F")
    nv, conv, nops, nit = (c_int64(0) for _ in range(4))
    ms = c_double(0.0)
    _check(lib().proxsdp_b200_lanczos(n, _dp(A), _dp(x0), howmany, krylovdim, maxiter, tol, _dp(vals), _dp(vecs),
                                      ctypes.byref(nv), ctypes.byref(conv), ctypes.byref(nops), ctypes.byref(nit),
                                      repeat, ctypes.byref(ms)))
    k = nv.value
    return vals[:k].copy(), vecs[:, :k].copy(), dict(converged=conv.value, numops=nops.value, numiter=nit.value,
                                                     ms=ms.value)


def eigh(A):
    """Full symmetric eigendecomposition on the device (ascending)."""
    A = np.asfortranarray(A, dtype=np.float64)
    n = A.shape[0]
    w = np.zeros(n)
    Z = np.zeros((n, n), order="F")
    _check(lib().proxsdp_b200_eigh(n, _dp(A), _dp(w), _dp(Z)))
    return w, Z
