"""ctypes binding of the CPU oracle (oracle/build/liboracle.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int64

import numpy as np

from proxsdp_b200._abi import bind_solve, call_solve
from proxsdp_b200.options import Options, OptionsPOD

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("proxsdp_oracle.c", "oracle_eig.c", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "proxsdp_b200_types.h"))
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _pd, _pi = POINTER(c_double), POINTER(c_int64)
        _lib.oracle_eigh.argtypes = [c_int64, _pd, _pd, _pd]
        _lib.oracle_eigh.restype = ctypes.c_int
        _lib.oracle_lanczos.argtypes = [c_int64, _pd, _pd, c_int64, c_int64, c_int64, c_double,
                                        _pd, _pd, _pi, _pi, _pi, _pi]
        _lib.oracle_lanczos.restype = ctypes.c_int
        _lib.oracle_eig_resid.argtypes = [c_int64, c_int64, c_int64, _pd]
        _lib.oracle_eig_resid.restype = None
        _lib.proxsdp_oracle_psd_project.argtypes = [c_int64, _pi, _pd, _pi, POINTER(OptionsPOD), c_int64,
                                                    c_int64, _pd, _pi, _pd, _pi, _pi]
        _lib.proxsdp_oracle_psd_project.restype = ctypes.c_int
        _lib.proxsdp_oracle_soc_project.argtypes = [c_int64, _pi, _pd]
        _lib.proxsdp_oracle_soc_project.restype = ctypes.c_int
        _lib.proxsdp_oracle_num_threads.restype = ctypes.c_int
        _lib.proxsdp_oracle_set_num_threads.argtypes = [ctypes.c_int]
    return _lib


def _pd(a):
    return a.ctypes.data_as(POINTER(c_double))


def _pi(a):
    return a.ctypes.data_as(POINTER(c_int64))


def chambolle_pock(aff, con, opt: Options, eig_resid=None):
    """Oracle counterpart of reference src/pdhg.jl:1 `chambolle_pock`."""
    fn = bind_solve(lib(), "proxsdp_oracle_solve")
    return call_solve(fn, aff, con, opt, eig_resid)


def eigh(A: np.ndarray):
    """Full symmetric eigendecomposition (upper triangle read), ascending."""
    n = A.shape[0]
    a = np.asfortranarray(A, dtype=np.float64).copy(order="F")
    w = np.zeros(n)
    Z = np.zeros((n, n), order="F")
    rc = lib().oracle_eigh(n, _pd(a), _pd(w), _pd(Z))
    assert rc == 0
    return w, Z


def lanczos(A: np.ndarray, x0: np.ndarray, howmany: int, krylovdim: int, maxiter: int = 100, tol: float = 1e-12):
    """KrylovKit-style eigsolve(:LR).  Returns vals, vecs, dict(converged, numops, numiter)."""
    n = A.shape[0]
    a = np.asfortranarray(A, dtype=np.float64)
    vals = np.zeros(krylovdim)
    vecs = np.zeros((n, krylovdim), order="F")
    nv, conv, nops, nit = (c_int64(0) for _ in range(4))
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    lib().oracle_lanczos(n, _pd(a), _pd(x0), howmany, krylovdim, maxiter, tol, _pd(vals), _pd(vecs),
                         ctypes.byref(nv), ctypes.byref(conv), ctypes.byref(nops), ctypes.byref(nit))
    k = nv.value
    return vals[:k].copy(), vecs[:, :k].copy(), dict(converged=conv.value, numops=nops.value, numiter=nit.value)


def eig_resid(n: int, seed: int = 1234, init: int = 3) -> np.ndarray:
    out = np.zeros(n)
    lib().oracle_eig_resid(n, seed, init, _pd(out))
    return out


def psd_project(sides, x, target_rank, opt: Options, iter: int = 1, mode: int = 0, resid=None):
    """One psd_projection! (reference src/prox_operators.jl:33-66) on concatenated svec blocks."""
    sides = np.ascontiguousarray(sides, dtype=np.int64)
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    tr = np.ascontiguousarray(target_rank, dtype=np.int64)
    k = len(sides)
    cur = np.zeros(k, dtype=np.int64)
    mineig = np.zeros(k)
    conv = np.zeros(k, dtype=np.int64)
    nops = c_int64(0)
    opod = opt.to_pod()
    rp = _pd(np.ascontiguousarray(resid, dtype=np.float64)) if resid is not None else ctypes.cast(None, POINTER(c_double))
    rc = lib().proxsdp_oracle_psd_project(k, _pi(sides), _pd(x), _pi(tr), ctypes.byref(opod), iter, mode, rp,
                                          _pi(cur), _pd(mineig), _pi(conv), ctypes.byref(nops))
    assert rc == 0
    return x, cur, mineig, conv, nops.value


def dual_step(A, G, n, p, m, opt: Options, **state):
    """Oracle counterpart of `proxsdp_b200_dual_step` (reference src/pdhg.jl:532-609)."""
    from proxsdp_b200._abi import call_dual_step
    return call_dual_step(lib().proxsdp_oracle_dual_step, A, G, n, p, m, opt, **state)


def residuals(n, p, m, opt: Options, **state):
    """Oracle counterpart of `proxsdp_b200_residuals` (reference src/residuals.jl:2-71)."""
    from proxsdp_b200._abi import call_residuals
    return call_residuals(lib().proxsdp_oracle_residuals, n, p, m, opt, **state)


def soc_project(lens, x):
    lens = np.ascontiguousarray(lens, dtype=np.int64)
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    lib().proxsdp_oracle_soc_project(len(lens), _pi(lens), _pd(x))
    return x


def num_threads() -> int:
    return lib().proxsdp_oracle_num_threads()


def set_num_threads(nt: int) -> None:
    lib().proxsdp_oracle_set_num_threads(nt)


def chambolle_pock_sharded(aff_loc, con_loc, opt: Options, info, reduce_fn):
    """Oracle counterpart of `proxsdp_b200_solve_sharded`: this rank's blocks, whole-problem scalars combined by
    `reduce_fn(vals: np.ndarray, op: int)` (op 0 = sum, 1 = max; in place).  Test infrastructure for the
    world_size > 1 host logic (gloo)."""
    from proxsdp_b200._abi import MarshalledProblem, MarshalledResult, ProblemPOD, ResultPOD
    from proxsdp_b200.sharding import REDUCE_FN, ShardPOD

    def _cb(ptr, count, op, _ctx):
        arr = np.ctypeslib.as_array(ptr, shape=(int(count),))
        reduce_fn(arr, int(op))

    cb = REDUCE_FN(_cb)
    shard = ShardPOD(info.rank, info.world, info.global_n, info.global_p, info.global_m, None, cb, None)
    mp = MarshalledProblem(aff_loc, con_loc)
    mr = MarshalledResult(mp.n, mp.p, mp.m, mp.n_sdp, int(opt.trace_cap))
    opod = opt.to_pod()
    fn = lib().proxsdp_oracle_solve_sharded
    fn.argtypes = [POINTER(ProblemPOD), POINTER(OptionsPOD), POINTER(ShardPOD), POINTER(ResultPOD)]
    fn.restype = ctypes.c_int
    rc = fn(ctypes.byref(mp.pod), ctypes.byref(opod), ctypes.byref(shard), ctypes.byref(mr.pod))
    assert rc == 0, rc
    return mr.to_result()
