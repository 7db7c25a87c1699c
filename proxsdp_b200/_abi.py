"""ctypes images of include/proxsdp_b200_types.h and marshalling helpers.

Used by the product binding (proxsdp_b200/solver.py) and by the oracle binding
(oracle/oracle.py) — both libraries export a `…_solve(problem*, options*, result*)`
with the same POD layouts.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char, c_double, c_int64

import numpy as np
import scipy.sparse as sp

from .options import Options, OptionsPOD
from .structs import AffineSets, ConicSets, Result, SparseMatrixCSC

STATUS_STRING_LEN = 256
TRACE_COLS = 14
TRACE_NAMES = [
    "iter", "prim_obj", "dual_obj", "gap", "feasibility", "primal_res", "dual_res",
    "primal_step", "beta", "sum_target_rank", "sum_current_rank", "min_eig", "matvecs", "ls_trials",
]

_pd = POINTER(c_double)
_pi = POINTER(c_int64)


class ProblemPOD(ctypes.Structure):
    _fields_ = [
        ("n", c_int64), ("p", c_int64), ("m", c_int64), ("index_base", c_int64),
        ("A_colptr", _pi), ("A_rowval", _pi), ("A_nzval", _pd),
        ("G_colptr", _pi), ("G_rowval", _pi), ("G_nzval", _pd),
        ("b", _pd), ("h", _pd), ("c", _pd),
        ("n_sdp", c_int64), ("sdp_side", _pi), ("sdp_ptr", _pi), ("sdp_idx", _pi),
        ("n_soc", c_int64), ("soc_ptr", _pi), ("soc_idx", _pi),
        ("eig_resid", _pd),
    ]


class ResultPOD(ctypes.Structure):
    _fields_ = [
        ("status", c_int64),
        ("status_string", c_char * STATUS_STRING_LEN),
        ("primal", _pd), ("dual_cone", _pd), ("dual_eq", _pd), ("dual_in", _pd),
        ("slack_eq", _pd), ("slack_in", _pd),
        ("primal_residual", c_double), ("dual_residual", c_double),
        ("objval", c_double), ("dual_objval", c_double), ("gap", c_double), ("time", c_double),
        ("iter", c_int64), ("final_rank", c_int64),
        ("primal_feasible_user_tol", c_int64), ("dual_feasible_user_tol", c_int64),
        ("certificate_found", c_int64), ("result_count", c_int64),
        ("final_primal_res", c_double), ("final_dual_res", c_double),
        ("time_setup", c_double), ("time_loop", c_double), ("time_psd_proj", c_double),
        ("n_psd_proj", c_int64), ("lanczos_matvecs", c_int64), ("lanczos_calls", c_int64),
        ("full_eig_calls", c_int64), ("linesearch_trials", c_int64), ("gpu_launches", c_int64),
        ("target_rank", _pi),
        ("trace", _pd), ("trace_len", c_int64),
        ("time_lanczos", c_double), ("time_rest", c_double), ("time_l2_flush", c_double),
        ("lanczos_timed_calls", c_int64), ("h2d_bytes", c_int64), ("d2h_bytes", c_int64),
        ("implicit_calls", c_int64),
    ]


class StepStatePOD(ctypes.Structure):
    """include/proxsdp_b200_types.h: proxsdp_step_state_t."""
    _fields_ = [
        ("n", c_int64), ("p", c_int64), ("m", c_int64),
        ("b", _pd), ("h", _pd), ("c", _pd),
        ("x", _pd), ("x_old", _pd), ("y", _pd), ("y_old", _pd),
        ("Mx", _pd), ("Mx_old", _pd), ("Mty", _pd), ("Mty_old", _pd),
        ("primal_step", c_double), ("primal_step_old", c_double), ("dual_step", c_double), ("theta", c_double),
        ("beta", c_double), ("norm_b", c_double), ("norm_h", c_double), ("norm_c", c_double),
    ]


class MarshalledStepState:
    """Owns the buffers a StepStatePOD points into.  Vectors that a seam does not read may be omitted (NULL)."""
    VECS = ("b", "h", "c", "x", "x_old", "y", "y_old", "Mx", "Mx_old", "Mty", "Mty_old")
    SCALARS = ("primal_step", "primal_step_old", "dual_step", "theta", "beta", "norm_b", "norm_h", "norm_c")

    def __init__(self, n: int, p: int, m: int, **kw):
        self.bufs = {}
        pod = StepStatePOD()
        pod.n, pod.p, pod.m = int(n), int(p), int(m)
        for name in self.VECS:
            v = kw.get(name)
            if v is None:
                setattr(pod, name, ctypes.cast(None, _pd))
            else:
                a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())
                self.bufs[name] = a
                setattr(pod, name, a.ctypes.data_as(_pd) if a.size else ctypes.cast(None, _pd))
        for name in self.SCALARS:
            setattr(pod, name, float(kw.get(name, 0.0)))
        self.pod = pod


def call_dual_step(fn, A, G, n, p, m, opt: Options, err_fn=None, **state):
    """Invoke `int dual_step(const proxsdp_problem_t* rows, options*, step_state*, y_new, Mty_new, scalars[4], trials*)`.
    Returns (y_new, Mty_new, dict(primal_step, theta, dual_step, primal_step_old), trials)."""
    aff = AffineSets(n, p, m, 0, A, G, np.zeros(p), np.zeros(m), np.zeros(n))
    mp = MarshalledProblem(aff, ConicSets())
    ms = MarshalledStepState(n, p, m, **state)
    y_new, Mty_new, sc = np.zeros(p + m), np.zeros(n), np.zeros(4)
    trials = c_int64(0)
    opod = opt.to_pod()
    fn.argtypes = [POINTER(ProblemPOD), POINTER(OptionsPOD), POINTER(StepStatePOD), _pd, _pd, _pd, _pi]
    fn.restype = ctypes.c_int
    rc = fn(ctypes.byref(mp.pod), ctypes.byref(opod), ctypes.byref(ms.pod), _ptr_d(y_new), _ptr_d(Mty_new), _ptr_d(sc),
            ctypes.byref(trials))
    if rc != 0:
        raise RuntimeError(f"dual_step failed with code {rc}: {err_fn().decode() if err_fn else ''}")
    return y_new, Mty_new, dict(primal_step=sc[0], theta=sc[1], dual_step=sc[2], primal_step_old=sc[3]), int(trials.value)


RESIDUAL_NAMES = ("primal_residual", "dual_residual", "comb_residual", "equa_feasibility", "ineq_feasibility",
                  "prim_obj", "dual_obj", "dual_gap")


def call_residuals(fn, n, p, m, opt: Options, err_fn=None, **state):
    """Invoke `int residuals(options*, step_state*, out[8])`; returns a dict keyed by RESIDUAL_NAMES."""
    ms = MarshalledStepState(n, p, m, **state)
    out = np.zeros(8)
    opod = opt.to_pod()
    fn.argtypes = [POINTER(OptionsPOD), POINTER(StepStatePOD), _pd]
    fn.restype = ctypes.c_int
    rc = fn(ctypes.byref(opod), ctypes.byref(ms.pod), _ptr_d(out))
    if rc != 0:
        raise RuntimeError(f"residuals failed with code {rc}: {err_fn().decode() if err_fn else ''}")
    return dict(zip(RESIDUAL_NAMES, (float(v) for v in out)))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64).ravel())


def _concat(arrs):
    """np.concatenate without the copy when there is nothing to join."""
    if len(arrs) == 0:
        return np.zeros(0, dtype=np.int64)
    if len(arrs) == 1:
        return arrs[0]
    return np.concatenate(arrs)


def _ptr_d(a: np.ndarray):
    return a.ctypes.data_as(_pd) if a.size else ctypes.cast(None, _pd)


def _ptr_i(a: np.ndarray):
    return a.ctypes.data_as(_pi) if a.size else ctypes.cast(None, _pi)


class MarshalledProblem:
    """Owns the numpy buffers a ProblemPOD points into."""

    def __init__(self, aff: AffineSets, con: ConicSets, eig_resid=None):
        n, p, m = int(aff.n), int(aff.p), int(aff.m)
        def csc_arrays(M, rows):
            """(colptr, rowval, nzval) as int64 / int64 / float64 without a copy when the matrix already is a
            SparseMatrixCSC; scipy matrices are canonicalised (sorted rows, duplicates summed — on a private copy)
            and widened."""
            if rows == 0:
                return np.zeros(n + 1, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0)
            if isinstance(M, SparseMatrixCSC):
                assert M.shape == (rows, n), (M.shape, rows, n)
                return M.colptr, M.rowval, M.nzval
            M = M if (sp.isspmatrix_csc(M) and M.shape == (rows, n)) else sp.csc_matrix(M, shape=(rows, n))
            if not M.has_canonical_format:
                M = M.copy()
                M.sum_duplicates()
            return _i64(M.indptr), _i64(M.indices), _f64(M.data)

        Acp, Arv, Anz = csc_arrays(aff.A, p)
        Gcp, Grv, Gnz = csc_arrays(aff.G, m)
        self.bufs = dict(
            A_colptr=Acp, A_rowval=Arv, A_nzval=Anz, G_colptr=Gcp, G_rowval=Grv, G_nzval=Gnz,
            b=_f64(aff.b), h=_f64(aff.h), c=_f64(aff.c),
        )
        sides = [int(s.sq_side) for s in con.sdpcone]
        sdp_ptr = np.zeros(len(sides) + 1, dtype=np.int64)
        for k, s in enumerate(con.sdpcone):
            assert len(s.vec_i) == s.tri_len == s.sq_side * (s.sq_side + 1) // 2
            sdp_ptr[k + 1] = sdp_ptr[k] + s.tri_len
        soc_ptr = np.zeros(len(con.socone) + 1, dtype=np.int64)
        for k, s in enumerate(con.socone):
            soc_ptr[k + 1] = soc_ptr[k] + len(s.idx)
        self.bufs.update(
            sdp_side=_i64(sides), sdp_ptr=sdp_ptr,
            sdp_idx=_i64(_concat([s.vec_i for s in con.sdpcone])),
            soc_ptr=soc_ptr,
            soc_idx=_i64(_concat([s.idx for s in con.socone])),
        )
        if eig_resid is not None:
            self.bufs["eig_resid"] = _f64(eig_resid)
        pod = ProblemPOD()
        pod.n, pod.p, pod.m, pod.index_base = n, p, m, 0
        for name in ("A_colptr", "A_rowval", "G_colptr", "G_rowval", "sdp_side", "sdp_ptr", "sdp_idx",
                     "soc_ptr", "soc_idx"):
            setattr(pod, name, _ptr_i(self.bufs[name]))
        for name in ("A_nzval", "G_nzval", "b", "h", "c"):
            setattr(pod, name, _ptr_d(self.bufs[name]))
        pod.n_sdp = len(sides)
        pod.n_soc = len(con.socone)
        pod.eig_resid = _ptr_d(self.bufs["eig_resid"]) if eig_resid is not None else ctypes.cast(None, _pd)
        self.pod = pod
        self.n, self.p, self.m, self.n_sdp = n, p, m, len(sides)


class MarshalledResult:
    def __init__(self, n: int, p: int, m: int, n_sdp: int, trace_cap: int = 0, empty=None):
        # `empty(count)`: allocator of the two n-long outputs (the product binding passes page-locked memory so that
        # the device->host copies need no staging and no first-touch page faults)
        self.primal = np.zeros(n) if empty is None else empty(n)
        self.dual_cone = np.zeros(n) if empty is None else empty(n)
        self.dual_eq = np.zeros(p)
        self.dual_in = np.zeros(m)
        self.slack_eq = np.zeros(p)
        self.slack_in = np.zeros(m)
        self.target_rank = np.zeros(max(n_sdp, 1), dtype=np.int64)
        self.trace = np.zeros((max(trace_cap, 1), TRACE_COLS))
        self.n_sdp = n_sdp
        pod = ResultPOD()
        pod.primal = _ptr_d(self.primal)
        pod.dual_cone = _ptr_d(self.dual_cone)
        pod.dual_eq = _ptr_d(self.dual_eq)
        pod.dual_in = _ptr_d(self.dual_in)
        pod.slack_eq = _ptr_d(self.slack_eq)
        pod.slack_in = _ptr_d(self.slack_in)
        pod.target_rank = _ptr_i(self.target_rank)
        pod.trace = _ptr_d(self.trace) if trace_cap > 0 else ctypes.cast(None, _pd)
        pod.trace_len = 0
        self.pod = pod

    def to_result(self) -> Result:
        q = self.pod
        return Result(
            status=int(q.status),
            status_string=q.status_string.decode("utf-8", "replace"),
            primal=self.primal, dual_cone=self.dual_cone, dual_eq=self.dual_eq, dual_in=self.dual_in,
            slack_eq=self.slack_eq, slack_in=self.slack_in,
            primal_residual=q.primal_residual, dual_residual=q.dual_residual,
            objval=q.objval, dual_objval=q.dual_objval, gap=q.gap, time=q.time,
            iter=int(q.iter), final_rank=int(q.final_rank),
            primal_feasible_user_tol=bool(q.primal_feasible_user_tol),
            dual_feasible_user_tol=bool(q.dual_feasible_user_tol),
            certificate_found=bool(q.certificate_found), result_count=int(q.result_count),
            final_primal_res=q.final_primal_res, final_dual_res=q.final_dual_res,
            time_setup=q.time_setup, time_loop=q.time_loop, time_psd_proj=q.time_psd_proj,
            n_psd_proj=int(q.n_psd_proj), lanczos_matvecs=int(q.lanczos_matvecs),
            lanczos_calls=int(q.lanczos_calls), full_eig_calls=int(q.full_eig_calls),
            linesearch_trials=int(q.linesearch_trials), gpu_launches=int(q.gpu_launches),
            target_rank=self.target_rank[: self.n_sdp].copy(),
            trace=self.trace[: int(q.trace_len)].copy(),
            time_lanczos=q.time_lanczos, time_rest=q.time_rest, time_l2_flush=q.time_l2_flush,
            lanczos_timed_calls=int(q.lanczos_timed_calls), h2d_bytes=int(q.h2d_bytes), d2h_bytes=int(q.d2h_bytes),
            implicit_calls=int(q.implicit_calls),
        )


def call_solve(fn, aff: AffineSets, con: ConicSets, opt: Options, eig_resid=None, err_fn=None, empty=None) -> Result:
    """Invoke a `int solve(const proxsdp_problem_t*, const proxsdp_options_t*, proxsdp_result_t*)`."""
    mp = MarshalledProblem(aff, con, eig_resid)
    mr = MarshalledResult(mp.n, mp.p, mp.m, mp.n_sdp, int(opt.trace_cap), empty=empty)
    opod = opt.to_pod()
    rc = fn(ctypes.byref(mp.pod), ctypes.byref(opod), ctypes.byref(mr.pod))
    if rc != 0:
        msg = err_fn().decode() if err_fn is not None else ""
        raise RuntimeError(f"solve failed with code {rc}: {msg}")
    return mr.to_result()


def bind_solve(lib, name: str):
    fn = getattr(lib, name)
    fn.argtypes = [POINTER(ProblemPOD), POINTER(OptionsPOD), POINTER(ResultPOD)]
    fn.restype = ctypes.c_int
    return fn
