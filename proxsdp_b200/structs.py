"""Python mirrors of the reference's data structs that cross the `chambolle_pock` seam.

AffineSets / SDPSet / SOCSet / ConicSets / Result follow reference
src/structs.jl:32-81 field for field.  Indices are 0-based on the Python side
(the C ABI accepts either base through `index_base`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import scipy.sparse as sp


class SparseMatrixCSC:
    """Julia's `SparseMatrixCSC{Float64,Int64}` (the type of `AffineSets.A` / `.G`, structs.jl:32-42) as three flat
    arrays: colptr (n + 1, int64), rowval (nnz, int64), nzval (nnz, float64), 0-based.  scipy down-casts index arrays
    to int32, so a scipy matrix has to be widened on every call of the C ABI; a problem held in this form crosses the
    boundary without a copy."""

    __slots__ = ("m", "n", "colptr", "rowval", "nzval")

    def __init__(self, m: int, n: int, colptr, rowval, nzval):
        self.m, self.n = int(m), int(n)
        self.colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        self.rowval = np.ascontiguousarray(rowval, dtype=np.int64)
        self.nzval = np.ascontiguousarray(nzval, dtype=np.float64)
        assert self.colptr.shape == (self.n + 1,) and self.rowval.shape == self.nzval.shape

    @property
    def shape(self):
        return (self.m, self.n)

    @property
    def nnz(self) -> int:
        return int(self.nzval.size)

    @classmethod
    def from_scipy(cls, M) -> "SparseMatrixCSC":
        M = sp.csc_matrix(M)
        if not M.has_canonical_format:
            M = M.copy()
            M.sum_duplicates()
        return cls(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)

    def to_scipy(self) -> sp.csc_matrix:
        return sp.csc_matrix((self.nzval, self.rowval, self.colptr), shape=(self.m, self.n))


@dataclass
class AffineSets:
    """structs.jl:32-42.  A is p x n, G is m x n (scipy CSC or SparseMatrixCSC), b (p), h (m), c (n)."""

    n: int
    p: int
    m: int
    extra: int
    A: sp.csc_matrix
    G: sp.csc_matrix
    b: np.ndarray
    h: np.ndarray
    c: np.ndarray


@dataclass
class SDPSet:
    """structs.jl:44-48.  vec_i: variable indices of the column-major upper triangle."""

    vec_i: np.ndarray
    tri_len: int
    sq_side: int


@dataclass
class SOCSet:
    """structs.jl:50-53.  idx[0] is the epigraph variable t."""

    idx: np.ndarray
    len: int


@dataclass
class ConicSets:
    """structs.jl:55-58."""

    sdpcone: List[SDPSet] = field(default_factory=list)
    socone: List[SOCSet] = field(default_factory=list)


@dataclass
class Result:
    """structs.jl:60-81 plus measurement extras (not part of the reference Result)."""

    status: int = 0
    status_string: str = "Problem not solved"
    primal: np.ndarray = field(default_factory=lambda: np.zeros(0))
    dual_cone: np.ndarray = field(default_factory=lambda: np.zeros(0))
    dual_eq: np.ndarray = field(default_factory=lambda: np.zeros(0))
    dual_in: np.ndarray = field(default_factory=lambda: np.zeros(0))
    slack_eq: np.ndarray = field(default_factory=lambda: np.zeros(0))
    slack_in: np.ndarray = field(default_factory=lambda: np.zeros(0))
    primal_residual: float = float("nan")
    dual_residual: float = float("nan")
    objval: float = float("nan")
    dual_objval: float = float("nan")
    gap: float = float("nan")
    time: float = float("nan")
    iter: int = -1
    final_rank: int = -1
    primal_feasible_user_tol: bool = False
    dual_feasible_user_tol: bool = False
    certificate_found: bool = False
    result_count: int = 0
    # extras
    final_primal_res: float = float("nan")
    final_dual_res: float = float("nan")
    time_setup: float = 0.0
    time_loop: float = 0.0
    time_psd_proj: float = 0.0
    n_psd_proj: int = 0
    lanczos_matvecs: int = 0
    lanczos_calls: int = 0
    full_eig_calls: int = 0
    linesearch_trials: int = 0
    gpu_launches: int = 0
    target_rank: Optional[np.ndarray] = None
    trace: Optional[np.ndarray] = None
    time_lanczos: float = 0.0
    time_rest: float = 0.0
    time_l2_flush: float = 0.0
    lanczos_timed_calls: int = 0
    h2d_bytes: int = 0
    d2h_bytes: int = 0
    implicit_calls: int = 0


def sympackedlen(n: int) -> int:
    """MOI_wrapper.jl:218."""
    return n * (n + 1) // 2


def sympackeddim(length: int) -> int:
    """MOI.Utilities.side_dimension_for_vectorized_dimension."""
    n = int((np.sqrt(8 * length + 1) - 1) // 2)
    assert n * (n + 1) // 2 == length
    return n


def ivech(v: np.ndarray) -> np.ndarray:
    """util.jl:18-36: svec (column-major upper triangle) -> upper-triangular matrix."""
    n = sympackeddim(len(v))
    out = np.zeros((n, n))
    c = 0
    for j in range(n):
        for i in range(j + 1):
            out[i, j] = v[c]
            c += 1
    return out


def ivec(v: np.ndarray) -> np.ndarray:
    """util.jl:38: full symmetric matrix from svec."""
    u = ivech(v)
    return u + np.triu(u, 1).T
