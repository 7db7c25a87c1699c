"""Problem builders for the BASELINE.json configs.

Each builder returns `(AffineSets, ConicSets)` exactly as the reference's
`_optimize!` (src/MOI_wrapper.jl:229-292) would assemble them from the JuMP model
the corresponding reference script builds:

  maxcut_problem        README.md:58-84                         (config C1 with README W)
  sdplib_problem        test/base_sdplib.jl:1-45 + test/jump_sdplib.jl:7-20   (C2, C3)
  mimo_problem          test/base_mimo.jl:3-17  + test/jump_mimo.jl:1-16      (C4)
  sensorloc_problem     test/base_sensorloc.jl:2-21 + test/jump_sensorloc.jl:9-53 (C5)
  randsdp_problem       test/base_randsdp.jl:4-25 + test/jump_randsdp.jl:1-9

Julia's MersenneTwister streams are not reproducible outside Julia, so the random
instances are drawn from numpy's PCG64 / a splitmix64 edge generator instead and are
meant to be exchanged as files (`write_sdpa`, `save_problem`).
"""
from __future__ import annotations

import io
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import scipy.sparse as sp

from .structs import AffineSets, ConicSets, SDPSet, SOCSet

README_W = np.array(
    [[18.0, -5.0, -7.0, -6.0], [-5.0, 6.0, 0.0, -1.0], [-7.0, 0.0, 8.0, -1.0], [-6.0, -1.0, -1.0, 8.0]]
)  # README.md:68-73


def svec_index(i, j):
    """0-based position of (i, j), i <= j, in the column-major upper triangle."""
    return j * (j + 1) // 2 + i


def _psd_only_cones(n: int, offset: int = 0) -> ConicSets:
    tri = n * (n + 1) // 2
    return ConicSets([SDPSet(np.arange(offset, offset + tri, dtype=np.int64), tri, n)], [])


def _sym_to_svec_coeffs(rows, cols, vals, n):
    """JuMP's `sum(F[i,j] * X[i,j] for (i,j) in findnz(F))` over a symmetric F that stores
    both triangles: coefficient of the svec variable (i<=j) is F[i,j] (+ F[j,i] if i != j)."""
    rows = np.asarray(rows)
    cols = np.asarray(cols)
    vals = np.asarray(vals, dtype=np.float64)
    i = np.minimum(rows, cols)
    j = np.maximum(rows, cols)
    k = svec_index(i, j)
    return k, vals


# --------------------------------------------------------------------------
# Max-Cut from a weight matrix (README example)
# --------------------------------------------------------------------------
def maxcut_problem(W: np.ndarray) -> Tuple[AffineSets, ConicSets, float]:
    """`@objective(model, Max, 0.25*dot(W, X)); @constraint(diag(X) .== 1)`.
    Returns (aff, con, obj_sign) with c already sign-flipped for Max."""
    n = W.shape[0]
    N = n * (n + 1) // 2
    c = np.zeros(N)
    for j in range(n):
        for i in range(j + 1):
            c[svec_index(i, j)] = -0.25 * (W[i, j] if i == j else W[i, j] + W[j, i])
    diag_idx = np.array([svec_index(i, i) for i in range(n)])
    A = sp.csc_matrix((np.ones(n), (np.arange(n), diag_idx)), shape=(n, N))
    aff = AffineSets(N, n, 0, 0, A, sp.csc_matrix((0, N)), np.ones(n), np.zeros(0), c)
    return aff, _psd_only_cones(n), -1.0


# --------------------------------------------------------------------------
# SDPA-sparse (.dat-s) reader / writer
# --------------------------------------------------------------------------
def read_sdpa(path: str):
    """Restates `sdplib_data` (test/base_sdplib.jl:1-45): returns (n, m, F, c) where
    F[k] = (rows, cols, vals) upper-triangle triplets (0-based), F[0] already negated,
    and — the loader's quirk — n = length(c) (= m), not the block size."""
    with open(path, "r") as fh:
        lines = [ln for ln in fh.read().splitlines() if ln.strip() != ""]
    m = int(float(lines[0].split()[0]))

    def parse_list(line):
        s = line.strip()
        if s[0] in "{(":
            s = s[1:-1]
        return [float(t) for t in s.replace(",", " ").split()]

    blks = [int(v) for v in parse_list(lines[2])]
    cum = np.concatenate([[0], np.cumsum(np.abs(blks))]).astype(np.int64)
    c = np.array(parse_list(lines[3]), dtype=np.float64)
    n = len(c)
    data = np.loadtxt(io.StringIO("\n".join(lines[4:])), ndmin=2)
    matno = data[:, 0].astype(np.int64)
    blk = data[:, 1].astype(np.int64)
    ii = data[:, 2].astype(np.int64) - 1 + cum[blk - 1]
    jj = data[:, 3].astype(np.int64) - 1 + cum[blk - 1]
    vv = data[:, 4].astype(np.float64)
    vv = np.where(matno == 0, -vv, vv)
    return n, m, (matno, ii, jj, vv), c


def sdplib_problem(path: str) -> Tuple[AffineSets, ConicSets]:
    """`jump_sdplib` (test/jump_sdplib.jl:5-20): X n x n PSD, Min <F0, X>, <Fk, X> == c[k]."""
    n, m, (matno, ii, jj, vv), c = read_sdpa(path)
    N = n * (n + 1) // 2
    lo = np.minimum(ii, jj)
    hi = np.maximum(ii, jj)
    k = svec_index(lo, hi)
    coef = np.where(lo == hi, vv, 2.0 * vv)      # F[i,j] + F[j,i]
    obj = matno == 0
    cvec = np.zeros(N)
    np.add.at(cvec, k[obj], coef[obj])
    con = ~obj
    A = sp.coo_matrix((coef[con], (matno[con] - 1, k[con])), shape=(m, N)).tocsc()
    A.sum_duplicates()
    aff = AffineSets(N, m, 0, 0, A, sp.csc_matrix((0, N)), c.copy(), np.zeros(0), cvec)
    return aff, _psd_only_cones(n)


def write_sdpa(path: str, n: int, C_tri, cons, rhs) -> None:
    """Write an SDPA-sparse file in the layout of test/data/mcp*.dat-s.
    C_tri / cons[k]: iterables of (i, j, val), 0-based, i <= j; the file stores the
    un-negated F0 (the loader negates it, base_sdplib.jl:36-38)."""
    with open(path, "w") as fh:
        fh.write(f" {len(rhs)}\n 1\n {n}\n")
        fh.write("{" + ",".join(f"{v:+.17g}" for v in rhs) + "}\n")
        for i, j, v in C_tri:
            fh.write(f"0 1 {i + 1} {j + 1} {v:.17g}\n")
        for kk, tri in enumerate(cons):
            for i, j, v in tri:
                fh.write(f"{kk + 1} 1 {i + 1} {j + 1} {v:.17g}\n")


# --------------------------------------------------------------------------
# Erdős–Rényi Max-Cut (config C2): language-neutral generator
# --------------------------------------------------------------------------
def _splitmix64_stream(seed: int, count: int) -> np.ndarray:
    """Vectorised splitmix64: uniform doubles in [0,1)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def er_graph_edges(n: int, p_edge: float, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Edges (i<j) of G(n, p): pair (i,j) in column-major upper-triangle order is kept when
    the corresponding splitmix64 draw is < p."""
    npairs = n * (n - 1) // 2
    u = _splitmix64_stream(seed, npairs)
    keep = np.nonzero(u < p_edge)[0]
    # invert pair index q = j(j-1)/2 + i  (i<j)
    j = np.floor((1.0 + np.sqrt(1.0 + 8.0 * keep.astype(np.float64))) / 2.0).astype(np.int64)
    j = np.where(j * (j - 1) // 2 > keep, j - 1, j)
    j = np.where((j + 1) * j // 2 <= keep, j + 1, j)
    i = keep - j * (j - 1) // 2
    return i.astype(np.int64), j.astype(np.int64)


def maxcut_er_problem(n: int = 2000, p_edge: float = 0.01, seed: int = 0,
                      sdpa_path: Optional[str] = None) -> Tuple[AffineSets, ConicSets]:
    """Max-Cut SDP relaxation of G(n, p_edge), unit weights, in SDPLIB `mcp` form:
    min -1/4 <L, X>  s.t. diag(X) = 1, X PSD  (L = D - W).  Optionally also written as
    an SDPA-sparse file so that `sdplib_data` can ingest identical bytes."""
    ei, ej = er_graph_edges(n, p_edge, seed)
    deg = np.bincount(ei, minlength=n) + np.bincount(ej, minlength=n)
    N = n * (n + 1) // 2
    c = np.zeros(N)
    # file stores C = L/4; loader negates; JuMP doubles the off-diagonals
    c[svec_index(np.arange(n), np.arange(n))] = -0.25 * deg
    c[svec_index(ei, ej)] = 2.0 * 0.25
    diag_idx = svec_index(np.arange(n), np.arange(n))
    A = sp.csc_matrix((np.ones(n), (np.arange(n), diag_idx)), shape=(n, N))
    aff = AffineSets(N, n, 0, 0, A, sp.csc_matrix((0, N)), np.ones(n), np.zeros(0), c)
    if sdpa_path is not None:
        C_tri = [(i, i, 0.25 * deg[i]) for i in range(n) if deg[i] > 0]
        C_tri += [(int(a), int(b), -0.25) for a, b in zip(ei, ej)]
        write_sdpa(sdpa_path, n, C_tri, [[(k, k, 1.0)] for k in range(n)], np.ones(n))
    return aff, _psd_only_cones(n)


# --------------------------------------------------------------------------
# MIMO detection (config C4)
# --------------------------------------------------------------------------
def mimo_data(seed: int, m: int, n: int):
    """base_mimo.jl:3-17 with numpy's PCG64 in place of MersenneTwister."""
    rng = np.random.default_rng(seed)
    H = rng.standard_normal((m, n))
    v = rng.standard_normal(m)
    s = rng.choice(np.array([-1.0, 1.0]), n)
    sigma = 1e-4
    y = H @ s + sigma * v
    L = np.block([[H.T @ H, -(H.T @ y)[:, None]], [-(y @ H)[None, :], np.array([[y @ y]])]])
    return s, H, y, L


def mimo_problem(seed: int, n: int, var_offset: int = 0, L: Optional[np.ndarray] = None):
    """jump_mimo.jl:1-16: X (n+1)x(n+1) PSD, -1 <= X_ij <= 1, diag(X) = 1, Min <L, X>."""
    if L is None:
        _, _, _, L = mimo_data(seed, 10 * n, n)
    side = n + 1
    N = side * (side + 1) // 2
    c = np.zeros(N)
    for j in range(side):
        for i in range(j + 1):
            c[svec_index(i, j)] = L[i, j] if i == j else L[i, j] + L[j, i]
    diag_idx = svec_index(np.arange(side), np.arange(side))
    A = sp.csc_matrix((np.ones(side), (np.arange(side), diag_idx)), shape=(side, N))
    # X_ij <= 1 rows, then -X_ij <= 1 rows (one pair per lower-triangle entry = per svec var)
    eye = sp.identity(N, format="csc")
    G = sp.vstack([eye, -eye]).tocsc()
    h = np.ones(2 * N)
    aff = AffineSets(N, side, 2 * N, 0, A, G, np.ones(side), h, c)
    return aff, _psd_only_cones(side)


def stack_problems(probs: List[Tuple[AffineSets, ConicSets]]) -> Tuple[AffineSets, ConicSets]:
    """Block-diagonal stack of independent problems into one (config C4 'stacked' form)."""
    n = sum(a.n for a, _ in probs)
    A = sp.block_diag([a.A for a, _ in probs], format="csc")
    G = sp.block_diag([a.G for a, _ in probs], format="csc")
    b = np.concatenate([a.b for a, _ in probs])
    h = np.concatenate([a.h for a, _ in probs])
    c = np.concatenate([a.c for a, _ in probs])
    con = ConicSets()
    off = 0
    for a, k in probs:
        for s in k.sdpcone:
            con.sdpcone.append(SDPSet(s.vec_i + off, s.tri_len, s.sq_side))
        for s in k.socone:
            con.socone.append(SOCSet(s.idx + off, s.len))
        off += a.n
    aff = AffineSets(n, A.shape[0], G.shape[0], 0, sp.csc_matrix(A, shape=(len(b), n)),
                     sp.csc_matrix(G, shape=(len(h), n)), b, h, c)
    return aff, con


# --------------------------------------------------------------------------
# Sensor-network localisation (config C5)
# --------------------------------------------------------------------------
def sensorloc_problem(seed: int, n: int, soc_variant: bool = False):
    """jump_sensorloc.jl:9-53.  With `soc_variant` an extra SOC block (t, u) with
    t = 2 and u = the first min(n,8) sensors' X[1, j+2] coordinates is appended — the
    reference generator has no SOC cone, this gives a mixed-cone workload."""
    rng = np.random.default_rng(seed)
    m = int(np.floor(0.1 * n))
    x_true = rng.random((2, n))
    anchors = rng.random((m, 2))
    side = n + 2
    N = side * (side + 1) // 2
    rows, cols, vals, rhs = [], [], [], []
    r = 0

    def add(row_terms, rh):
        nonlocal r
        for (i, j, v) in row_terms:
            lo, hi = (i, j) if i <= j else (j, i)
            rows.append(r)
            cols.append(svec_index(lo, hi))
            vals.append(v)
        rhs.append(rh)
        r += 1

    for j in range(n):
        for k in range(m):
            a = anchors[k]
            dbar2 = float(np.sum((x_true[:, j] - a) ** 2))
            add([(0, 0, a[0] * a[0]), (1, 1, a[1] * a[1]), (0, j + 2, -2 * a[0]), (1, j + 2, -2 * a[1]),
                 (j + 2, j + 2, 1.0)], dbar2)
    rng2 = np.random.default_rng(seed)
    for i in range(n):
        for j in range(i):
            if rng2.random() > 0.9:
                d2 = float(np.sum((x_true[:, i] - x_true[:, j]) ** 2))
                add([(i + 2, i + 2, 1.0), (j + 2, j + 2, 1.0), (i + 2, j + 2, -2.0)], d2)
    add([(0, 0, 1.0)], 1.0)
    add([(0, 1, 1.0)], 0.0)
    add([(1, 0, 1.0)], 0.0)
    add([(1, 1, 1.0)], 1.0)
    nvar = N
    con = _psd_only_cones(side)
    if soc_variant:
        ns = min(n, 8)
        # new variables: t, u_1..u_ns ; u_q == X[0, q+2] ; t == 2  (norm of ns coordinates in [0,1] is <= sqrt(8) ...)
        t_idx = nvar
        u_idx = np.arange(nvar + 1, nvar + 1 + ns)
        nvar += 1 + ns
        for q in range(ns):
            rows.append(r); cols.append(svec_index(0, q + 2)); vals.append(1.0)
            rows.append(r); cols.append(int(u_idx[q])); vals.append(-1.0)
            rhs.append(0.0); r += 1
        rows.append(r); cols.append(t_idx); vals.append(1.0); rhs.append(3.0); r += 1
        con.socone.append(SOCSet(np.concatenate([[t_idx], u_idx]).astype(np.int64), 1 + ns))
    A = sp.coo_matrix((vals, (rows, cols)), shape=(r, nvar)).tocsc()
    A.sum_duplicates()
    aff = AffineSets(nvar, r, 0, 0, A, sp.csc_matrix((0, nvar)), np.array(rhs), np.zeros(0), np.zeros(nvar))
    return aff, con


# --------------------------------------------------------------------------
# Random SDP (run_mini_benchmark.jl)
# --------------------------------------------------------------------------
def randsdp_problem(seed: int, n: int, m: int):
    """base_randsdp.jl:4-25 + jump_randsdp.jl:1-9."""
    rng = np.random.default_rng(seed)
    c_sqrt = rng.random((n, n))
    C = c_sqrt @ c_sqrt.T
    X_ = rng.standard_normal((n, n))
    X_ = X_ @ X_.T
    N = n * (n + 1) // 2
    iu = [(i, j) for j in range(n) for i in range(j + 1)]
    ii = np.array([t[0] for t in iu])
    jj = np.array([t[1] for t in iu])
    w = np.where(ii == jj, 1.0, 2.0)
    A = np.zeros((m, N))
    b = np.zeros(m)
    for k in range(m):
        Ak = rng.random((n, n))
        Ak = Ak @ Ak.T
        A[k] = Ak[ii, jj] * w
        b[k] = np.trace(Ak @ X_)
    c = C[ii, jj] * w
    aff = AffineSets(N, m, 0, 0, sp.csc_matrix(A), sp.csc_matrix((0, N)), b, np.zeros(0), c)
    return aff, _psd_only_cones(n)


# --------------------------------------------------------------------------
# flat binary dump (exactly the arguments of chambolle_pock)
# --------------------------------------------------------------------------
def save_problem(path: str, aff: AffineSets, con: ConicSets) -> None:
    A = sp.csc_matrix(aff.A)
    G = sp.csc_matrix(aff.G)
    np.savez_compressed(
        path, n=aff.n, p=aff.p, m=aff.m,
        A_colptr=A.indptr.astype(np.int64), A_rowval=A.indices.astype(np.int64), A_nzval=A.data,
        G_colptr=G.indptr.astype(np.int64), G_rowval=G.indices.astype(np.int64), G_nzval=G.data,
        b=aff.b, h=aff.h, c=aff.c,
        sdp_side=np.array([s.sq_side for s in con.sdpcone], dtype=np.int64),
        sdp_idx=np.concatenate([s.vec_i for s in con.sdpcone]) if con.sdpcone else np.zeros(0, np.int64),
        soc_len=np.array([s.len for s in con.socone], dtype=np.int64),
        soc_idx=np.concatenate([s.idx for s in con.socone]) if con.socone else np.zeros(0, np.int64),
    )


def load_problem(path: str) -> Tuple[AffineSets, ConicSets]:
    z = np.load(path)
    n, p, m = int(z["n"]), int(z["p"]), int(z["m"])
    A = sp.csc_matrix((z["A_nzval"], z["A_rowval"], z["A_colptr"]), shape=(p, n))
    G = sp.csc_matrix((z["G_nzval"], z["G_rowval"], z["G_colptr"]), shape=(m, n))
    con = ConicSets()
    off = 0
    for side in z["sdp_side"]:
        tri = int(side) * (int(side) + 1) // 2
        con.sdpcone.append(SDPSet(z["sdp_idx"][off:off + tri].astype(np.int64), tri, int(side)))
        off += tri
    off = 0
    for ln in z["soc_len"]:
        con.socone.append(SOCSet(z["soc_idx"][off:off + int(ln)].astype(np.int64), int(ln)))
        off += int(ln)
    return AffineSets(n, p, m, 0, A, G, z["b"], z["h"], z["c"]), con
