"""CPU tests that pin the oracle from OUTSIDE (no golden file made by the oracle itself, no number quoted from memory):

* the truncated projection against ARPACK's dsaupd / dseupd (scipy `eigsh`), i.e. the library the reference itself
  calls for `eigsolver = 1` (reference src/eigsolver.jl:668-770, parameters of `arpack_init!` :430-483), and against
  LAPACK dsyevr (`scipy.linalg.eigh(driver="evr")`, the routine behind `eigen!`, reference src/prox_operators.jl:113);
* full solves against a weak-duality certificate recomputed in numpy from the returned primal and dual points: for
  min c'x, Ax = b, X PSD any y with S = C + A'y PSD gives c'x >= -b'y for every feasible x, so a returned pair that
  is feasible / dual feasible (checked here with LAPACK, not by the solver) and has a small gap brackets the optimum.

KrylovKit itself is not on this machine: what can be pinned externally is that the oracle's Krylov path
returns what ARPACK and LAPACK return on the same matrix, and that its full solves end at certified optima.
"""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse.linalg as spla

from proxsdp_b200 import Options
from proxsdp_b200.problems import load_problem, maxcut_er_problem, mimo_problem
from proxsdp_b200.structs import ivec

from certificates import certificate as _certificate

SQ2 = np.sqrt(2.0)


def _svec_scaled(X):
    """Column-major upper triangle with the off-diagonal entries multiplied by sqrt(2) (reference src/scaling.jl:40-58)."""
    n = X.shape[0]
    ii, jj = np.triu_indices(n)
    order = np.lexsort((ii, jj))
    ii, jj = ii[order], jj[order]
    return np.where(ii != jj, X[ii, jj] * SQ2, X[ii, jj])


def _smat_scaled(x):
    X = ivec(x)
    d = np.diag(X).copy()
    X = X / SQ2
    np.fill_diagonal(X, d)
    return X


def _iterate_like_matrix(n, rank, seed):
    """What the projection sees along a solve: a few dominant positive eigenvalues and a negative bulk."""
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.concatenate([np.linspace(40.0, 8.0, rank), -np.abs(rng.standard_normal(n - rank)) * 3.0 - 0.05])
    A = (Q * lam) @ Q.T
    return 0.5 * (A + A.T)


@pytest.mark.parametrize("n,rank,seed", [(150, 3, 0), (260, 6, 1), (400, 9, 2)])
@pytest.mark.parametrize("eigsolver", [1, 2])
def test_truncated_projection_vs_arpack_and_lapack(oracle_mod, n, rank, seed, eigsolver):
    A = _iterate_like_matrix(n, rank, seed)
    # target rank above the true positive rank, as convergedrank() leaves it (reference src/pdhg.jl:486-506)
    nev = rank + 2
    opt = Options(eigsolver=eigsolver)
    xo, cur, mineig, conv, nops = oracle_mod.psd_project([n], _svec_scaled(A), [nev], opt)
    assert conv[0] >= 1 and cur[0] == rank and nops > 0

    # ARPACK with the reference's parameters: which = "LA", ncv = max(2 nev + 1, eigsolver_min_lanczos), tol = arpack_tol
    ncv = max(2 * nev + 1, int(opt.eigsolver_min_lanczos))
    w, V = spla.eigsh(A, k=nev, which="LA", ncv=ncv, tol=float(opt.arpack_tol), maxiter=int(opt.arpack_max_iter),
                      v0=oracle_mod.eig_resid(n))
    keep = w > 0
    X_arpack = (V[:, keep] * w[keep]) @ V[:, keep].T
    # LAPACK dsyevr
    wl, Vl = sla.eigh(A, driver="evr")
    keep_l = wl > 0
    X_lapack = (Vl[:, keep_l] * wl[keep_l]) @ Vl[:, keep_l].T

    Xo = _smat_scaled(xo)
    scale = np.abs(X_lapack).max()
    assert np.abs(Xo - X_lapack).max() <= 1e-9 * scale
    assert np.abs(Xo - X_arpack).max() <= 1e-8 * scale       # ARPACK stops at arpack_tol = 1e-10 relative
    # min_eig of the computed block = the smallest of the nev Ritz values (reference src/prox_operators.jl:92-99)
    assert abs(mineig[0] - np.sort(wl)[::-1][nev - 1]) <= 1e-8 * scale


def test_krylov_eigenpairs_vs_arpack(oracle_mod):
    """The eigsolve alone: the leading pairs KrylovKit's restated Lanczos returns are ARPACK's."""
    n, nev = 300, 5
    A = _iterate_like_matrix(n, 7, 11)
    x0 = oracle_mod.eig_resid(n)
    vals, vecs, info = oracle_mod.lanczos(np.triu(A), x0, nev, 25)
    assert info["converged"] >= nev
    w, V = spla.eigsh(A, k=nev, which="LA", ncv=25, tol=1e-12, v0=x0)
    order = np.argsort(w)[::-1]
    w, V = w[order], V[:, order]
    assert np.abs(vals[:nev] - w).max() <= 1e-9 * abs(w[0])
    # same invariant subspace, vector by vector (the spectrum is simple): |<v_oracle, v_arpack>| = 1
    cosines = np.abs(np.sum(vecs[:, :nev] * V, axis=0))
    assert np.abs(cosines - 1.0).max() <= 1e-9


@pytest.mark.parametrize("name,n", [("mcp124-1", 124)])      # mcp250-1: GPU suite (155 s on the CPU)
def test_maxcut_full_solve_certified_optimum(oracle_mod, golden_dir, name, n):
    """Max-Cut relaxations (diag(X) = 1, so trace(X) = n is known): with lam = lambda_min of the dual slack matrix,
    every feasible X has <C, X> >= dual objective + lam * n.  The returned objective must sit in that bracket, which
    pins the optimum of the full Krylov-path solve without trusting the solver's own residuals."""
    aff, con = load_problem(f"{golden_dir}/sdplib_{name}.npz")
    r = oracle_mod.chambolle_pock(aff, con, Options(tol_gap=1e-5, tol_feasibility=1e-5))
    assert r.status == 1 and r.lanczos_calls > 0
    k = _certificate(aff, con, r)
    assert k["eq"] <= 2e-5 * (1.0 + np.sqrt(n)) and k["lam_x"] >= -1e-6
    assert abs(k["trace"] - n) <= 1e-2
    lower = k["dual"] + min(k["lam_s"], 0.0) * n            # certified lower bound on the optimum
    assert lower <= k["primal"] + 1e-5 * abs(k["primal"])   # X is feasible to 1e-5 only, so it may undercut the bound by that
    assert k["primal"] - lower <= 2e-3 * abs(k["primal"]), k
    assert abs(r.objval - k["primal"]) <= 1e-9 * abs(k["primal"])
    assert abs(r.dual_objval - k["dual"]) <= 1e-6 * abs(k["dual"])


def test_er_maxcut_certified_optimum(oracle_mod):
    """The headline family (Erdos-Renyi Max-Cut, BASELINE.json C2) at a side the CPU finishes in seconds."""
    n = 300
    aff, con = maxcut_er_problem(n, 0.05, seed=0)
    r = oracle_mod.chambolle_pock(aff, con, Options(tol_gap=1e-5, tol_feasibility=1e-5))
    assert r.status == 1 and r.lanczos_calls > 0
    k = _certificate(aff, con, r)
    assert k["eq"] <= 1e-4 and k["lam_x"] >= -1e-6
    lower = k["dual"] + min(k["lam_s"], 0.0) * n
    assert lower <= k["primal"] + 1e-5 * abs(k["primal"])
    assert k["primal"] - lower <= 2e-3 * abs(k["primal"]), k


def test_mimo_certified_optimum(oracle_mod):
    """A problem with inequality rows (reference test/base_mimo.jl:19-60): -1 <= X_ij <= 1 next to diag(X) = 1."""
    aff, con = mimo_problem(5, 16)
    r = oracle_mod.chambolle_pock(aff, con, Options(tol_gap=1e-6, tol_feasibility=1e-6))
    assert r.status == 1
    k = _certificate(aff, con, r)
    side = con.sdpcone[0].sq_side
    assert k["eq"] <= 1e-5 and k["ineq"] <= 1e-5 and k["lam_x"] >= -1e-6 and k["y_in_min"] >= -1e-9
    assert abs(k["trace"] - side) <= 1e-3
    lower = k["dual"] + min(k["lam_s"], 0.0) * side
    scale = 1.0 + abs(k["primal"])
    assert lower <= k["primal"] + 1e-5 * scale
    assert k["primal"] - lower <= 1e-3 * scale, k


def test_soc_projection_vs_closed_form(oracle_mod):
    """soc_projection! (reference src/prox_operators.jl:138-158) against the textbook projection onto {(t, v): ||v|| <= t}."""
    rng = np.random.default_rng(3)
    lens = [1, 2, 3, 7, 50, 300, 4, 4, 4]
    x = rng.standard_normal(sum(lens)) * 3.0
    # force the three branches on the last three cones: inside, polar, outside
    off = sum(lens[:-3])
    x[off:off + 4] = [5.0, 1.0, 1.0, 1.0]
    x[off + 4:off + 8] = [-5.0, 1.0, 1.0, 1.0]
    x[off + 8:off + 12] = [0.5, 2.0, -1.0, 0.5]
    out = oracle_mod.soc_project(lens, x)
    o = 0
    for ln in lens:
        t, v = x[o], x[o + 1:o + ln]
        nv = np.linalg.norm(v)
        if nv <= -t:
            want = np.zeros(ln)
        elif nv <= t:
            want = x[o:o + ln]
        else:
            a = 0.5 * (1.0 + t / nv)
            want = np.concatenate([[a * nv], a * v])
        assert np.abs(out[o:o + ln] - want).max() <= 1e-14 * max(1.0, np.abs(want).max())
        o += ln


@pytest.mark.parametrize("sides", [[1, 1], [2, 3, 7], [40, 65], [130]])
def test_exact_projection_vs_lapack(oracle_mod, sides):
    """full_eig! (reference src/prox_operators.jl:111-126) on ragged batches: X+ = V max(lam, 0) V' from LAPACK dsyevr,
    current_rank = #{lam > tol_psd}, min_eig = 0."""
    rng = np.random.default_rng(sum(sides))
    xs, wants, ranks = [], [], []
    for n in sides:
        B = rng.standard_normal((n, n))
        A = B + B.T
        xs.append(_svec_scaled(A))
        w, V = sla.eigh(A, driver="evr")
        keep = w > 0
        wants.append(_svec_scaled((V[:, keep] * w[keep]) @ V[:, keep].T))
        ranks.append(int((w > Options().tol_psd).sum()))
    x = np.concatenate(xs)
    out, cur, mineig, conv, nops = oracle_mod.psd_project(sides, x, [2] * len(sides), Options(full_eig_decomp=True))
    want = np.concatenate(wants)
    assert np.abs(out - want).max() <= 1e-11 * max(1.0, np.abs(want).max())
    assert [int(c) for c in cur] == [r if n > 1 else int(c) for r, n, c in zip(ranks, sides, cur)]
    assert nops == 0
