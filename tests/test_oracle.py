"""CPU tests: the oracle against the reference's own known answers and a LAPACK mirror."""
import numpy as np
import pytest

from proxsdp_b200 import MAX_SENSE, MIN_SENSE, Optimizer, Options
from proxsdp_b200.problems import (README_W, load_problem, maxcut_problem, mimo_problem, randsdp_problem,
                                   sensorloc_problem)
from proxsdp_b200.structs import ivec

import problems_ref


def _opt(oracle_mod, **kw):
    return Optimizer(backend=oracle_mod.chambolle_pock, **kw)


# ---------------------------------------------------------------- eigen routines
@pytest.mark.parametrize("n", [1, 2, 3, 10, 65, 150])
def test_oracle_eigh_vs_lapack(oracle_mod, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    A = A + A.T
    w, Z = oracle_mod.eigh(np.triu(A))
    assert np.allclose(w, np.linalg.eigvalsh(A), atol=1e-11 * max(1, n))
    assert np.abs(Z @ np.diag(w) @ Z.T - A).max() < 1e-11 * max(1, n)
    assert np.abs(Z.T @ Z - np.eye(n)).max() < 1e-12 * max(1, n)


@pytest.mark.parametrize("n,rank,nev", [(120, 3, 2), (300, 6, 4), (300, 6, 8)])
def test_oracle_lanczos_vs_lapack(oracle_mod, n, rank, nev):
    rng = np.random.default_rng(7)
    B = rng.standard_normal((n, rank))
    S = rng.standard_normal((n, n))
    A = B @ B.T - 0.1 * np.eye(n) + 0.01 * (S + S.T)
    vals, vecs, info = oracle_mod.lanczos(np.triu(A), oracle_mod.eig_resid(n), nev, max(2 * nev + 1, 25))
    assert info["converged"] >= nev
    w = np.linalg.eigvalsh(A)[::-1]
    assert np.allclose(vals[:nev], w[:nev], atol=1e-10)
    assert np.abs(A @ vecs - vecs * vals).max() < 1e-10


def test_oracle_lanczos_breakdown_zero_matrix(oracle_mod):
    """Iteration 1 of the solver projects the zero matrix: exact breakdown."""
    n = 130
    vals, vecs, info = oracle_mod.lanczos(np.zeros((n, n)), oracle_mod.eig_resid(n), 2, 25)
    assert info["converged"] == 1 and info["numops"] == 1 and len(vals) == 1 and vals[0] == 0.0


def test_oracle_lanczos_restart_cluster(oracle_mod):
    n = 124
    rng = np.random.default_rng(1)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.concatenate([[10.0, 5.0, 5.0 - 1e-5, 5.0 - 2e-5, 4.0], np.linspace(3.0, -8.0, n - 5)])
    A = (Q * lam) @ Q.T
    A = 0.5 * (A + A.T)
    vals, vecs, info = oracle_mod.lanczos(np.triu(A), oracle_mod.eig_resid(n), 2, 25)
    assert info["numiter"] > 1 and info["converged"] >= 2
    assert np.allclose(vals[:2], lam[:2], atol=1e-10)


# ---------------------------------------------------------------- reference known answers
@pytest.mark.parametrize("fn", problems_ref.ALL, ids=lambda f: f.__name__)
def test_reference_unit_problems(oracle_mod, fn):
    problems_ref.check(fn(_opt(oracle_mod)))


@pytest.mark.parametrize("kw", [dict(), dict(eigsolver=1, min_size_krylov_eigs=1), dict(eigsolver=2, min_size_krylov_eigs=1),
                                dict(full_eig_decomp=True)], ids=str)
def test_sdp_wiki_all_eig_paths(oracle_mod, kw):
    """test/moi_proxsdp_unit.jl:302-338 and the re-runs at :358-370."""
    problems_ref.check(problems_ref.sdp_wiki(_opt(oracle_mod, **kw), MIN_SENSE))
    problems_ref.check(problems_ref.sdp_wiki(_opt(oracle_mod, **kw), MAX_SENSE))


def test_readme_maxcut(oracle_mod):
    """README.md:58-84 (config C1)."""
    aff, con, sgn = maxcut_problem(README_W)
    r = oracle_mod.chambolle_pock(aff, con, Options(tol_gap=1e-4, tol_feasibility=1e-4))
    assert r.status == 1
    assert abs(sgn * r.objval - 18.0) < 2e-2
    X = ivec(r.primal)
    assert np.allclose(np.diag(X), 1.0, atol=1e-3)
    assert np.linalg.eigvalsh(X).min() > -1e-6


@pytest.mark.parametrize("n", [2, 3, 4, 5])
def test_mimo_small(oracle_mod, n):
    """test/moi_mimo.jl:71-75: every |X_ij| in (0.99, 1.01)."""
    aff, con = mimo_problem(123 + n, n)
    r = oracle_mod.chambolle_pock(aff, con, Options(tol_gap=1e-6, tol_feasibility=1e-6))
    X = ivec(r.primal)
    assert np.all((np.abs(X) > 0.99) & (np.abs(X) < 1.01))


@pytest.mark.parametrize("name,optimum", [("mcp124-1", -141.9905), ("gpp124-2", 46.8623)])
def test_sdplib_small(oracle_mod, golden_dir, name, optimum):
    """test/moi_sdplib.jl:53-56 (minus_rank == 0) plus the SDPLIB optimal value as an anchor."""
    aff, con = load_problem(f"{golden_dir}/sdplib_{name}.npz")
    r = oracle_mod.chambolle_pock(aff, con, Options(tol_gap=1e-3, tol_feasibility=1e-3))
    X = ivec(r.primal)
    assert (np.linalg.eigvalsh(X) < -1e-4).sum() == 0
    assert r.status == 1
    assert abs(r.objval - optimum) / abs(optimum) < 5e-3
    assert r.lanczos_calls > 0            # the Krylov path ran (n > 100)


def test_sensorloc_runs(oracle_mod):
    """test/moi_sensorloc.jl: runs to a status without error; X11 = X22 = 1."""
    aff, con = sensorloc_problem(0, 10)
    r = oracle_mod.chambolle_pock(aff, con, Options())
    assert r.status in (1, 3)
    X = ivec(r.primal)
    assert abs(X[0, 0] - 1) < 1e-2 and abs(X[1, 1] - 1) < 1e-2 and abs(X[0, 1]) < 1e-2


def test_soc_mixed_cones(oracle_mod):
    aff, con = sensorloc_problem(0, 10, soc_variant=True)
    assert len(con.socone) == 1
    r = oracle_mod.chambolle_pock(aff, con, Options())
    assert r.status == 1
    t, u = r.primal[con.socone[0].idx[0]], r.primal[con.socone[0].idx[1:]]
    assert np.linalg.norm(u) <= t + 1e-4


# ---------------------------------------------------------------- statuses / options
def test_termination_statuses(oracle_mod):
    """test/test_terminationstatus.jl:40-72."""
    o = _opt(oracle_mod)
    problems_ref.sdp_from_moi(o)
    assert o.termination_status() == "OPTIMAL"
    o = _opt(oracle_mod, max_iter=1)
    problems_ref.sdp_wiki(o)
    assert o.termination_status() == "ITERATION_LIMIT"
    o = _opt(oracle_mod, time_limit=0.0)
    problems_ref.sdp_wiki(o)
    assert o.termination_status() == "TIME_LIMIT"


def test_option_errors_and_time_limit_attr():
    """test/moitest.jl:153-171."""
    with pytest.raises(ValueError):
        Options(not_an_option=1)
    o = Optimizer(backend=lambda *a: None)
    assert o.get_time_limit_sec() is None
    o.set_time_limit_sec(0.0)
    assert o.get_time_limit_sec() == 0.0
    o.set_time_limit_sec(None)
    assert o.get_time_limit_sec() is None
    o.set_silent(False)
    assert o.get_attribute("log_verbose") is True


def test_infeasible_and_unbounded_status(oracle_mod):
    """MOI.Test conic infeasible/unbounded cases (test/moitest.jl:34-91) in miniature."""
    o = _opt(oracle_mod)
    x = o.add_variables(1)
    o.add_equal_to([(1.0, x[0])], 1.0)
    o.add_equal_to([(1.0, x[0])], 2.0)
    o.set_objective(MIN_SENSE, [(1.0, x[0])])
    o.optimize()
    assert o.termination_status() == "INFEASIBLE"
    o = _opt(oracle_mod)
    x = o.add_variables(1)
    o.add_greater_than([(1.0, x[0])], 0.0)
    o.set_objective(MAX_SENSE, [(1.0, x[0])])
    o.optimize()
    assert o.termination_status() == "DUAL_INFEASIBLE"


# ---------------------------------------------------------------- oracle pinned by the LAPACK mirror
@pytest.mark.parametrize("which", ["C1", "mimo8", "mcp124-1"])
def test_oracle_vs_numpy_mirror(oracle_mod, golden_dir, which):
    from oracle import oracle_np
    if which == "C1":
        aff, con = maxcut_problem(README_W)[:2]
    elif which == "mimo8":
        aff, con = mimo_problem(1, 8)
    else:
        aff, con = load_problem(f"{golden_dir}/sdplib_mcp124-1.npz")
    opt = Options(full_eig_decomp=True, trace_cap=400, max_iter=400)
    ro = oracle_mod.chambolle_pock(aff, con, opt)
    rn = oracle_np.solve_exact(aff, con, opt, 400)
    assert ro.iter == rn["iter"]
    k = len(rn["trace"])
    a, b = ro.trace[:k, 1:9], rn["trace"][:k, 1:9]
    assert np.abs(a - b).max() <= 1e-6 * max(1.0, np.abs(b).max())
    assert np.abs(ro.primal - rn["primal"]).max() < 1e-8


def test_oracle_matches_golden_trace(oracle_mod, golden_dir):
    """The committed golden trace (tests/golden/make_golden_traces.py) pins the oracle."""
    z = np.load(f"{golden_dir}/trace_mcp124-1_exact.npz")
    aff, con = load_problem(f"{golden_dir}/sdplib_mcp124-1.npz")
    r = oracle_mod.chambolle_pock(aff, con, Options(full_eig_decomp=True, max_iter=int(z["iters"]), trace_cap=int(z["iters"])))
    assert np.abs(r.trace[:, 1:9] - z["trace"][:, 1:9]).max() <= 1e-7 * max(1.0, np.abs(z["trace"][:, 1:9]).max())


def test_randsdp_mini_benchmark(oracle_mod):
    """test/run_mini_benchmark.jl:37-40 (randsdp 10x10)."""
    aff, con = randsdp_problem(0, 10, 10)
    r = oracle_mod.chambolle_pock(aff, con, Options(max_iter=20000))
    assert r.status in (1, 3)
    X = ivec(r.primal)
    assert np.linalg.eigvalsh(X).min() > -1e-5


def test_step_seams_against_numpy_restatement(oracle_mod):
    """The oracle's linesearch! / compute_residual! / compute_gap! seams against a direct numpy restatement of
    reference src/pdhg.jl:532-582 and src/residuals.jl:2-71."""
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    n, p, m = 40, 7, 9
    A = sp.random(p, n, 0.3, random_state=1, format="csc")
    G = sp.random(m, n, 0.3, random_state=2, format="csc")
    M = sp.vstack([A, G]).tocsc()
    st = dict(b=rng.standard_normal(p), h=rng.standard_normal(m), c=rng.standard_normal(n), x=rng.standard_normal(n),
              x_old=rng.standard_normal(n), y=rng.standard_normal(p + m), y_old=rng.standard_normal(p + m),
              Mx=rng.standard_normal(p + m), Mx_old=rng.standard_normal(p + m), Mty=rng.standard_normal(n),
              Mty_old=rng.standard_normal(n), primal_step=0.3, primal_step_old=0.25, dual_step=0.2, theta=1.0, beta=0.8,
              norm_b=1.5, norm_h=0.7, norm_c=2.0)
    # residuals.jl:37-71
    r = oracle_mod.residuals(n, p, m, Options(), **st)
    tau, sig = st["primal_step"], st["dual_step"]
    pold = st["x_old"] - tau * st["Mty_old"]
    pr = np.sqrt(n) * np.abs((st["x"] - tau * st["Mty"]) - pold).max() / max(np.abs(pold).max(), st["norm_b"], st["norm_h"], 1.0)
    dold = st["y_old"] - sig * st["Mx_old"]
    dr = np.sqrt(p + m) * np.abs((st["y"] - sig * st["Mx"]) - dold).max() / max(np.abs(dold).max(), st["norm_c"], 1.0)
    assert abs(r["primal_residual"] - pr) <= 1e-13 * pr and abs(r["dual_residual"] - dr) <= 1e-13 * dr
    # residuals.jl:2-35
    eq = np.abs(st["Mx"][:p] - st["b"]).max() / (1 + st["norm_b"])
    ineq = max(0.0, (st["Mx"][p:] - st["h"]).max()) / (1 + st["norm_h"])
    po = st["c"] @ st["x"]
    do = -(st["b"] @ st["y"][:p]) - (st["h"] @ st["y"][p:])
    assert abs(r["equa_feasibility"] - eq) <= 1e-14 and abs(r["ineq_feasibility"] - ineq) <= 1e-14
    assert abs(r["prim_obj"] - po) <= 1e-12 and abs(r["dual_obj"] - do) <= 1e-12
    assert abs(r["dual_gap"] - abs(po - do) / (1 + abs(po) + abs(do))) <= 1e-14
    # pdhg.jl:532-582
    opt = Options()
    yn, Mn, sc, trials = oracle_mod.dual_step(A, G, n, p, m, opt, **st)
    t = st["primal_step"] * np.sqrt(1 + st["theta"])
    for it in range(int(opt.max_linsearch_steps)):
        theta = t / st["primal_step_old"]
        bt = st["beta"] * t
        yh = st["y"] + bt * ((1 + theta) * st["Mx"] - theta * st["Mx_old"])
        proj = np.concatenate([st["b"], np.minimum(yh[p:] / bt, st["h"])])
        yt = yh - bt * proj
        Mt = M.T @ yt
        if np.sqrt(st["beta"]) * t * np.linalg.norm(Mt - st["Mty"]) <= opt.delta * np.linalg.norm(yt - st["y"]):
            break
        t *= opt.linsearch_decay
    assert trials == it + 1 and abs(sc["primal_step"] - t) <= 1e-15 and abs(sc["theta"] - theta) <= 1e-15
    assert np.abs(yn - yt).max() <= 1e-13 and np.abs(Mn - Mt).max() <= 1e-13


@pytest.mark.parametrize("which", ["C1", "mimo8", "sdp_badly_scaled"])
def test_equilibration_vs_numpy_mirror(oracle_mod, which):
    """equilibrate! (reference src/equilibration.jl:1-71) and its use in setup / result assembly (src/pdhg.jl:64-93,
    751-755): the C oracle against the numpy mirror, exact-projection mode, forced equilibration."""
    from oracle import oracle_np
    if which == "C1":
        aff, con = maxcut_problem(README_W)[:2]
    elif which == "mimo8":
        aff, con = mimo_problem(1, 8)
    else:
        # rows of very different magnitude: the case the preconditioner is for
        aff, con = mimo_problem(2, 6)
        scale = np.logspace(-2, 2, aff.p)
        import scipy.sparse as sp
        aff.A = sp.csc_matrix(sp.diags(scale) @ aff.A)
        aff.b = scale * aff.b
    opt = Options(full_eig_decomp=True, trace_cap=300, max_iter=300, equilibration_force=True, equilibration_iters=200)
    ro = oracle_mod.chambolle_pock(aff, con, opt)
    rn = oracle_np.solve_exact(aff, con, opt, 300)
    plain = oracle_mod.chambolle_pock(aff, con, Options(full_eig_decomp=True, trace_cap=300, max_iter=300))
    assert ro.iter == rn["iter"]
    k = len(rn["trace"])
    a, b = ro.trace[:k, 1:9], rn["trace"][:k, 1:9]
    assert np.abs(a - b).max() <= 1e-6 * max(1.0, np.abs(b).max())
    assert np.abs(ro.primal - rn["primal"]).max() < 1e-8 * max(1.0, np.abs(rn["primal"]).max())
    assert np.abs(np.concatenate([ro.dual_eq, ro.dual_in]) - rn["y"]).max() < 1e-8 * max(1.0, np.abs(rn["y"]).max())
    # the preconditioner changed the trajectory (the test would be vacuous otherwise)
    kk = min(len(plain.trace), k)
    assert np.abs(plain.trace[:kk, 1:9] - ro.trace[:kk, 1:9]).max() > 1e-6


def test_equilibration_switches_itself_off(oracle_mod):
    """pdhg.jl:66-73: with min(M) / max(M) <= equilibration_limit (any sparse M: min = 0) the option is dropped."""
    aff, con = mimo_problem(1, 6)
    a = oracle_mod.chambolle_pock(aff, con, Options(equilibration=True, max_iter=50, trace_cap=50))
    b = oracle_mod.chambolle_pock(aff, con, Options(max_iter=50, trace_cap=50))
    assert np.array_equal(a.trace[:, 1:9], b.trace[:, 1:9]) and np.array_equal(a.primal, b.primal)


@pytest.mark.parametrize("which", ["C1", "mimo8", "sensorloc"])
def test_exact_spectral_norm_vs_arpack_svds(oracle_mod, which):
    """approx_norm = false (reference src/pdhg.jl:107-118): the step sizes start from 1 / sigma_max(M) — the numpy mirror
    calls ARPACK svds like the reference does; exact-projection mode."""
    from oracle import oracle_np
    if which == "C1":
        aff, con = maxcut_problem(README_W)[:2]
    elif which == "mimo8":
        aff, con = mimo_problem(1, 8)
    else:
        aff, con = sensorloc_problem(0, 10)
    opt = Options(full_eig_decomp=True, trace_cap=200, max_iter=200, approx_norm=False)
    ro = oracle_mod.chambolle_pock(aff, con, opt)
    rn = oracle_np.solve_exact(aff, con, opt, 200)
    assert ro.iter == rn["iter"]
    k = len(rn["trace"])
    a, b = ro.trace[:k, 1:9], rn["trace"][:k, 1:9]
    assert np.abs(a - b).max() <= 1e-6 * max(1.0, np.abs(b).max())
    assert np.abs(ro.primal - rn["primal"]).max() < 1e-8 * max(1.0, np.abs(rn["primal"]).max())
    plain = oracle_mod.chambolle_pock(aff, con, Options(full_eig_decomp=True, trace_cap=200, max_iter=200))
    if which != "C1":        # C1: M = 4 unit rows, sigma_max = 1 but ||M||_F = 2
        assert plain.trace[0, 7] != ro.trace[0, 7]


@pytest.mark.parametrize("n,rank,nev", [(150, 3, 5), (260, 6, 2), (400, 9, 11)])
def test_krylovkit_eager_schedule_oracle(oracle_mod, n, rank, nev):
    """`krylovkit_eager` (reference src/eigsolver.jl:809): never more mat-vecs than the plain schedule, the same projection
    (both stop at tol = 1e-12), and at most as many converged pairs come back."""
    rng = np.random.default_rng(n)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.concatenate([np.linspace(40.0, 8.0, rank), -np.abs(rng.standard_normal(n - rank)) * 3.0 - 0.05])
    A = (Q * lam) @ Q.T
    A = 0.5 * (A + A.T)
    ii, jj = np.triu_indices(n)
    order = np.lexsort((ii, jj))
    ii, jj = ii[order], jj[order]
    x = np.where(ii != jj, A[ii, jj] * np.sqrt(2.0), A[ii, jj])
    x0, c0, m0, cv0, n0 = oracle_mod.psd_project([n], x, [nev], Options())
    x1, c1, m1, cv1, n1 = oracle_mod.psd_project([n], x, [nev], Options(krylovkit_eager=True))
    assert n1 <= n0 and cv1[0] <= cv0[0] and cv1[0] >= min(nev, cv0[0]) and list(c0) == list(c1)
    assert np.abs(x0 - x1).max() <= 1e-9 * np.abs(x0).max()


def test_oracle_vs_numpy_mirror_permuted_mixed_cones(oracle_mod):
    """preprocess! with a real permutation (reference src/scaling.jl:2-26: [PSD | SOC | free] order, undone for the
    result at src/pdhg.jl:768-769) on a problem with a PSD cone, an SOC cone and free variables: the C oracle against the
    numpy mirror, and against its own solve of the un-permuted problem."""
    from oracle import oracle_np
    from test_gpu_parity_full import _scrambled
    aff, con = sensorloc_problem(1, 8, soc_variant=True)
    aff2, con2, new_of_old = _scrambled(aff, con, 5)
    opt = Options(full_eig_decomp=True, trace_cap=300, max_iter=300)
    ro = oracle_mod.chambolle_pock(aff2, con2, opt)
    rn = oracle_np.solve_exact(aff2, con2, opt, 300)
    r1 = oracle_mod.chambolle_pock(aff, con, opt)
    assert ro.iter == rn["iter"] == r1.iter
    k = len(rn["trace"])
    assert np.abs(ro.trace[:k, 1:9] - rn["trace"][:k, 1:9]).max() <= 1e-6 * max(1.0, np.abs(rn["trace"][:k, 1:9]).max())
    assert np.abs(ro.primal - rn["primal"]).max() < 1e-8
    assert np.abs(ro.primal[new_of_old] - r1.primal).max() < 1e-9
    assert np.abs(ro.dual_cone[new_of_old] - r1.dual_cone).max() < 1e-9
