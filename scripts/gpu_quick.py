import sys, time, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
from oracle import oracle
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import maxcut_problem, README_W, sdplib_problem, maxcut_er_problem, mimo_problem
print("devices", solver.device_count(), flush=True)
rng = np.random.default_rng(0)
for n in (5, 65, 130, 300):
    A = rng.standard_normal((n, n)); A = A + A.T
    t=time.time(); w, Z = solver.eigh(A); dt=time.time()-t
    w0 = np.linalg.eigvalsh(A)
    print("eigh", n, "val err", np.abs(w - w0).max(), "recon", np.abs(Z @ np.diag(w) @ Z.T - A).max(), "orth", np.abs(Z.T @ Z - np.eye(n)).max(), "t %.3f"%dt, flush=True)
for n, r, nev in ((150, 4, 2), (300, 6, 4), (2000, 8, 6)):
    B = rng.standard_normal((n, r)); A = B @ B.T - 0.1 * np.eye(n) + 0.01 * (lambda S: S + S.T)(rng.standard_normal((n, n)))
    x0 = oracle.eig_resid(n)
    v1, V1, i1 = oracle.lanczos(np.triu(A), x0, nev, 25)
    v2, V2, i2 = solver.lanczos(A, x0, nev, 25, repeat=3)
    print("lanczos", n, i1, i2, "vals diff", np.abs(v1[:nev]-v2[:nev]).max(), "resid", np.abs(A @ V2 - V2 * v2).max(), flush=True)
# psd project
for n, tr in ((40, 2), (150, 3), (600, 5)):
    N = n*(n+1)//2
    x = rng.standard_normal(N)
    for mode in (0, 1):
        xo, co, mo, cvo, no = oracle.psd_project([n], x, [tr], Options(), mode=mode)
        xg, cg, mg, cvg, ng, ms = solver.psd_project([n], x, [tr], Options(), mode=mode)
        print("psd", n, "mode", mode, "diff", np.abs(xo-xg).max(), co, cg, mo, mg, cvo, cvg, no, ng, "ms %.3f"%ms, flush=True)
aff, con, sgn = maxcut_problem(README_W)
ro = oracle.chambolle_pock(aff, con, Options())
rg = solver.chambolle_pock(aff, con, Options())
print("C1 oracle", ro.status, ro.objval, ro.iter, "gpu", rg.status, rg.objval, rg.iter, np.abs(ro.primal-rg.primal).max(), flush=True)
D = os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests/golden/")
from proxsdp_b200.problems import load_problem
aff, con = load_problem(D+"sdplib_mcp124-1.npz")
for kw in (dict(), dict(full_eig_decomp=True, max_iter=300)):
    t=time.time(); ro = oracle.chambolle_pock(aff, con, Options(**kw)); to=time.time()-t
    t=time.time(); rg = solver.chambolle_pock(aff, con, Options(**kw)); tg=time.time()-t
    print("mcp124", kw, "oracle", ro.status, ro.objval, ro.iter, "%.2fs"%to, "gpu", rg.status, rg.objval, rg.iter, "%.2fs"%tg, "launches", rg.gpu_launches, "mv", ro.lanczos_matvecs, rg.lanczos_matvecs, "ls", ro.linesearch_trials, rg.linesearch_trials, flush=True)
aff, con = mimo_problem(0, 16)
ro = oracle.chambolle_pock(aff, con, Options()); rg = solver.chambolle_pock(aff, con, Options())
print("mimo16 oracle", ro.status, ro.objval, ro.iter, "gpu", rg.status, rg.objval, rg.iter, np.abs(ro.primal-rg.primal).max(), flush=True)
aff, con = maxcut_er_problem(2000, 0.01, 0)
t=time.time(); rg = solver.chambolle_pock(aff, con, Options(max_iter=300)); tg=time.time()-t
print("C2 300 it gpu", rg.status, rg.objval, rg.dual_objval, "t %.2f"%tg, "loop %.3f"%rg.time_loop, "psd %.3f"%rg.time_psd_proj, "mv", rg.lanczos_matvecs, "ls", rg.linesearch_trials, "launches", rg.gpu_launches, flush=True)
print("expected oracle: -9721.954448502762 33806.55870589523 mv 7546 ls 653")
