import sys, time, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
from oracle import oracle
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import load_problem
D = os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests/golden/")
aff, con = load_problem(D+"sdplib_mcp124-1.npz")
opt = Options(trace_cap=4000)
ro = oracle.chambolle_pock(aff, con, opt)
rg = solver.chambolle_pock(aff, con, opt)
to, tg = ro.trace, rg.trace
n = min(len(to), len(tg))
print("iters", ro.iter, rg.iter, "mv", ro.lanczos_matvecs, rg.lanczos_matvecs, "final tr", ro.target_rank, rg.target_rank)
rel = np.abs(to[:n, 1:] - tg[:n, 1:]) / (1e-300 + np.abs(to[:n, 1:]))
for k in (0, 1, 2, 5, 10, 50, 100, 200, 201, 202, 250, 300, 400, 500, 800, 1000, 1500, 2000, 2500):
    if k < n:
        print(k + 1, "prim", to[k, 1], tg[k, 1], "gap", to[k, 3], tg[k, 3], "tr", to[k, 9], tg[k, 9], "cr", to[k, 10], tg[k, 10], "mineig", to[k, 11], tg[k, 11], "maxrel", rel[k, :7].max())
bad = np.nonzero(rel[:, :7].max(axis=1) > 1e-6)[0]
print("first iter with rel diff > 1e-6:", bad[:5] + 1)
bad = np.nonzero(to[:n, 9] != tg[:n, 9])[0]
print("first iter with different target rank:", bad[:5] + 1)
print("matvecs per iter oracle", to[:60, 12].astype(int))
print("matvecs per iter gpu   ", tg[:60, 12].astype(int))
print("ls oracle", to[:60, 13].astype(int))
print("ls gpu   ", tg[:60, 13].astype(int))
print("mean mv oracle/gpu", to[:,12].mean(), tg[:,12].mean(), "hist gpu", np.bincount(tg[:,12].astype(int))[:200].nonzero()[0])
np.set_printoptions(linewidth=250)
for a in range(60, 420, 60):
    print(a, "oracle", to[a:a+60, 12].astype(int))
    print(a, "gpu   ", tg[a:a+60, 12].astype(int))
print("oracle mean mv by 500-blocks", [round(to[i:i+500,12].mean(),1) for i in range(0, len(to), 500)])
print("gpu    mean mv by 500-blocks", [round(tg[i:i+500,12].mean(),1) for i in range(0, len(tg), 500)])
print("---- tails: iter, prim, gap, feas, tr, cr, mineig, mv")
for name, t in (("oracle", to), ("gpu", tg)):
    print(name)
    for k in range(max(0, len(t) - 260), len(t), 12):
        print("  ", int(t[k, 0]), "%.6f" % t[k, 1], "%.3e" % t[k, 3], "%.3e" % t[k, 4], int(t[k, 9]), int(t[k, 10]), "%.3e" % t[k, 11], int(t[k, 12]))
