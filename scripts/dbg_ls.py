import os, sys
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from oracle import oracle
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import mimo_problem
np.set_printoptions(linewidth=250, precision=6)
aff, con = mimo_problem(7, 16)
for kw in (dict(max_linsearch_steps=1), dict(linsearch_decay=0.97, max_linsearch_steps=3)):
    opt = Options(trace_cap=5000, **kw)
    rg = solver.chambolle_pock(aff, con, opt)
    ro = oracle.chambolle_pock(aff, con, opt)
    print(kw, "GPU:", rg.status, rg.status_string, rg.iter, "| oracle:", ro.status, ro.status_string, ro.iter)
    k = min(len(rg.trace), len(ro.trace))
    d = np.abs(rg.trace[:k, 1:9] - ro.trace[:k, 1:9]) / np.maximum(1.0, np.abs(ro.trace[:k, 1:9]))
    bad = np.nonzero(d.max(axis=1) > 1e-6)[0]
    print(" first differing iteration:", bad[:3], " max rel diff:", d.max())
    i0 = max(0, (bad[0] if len(bad) else k) - 2)
    print(" gpu   ", rg.trace[i0:i0 + 4, [0, 1, 2, 3, 4, 5, 6, 7, 8, 13]])
    print(" oracle", ro.trace[i0:i0 + 4, [0, 1, 2, 3, 4, 5, 6, 7, 8, 13]])
    print(" last gpu rows", rg.trace[-2:, [0, 1, 2, 3, 4, 5, 6, 7, 8, 13]])
