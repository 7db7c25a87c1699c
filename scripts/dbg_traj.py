"""Where do two settings of an environment switch send a full solve apart?  usage: dbg_traj.py PROBLEM VAR VALUE_A VALUE_B"""
import os, sys
import numpy as np
ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import load_problem
name, var, va, vb = sys.argv[1:5]
aff, con = load_problem(os.path.join(ROOT, "tests", "golden", f"sdplib_{name}.npz"))
out = {}
for v in (va, vb):
    os.environ[var] = v
    r = solver.chambolle_pock(aff, con, Options(trace_cap=30000))
    out[v] = r
    mv = r.trace[:, 12]
    print(f"{var}={v}: status {r.status} iters {r.iter} obj {r.objval:.6f} time_loop {r.time_loop:.2f} s matvecs {int(r.lanczos_matvecs)} full eigs {r.full_eig_calls}", flush=True)
a, b = out[va].trace, out[vb].trace
k = min(len(a), len(b))
same = np.abs(a[:k, 1:7] - b[:k, 1:7]) <= 1e-9 * np.maximum(1.0, np.abs(b[:k, 1:7]))
first = int(np.argmin(same.all(axis=1))) if not same.all() else k
print("traces agree to 1e-9 for the first", first, "iterations")
lo, hi = max(0, first - 3), first + 8
print("cols: iter, prim_obj, dual_obj, gap, feas, res_p, res_d, tau, beta, rank_target, rank_cur, min_eig, matvecs, ls")
for t in (a, b):
    for row in t[lo:hi]:
        print("  ", " ".join(f"{x:.6g}" for x in row))
    print()
