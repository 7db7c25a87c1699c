"""Small driver for compute-sanitizer: one Krylov projection (third-generation Lanczos kernel incl. a thick restart),
one batch of small cones (warm + cold Jacobi), one short README Max-Cut solve."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import README_W, maxcut_problem, mimo_problem, stack_problems

rng = np.random.default_rng(0)
for side, tr in ((160, 3), (300, 8)):
    x = rng.standard_normal(side * (side + 1) // 2)
    xg, cg, mg, cvg, ng, ms = solver.psd_project([side], x, [tr], Options())
    print("psd_project", side, tr, "rank", cg, "converged", cvg, "matvecs", ng, flush=True)
aff, con = stack_problems([mimo_problem(s, 8, 0) for s in range(4)])
r = solver.chambolle_pock(aff, con, Options(max_iter=40))
print("mimo 4 x n=8:", r.status, r.iter, r.objval, flush=True)
aff, con, sgn = maxcut_problem(README_W)
r = solver.chambolle_pock(aff, con, Options(tol_gap=1e-4, tol_feasibility=1e-4))
print("readme maxcut:", r.status, r.iter, sgn * r.objval, flush=True)
