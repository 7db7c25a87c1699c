"""Reduced driver for `compute-sanitizer --tool racecheck` (shared-memory hazards; ~100x slower than a plain run on the
spin-waiting cluster kernel): one Krylov projection of a side-128 cone (sides <= 100 = min_size_krylov_eigs take the full
eigendecomposition) through the single-cluster resident variant of k_lanczos_cl3 and one through the grid-wide variant
(TMA staging, DSMEM exchanges, flagged grid exchange, Ritz bisection), one exact projection of a side-130 cone (block-Jacobi), a batch of
four small cones, and 12 iterations of the README Max-Cut solve (fused line-search ladder, residual kernels, record)."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import README_W, maxcut_problem, mimo_problem, stack_problems

rng = np.random.default_rng(0)
x = rng.standard_normal(128 * 129 // 2)
xg, cg, mg, cvg, ng, ms = solver.psd_project([128], x, [2], Options())
print("krylov projection side 128, resident single-cluster kernel: rank", cg, "converged", cvg, "matvecs", ng, flush=True)
os.environ["PROXSDP_B200_LZ_RESIDENT"] = "0"
xg, cg, mg, cvg, ng, ms = solver.psd_project([128], x, [2], Options())
print("krylov projection side 128, grid-wide kernel: rank", cg, "converged", cvg, "matvecs", ng, flush=True)
os.environ.pop("PROXSDP_B200_LZ_RESIDENT")
x = rng.standard_normal(130 * 131 // 2)
xg, cg, mg, cvg, ng, ms = solver.psd_project([130], x, [2], Options(full_eig_decomp=True))
print("exact projection side 130: rank", cg, flush=True)
aff, con = stack_problems([mimo_problem(s, 6, 0) for s in range(4)])
r = solver.chambolle_pock(aff, con, Options(max_iter=10))
print("mimo 4 x n=6:", r.status, r.iter, flush=True)
aff, con, sgn = maxcut_problem(README_W)
r = solver.chambolle_pock(aff, con, Options(max_iter=12))
print("readme maxcut, 12 iterations:", r.status, r.iter, sgn * r.objval, flush=True)
