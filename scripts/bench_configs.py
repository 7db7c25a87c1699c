"""Secondary measurements for BASELINE.json configs C1, C3, C4 (1 GPU), C5 — parity-test workloads, timed for
DESIGN.md (the bench.py line is C2 only).  Prints one JSON object per config."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from proxsdp_b200 import Options, solver  # noqa: E402
from proxsdp_b200.problems import (README_W, load_problem, maxcut_problem, mimo_problem, sensorloc_problem,  # noqa: E402
                                   stack_problems)

G = os.path.join(ROOT, "tests", "golden")
CPU = "--no-cpu" not in sys.argv


def emit(**kw):
    print(json.dumps(kw), flush=True)


# C1
aff, con, sgn = maxcut_problem(README_W)
opt = Options(tol_gap=1e-4, tol_feasibility=1e-4)
rg = solver.chambolle_pock(aff, con, opt); ro = oracle.chambolle_pock(aff, con, opt)
emit(config="C1 README max-cut n=4", gpu_iter=rg.iter, cpu_iter=ro.iter, obj_gpu=sgn * rg.objval, obj_cpu=sgn * ro.objval,
     gpu_loop_s=rg.time_loop, cpu_loop_s=ro.time_loop)

# C3: rank sweep, fixed target rank for a 200-iteration window
for name in ("mcp500-1", "gpp500-1"):
    aff, con = load_problem(os.path.join(G, f"sdplib_{name}.npz"))
    for r in (5, 10, 16, 25, 50):
        opt = Options(max_iter=200, initial_target_rank=r, freeze_target_rank=1, max_target_rank_krylov_eigs=50)
        solver.chambolle_pock(aff, con, Options(max_iter=5, initial_target_rank=r, freeze_target_rank=1, max_target_rank_krylov_eigs=50))
        rg = solver.chambolle_pock(aff, con, opt)
        row = dict(config=f"C3 {name} rank sweep", target_rank=r, krylov_dim=max(2 * r + 1, 25), iters=rg.iter,
                   gpu_ms_per_eig_projection=1e3 * rg.time_psd_proj / max(rg.n_psd_proj, 1), gpu_ms_per_iter=1e3 * rg.time_loop / rg.iter,
                   gpu_matvecs_per_projection=rg.lanczos_matvecs / max(rg.lanczos_calls, 1), gpu_full_eig_calls=rg.full_eig_calls)
        if CPU:
            ro = oracle.chambolle_pock(aff, con, Options(max_iter=40, initial_target_rank=r, freeze_target_rank=1, max_target_rank_krylov_eigs=50))
            row.update(cpu_ms_per_eig_projection=1e3 * ro.time_psd_proj / max(ro.n_psd_proj, 1), cpu_ms_per_iter=1e3 * ro.time_loop / ro.iter,
                       cpu_threads=oracle.num_threads())
        emit(**row)
    rg = solver.chambolle_pock(aff, con, Options())
    emit(config=f"C3 {name} full solve, default options", status=rg.status, iters=rg.iter, objval=rg.objval, gpu_total_s=rg.time,
         gpu_ms_per_iter=1e3 * rg.time_loop / rg.iter, final_target_rank=[int(t) for t in rg.target_rank])

# C4 on one GPU: 256 stacked MIMO n=64 cones
probs = [mimo_problem(1000 + s, 64) for s in range(256)]
aff, con = stack_problems(probs)
solver.chambolle_pock(aff, con, Options(max_iter=5))
rg = solver.chambolle_pock(aff, con, Options(max_iter=300))
row = dict(config="C4 256 x MIMO n=64 stacked, 1 GPU", iters=rg.iter, status=rg.status, gpu_ms_per_iter=1e3 * rg.time_loop / rg.iter,
           gpu_ms_per_eig_projection_batch=1e3 * rg.time_psd_proj / max(rg.n_psd_proj, 1),
           gpu_cone_projections_per_s=256 * rg.iter / rg.time_loop)
if CPU:
    ro = oracle.chambolle_pock(aff, con, Options(max_iter=10))
    row.update(cpu_ms_per_iter=1e3 * ro.time_loop / ro.iter, cpu_threads=oracle.num_threads())
emit(**row)

# C5: sensor localisation n = 1000 (PSD side 1002, ~150 k equality rows)
t0 = time.time(); aff, con = sensorloc_problem(0, 1000); tb = time.time() - t0
solver.chambolle_pock(aff, con, Options(max_iter=5))
rg = solver.chambolle_pock(aff, con, Options(max_iter=300))
row = dict(config="C5 sensorloc n=1000", n=int(aff.n), p=int(aff.p), nnzA=int(aff.A.nnz), build_s=tb, iters=rg.iter,
           gpu_ms_per_iter=1e3 * rg.time_loop / rg.iter, gpu_ms_per_eig_projection=1e3 * rg.time_psd_proj / max(rg.n_psd_proj, 1),
           gpu_setup_s=rg.time_setup)
if CPU:
    ro = oracle.chambolle_pock(aff, con, Options(max_iter=20))
    row.update(cpu_ms_per_iter=1e3 * ro.time_loop / ro.iter, cpu_ms_per_eig_projection=1e3 * ro.time_psd_proj / max(ro.n_psd_proj, 1))
emit(**row)
