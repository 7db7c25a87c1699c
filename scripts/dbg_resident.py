"""Resident single-cluster kernel vs grid-wide kernel on mcp500-1: counters of a full solve and the per-phase profile."""
import os, sys, time
import numpy as np
ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import load_problem
aff, con = load_problem(os.path.join(ROOT, "tests", "golden", "sdplib_" + os.environ.get("PROBLEM", "mcp500-1") + ".npz"))
mode = sys.argv[1] if len(sys.argv) > 1 else "full"
if mode == "full":
    solver.chambolle_pock(aff, con, Options(max_iter=5))
    r = solver.chambolle_pock(aff, con, Options(trace_cap=20000))
    print("status", r.status, "iters", r.iter, "time_loop %.2f s" % r.time_loop, "psd %.2f s" % r.time_psd_proj, "lanczos calls", r.lanczos_calls,
          "matvecs", r.lanczos_matvecs, "full eig calls", r.full_eig_calls, "time_lanczos %.2f" % r.time_lanczos, "rest %.2f" % r.time_rest, flush=True)
    tr = r.trace
    mv = tr[:, 12]
    print("matvecs per iteration: mean %.1f max %.0f; iterations with > 500 mat-vecs: %d" % (mv.mean(), mv.max(), int((mv > 500).sum())))
    print("target rank sum over time (every 500 its):", tr[::500, 9].astype(int).tolist())
else:
    r_ = int(mode)
    opt = Options(max_iter=200, initial_target_rank=r_, freeze_target_rank=1, max_target_rank_krylov_eigs=50)
    solver.chambolle_pock(aff, con, Options(max_iter=5, initial_target_rank=r_, freeze_target_rank=1, max_target_rank_krylov_eigs=50))
    r = solver.chambolle_pock(aff, con, opt)
    print("rank", r_, "ms per eig projection %.3f" % (1e3 * r.time_psd_proj / r.n_psd_proj), "matvecs per projection %.1f" % (r.lanczos_matvecs / r.lanczos_calls),
          "lanczos kernel ms per call %.3f" % (1e3 * r.time_lanczos / max(r.lanczos_calls, 1)), flush=True)
