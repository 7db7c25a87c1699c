"""GPU solves of the headline instances to termination: iterations, time, objective (feeds the golden comparison)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import load_problem, maxcut_er_problem
G = os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests", "golden")
for name in (sys.argv[1:] or ["c2", "maxG32", "mcp500-1_exact"]):
    kw = {}
    if name == "c2":
        aff, con = maxcut_er_problem(2000, 0.01, 0)
    elif name == "maxG32":
        aff, con = load_problem(f"{G}/sdplib_maxG32.npz")
    else:
        aff, con = load_problem(f"{G}/sdplib_mcp500-1.npz"); kw = dict(full_eig_decomp=True)
    t0 = time.time()
    r = solver.chambolle_pock(aff, con, Options(trace_cap=200000, **kw))
    print(f"{name}: GPU status {r.status} ({r.status_string}) iter {r.iter} obj {r.objval:.9g} gap {r.gap:.3e} feas {r.primal_residual:.3e} "
          f"rank {r.final_rank} target {r.target_rank} mat-vecs {r.lanczos_matvecs} full-eig {r.full_eig_calls} "
          f"wall {time.time()-t0:.2f} s loop {r.time_loop:.2f} s  ({r.iter / max(r.time_loop, 1e-9):.0f} it/s)", flush=True)
    np.savez_compressed(os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "gpurun_out", f"gpu_full_{name}.npz"),
                        trace=r.trace, iters=r.iter, objval=r.objval, status=r.status)
