import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import sensorloc_problem
aff, con = sensorloc_problem(0, 1000)
A = aff.A.tocsr()
rl = np.diff(A.indptr); print("rows", A.shape, "row nnz min/mean/max", rl.min(), rl.mean(), rl.max())
cl = np.diff(aff.A.tocsc().indptr); print("col nnz max", cl.max(), "mean", cl.mean(), "nonempty cols", (cl > 0).sum())
rg = solver.chambolle_pock(aff, con, Options(max_iter=100, trace_cap=100))
print("iters", rg.iter, "loop", rg.time_loop, "psd", rg.time_psd_proj, "rest(dev)", rg.time_rest, "ls trials", rg.linesearch_trials, "launches", rg.gpu_launches, "mv", rg.lanczos_matvecs)
print("ls per iter", rg.trace[:20, 13])
