import os, sys
import numpy as np
ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import load_problem
aff, con = load_problem(os.path.join(ROOT, "tests", "golden", "sdplib_mcp500-1.npz"))
its = int(sys.argv[1])
opt = Options(max_iter=its, initial_target_rank=10, freeze_target_rank=1, max_target_rank_krylov_eigs=50, trace_cap=its + 5)
with solver.Solve(aff, con, opt) as s:
    s.iterate(its - 1, False)
    print("=== last iteration", flush=True)
    os.environ["PROXSDP_B200_LZ_DEBUG"] = "1"
    s.iterate(1, False)
    r = s.finish()
print("matvecs", r.trace[:, 12].astype(int).tolist()[-4:], flush=True)
