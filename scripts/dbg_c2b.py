import os, sys
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import maxcut_er_problem
aff, con = maxcut_er_problem(2000, 0.01, 0)
K = int(sys.argv[1])
r = solver.chambolle_pock(aff, con, Options(max_iter=K, trace_cap=K, max_target_rank_krylov_eigs=32))
print("iter", r.iter, "full", r.full_eig_calls, "last rows", r.trace[-3:, [0, 1, 4, 9, 10, 11, 12]])
