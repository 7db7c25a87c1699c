import os, sys
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import maxcut_er_problem
np.set_printoptions(linewidth=250, precision=6)
aff, con = maxcut_er_problem(2000, 0.01, 0)
z = np.load(os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests/golden/full_c2.npz"))
tr_o = z["trace"]
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
r = solver.chambolle_pock(aff, con, Options(max_iter=K, trace_cap=K, max_target_rank_krylov_eigs=32))
tr = r.trace
k = min(len(tr), len(tr_o))
d = np.abs(tr[:k, 1:7] - tr_o[:k, 1:7]) / np.maximum(1.0, np.abs(tr_o[:k, 1:7]))
bad = np.nonzero(d.max(axis=1) > 1e-6)[0]
mvbad = np.nonzero(tr[:k, 12] != tr_o[:k, 12])[0]
print(os.environ.get("TAGX", ""), "status", r.status, r.status_string, "iter", r.iter, "first trace diff at", bad[:3], "first matvec-count diff at", mvbad[:5], "full eig calls", r.full_eig_calls)
cols = [0, 1, 2, 3, 4, 9, 10, 11, 12, 13]
i0 = max(0, (min(bad[0] if len(bad) else k, mvbad[0] if len(mvbad) else k)) - 2)
print(" gpu   \n", tr[i0:i0 + 6, cols]); print(" oracle\n", tr_o[i0:i0 + 6, cols])
