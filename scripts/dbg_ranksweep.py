import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
from oracle import oracle
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import load_problem
aff, con = load_problem("tests/golden/sdplib_mcp250-1.npz")
for r in (5, 25):
    opt = Options(max_iter=25, trace_cap=25, initial_target_rank=r, freeze_target_rank=1, max_target_rank_krylov_eigs=50)
    rg = solver.chambolle_pock(aff, con, opt); ro = oracle.chambolle_pock(aff, con, opt)
    d = np.abs(rg.trace[:, 1:9] - ro.trace[:, 1:9]).max(axis=1)
    print("r", r, "per-iter max diff", np.array2string(d, precision=1, max_line_width=200))
    print(" mv gpu", rg.trace[:, 12].astype(int).tolist()); print(" mv cpu", ro.trace[:, 12].astype(int).tolist())
    print(" cr gpu", rg.trace[:, 10].astype(int).tolist()); print(" cr cpu", ro.trace[:, 10].astype(int).tolist())
    print(" mineig gpu", np.array2string(rg.trace[:8, 11], precision=3)); print(" mineig cpu", np.array2string(ro.trace[:8, 11], precision=3))
