import os, sys
import numpy as np
ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import load_problem
aff, con = load_problem(os.path.join(ROOT, "tests", "golden", "sdplib_mcp500-1.npz"))
r_ = int(sys.argv[1]); its = int(sys.argv[2])
out = {}
for mode in ("resident", "0", "8"):
    if mode == "resident": os.environ.pop("PROXSDP_B200_LZ_RESIDENT", None)
    else: os.environ["PROXSDP_B200_LZ_RESIDENT"] = mode
    opt = Options(max_iter=its, initial_target_rank=r_, freeze_target_rank=1, max_target_rank_krylov_eigs=50, trace_cap=its + 5)
    r = solver.chambolle_pock(aff, con, opt)
    out[mode] = r.trace[:, 12].astype(int)
    print(mode, "matvecs per iteration:", out[mode].tolist()[:its], flush=True)
    print(mode, "objective trace:", np.round(r.trace[:its:5, 1], 6).tolist(), flush=True)
a, b = out["resident"], out["0"]
k = min(len(a), len(b))
diff = np.nonzero(a[:k] != b[:k])[0]
print("first differing iteration:", int(diff[0]) if len(diff) else None)
if len(diff):
    i = int(diff[0])
    print("resident around it:", a[max(0, i - 3):i + 12].tolist())
    print("grid-wide around it:", b[max(0, i - 3):i + 12].tolist())
print("iterations with >= 500 mat-vecs: resident", int((a >= 500).sum()), "grid-wide", int((b >= 500).sum()))
