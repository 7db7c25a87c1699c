import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import load_problem
aff, con = load_problem(os.path.join(ROOT, "tests", "golden", f"sdplib_{sys.argv[1]}.npz"))
r = solver.chambolle_pock(aff, con, Options())
print(" ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("PROXSDP_B200_")), "->", f"status {r.status} iters {r.iter} obj {r.objval:.6f} time_loop {r.time_loop:.2f} s matvecs {int(r.lanczos_matvecs)} full eigs {r.full_eig_calls}", flush=True)
