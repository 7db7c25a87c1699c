import os, sys
os.environ["PROXSDP_B200_LZ_PROF"] = "1"
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import maxcut_er_problem
aff, con = maxcut_er_problem(2000, 0.01, 0)
with solver.Solve(aff, con, Options()) as s:
    s.iterate(int(sys.argv[1]) if len(sys.argv) > 1 else 200, False)
    c = s.counters()
    print({k: c[k] for k in ("iterations", "lanczos_matvecs", "lanczos_ms", "psd_proj_ms", "rest_ms")})
    r = s.finish()
