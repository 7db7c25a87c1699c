import os, sys, subprocess
code = r'''
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import maxcut_er_problem
aff, con = maxcut_er_problem(2000, 0.01, 0)
with solver.Solve(aff, con, Options()) as s:
    s.iterate(300, False)
    c = s.counters()
    print(os.environ.get("PROXSDP_B200_CLUSTER"), os.environ.get("PROXSDP_B200_LANCZOS"), "lanczos ms/launch %.4f" % (c["lanczos_ms"]/c["lanczos_timed_calls"]), "matvecs", c["lanczos_matvecs"], flush=True)
'''
for C, mode in (("2", ""), ("4", ""), ("8", ""), ("16", ""), ("8", "rows")):
    env = dict(os.environ, PROXSDP_B200_CLUSTER=C, PROXSDP_B200_LANCZOS=mode)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print((r.stdout + r.stderr).strip().splitlines()[-1])
