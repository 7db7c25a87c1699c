"""Lanczos kernel time per mat-vec for several cluster sizes (PROXSDP_B200_CLUSTER), Max-Cut n=2000, 200 iterations."""
import os, subprocess, sys, json
root = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
code = r'''
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import maxcut_er_problem
aff, con = maxcut_er_problem(2000, 0.01, 0)
with solver.Solve(aff, con, Options()) as s:
    s.iterate(200, False)
    c = s.counters()
    r = s.finish()
print("RESULT", c["lanczos_matvecs"], c["lanczos_ms"], c["psd_proj_ms"], r.objval)
'''
for C in sys.argv[1:] or ["4", "5", "6", "7", "8"]:
    env = dict(os.environ, PROXSDP_B200_CLUSTER=C)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=root)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    if not line:
        print(C, "failed", out.stderr[-300:])
        continue
    _, mv, ms, psd, obj = line[0].split()
    print(f"cluster {C}: {float(ms) / int(mv) * 1e3:.2f} us per mat-vec, lanczos {float(ms):.2f} ms, psd {float(psd):.2f} ms, obj {obj}  {out.stderr.strip()[-200:]}", flush=True)
