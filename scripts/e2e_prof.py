import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import maxcut_er_problem
aff, con = maxcut_er_problem(2000, 0.01, 0)
def e2e(tag):
    t0 = time.perf_counter()
    r = solver.chambolle_pock(aff, con, Options(max_iter=300))
    w = time.perf_counter() - t0
    print(f"{tag}: wall {w:.3f} setup {r.time_setup:.3f} loop {r.time_loop:.3f} result.time {r.time:.3f} psd {r.time_psd_proj:.3f}", flush=True)
e2e("cold")
e2e("warm1")
e2e("warm2")
for flush in (False, True):
    with solver.Solve(aff, con, Options()) as s:
        t0 = time.perf_counter(); s.iterate(50, flush); t1 = time.perf_counter()
        r = s.finish(); t2 = time.perf_counter()
    t3 = time.perf_counter()
    print(f"stepwise flush={flush}: iterate {t1-t0:.3f} finish {t2-t1:.3f} destroy {t3-t2:.3f}", flush=True)
    e2e(f"after stepwise flush={flush}")
    e2e(f"again")
