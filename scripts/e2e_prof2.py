import os, sys, time
os.environ["PROXSDP_B200_TIMING"] = "1"
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import maxcut_er_problem
from proxsdp_b200._abi import MarshalledProblem
aff, con = maxcut_er_problem(2000, 0.01, 0)
for i in range(3):
    t0 = time.perf_counter(); mp = MarshalledProblem(aff, con); t1 = time.perf_counter()
    print(f"--- run {i}: marshal {1e3*(t1-t0):.1f} ms", flush=True)
    t0 = time.perf_counter()
    r = solver.chambolle_pock(aff, con, Options(max_iter=300))
    print(f"wall {time.perf_counter()-t0:.3f} setup {r.time_setup:.3f} loop {r.time_loop:.3f}", flush=True)
