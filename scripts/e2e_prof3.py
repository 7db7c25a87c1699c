"""Stage timers of one chambolle_pock call (PROXSDP_B200_TIMING=1) on the pinned / SparseMatrixCSC form of C2."""
import os, sys, time
os.environ["PROXSDP_B200_TIMING"] = "1"
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import maxcut_er_problem
from proxsdp_b200._abi import MarshalledProblem, MarshalledResult
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
aff, con = maxcut_er_problem(2000, 0.01, 0)
aff_p, con_p = solver.pin_problem(aff, con)
for tag, (a, c) in (("scipy/pageable", (aff, con)), ("pinned int64", (aff_p, con_p))):
    for i in range(4):
        t0 = time.perf_counter(); mp = MarshalledProblem(a, c); t1 = time.perf_counter()
        mr = MarshalledResult(mp.n, mp.p, mp.m, mp.n_sdp, 0, empty=solver.pinned_empty); t2 = time.perf_counter()
        print(f"--- {tag} run {i}: marshal problem {1e3*(t1-t0):.2f} ms, result buffers {1e3*(t2-t1):.2f} ms", flush=True)
        del mr
        t0 = time.perf_counter()
        r = solver.chambolle_pock(a, c, Options(max_iter=K))
        w = time.perf_counter() - t0
        print(f"wall {1e3*w:.2f} ms  setup {1e3*r.time_setup:.2f}  loop {1e3*r.time_loop:.2f}  rest {1e3*(w - r.time_setup - r.time_loop):.2f}  "
              f"h2d {r.h2d_bytes/1e6:.1f} MB d2h {r.d2h_bytes/1e6:.1f} MB  launches {r.gpu_launches}", flush=True)
