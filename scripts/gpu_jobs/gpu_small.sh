#!/bin/bash
# small-cone (full eigendecomposition) path: parity tests + C4 timing: warm / cold double-buffered Jacobi, in-place Jacobi
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests -m gpu -q -x -k "eigh or psd_projection or mimo or unit_problems or sdp_wiki or termination or readme or mixed_soc or fixed_step or exact_mode" 2>&1 | tail -4
for cfg in "1 1" "1 0" "0 0"; do
set -- $cfg
PROXSDP_B200_SMALL_FAST=$1 PROXSDP_B200_SMALL_WARM=$2 python - <<'PY'
import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from proxsdp_b200 import Options, solver
from proxsdp_b200.problems import mimo_problem, stack_problems
probs = [mimo_problem(s, 64, 0) for s in range(256)]
aff, con = stack_problems(probs)
solver.chambolle_pock(aff, con, Options(max_iter=3))
r = solver.chambolle_pock(aff, con, Options())
print("SMALL_FAST", os.environ["PROXSDP_B200_SMALL_FAST"], "WARM", os.environ["PROXSDP_B200_SMALL_WARM"], "status", r.status, "iters", r.iter, "obj", r.objval, "ms/iter", 1e3 * r.time_loop / r.iter, "ms/proj", 1e3 * r.time_psd_proj / max(r.n_psd_proj, 1))
PY
done
