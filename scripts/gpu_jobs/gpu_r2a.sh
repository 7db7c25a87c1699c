#!/bin/bash
# round 2, first visit: does the refactor (device ingest, device result assembly, spin sync, fused ladder) hold?
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -2
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== e2e stage timers"; timeout 300 python scripts/e2e_prof3.py 20 > gpurun_out/e2e_prof_$TAG.txt 2>&1; grep -v "^\[bj\]" gpurun_out/e2e_prof_$TAG.txt | tail -40
echo "== bench 20/5"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -2 | tee gpurun_out/bench20_$TAG.json | cut -c1-3000
echo "== pytest gpu (new file first)"
timeout 1500 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -x --deselect tests/test_gpu_parity_full.py::test_full_solve_c2_headline --deselect tests/test_gpu_parity_full.py::test_full_solve_maxG32 --deselect tests/test_gpu_parity_full.py::test_full_solve_exact_mode_mcp500 2>&1 | tail -40 > gpurun_out/pytest_new_$TAG.log; tail -40 gpurun_out/pytest_new_$TAG.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_old_$TAG.log; tail -25 gpurun_out/pytest_old_$TAG.log
