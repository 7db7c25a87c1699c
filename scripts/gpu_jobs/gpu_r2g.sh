#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2g}
echo "== eigh / exact-mode tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -m gpu -q -x -k "eigh or exact_mode_trace or krylov_failure or psd_projection_parity or seam" 2>&1 | tail -8
for inner in 1 2 3; do echo "== eigh bench, inner sweeps $inner"; PROXSDP_B200_BJ_INNER=$inner timeout 200 python scripts/eigh_bench.py 500 1000 2000 2>&1 | grep -v "^\[bj\]" | tail -3 | tee -a gpurun_out/eigh_bench_$TAG.txt; done
echo "== full solves vs golden"; timeout 900 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -x -s -k "full_solve_c2 or full_solve_exact" --durations=5 2>&1 | tail -25
echo "== bench 20/5"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench20_$TAG.json | cut -c1-1500
