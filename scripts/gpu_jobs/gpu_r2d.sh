#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2d}
echo "== lanczos tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -m gpu -q -x -k "lanczos or psd_projection or c2_fullsize or rank_sweep or krylov or min_size" 2>&1 | tail -15
for ns in 0 50 150 400; do
  echo "== poll back-off $ns ns"; PROXSDP_B200_LZ_POLL_NS=$ns PROXSDP_B200_LZ_PROF=1 timeout 120 python scripts/lz_prof.py 20 2>&1 | grep -E "iterations|ritz:|symv  |gridxchg|ritz   |symv\+reduce|exchange\+alpha" | tee -a gpurun_out/poll_sweep_$TAG.txt
done
for ns in 0 150; do
  echo "== poll back-off $ns ns, 200 iterations"; PROXSDP_B200_LZ_POLL_NS=$ns PROXSDP_B200_LZ_PROF=1 timeout 120 python scripts/lz_prof.py 200 2>&1 | grep -E "iterations|ritz:|symv  |gridxchg|ritz   " | tee -a gpurun_out/poll_sweep_$TAG.txt
done
echo "== bench 20/5"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench20_$TAG.json | cut -c1-2200
