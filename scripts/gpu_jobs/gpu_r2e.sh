#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2e}
echo "== lanczos tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -m gpu -q -x -k "lanczos or psd_projection or c2_fullsize or rank_sweep or krylov or min_size" 2>&1 | tail -8
for rb in 3 5 6 9; do
  for it in 20 200; do
  echo "== RB $rb, $it iterations"; PROXSDP_B200_LZ_RB=$rb PROXSDP_B200_LZ_PROF=1 timeout 120 python scripts/lz_prof.py $it 2>&1 | grep -E "iterations|ritz:|ritz_top_bi|symv  |gridxchg|ritz   |symv: " | tee -a gpurun_out/rb_sweep_$TAG.txt
  done
done
echo "== bench 20/5"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench20_$TAG.json | cut -c1-300
