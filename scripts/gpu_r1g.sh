#!/bin/bash
# GPU visit r1g: strip symv / barrier modes of the third-generation Lanczos kernel.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== lanczos parity tests"; timeout 600 python -m pytest tests -m gpu -q -x -k "lanczos or psd_projection" 2>&1 | tail -8
for cfg in ${CFGS:-0,1,0 0,1,1 2,1,1}; do
  IFS=, read b s x <<< "$cfg"
  echo "== lz_prof bar=$b symv=$s xres=$x"
  PROXSDP_B200_LZ_BAR=$b PROXSDP_B200_LZ_SYMV=$s PROXSDP_B200_LZ_XRES=$x timeout 120 python scripts/lz_prof.py 2>&1 | tail -22
done
echo "== bench gen3 noflush"; timeout 300 python bench.py --steps 300 --warmup 10 --no-flush-l2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_noflush_r1g.json | cut -c1-400
