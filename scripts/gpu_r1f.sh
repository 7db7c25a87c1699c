#!/bin/bash
# GPU visit r1f: validate + time the third-generation Lanczos kernel against the second generation.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== quick parity, kernel gen 3"; timeout 300 python scripts/gpu_quick.py 2>&1 | tail -25
for cfg in "2 1" "3 0" "3 1"; do
  set -- $cfg
  echo "== lz_prof kernel=$1 xres=$2"
  PROXSDP_B200_LZ_KERNEL=$1 PROXSDP_B200_LZ_XRES=$2 timeout 120 python scripts/lz_prof.py 2>&1 | tail -14
done
echo "== bench gen3"; timeout 300 python bench.py --steps 300 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_r1f.json
echo "== bench gen3 noflush"; timeout 300 python bench.py --steps 300 --warmup 10 --no-flush-l2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_noflush_r1f.json
echo "== bench gen2 noflush"; PROXSDP_B200_LZ_KERNEL=2 timeout 300 python bench.py --steps 300 --warmup 10 --no-flush-l2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_noflush_gen2_r1f.json
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r1f.log
