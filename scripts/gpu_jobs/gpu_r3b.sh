#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
scripts/ritz_bench.bin 25 | grep "reps 20" | tee gpurun_out/ritz_bench_v3.txt | cut -c1-200
echo "== lanczos tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lanczos or psd_projection or c2_fullsize or rank_sweep or krylov_sdplib or smoke" 2>&1 | tail -4
for v in default 0; do
  if [ "$v" = "default" ]; then unset PROXSDP_B200_LZ_RESIDENT; else export PROXSDP_B200_LZ_RESIDENT=$v; fi
  echo "== PROXSDP_B200_LZ_RESIDENT=$v: rank 5 sweep with phase profile"
  PROXSDP_B200_LZ_PROF=1 PROXSDP_B200_DEBUG=1 timeout 300 python scripts/dbg_resident.py 5 2>&1 | grep -v "^\[bj\]" | tail -24
  echo "== PROXSDP_B200_LZ_RESIDENT=$v: rank 10 sweep"
  timeout 300 python scripts/dbg_resident.py 10 2>&1 | grep -v "^\[bj\]" | tail -1
  echo "== PROXSDP_B200_LZ_RESIDENT=$v: full solve"
  PROXSDP_B200_DEBUG=1 timeout 300 python scripts/dbg_resident.py full 2>&1 | grep -v "^\[bj\]" | sort | uniq -c | sort -rn | head -8
done
unset PROXSDP_B200_LZ_RESIDENT
echo "== bench 20/3"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-large-cone 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench: it/s %.1f' % d['value'], 'roofline frac %.3f' % d['roofline']['frac'], 'avg launch ms %.4f' % d['roofline']['avg_launch_ms'], 'e2e %.1f' % d['e2e']['value'])
"
