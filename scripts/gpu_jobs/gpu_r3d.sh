#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== lanczos tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lanczos or psd_projection or c2_fullsize or rank_sweep or krylov_sdplib or smoke" 2>&1 | tail -4
echo "== rank 10, 205 iterations: resident vs grid-wide"
python scripts/dbg_resident2.py 10 205 2>&1 | grep -v "^\[bj\]" | grep -v "matvecs per iteration\|objective trace" | cut -c1-300
for v in default 0; do
  if [ "$v" = "default" ]; then unset PROXSDP_B200_LZ_RESIDENT; else export PROXSDP_B200_LZ_RESIDENT=$v; fi
  for r in 5 10 16; do timeout 300 python scripts/dbg_resident.py $r 2>&1 | grep -v "^\[bj\]" | tail -1 | sed "s/^/RESIDENT=$v: /"; done
  timeout 300 python scripts/dbg_resident.py full 2>&1 | grep -v "^\[bj\]" | grep "status\|per iteration" | sed "s/^/RESIDENT=$v: /"
done
unset PROXSDP_B200_LZ_RESIDENT
echo "== C2 full solve + implicit operator tests"
timeout 900 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -x -k "c2_headline or implicit or fallback or arpack or small_sides" 2>&1 | tail -4
echo "== bench 20/3"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-large-cone 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench: it/s %.1f' % d['value'], 'roofline frac %.3f' % d['roofline']['frac'], 'avg launch ms %.4f' % d['roofline']['avg_launch_ms'], 'e2e %.1f' % d['e2e']['value'])
"
