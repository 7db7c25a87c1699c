#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== lanczos tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lanczos or rank_sweep or krylov_sdplib or c2_fullsize or gpp500" 2>&1 | tail -3
for p in gpp500-1 mcp500-1; do
  timeout 300 python scripts/dbg_hyp.py $p 2>&1 | grep -v "^\[bj\]" | tail -1 | sed "s/^/$p default: /"
  PROXSDP_B200_LZ_RESIDENT=0 timeout 300 python scripts/dbg_hyp.py $p 2>&1 | grep -v "^\[bj\]" | tail -1 | sed "s/^/$p: /"
done
echo "== C2 full solve"
timeout 900 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -x -k "c2_headline or implicit or fallback" 2>&1 | tail -3
echo "== bench 20/3"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-large-cone 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench: it/s %.1f' % d['value'], 'roofline frac %.3f' % d['roofline']['frac'], 'avg launch ms %.4f' % d['roofline']['avg_launch_ms'], 'e2e %.1f' % d['e2e']['value'])
"
