#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== parity tests touching the streaming kernels"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -m gpu -q -x -k "readme or mimo or exact_mode_trace or krylov_sdplib or c2_fullsize or stepwise or mixed_soc or gpp500 or reference_unit or c2_headline or seam or implicit or permuted or pinned or duals" 2>&1 | tail -4
for v in 0 1; do
  if [ $v = 1 ]; then export PROXSDP_B200_NO_NZMASK=1; else unset PROXSDP_B200_NO_NZMASK; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-large-cone 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('NO_NZMASK=$v: it/s %.1f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], {k: round(v, 4) for k, v in d['sections_ms_per_step'].items()}, 'obj', d['objective_after_steps'])
"
done
