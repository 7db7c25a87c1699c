#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for PF in 0 2 4 8 16; do
  echo "== prefetch rows $PF"; PROXSDP_B200_LZ_PF=$PF timeout 300 python scripts/lz_large.py 1000 1500 2000 2500 2>&1 | grep -v "^\[bj\]" | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['n'], 'us/matvec %.2f' % d['us_per_matvec'], 'frac %.3f' % d['frac_of_hbm_peak'], 'equal', d['counts_equal_oracle'], 'res %.1e' % d['residual_rel'])
"
  PROXSDP_B200_LZ_PF=$PF timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-large-cone 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench: it/s %.1f' % d['value'], 'roofline frac %.3f' % d['roofline']['frac'], 'avg launch ms %.4f' % d['roofline']['avg_launch_ms'], 'e2e %.1f' % d['e2e']['value'])
"
done | tee gpurun_out/lz_pf_small_r2v.txt
