#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for PF in 0 4 8 16 32; do
  echo "== prefetch rows $PF"; PROXSDP_B200_LZ_PF=$PF timeout 300 python scripts/lz_large.py 3000 4000 5000 2>&1 | grep -v "^\[bj\]" | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['n'], 'us/matvec %.2f' % d['us_per_matvec'], 'frac %.3f' % d['frac_of_hbm_peak'], 'equal', d['counts_equal_oracle'], 'res %.1e' % d['residual_rel'])
"
done | tee gpurun_out/lz_large_pf_r2u.txt
