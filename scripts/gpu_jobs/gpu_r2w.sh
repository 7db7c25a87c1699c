#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== lanczos tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lanczos or psd_projection or c2_fullsize or rank_sweep or krylov_sdplib or smoke" 2>&1 | tail -4
echo "== large cones, default settings"; timeout 400 python scripts/lz_large.py 2000 2500 3000 4000 5000 6000 2>&1 | grep -v "^\[bj\]" | tee gpurun_out/lz_large_r2w.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['n'], 'us/matvec %.2f' % d['us_per_matvec'], 'frac %.3f' % d['frac_of_hbm_peak'], 'equal', d['counts_equal_oracle'], 'res %.1e' % d['residual_rel'])
"
for i in 1 2; do
echo "== bench 20/5 run $i"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench20_r2w_$i.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench: it/s %.1f' % d['value'], 'roofline frac %.3f' % d['roofline']['frac'], 'avg launch ms %.4f' % d['roofline']['avg_launch_ms'], 'e2e %.1f' % d['e2e']['value']); print(json.dumps(d.get('large_cone')))
"
done
