#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2h}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench20_ref_$TAG.json
timeout 600 python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench20_$TAG.json
timeout 900 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_default_$TAG.json
for f in bench20_ref_$TAG bench20_$TAG bench_default_$TAG; do python - gpurun_out/$f.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read())
print(sys.argv[1], 'value %.1f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], 'frac', (d.get('roofline') or {}).get('frac'), 'launch ms', (d.get('roofline') or {}).get('avg_launch_ms'),
      'cpu', (d.get('cpu_baseline') or {}).get('value'), 'large', ((d.get('large_cone') or {}).get('roofline') or {}).get('frac'), 'launches', d.get('gpu_launches'))
PY
done
