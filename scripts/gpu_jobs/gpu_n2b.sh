#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nproc
for i in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus 2 --steps 20 --warmup 3 --no-sharded-batch 2>/dev/null | tail -1 | tee gpurun_out/bench_n2_r2g_$i.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('N=2: value %.1f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'sections', {k: round(v, 4) for k, v in d['sections_ms_per_step'].items()})
"
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-large-cone 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('N=1: value %.1f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'sections', {k: round(v, 4) for k, v in d['sections_ms_per_step'].items()})
"
