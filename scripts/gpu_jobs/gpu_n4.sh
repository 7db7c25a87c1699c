#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 20 --warmup 3 2>gpurun_out/bench_n4.err | tail -1 > gpurun_out/bench_n4_r2h.json
echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n4_r2h.json').read())
print('N=4: value %.1f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'])
print(json.dumps(d.get('sharded_batch'))[:900])
PY
tail -3 gpurun_out/bench_n4.err | cut -c1-300
