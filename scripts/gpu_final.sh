#!/bin/bash
# end-of-round visit: whole GPU suite, secondary configs, ncu evidence of the final build, bench lines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2f}
echo "== pytest -m gpu (whole suite)"
timeout 2700 python -m pytest tests/ -x -q -m gpu --durations=6 2>&1 | tail -14 | tee gpurun_out/pytest_gpu_$TAG.log
echo "== secondary configs"
timeout 900 python scripts/bench_configs.py --no-cpu 2>&1 | grep -v "^\[bj\]" | grep "^{" | tee gpurun_out/configs_$TAG.jsonl | cut -c1-260
echo "== ncu evidence"
bash scripts/gpu_ncu_r2.sh $TAG 2>&1 | tail -12
echo "== bench default flags"
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default_$TAG.json | cut -c1-300
echo "== bench 20/3"
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench20_$TAG.json | cut -c1-300
echo "== bench reference 20/3"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench20_ref_$TAG.json | cut -c1-300
