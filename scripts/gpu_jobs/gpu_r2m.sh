#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2m}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --durations=12 2>&1 | grep -v "^\[bj\]" | tail -45 > gpurun_out/pytest_gpu_$TAG.log; tail -45 gpurun_out/pytest_gpu_$TAG.log
echo "== bench 20/5"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench20_$TAG.json | cut -c1-2600
