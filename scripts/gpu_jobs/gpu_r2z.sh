#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2z}
bash scripts/sanitize.sh $TAG
echo "== pytest -m gpu tests/test_gpu_parity_full.py"
timeout 2400 python -m pytest tests/test_gpu_parity_full.py -x -q -m gpu --durations=8 2>&1 | tail -20 | tee gpurun_out/pytest_gpu_full_$TAG.log
