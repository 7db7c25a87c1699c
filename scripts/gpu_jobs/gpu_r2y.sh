#!/bin/bash
# full GPU suite + bench + a check that the library still launches under ncu (cooperative attribute dropped there)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2y}
echo "== env seen by a process under ncu"
timeout 120 ncu --metrics gpu__time_duration.sum -c 1 python -c "
import os
print(sorted(k for k in os.environ if 'INJECT' in k or 'NV_' in k or 'NSIGHT' in k or 'CUPTI' in k))" 2>&1 | tail -3 | cut -c1-400
echo "== smoke under ncu launch list (no env override)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/smoke_launches_$TAG.csv python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | cut -c1-300
grep -c "k_lanczos" gpurun_out/smoke_launches_$TAG.csv
echo "== pytest -m gpu"
timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_$TAG.log
echo "== bench default flags"
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default_$TAG.json | cut -c1-600
echo "== bench 20/3"
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench20_$TAG.json | cut -c1-300
echo "== bench reference 20/3"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench20_ref_$TAG.json | cut -c1-300
