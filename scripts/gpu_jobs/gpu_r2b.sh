#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2b}
echo "== symv microbench"; timeout 120 ./scripts/symv_bench2.bin 2>&1 | tee gpurun_out/symv_bench2_$TAG.txt
echo "== linesearch debug"; timeout 300 python scripts/dbg_ls.py 2>&1 | tail -30
echo "== GPU full solves"; timeout 900 python scripts/full_probe.py 2>&1 | grep -v "^\[bj\]" | tail -5
echo "== oracle full solves (golden)"
for inst in mcp500-1_exact c2 maxG32; do
  timeout 1200 python tests/golden/make_full_solves.py $inst 2>&1 | tail -2
  cp tests/golden/full_$inst.npz gpurun_out/ 2>/dev/null
done
echo "== bench 20/5"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench20_$TAG.json | cut -c1-2500
echo "== e2e stage timers"; timeout 300 python scripts/e2e_prof3.py 20 2>&1 | grep -v "^\[bj\]" | tail -12
