#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2f}
echo "== gpp500-1 full solve: counters"
for v in default 0; do
  if [ "$v" = "default" ]; then unset PROXSDP_B200_LZ_RESIDENT; else export PROXSDP_B200_LZ_RESIDENT=$v; fi
  PROBLEM=gpp500-1 timeout 300 python scripts/dbg_resident.py full 2>&1 | grep -v "^\[bj\]" | grep "status\|per iteration" | sed "s/^/RESIDENT=$v: /"
done
unset PROXSDP_B200_LZ_RESIDENT
echo "== ncu evidence"
bash scripts/gpu_ncu_r2.sh $TAG 2>&1 | tail -8
echo "== bench default flags"
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default_$TAG.json | cut -c1-200
echo "== bench 20/3"
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench20_$TAG.json | cut -c1-200
echo "== bench reference 20/3"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench20_ref_$TAG.json | cut -c1-200
echo "== racecheck, reduced driver"
PROXSDP_B200_LZ_SPIN_S=400 timeout 700 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_race.py 2>&1 | grep -v "^\[bj\]" | tail -14 | tee gpurun_out/racecheck_$TAG.txt
du -sh gpurun_out
