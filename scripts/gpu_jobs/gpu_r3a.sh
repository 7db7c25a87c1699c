#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r3a}
echo "== lanczos tests (resident single-cluster kernel is the default for sides 64..~580)"
PROXSDP_B200_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lanczos or psd_projection or c2_fullsize or rank_sweep or krylov_sdplib or smoke or eigh" 2>&1 | grep -v "^\[bj\]" | tail -12
echo "== C3 configs, resident kernel (default)"
timeout 600 python scripts/bench_configs.py --no-cpu 2>&1 | grep -v "^\[bj\]" | grep "C3" | tee gpurun_out/configs_resident_$TAG.jsonl | cut -c1-330
echo "== C3 configs, grid-wide kernel"
PROXSDP_B200_LZ_RESIDENT=0 timeout 600 python scripts/bench_configs.py --no-cpu 2>&1 | grep -v "^\[bj\]" | grep "C3" | tee gpurun_out/configs_gridwide_$TAG.jsonl | cut -c1-330
echo "== racecheck, reduced driver, long spin limit"
PROXSDP_B200_LZ_SPIN_S=400 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_race.py 2>&1 | grep -v "^\[bj\]" | tail -12 | tee gpurun_out/racecheck_$TAG.txt
