#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2q}
echo "== lanczos tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lanczos or psd_projection or c2_fullsize or rank_sweep or krylov_sdplib or smoke" 2>&1 | tail -8
echo "== new parity tests (no full solves)"; timeout 900 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -k "not full_solve" 2>&1 | tail -8
echo "== lz prof 20 iterations"; PROXSDP_B200_LZ_PROF=1 PROXSDP_B200_DEBUG=1 timeout 120 python scripts/lz_prof.py 20 2>&1 | grep -v "^\[bj\]" | tail -24 | tee gpurun_out/lz_prof_$TAG.txt
echo "== bench 20/5"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench20_$TAG.json | cut -c1-400
