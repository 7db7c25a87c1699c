#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2p}
echo "== lz prof 20 iterations"; PROXSDP_B200_LZ_PROF=1 PROXSDP_B200_DEBUG=1 timeout 120 python scripts/lz_prof.py 20 2>&1 | grep -v "^\[bj\]" | tail -32 | tee gpurun_out/lz_prof_$TAG.txt
