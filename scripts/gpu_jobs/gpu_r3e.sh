#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== lanczos tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lanczos or rank_sweep or krylov_sdplib or c2_fullsize" 2>&1 | tail -3
for p in gpp500-1 mcp500-1; do
for v in default 0; do
  if [ "$v" = "default" ]; then unset PROXSDP_B200_LZ_RESIDENT; else export PROXSDP_B200_LZ_RESIDENT=$v; fi
  PROBLEM=$p timeout 300 python scripts/dbg_resident.py full 2>&1 | grep -v "^\[bj\]" | grep "status\|per iteration" | sed "s/^/$p RESIDENT=$v: /"
done; done
unset PROXSDP_B200_LZ_RESIDENT
for r in 5 10 16 25; do PROBLEM=gpp500-1 python scripts/dbg_resident.py $r 2>&1 | grep -v "^\[bj\]" | tail -1 | sed "s/^/gpp500-1: /"; done
