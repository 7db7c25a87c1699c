#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== dbg c2"; PROXSDP_B200_DEBUG=1 timeout 300 python scripts/dbg_c2.py 2>&1 | grep -v "^\[bj\]" | tail -50
echo "== seam tests"; timeout 600 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -k "seam" 2>&1 | tail -6
echo "== exact mode full solve mcp500"; timeout 900 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -x -s -k "full_solve_exact" --durations=3 2>&1 | tail -12
