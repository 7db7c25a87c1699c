#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAGX=default timeout 200 python scripts/dbg_c2.py 1500 2>&1 | grep -v "^\[bj\]" | tail -16
echo "== lanczos tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -m gpu -q -x -k "lanczos or psd_projection or c2_fullsize or rank_sweep or krylov or min_size" 2>&1 | tail -6
echo "== full solves vs golden"; timeout 900 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -x -s -k "full_solve_c2" --durations=5 2>&1 | tail -12
