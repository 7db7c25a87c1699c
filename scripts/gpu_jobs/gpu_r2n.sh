#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== implicit operator + arpack tests"; PROXSDP_B200_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -x -s -k "implicit" 2>&1 | grep -v "^\[bj\]" | tail -30
echo "== lanczos tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lanczos or c2_fullsize" 2>&1 | tail -4
