#!/bin/bash
# One short GPU visit for Lanczos kernel work: eigsolve parity tests, per-phase profile of four CTAs, un-flushed bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== lanczos parity tests"; timeout 600 python -m pytest tests -m gpu -q -x -k "lanczos or psd_projection" 2>&1 | tail -4
echo "== per-phase profile (PROXSDP_B200_LZ_PROF=1)"; timeout 120 python scripts/lz_prof.py 2>&1 | tail -24
echo "== bench, L2 not flushed"; timeout 300 python bench.py --steps 300 --warmup 10 --no-flush-l2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_noflush_quick.json | cut -c1-300
