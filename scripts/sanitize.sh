#!/bin/bash
# compute-sanitizer over the small driver: memcheck, synccheck, racecheck (shared-memory hazards).  The Lanczos kernel
# spins on flagged words written by other CTAs; all spin loops are bounded, and each tool gets its own timeout.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 240 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py 2>&1 | grep -v "^\[bj\]" | tail -14
done
