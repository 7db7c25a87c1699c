#!/bin/bash
# compute-sanitizer: memcheck and synccheck over the small driver, racecheck (shared-memory hazards) over the reduced one.
# The Lanczos kernel spins on flagged words written by other CTAs; all spin loops are bounded, and each tool gets its own timeout.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2}
{
for tool in memcheck synccheck; do
  echo "== compute-sanitizer --tool $tool python scripts/sanitize_small.py"
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py 2>&1 | grep -v "^\[bj\]" | tail -14
done
echo "== compute-sanitizer --tool racecheck python scripts/sanitize_race.py"
PROXSDP_B200_LZ_SPIN_S=400 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_race.py 2>&1 | grep -v "^\[bj\]" | tail -24
echo "rc=$?"
} | tee gpurun_out/sanitizer_$TAG.txt
