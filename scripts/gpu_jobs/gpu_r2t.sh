#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for C in 8 4 6; do
  echo "== cluster $C"; PROXSDP_B200_CLUSTER=$C PROXSDP_B200_DEBUG=1 timeout 300 python scripts/lz_large.py 2000 5000 2>&1 | grep -v "^\[bj\]" | cut -c1-420
done | tee gpurun_out/lz_large_clusters_r2t.txt
