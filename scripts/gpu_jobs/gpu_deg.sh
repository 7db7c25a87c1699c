#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for env in "X=1" "PROXSDP_B200_LZ_ARROW=1" "PROXSDP_B200_LZ_STRICT=1" "PROXSDP_B200_RITZ_BI=0" "PROXSDP_B200_LZ_RESIDENT=0" "PROXSDP_B200_LZ_KERNEL=2"; do
  for cfg in "300 6" "520 10" "900 8"; do
    echo "## $env  n,nev = $cfg"
    env $env python scripts/dbg_degenerate.py $cfg 2>&1 | grep -v "^\[bj\]\|^\[lz\]" | grep "gpu:\|oracle:" | cut -c1-260
  done
done
