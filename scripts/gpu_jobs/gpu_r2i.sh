#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAGX=default timeout 200 python scripts/dbg_c2.py 1500 2>&1 | grep -v "^\[bj\]" | tail -18
TAGX=arrow PROXSDP_B200_LZ_ARROW=1 timeout 200 python scripts/dbg_c2.py 1500 2>&1 | grep -v "^\[bj\]" | tail -18
TAGX=strict PROXSDP_B200_LZ_STRICT=1 timeout 200 python scripts/dbg_c2.py 1500 2>&1 | grep -v "^\[bj\]" | tail -18
