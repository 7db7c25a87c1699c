#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PROXSDP_B200_LZ_RESIDENT=0
PROXSDP_B200_RITZ_MEM=1 PROXSDP_B200_RITZ_WARM=0 timeout 200 python scripts/dbg_hyp.py gpp500-1 2>&1 | grep -v "^\[bj\]" | tail -1
PROXSDP_B200_RITZ_MEM=1 PROXSDP_B200_LZ_ARROW=1 timeout 200 python scripts/dbg_hyp.py gpp500-1 2>&1 | grep -v "^\[bj\]" | tail -1
PROXSDP_B200_RITZ_BI=0 timeout 200 python scripts/dbg_hyp.py gpp500-1 2>&1 | grep -v "^\[bj\]" | tail -1
PROXSDP_B200_RITZ_BI=0 PROXSDP_B200_RITZ_WARM=0 timeout 200 python scripts/dbg_hyp.py gpp500-1 2>&1 | grep -v "^\[bj\]" | tail -1
PROXSDP_B200_RITZ_BI=0 PROXSDP_B200_LZ_ARROW=1 timeout 200 python scripts/dbg_hyp.py gpp500-1 2>&1 | grep -v "^\[bj\]" | tail -1
PROXSDP_B200_RITZ_MEM=1 PROXSDP_B200_LZ_STRICT=1 timeout 200 python scripts/dbg_hyp.py gpp500-1 2>&1 | grep -v "^\[bj\]" | tail -1
