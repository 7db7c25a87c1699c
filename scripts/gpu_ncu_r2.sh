#!/bin/bash
# ncu evidence of round 2: launch list of the bench command, --set full capture of two steady-state PDHG iterations (every
# kernel of the step), and of the eigsolve kernel on a side-5000 cone (matrix larger than L2).  ncu cannot replay a launch
# that carries the cooperative attribute together with a cluster dimension ("LaunchFailed"), hence PROXSDP_B200_LZ_COOP=0.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2x}
echo "== ncu launch list"
PROXSDP_B200_LZ_COOP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-large-cone > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_launch_$TAG.log | cut -c1-300
echo "== ncu full: 26 consecutive launches of the timed region"
PROXSDP_B200_LZ_COOP=0 timeout 1500 ncu --set full --clock-control none --import-source on -s 200 -c 26 -o gpurun_out/prof_step_$TAG -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-large-cone > gpurun_out/ncu_full_step_$TAG.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full_step_$TAG.log | cut -c1-300
# (gpurun brings back at most 64 MiB: keep the raw-page CSV, drop the 50 MB report)
ncu -i gpurun_out/prof_step_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_step_${TAG}_raw.csv 2>/dev/null && rm -f gpurun_out/prof_step_$TAG.ncu-rep
echo "== ncu full: eigsolve kernel, side 5000"
PROXSDP_B200_LZ_COOP=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lanczos_cl3 -s 2 -c 1 -o gpurun_out/prof_lanczos5000_$TAG -f python scripts/lz_large.py 5000 > gpurun_out/ncu_full_5000_$TAG.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full_5000_$TAG.log | cut -c1-300
ncu -i gpurun_out/prof_lanczos5000_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_lanczos5000_${TAG}_raw.csv 2>/dev/null && rm -f gpurun_out/prof_lanczos5000_$TAG.ncu-rep
ls -la gpurun_out/*$TAG*
