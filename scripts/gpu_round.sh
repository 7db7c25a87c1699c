#!/bin/bash
# One GPU-box visit: smoke, remaining parity tests, bench, ncu launch list + full capture of the Lanczos kernel.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r1}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks_$TAG.csv &
SMI=$!
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== bench (flush)"; python bench.py --steps 300 --warmup 10 2>&1 | tail -3 | tee gpurun_out/bench_$TAG.json
echo "== bench (no flush)"; python bench.py --steps 300 --warmup 10 --no-flush-l2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_noflush_$TAG.json
echo "== bench reference"; python bench.py --impl reference --steps 300 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_ref_$TAG.json
kill $SMI
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== pytest gpu"; python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_$TAG.log; tail -15 gpurun_out/pytest_gpu_$TAG.log
fi
echo "== lanczos on matrices larger than L2"; timeout 600 python scripts/lz_large.py 2>&1 | tee gpurun_out/lz_large_$TAG.jsonl | cut -c1-200
echo "== lanczos per-phase profile"; timeout 120 python scripts/lz_prof.py > gpurun_out/lz_prof_$TAG.txt 2>&1
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-flush-l2 > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "== ncu full (lanczos)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lanczos_cl3 -s 20 -c 1 -o gpurun_out/prof_lanczos_$TAG -f python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-flush-l2 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out
fi
