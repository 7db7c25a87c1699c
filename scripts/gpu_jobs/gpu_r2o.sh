#!/bin/bash
# two-GPU validation of the sharded path: the NCCL parity test and the bench at N = 2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
echo "== sharded two-GPU test"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sharded_two_gpus" 2>&1 | tail -15
echo "== bench --gpus 2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_r2f.json 2> gpurun_out/bench_n2_r2f.err
echo "rc=$?"; tail -c 6000 gpurun_out/bench_n2_r2f.json; tail -20 gpurun_out/bench_n2_r2f.err
echo "== bench --impl reference --gpus 2 (rank 0 only)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>&1 | tail -3 | cut -c1-600
