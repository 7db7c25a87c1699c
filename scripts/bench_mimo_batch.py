"""Config C4 (BASELINE.json): a batch of independent MIMO detection SDPs (PSD side n+1) stacked into one problem
and sharded across the GPUs of one node, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_mimo_batch.py --batch 256 --side 64 --iters 300 [--check]

Prints one JSON line: PDHG iterations/s of the WHOLE stacked problem (all ranks advance the same iteration),
and with --check the parity of the merged result against the un-sharded single-GPU solve on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from proxsdp_b200 import Options, solver  # noqa: E402
from proxsdp_b200.problems import mimo_problem, stack_problems  # noqa: E402
from proxsdp_b200.sharding import chambolle_pock_sharded  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--side", type=int, default=64, dest="n")
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    probs = [mimo_problem(1000 + s, args.n) for s in range(args.batch)]
    aff, con = stack_problems(probs)
    opt = Options(max_iter=args.iters)
    chambolle_pock_sharded(aff, con, Options(max_iter=20), device_id=local_rank)        # warm-up (communicator, modules)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = chambolle_pock_sharded(aff, con, opt, device_id=local_rank)
    torch.cuda.synchronize(); dist.barrier()
    wall = time.perf_counter() - t0
    t = torch.tensor([res.time_loop, wall], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    line = {
        "metric": "pdhg_iterations_per_sec_mimo_batch", "workload": f"{args.batch} x MIMO n={args.n} (PSD side {args.n + 1}) stacked",
        "n_gpus": world, "iterations": int(res.iter), "loop_s": float(t[0]), "wall_s": float(t[1]),
        "value": res.iter / float(t[0]), "unit": "iterations/s (whole batch)", "cone_projections_per_s": res.iter * args.batch / float(t[0]),
        "status": int(res.status), "objval": res.objval, "gap": res.gap,
    }
    if args.check and rank == 0:
        ref = solver.chambolle_pock(aff, con, Options(max_iter=args.iters, device_id=local_rank))
        line["check"] = {
            "iter_equal": bool(ref.iter == res.iter and ref.status == res.status),
            "objval_rel_diff": abs(ref.objval - res.objval) / max(1.0, abs(ref.objval)),
            "primal_max_abs_diff": float(np.abs(ref.primal - res.primal).max()),
            "dual_eq_max_abs_diff": float(np.abs(ref.dual_eq - res.dual_eq).max()),
            "single_gpu_loop_s": ref.time_loop,
        }
        assert line["check"]["iter_equal"], line
        assert line["check"]["objval_rel_diff"] <= 1e-8 and line["check"]["primal_max_abs_diff"] <= 1e-7, line
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
