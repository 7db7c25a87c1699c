"""Worker of tests/test_sharding.py: one rank of a world_size-N gloo group solving a stacked MIMO batch with the
sharded host logic, the CPU oracle as the per-rank engine.  Rank 0 checks the merged result against the
un-sharded oracle solve of the whole problem."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from proxsdp_b200 import Options  # noqa: E402
from proxsdp_b200.problems import mimo_problem, sensorloc_problem, stack_problems  # noqa: E402
from proxsdp_b200.sharding import chambolle_pock_sharded, partition_blocks  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    case = sys.argv[1] if len(sys.argv) > 1 else "mimo"
    if case == "mimo":
        probs = [mimo_problem(200 + s, 6 + (s % 3)) for s in range(7)]          # ragged batch, 7 blocks on 2-3 ranks
        opt = Options(trace_cap=50)
    elif case == "mimo_exact_norm":
        probs = [mimo_problem(200 + s, 6 + (s % 3)) for s in range(5)]
        opt = Options(trace_cap=50, approx_norm=False)          # sigma_max of the block-diagonal M = max over the ranks
    else:
        probs = [sensorloc_problem(3, 8, soc_variant=True), mimo_problem(11, 5), sensorloc_problem(4, 6)]   # SOC + PSD blocks
        opt = Options(max_iter=400, trace_cap=50)
    aff, con = stack_problems(probs)

    def reduce_fn(arr, op):
        t = torch.from_numpy(arr)
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX)

    def local_solve(a, c, o, info):
        return oracle.chambolle_pock_sharded(a, c, o, info, reduce_fn)

    res = chambolle_pock_sharded(aff, con, opt, local_solve=local_solve)
    parts = partition_blocks(aff, con, world)
    assert sum(len(p["vars"]) for p in parts) == aff.n and sum(len(p["sdp_ids"]) for p in parts) == len(con.sdpcone)
    assert all(len(p["sdp_ids"]) + len(p["soc_ids"]) > 0 for p in parts), "every rank should own a block"
    if rank == 0:
        ref = oracle.chambolle_pock(aff, con, opt)
        assert res.status == ref.status and res.iter == ref.iter, (res.status, ref.status, res.iter, ref.iter)
        for name in ("objval", "dual_objval", "gap", "primal_residual", "dual_residual", "final_primal_res", "final_dual_res"):
            a, b = getattr(res, name), getattr(ref, name)
            assert abs(a - b) <= 1e-8 * max(1.0, abs(b)), (name, a, b)
        assert res.final_rank == ref.final_rank
        assert np.abs(res.primal - ref.primal).max() <= 1e-8 * max(1.0, np.abs(ref.primal).max())
        assert np.abs(res.dual_cone - ref.dual_cone).max() <= 1e-8 * max(1.0, np.abs(ref.dual_cone).max())
        assert np.abs(res.dual_eq - ref.dual_eq).max() <= 1e-8 * max(1.0, np.abs(ref.dual_eq).max())
        if aff.m:
            assert np.abs(res.dual_in - ref.dual_in).max() <= 1e-8 * max(1.0, np.abs(ref.dual_in).max())
            assert np.abs(res.slack_in - ref.slack_in).max() <= 1e-8 * max(1.0, np.abs(ref.slack_in).max())
        k = min(len(res.trace), len(ref.trace))
        assert np.abs(res.trace[:k, 1:9] - ref.trace[:k, 1:9]).max() <= 1e-8 * max(1.0, np.abs(ref.trace[:k, 1:9]).max())
        print(f"sharded({world}) == whole: status {res.status}, iter {res.iter}, obj {res.objval:.9f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
