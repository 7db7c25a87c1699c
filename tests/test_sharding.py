"""CPU tests of the multi-GPU host logic: partitioning, sub-problem extraction, merge, and the
whole-problem scalar reductions — world_size 2 and 3 over gloo, with the CPU oracle as the per-rank engine."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from proxsdp_b200.problems import mimo_problem, sensorloc_problem, stack_problems
from proxsdp_b200.sharding import merge_results, partition_blocks, shard_problem
from proxsdp_b200.structs import Result

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_partition_covers_everything_once():
    probs = [mimo_problem(s, 4 + s % 3) for s in range(9)]
    aff, con = stack_problems(probs)
    for world in (1, 2, 4):
        parts = partition_blocks(aff, con, world)
        allv = np.concatenate([p["vars"] for p in parts])
        assert sorted(allv.tolist()) == list(range(aff.n))
        assert sorted(np.concatenate([p["eq_rows"] for p in parts]).tolist()) == list(range(aff.p))
        assert sorted(np.concatenate([p["in_rows"] for p in parts]).tolist()) == list(range(aff.m))
        assert sorted(k for p in parts for k in p["sdp_ids"]) == list(range(len(con.sdpcone)))
        # blocks are independent: no row of a rank touches another rank's variables
        A = aff.A.tocsr()
        for r, p in enumerate(parts):
            cols = set(p["vars"].tolist())
            for i in p["eq_rows"]:
                assert set(A[i].indices.tolist()) <= cols


def test_shard_and_merge_roundtrip():
    probs = [mimo_problem(s, 5) for s in range(4)] + [sensorloc_problem(1, 6, soc_variant=True)]
    aff, con = stack_problems(probs)
    world = 3
    rng = np.random.default_rng(0)
    x = rng.standard_normal(aff.n)
    pieces = []
    for r in range(world):
        a, c, info = shard_problem(aff, con, r, world)
        assert a.n == len(info.var_idx) and a.p == len(info.eq_rows) and a.m == len(info.in_rows)
        # local cone index lists point at the same whole-problem variables
        for s_loc, k in zip(c.sdpcone, info.sdp_ids):
            assert np.array_equal(info.var_idx[s_loc.vec_i], np.asarray(con.sdpcone[k].vec_i))
        assert np.allclose(a.c, np.asarray(aff.c)[info.var_idx])
        res = Result(status=1, primal=x[info.var_idx], dual_cone=x[info.var_idx], dual_eq=np.zeros(a.p), dual_in=np.zeros(a.m),
                     slack_eq=np.zeros(a.p), slack_in=np.zeros(a.m), target_rank=np.full(len(info.sdp_ids), 2), trace=np.zeros((0, 14)))
        pieces.append((info, res))
    merged = merge_results(pieces)
    assert np.array_equal(merged.primal, x)


@pytest.mark.parametrize("world,case", [(2, "mimo"), (3, "mimo"), (2, "mixed"), (2, "mimo_exact_norm")])
def test_sharded_oracle_equals_whole_solve(world, case):
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_shard_worker.py"), case], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for pr in procs:
        try:
            out, _ = pr.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for rank, (pr, out) in enumerate(zip(procs, outs)):
        assert pr.returncode == 0, f"rank {rank} failed:\n{out[-3000:]}"
    assert "== whole" in outs[0]
