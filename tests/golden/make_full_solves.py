"""Writes tests/golden/full_{c2,maxG32,mcp500-1_exact}.npz: the CPU oracle's solve TO TERMINATION of the headline
instances (status, iteration count, final scalars, per-iteration trace).  The oracle (oracle/: C restatement of the
reference) needs minutes per instance on 16 cores, which is why its output is committed as a fixture instead of
being recomputed inside `pytest -m gpu`; tests/test_gpu_parity_full.py compares the CUDA path with these files.

    python tests/golden/make_full_solves.py [c2] [maxG32] [mcp500-1_exact]      (default: all three)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from proxsdp_b200 import Options  # noqa: E402
from proxsdp_b200.problems import load_problem, maxcut_er_problem  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SCALARS = ("objval", "dual_objval", "gap", "primal_residual", "dual_residual", "final_primal_res", "final_dual_res")


# Options of the two n = 2000 Krylov solves: the reference's own `max_target_rank_krylov_eigs` raised from 16 to 32.  With
# the default, both instances bump their target rank to 17 near the end and the reference then takes a FULL 2000 x 2000
# eigendecomposition in each of the last few hundred iterations (prox_operators.jl:47): ~30 minutes for the CPU oracle and
# minutes for the GPU's block-Jacobi fallback, which no test run can afford; the option keeps the truncated projection.
KRYLOV_OPTS = dict(max_target_rank_krylov_eigs=32)


def problem(name):
    if name == "c2":
        return maxcut_er_problem(2000, 0.01, 0), dict(KRYLOV_OPTS)
    if name == "maxG32":
        return load_problem(os.path.join(HERE, "sdplib_maxG32.npz")), dict(KRYLOV_OPTS)
    if name == "mcp500-1_exact":
        return load_problem(os.path.join(HERE, "sdplib_mcp500-1.npz")), dict(full_eig_decomp=True)
    raise SystemExit(f"unknown instance {name}")


def main():
    oracle.build()
    try:
        oracle.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for name in (sys.argv[1:] or ["c2", "maxG32", "mcp500-1_exact"]):
        (aff, con), kw = problem(name)
        t0 = time.time()
        r = oracle.chambolle_pock(aff, con, Options(trace_cap=200000, **kw))
        out = dict(status=r.status, iters=r.iter, trace=r.trace.astype(np.float64), final_rank=r.final_rank,
                   target_rank=r.target_rank, lanczos_matvecs=r.lanczos_matvecs, options=repr(kw))
        out.update({k: getattr(r, k) for k in SCALARS})
        np.savez_compressed(os.path.join(HERE, f"full_{name}.npz"), **out)
        print(f"{name}: status {r.status} ({r.status_string}) after {r.iter} iterations, objective {r.objval:.9g}, "
              f"gap {r.gap:.3e}, {r.lanczos_matvecs} mat-vecs, {time.time() - t0:.1f} s on {oracle.num_threads()} threads",
              flush=True)


if __name__ == "__main__":
    main()
