"""Generates the committed problem fixtures under tests/golden/.

Run in the build container only (it reads the SDPLIB instances that ship with the
reference under /root/reference/test/data; that path does not exist on the GPU box).
Each fixture is the exact argument set of `chambolle_pock` (AffineSets + ConicSets) as
the reference's `jump_sdplib` (test/jump_sdplib.jl:5-20) + `_optimize!`
(src/MOI_wrapper.jl:229-292) assemble it, stored with proxsdp_b200.problems.save_problem.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from proxsdp_b200.problems import save_problem, sdplib_problem  # noqa: E402

REF_DATA = "/root/reference/test/data"
NAMES = ["mcp124-1", "gpp124-2", "mcp250-1", "mcp500-1", "gpp500-1", "maxG32"]

if __name__ == "__main__":
    for name in NAMES:
        aff, con = sdplib_problem(os.path.join(REF_DATA, name + ".dat-s"))
        out = os.path.join(HERE, f"sdplib_{name}.npz")
        save_problem(out, aff, con)
        print(name, "n", aff.n, "p", aff.p, "nnz", aff.A.nnz, "->", os.path.getsize(out), "bytes")
