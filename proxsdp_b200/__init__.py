"""proxsdp_b200 — B200-native hot path of ProxSDP behind the `chambolle_pock` seam.

Host-side mirror of the reference's interface for this path:
  Options, AffineSets, ConicSets, SDPSet, SOCSet, Result  (reference src/options.jl, src/structs.jl)
  chambolle_pock(aff, con, opt) -> Result                 (reference src/pdhg.jl:1)
  Optimizer                                               (reference src/MOI_wrapper.jl:54-74)
The arithmetic lives in the CUDA library `libproxsdp_b200.so` (csrc/), reached through the
C ABI declared in include/proxsdp_b200.h.
"""
from .options import Options
from .structs import AffineSets, ConicSets, Result, SDPSet, SOCSet, SparseMatrixCSC, ivec, ivech, sympackeddim, sympackedlen
from .model import MAX_SENSE, MIN_SENSE, Optimizer

__all__ = [
    "Options", "AffineSets", "ConicSets", "Result", "SDPSet", "SOCSet", "ivec", "ivech",
    "sympackeddim", "sympackedlen", "Optimizer", "MIN_SENSE", "MAX_SENSE", "chambolle_pock",
]


def chambolle_pock(aff, con, opt, **kw):
    from .solver import chambolle_pock as _cp
    return _cp(aff, con, opt, **kw)
