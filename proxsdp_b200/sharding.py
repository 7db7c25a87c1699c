"""Host side of the multi-GPU path: independent blocks of a stacked problem are
sharded across ranks, one process per GPU.

The reference is single-process, so there is nothing to mirror; the semantics are fixed by requiring that a
sharded solve walks through exactly the iterations of the un-sharded `chambolle_pock` on the whole problem:

  * `partition_blocks` finds the connected components of the variable / row / cone incidence graph (a MIMO
    batch stacks one component per instance) and deals them to ranks, heaviest first;
  * `shard_problem` extracts this rank's sub-problem (`AffineSets`, `ConicSets`) and the index maps;
  * the solver (`proxsdp_b200_solve_sharded`, or the oracle's counterpart in the CPU tests) combines every
    whole-problem scalar across ranks — norms, step size, line-search norms, residual / feasibility maxima,
    objective dot products, convergence flags — so all ranks take identical control decisions;
  * `merge_results` scatters the per-rank pieces back into whole-problem vectors.

torch.distributed is only plumbing here: rendezvous, the 128-byte NCCL id broadcast and the final gather of
result pieces.  No vector crosses ranks during the iterations.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_double, c_int64, c_void_p
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components

from .options import Options
from .structs import AffineSets, ConicSets, Result, SDPSet, SOCSet

REDUCE_FN = ctypes.CFUNCTYPE(None, POINTER(c_double), c_int64, c_int64, c_void_p)


class ShardPOD(ctypes.Structure):
    """include/proxsdp_b200_types.h: proxsdp_shard_t."""
    _fields_ = [
        ("rank", c_int64), ("nranks", c_int64),
        ("global_n", c_int64), ("global_p", c_int64), ("global_m", c_int64),
        ("comm", c_void_p), ("reduce", REDUCE_FN), ("reduce_ctx", c_void_p),
    ]


@dataclass
class ShardInfo:
    rank: int
    world: int
    var_idx: np.ndarray            # whole-problem variable ids of the local variables (ascending)
    eq_rows: np.ndarray            # whole-problem equality rows owned by this rank
    in_rows: np.ndarray
    sdp_ids: List[int] = field(default_factory=list)
    soc_ids: List[int] = field(default_factory=list)
    global_n: int = 0
    global_p: int = 0
    global_m: int = 0
    global_n_sdp: int = 0
    global_n_soc: int = 0


def _csr(mat, shape) -> sp.csr_matrix:
    if shape[0] == 0:
        return sp.csr_matrix(shape)
    return sp.csr_matrix(mat, shape=shape)


def partition_blocks(aff: AffineSets, con: ConicSets, world: int):
    """Connected components of the incidence graph, dealt to `world` ranks (heaviest first, least-loaded rank).
    Returns a list of `world` dicts with keys vars, eq_rows, in_rows, sdp_ids, soc_ids."""
    n, p, m = int(aff.n), int(aff.p), int(aff.m)
    nsd, nso = len(con.sdpcone), len(con.socone)
    A = _csr(aff.A, (p, n)).tocoo()
    G = _csr(aff.G, (m, n)).tocoo()
    # nodes: variables [0, n), eq rows [n, n+p), in rows [n+p, n+p+m), cones after that
    N = n + p + m + nsd + nso
    src = [A.col, G.col]
    dst = [A.row + n, G.row + n + p]
    for k, s in enumerate(con.sdpcone):
        src.append(np.asarray(s.vec_i, dtype=np.int64))
        dst.append(np.full(len(s.vec_i), n + p + m + k, dtype=np.int64))
    for k, s in enumerate(con.socone):
        src.append(np.asarray(s.idx, dtype=np.int64))
        dst.append(np.full(len(s.idx), n + p + m + nsd + k, dtype=np.int64))
    src = np.concatenate(src) if src else np.zeros(0, dtype=np.int64)
    dst = np.concatenate(dst) if dst else np.zeros(0, dtype=np.int64)
    graph = sp.coo_matrix((np.ones(len(src), dtype=np.int8), (src, dst)), shape=(N, N)).tocsr()
    ncomp, label = connected_components(graph, directed=False)
    # cost per component: eigen work ~ side^3, everything else ~ size
    cost = np.bincount(label, minlength=ncomp).astype(np.float64)
    for k, s in enumerate(con.sdpcone):
        cost[label[n + p + m + k]] += float(s.sq_side) ** 3
    order = sorted(range(ncomp), key=lambda c: (-cost[c], c))
    load = np.zeros(world)
    owner = np.zeros(ncomp, dtype=np.int64)
    for c in order:
        r = int(np.argmin(load))       # ties -> lowest rank: deterministic on every process
        owner[c] = r
        load[r] += cost[c]
    node_owner = owner[label]
    parts = []
    for r in range(world):
        parts.append(dict(
            vars=np.nonzero(node_owner[:n] == r)[0].astype(np.int64),
            eq_rows=np.nonzero(node_owner[n:n + p] == r)[0].astype(np.int64),
            in_rows=np.nonzero(node_owner[n + p:n + p + m] == r)[0].astype(np.int64),
            sdp_ids=[k for k in range(nsd) if node_owner[n + p + m + k] == r],
            soc_ids=[k for k in range(nso) if node_owner[n + p + m + nsd + k] == r],
        ))
    return parts


def shard_problem(aff: AffineSets, con: ConicSets, rank: int, world: int,
                  parts=None) -> Tuple[AffineSets, ConicSets, ShardInfo]:
    """This rank's sub-problem and its index maps."""
    n, p, m = int(aff.n), int(aff.p), int(aff.m)
    parts = parts if parts is not None else partition_blocks(aff, con, world)
    me = parts[rank]
    v, er, ir = me["vars"], me["eq_rows"], me["in_rows"]
    A = _csr(aff.A, (p, n))
    G = _csr(aff.G, (m, n))
    A_loc = A[er][:, v].tocsc() if len(er) else sp.csc_matrix((0, len(v)))
    G_loc = G[ir][:, v].tocsc() if len(ir) else sp.csc_matrix((0, len(v)))
    b = np.asarray(aff.b, dtype=np.float64)[er] if p else np.zeros(0)
    h = np.asarray(aff.h, dtype=np.float64)[ir] if m else np.zeros(0)
    c = np.asarray(aff.c, dtype=np.float64)[v]
    aff_loc = AffineSets(len(v), len(er), len(ir), 0, A_loc, G_loc, b, h, c)
    sdp = []
    for k in me["sdp_ids"]:
        s = con.sdpcone[k]
        loc = np.searchsorted(v, np.asarray(s.vec_i, dtype=np.int64))
        sdp.append(SDPSet(loc.astype(np.int64), s.tri_len, s.sq_side))
    soc = []
    for k in me["soc_ids"]:
        s = con.socone[k]
        loc = np.searchsorted(v, np.asarray(s.idx, dtype=np.int64))
        soc.append(SOCSet(loc.astype(np.int64), s.len))
    info = ShardInfo(rank, world, v, er, ir, list(me["sdp_ids"]), list(me["soc_ids"]), n, p, m,
                     len(con.sdpcone), len(con.socone))
    return aff_loc, ConicSets(sdp, soc), info


def merge_results(pieces: Sequence[Tuple[ShardInfo, Result]]) -> Result:
    """Scatter the per-rank result pieces into whole-problem vectors.  Scalars are identical on all ranks."""
    info0, r0 = pieces[0]
    n, p, m = info0.global_n, info0.global_p, info0.global_m
    out = Result(
        status=r0.status, status_string=r0.status_string,
        primal=np.zeros(n), dual_cone=np.zeros(n), dual_eq=np.zeros(p), dual_in=np.zeros(m),
        slack_eq=np.zeros(p), slack_in=np.zeros(m),
        primal_residual=r0.primal_residual, dual_residual=r0.dual_residual, objval=r0.objval,
        dual_objval=r0.dual_objval, gap=r0.gap, time=max(r.time for _, r in pieces), iter=r0.iter,
        final_rank=r0.final_rank, primal_feasible_user_tol=r0.primal_feasible_user_tol,
        dual_feasible_user_tol=r0.dual_feasible_user_tol, certificate_found=r0.certificate_found,
        result_count=r0.result_count, final_primal_res=r0.final_primal_res, final_dual_res=r0.final_dual_res,
        time_setup=max(r.time_setup for _, r in pieces), time_loop=max(r.time_loop for _, r in pieces),
        time_psd_proj=max(r.time_psd_proj for _, r in pieces), n_psd_proj=r0.n_psd_proj,
        lanczos_matvecs=sum(r.lanczos_matvecs for _, r in pieces), lanczos_calls=sum(r.lanczos_calls for _, r in pieces),
        full_eig_calls=sum(r.full_eig_calls for _, r in pieces), linesearch_trials=r0.linesearch_trials,
        gpu_launches=sum(r.gpu_launches for _, r in pieces),
        target_rank=np.zeros(info0.global_n_sdp, dtype=np.int64), trace=r0.trace,
    )
    for info, r in pieces:
        out.primal[info.var_idx] = r.primal
        out.dual_cone[info.var_idx] = r.dual_cone
        out.dual_eq[info.eq_rows] = r.dual_eq
        out.slack_eq[info.eq_rows] = r.slack_eq
        out.dual_in[info.in_rows] = r.dual_in
        out.slack_in[info.in_rows] = r.slack_in
        if r.target_rank is not None and len(info.sdp_ids):
            out.target_rank[np.asarray(info.sdp_ids, dtype=np.int64)] = r.target_rank[: len(info.sdp_ids)]
    return out


# ------------------------------------------------------------------------------------------------
# drivers
# ------------------------------------------------------------------------------------------------
_comm_cache = {}


def _nccl_comm(device_id: int, group=None):
    """One NCCL communicator per (process, group): rank 0 draws the id, torch.distributed broadcasts it."""
    import torch.distributed as dist
    from . import solver
    key = (id(group), device_id)
    if key in _comm_cache:
        return _comm_cache[key]
    L = solver.lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        solver._check(L.proxsdp_b200_comm_unique_id(buf))
    obj = [bytes(buf.raw) if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0, group=group)
    comm = c_void_p()
    L.proxsdp_b200_comm_create.argtypes = [ctypes.c_char_p, c_int64, c_int64, c_int64, POINTER(c_void_p)]
    solver._check(L.proxsdp_b200_comm_create(obj[0], rank, world, device_id, ctypes.byref(comm)))
    _comm_cache[key] = comm
    return comm


def solve_local_shard(aff_loc: AffineSets, con_loc: ConicSets, opt: Options, rank: int, world: int,
                      global_n: int, global_p: int, global_m: int, group=None, device_id: Optional[int] = None) -> Result:
    """`proxsdp_b200_solve_sharded` on this rank's blocks: the local sub-problem plus the sizes of the whole problem.
    For callers that generate their shard directly (e.g. a weak-scaling batch: every rank builds its own instances)."""
    from . import solver
    from ._abi import MarshalledProblem, MarshalledResult, OptionsPOD, ProblemPOD, ResultPOD
    L = solver.lib()
    if device_id is None:
        import torch
        device_id = torch.cuda.current_device()
    comm = _nccl_comm(device_id, group) if world > 1 else c_void_p()
    shard = ShardPOD(rank, world, global_n, global_p, global_m, comm, ctypes.cast(None, REDUCE_FN), None)
    mp = MarshalledProblem(aff_loc, con_loc)
    mr = MarshalledResult(mp.n, mp.p, mp.m, mp.n_sdp, int(opt.trace_cap))
    o = opt.copy()
    o.device_id = device_id
    opod = o.to_pod()
    L.proxsdp_b200_solve_sharded.argtypes = [POINTER(ProblemPOD), POINTER(OptionsPOD), POINTER(ShardPOD), POINTER(ResultPOD)]
    L.proxsdp_b200_solve_sharded.restype = ctypes.c_int
    solver._check(L.proxsdp_b200_solve_sharded(ctypes.byref(mp.pod), ctypes.byref(opod), ctypes.byref(shard),
                                               ctypes.byref(mr.pod)))
    return mr.to_result()


def chambolle_pock_sharded(aff: AffineSets, con: ConicSets, opt: Options, group=None, device_id: Optional[int] = None,
                           local_solve: Optional[Callable] = None, gather: bool = True) -> Result:
    """`chambolle_pock` on the whole problem, executed by all ranks of a torch.distributed group.

    Every rank passes the SAME whole problem (it is small on the host); each solves its blocks on its GPU.
    `local_solve(aff_loc, con_loc, opt, info)` overrides the per-rank engine (the CPU tests plug the oracle in).
    Returns the merged whole-problem Result on every rank (or this rank's piece when gather=False)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    aff_loc, con_loc, info = shard_problem(aff, con, rank, world)
    if local_solve is not None:
        res = local_solve(aff_loc, con_loc, opt, info)
    else:
        res = solve_local_shard(aff_loc, con_loc, opt, rank, world, info.global_n, info.global_p, info.global_m,
                                group=group, device_id=device_id)
    if not gather:
        return res
    pieces = [None] * world
    if world > 1:
        dist.all_gather_object(pieces, (info, res), group=group)
    else:
        pieces = [(info, res)]
    return merge_results(pieces)
