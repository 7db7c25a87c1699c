"""numpy / LAPACK mirror of the oracle's exact-projection path.

TEST INFRASTRUCTURE ONLY.  An independent, deliberately plain restatement of the
reference loop (src/pdhg.jl:145-332, 532-637; src/prox_operators.jl; src/residuals.jl)
in which every PSD projection is a full `numpy.linalg.eigh` (LAPACK, the same family
as the reference's `LinearAlgebra.eigen!`).  It exists to pin the C oracle
(oracle/proxsdp_oracle.c): the two share no code and no eigensolver.

Covers: preprocess!/norm_scaling, advanced initialisation, primal step with full-eig
projection, SOC projection, linesearch / fixed dual step, residuals, gap, the optimality
test and the adaptive step logic.  Not covered (the C oracle restates them, this mirror
stops instead): certificate search and the infeasibility heuristics.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from proxsdp_b200.options import Options
from proxsdp_b200.structs import AffineSets, ConicSets


def _svec_maps(side):
    ii, jj = [], []
    for j in range(side):
        for i in range(j + 1):
            ii.append(i)
            jj.append(j)
    return np.array(ii), np.array(jj)


def equilibrate(M: sp.spmatrix, opt: Options):
    """equilibration.jl:1-71 with scipy sparse products in place of the reference's loops over nzrange."""
    R, n = M.shape
    alpha2, beta2, gamma = np.sqrt(n / R), np.sqrt(R / n), 0.1
    u, v, u_, v_ = np.zeros(R), np.zeros(n), np.zeros(R), np.zeros(n)
    M = sp.csr_matrix(M)
    for it in range(1, int(opt.equilibration_iters) + 1):
        M_ = sp.diags(np.exp(u)) @ (M @ sp.diags(np.exp(v)))
        sq = M_.multiply(M_)
        row_norms, col_norms = np.asarray(sq.sum(axis=1)).ravel(), np.asarray(sq.sum(axis=0)).ravel()
        step = 2.0 / (gamma * (it + 1.0))
        u = np.clip(u - step * (row_norms - alpha2 + gamma * u), opt.equilibration_lb, opt.equilibration_ub)
        v = v - step * (col_norms - beta2 + gamma * v)
        v = np.clip(np.full(n, v.sum() / n), 0.0, opt.equilibration_ub)
        u_ = 2.0 * u / (it + 2.0) + it * u_ / (it + 2.0)
        v_ = 2.0 * v / (it + 2.0) + it * v_ / (it + 2.0)
    return np.exp(u_), np.exp(v_)


def solve_exact(aff: AffineSets, con: ConicSets, opt: Options, max_iter: int):
    n, p, m = aff.n, aff.p, aff.m
    norm_b, norm_h, norm_c = np.linalg.norm(aff.b), np.linalg.norm(aff.h), np.linalg.norm(aff.c)
    # preprocess! (scaling.jl:2-26)
    cone_vars = [v for s in con.sdpcone for v in s.vec_i] + [v for s in con.socone for v in s.idx]
    rest = sorted(set(range(n)) - set(cone_vars))
    ord_ = np.array(cone_vars + rest, dtype=np.int64)
    A = sp.csc_matrix(aff.A)[:, ord_] if p else sp.csc_matrix((0, n))
    G = sp.csc_matrix(aff.G)[:, ord_] if m else sp.csc_matrix((0, n))
    c = np.asarray(aff.c, dtype=float)[ord_].copy()
    b, h = np.asarray(aff.b, float).copy(), np.asarray(aff.h, float).copy()
    # diagonal preconditioning (pdhg.jl:64-93)
    use_eq = bool(opt.equilibration)
    M0 = sp.vstack([A, G]).tocsc()
    if use_eq:
        dense_min = min(M0.data.min(initial=np.inf), 0.0 if M0.nnz < M0.shape[0] * M0.shape[1] else np.inf)
        dense_max = max(M0.data.max(initial=-np.inf), 0.0 if M0.nnz < M0.shape[0] * M0.shape[1] else -np.inf)
        if dense_min / dense_max <= opt.equilibration_limit:
            use_eq = False
    if opt.equilibration_force:
        use_eq = True
    E_eq = D_eq = None
    if use_eq:
        E_eq, D_eq = equilibrate(M0, opt)
        Ms = sp.diags(E_eq) @ M0 @ sp.diags(D_eq)
        A, G = sp.csc_matrix(Ms[:p]), sp.csc_matrix(Ms[p:])
        b, h = E_eq[:p] * b, E_eq[p:] * h
        c = D_eq * c
    # norm_scaling (scaling.jl:28-58)
    scale = np.ones(n)
    off = 0
    blocks = []
    for s in con.sdpcone:
        ii, jj = _svec_maps(s.sq_side)
        scale[off:off + s.tri_len][ii != jj] = np.sqrt(2.0) / 2.0
        blocks.append((off, s.sq_side, ii, jj))
        off += s.tri_len
    socs = []
    for s in con.socone:
        socs.append((off, s.len))
        off += s.len
    D = sp.diags(scale)
    M = sp.vstack([A @ D, G @ D]).tocsr()
    Mt = M.T.tocsr()
    c = c * scale
    fro = np.sqrt((M.data ** 2).sum())
    if not opt.approx_norm:                      # pdhg.jl:107-118: Arpack.svds(M, nsv = 1), dense svd below two rows / columns
        import scipy.sparse.linalg as spla
        if min(M.shape) >= 2:
            fro = float(spla.svds(M.tocsc(), k=1, tol=0, return_singular_vectors=False)[0])
        else:
            fro = float(np.linalg.svd(M.toarray(), compute_uv=False).max(initial=0.0))
    tau = 1.0 / (fro if fro >= 1e-10 else 1.0)
    tau_old, sigma, theta, beta, adapt = tau, tau, opt.initial_theta, opt.initial_beta, opt.initial_adapt_level
    R = p + m
    x = tau * c if opt.advanced_initialization else np.zeros(n)
    x_old = np.zeros(n)
    y = np.zeros(R); y_old = np.zeros(R)
    Mty = np.zeros(n); Mty_old = np.zeros(n)
    Mx = M @ x; Mx_old = np.zeros(R)
    sqrt2 = np.sqrt(2.0)
    window = opt.convergence_window
    trace = []
    status = 3
    ada_count = 0
    equa, ineq = 0.0, 0.0
    nsd = len(con.sdpcone)
    target_rank = [2] * nsd
    current_rank = [2] * nsd
    min_eig = [0.0] * nsd
    rank_update, update_cont = 0, 0
    comb = {}

    def rank_rule(idx):   # pdhg.jl:271-279
        if current_rank[idx] + opt.rank_slack >= target_rank[idx] and min_eig[idx] > opt.tol_psd:
            t = opt.rank_increment_factor * target_rank[idx] if opt.rank_increment == 0 else opt.rank_increment_factor + target_rank[idx]
            target_rank[idx] = min(t, con.sdpcone[idx].sq_side)
    for k in range(1, max_iter + 1):
        # primal_step!
        x = x - tau * (Mty + c)
        for idx, (o, side, ii, jj) in enumerate(blocks):
            v = x[o:o + len(ii)]
            current_rank[idx] = 0
            min_eig[idx] = 0.0
            if side == 1:
                x[o] = max(0.0, v[0])
                min_eig[idx] = x[o]
                continue
            X = np.zeros((side, side))
            vals = np.where(ii != jj, v / sqrt2, v)
            X[ii, jj] = vals
            X[jj, ii] = vals
            w, V = np.linalg.eigh(X)
            pos = w > 0
            current_rank[idx] = int((w > opt.tol_psd).sum())
            Xp = (V[:, pos] * w[pos]) @ V[:, pos].T
            out = Xp[ii, jj]
            x[o:o + len(ii)] = np.where(ii != jj, out * sqrt2, out)
        for (o, ln) in socs:
            t, v = x[o], x[o + 1:o + ln]
            nv = np.linalg.norm(v)
            if nv <= -t:
                x[o:o + ln] = 0.0
            elif nv <= t:
                pass
            else:
                val = 0.5 * (1.0 + t / nv)
                x[o + 1:o + ln] = v * val
                x[o] = val * nv
        Mx = M @ x
        # linesearch! / dual_step!
        if opt.line_search_flag:
            tau = tau * np.sqrt(1.0 + theta)
            for _ in range(opt.max_linsearch_steps):
                theta = tau / tau_old
                bt = beta * tau
                y_half = y + bt * ((1.0 + theta) * Mx - theta * Mx_old)
                proj = np.concatenate([b, np.minimum(y_half[p:] / bt, h)])
                y_temp = y_half - bt * proj
                Mty = Mt @ y_temp
                if np.sqrt(beta) * tau * np.linalg.norm(Mty - Mty_old) <= opt.delta * np.linalg.norm(y_temp - y_old):
                    break
                tau *= opt.linsearch_decay
            y = y_temp
            tau_old = tau
            sigma = beta * tau
        else:
            y_half = y + sigma * (2.0 * Mx - Mx_old)
            proj = np.concatenate([b, np.minimum(y_half[p:] / sigma, h)])
            y = y_half - sigma * proj
            Mty = Mt @ y
            tau_old = tau
        # compute_residual!
        Pold = x_old - tau * Mty_old
        Pnew = x - tau * Mty
        pr = np.sqrt(n) * np.abs(Pnew - Pold).max(initial=0.0) / max(np.abs(Pold).max(initial=0.0), norm_b, norm_h, 1.0)
        Dold = y_old - sigma * Mx_old
        Dnew = y - sigma * Mx
        dr = np.sqrt(R) * np.abs(Dnew - Dold).max(initial=0.0) / max(np.abs(Dold).max(initial=0.0), norm_c, 1.0)
        x_old, y_old, Mty_old, Mx_old = x.copy(), y.copy(), Mty.copy(), Mx.copy()
        # compute_gap!
        if p:
            equa = np.abs(Mx[:p] - b).max() / (1.0 + norm_b)
        if m:
            ineq = max(0.0, (Mx[p:] - h).max()) / (1.0 + norm_h)
        feas = max(equa, ineq)
        po = float(c @ x)
        do = -float(b @ y[:p]) - float(h @ y[p:])
        gap = abs(po - do) / (1.0 + abs(po) + abs(do))
        trace.append((k, po, do, gap, feas, pr, dr, tau, beta))
        comb[k] = max(pr, dr)
        # control (pdhg.jl:247-332)
        rank_update += 1
        if gap <= opt.tol_gap and feas <= opt.tol_feasibility:
            soc_ok = all(np.linalg.norm(x[o + 1:o + ln]) - x[o] < opt.tol_soc for (o, ln) in socs)
            rank_ok = all(s_.sq_side < opt.min_size_krylov_eigs or target_rank[i_] > opt.max_target_rank_krylov_eigs
                          or min_eig[i_] < opt.tol_psd for i_, s_ in enumerate(con.sdpcone))
            if rank_ok and soc_ok and k > opt.min_iter:
                status = 1
                break
            elif rank_update > window:
                for i_ in range(nsd):
                    rank_rule(i_)
                rank_update, update_cont = 0, 0
        elif k > window and comb[k - window] < comb[k] and rank_update > window:
            update_cont += 1
            if update_cont > opt.divergence_min_update:
                for i_ in range(nsd):
                    if target_rank[i_] < con.sdpcone[i_].sq_side:
                        rank_update, update_cont = 0, 0
                    rank_rule(i_)
        elif pr > opt.tol_primal and dr < opt.tol_dual and k > window:
            ada_count += 1
            if ada_count > opt.adapt_window:
                ada_count = 0
                if opt.line_search_flag:
                    beta *= (1.0 - adapt); tau /= np.sqrt(1.0 - adapt)
                else:
                    tau /= (1.0 - adapt); sigma *= (1.0 - adapt)
                adapt *= opt.adapt_decay
        elif pr < opt.tol_primal and dr > opt.tol_dual and k > window:
            ada_count += 1
            if ada_count > opt.adapt_window:
                ada_count = 0
                if opt.line_search_flag:
                    beta /= (1.0 - adapt); tau *= np.sqrt(1.0 - adapt)
                else:
                    tau *= (1.0 - adapt); sigma /= (1.0 - adapt)
                adapt *= opt.adapt_decay
    # undo scaling / permutation
    xs = x.copy()
    for (o, side, ii, jj) in blocks:
        seg = xs[o:o + len(ii)]
        xs[o:o + len(ii)] = np.where(ii != jj, seg / sqrt2, seg)
    if use_eq:                                   # pdhg.jl:751-755
        xs = D_eq * xs
        y = E_eq * y
    primal = np.zeros(n)
    primal[ord_] = xs
    return dict(status=status, iter=k, primal=primal, y=y, trace=np.array(trace))
