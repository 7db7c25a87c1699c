"""Weak-duality certificates recomputed in numpy / LAPACK from a returned primal-dual pair (test infrastructure shared by
the CPU pins of the oracle and the GPU tests of the product: neither the oracle nor the device code is called here)."""
import numpy as np

from proxsdp_b200.structs import ivec


def certificate(aff, con, r):
    """Recompute, from the returned vectors only, everything that makes (x, y) a certified primal-dual pair of
    min c'x, Ax = b, Gx <= h, X PSD: residuals of the constraints, eigenvalues of X and of the dual slack matrix
    (LAPACK), complementarity-free weak-duality bracket [dual objective, primal objective]."""
    x = r.primal
    y_eq, y_in = r.dual_eq, r.dual_in
    eq = np.abs(aff.A @ x - aff.b).max() if aff.p else 0.0
    ineq = max(0.0, float((aff.G @ x - aff.h).max())) if aff.m else 0.0
    s = aff.c + (aff.A.T @ y_eq if aff.p else 0.0) + (aff.G.T @ y_in if aff.m else 0.0)
    lam_x, lam_s, trace = 0.0, 0.0, 0.0
    covered = np.zeros(aff.n, dtype=bool)
    for blk in con.sdpcone:
        idx = np.asarray(blk.vec_i)
        covered[idx] = True
        X = ivec(x[idx])
        # the dual slack of <C, X> = sum_ij C_ij X_ij with the objective stored on the upper triangle: off-diagonal
        # coefficients carry both (i, j) and (j, i), so the symmetric dual matrix has half of them on either side
        Sm = ivec(s[idx])
        d = np.diag(Sm).copy()
        Sm = Sm / 2.0
        np.fill_diagonal(Sm, d)
        lam_x = min(lam_x, float(np.linalg.eigvalsh(X).min()))
        lam_s = min(lam_s, float(np.linalg.eigvalsh(Sm).min()))
        trace += float(np.trace(X))
    free_dual = np.abs(s[~covered]).max(initial=0.0)          # free variables: the dual slack must vanish
    primal = float(aff.c @ x)
    dual = -float(aff.b @ y_eq) - float(aff.h @ y_in)
    return dict(eq=eq, ineq=ineq, lam_x=lam_x, lam_s=lam_s, trace=trace, free_dual=free_dual, primal=primal, dual=dual,
                y_in_min=float(y_in.min(initial=0.0)))


def maxcut_bracket(aff, con, r, n):
    """Max-Cut relaxations (diag(X) = 1, so trace(X) = n is known): with lam = lambda_min of the dual slack matrix every
    feasible X has <C, X> >= dual objective + min(lam, 0) n.  Returns (certified lower bound, primal objective, dict)."""
    k = certificate(aff, con, r)
    return k["dual"] + min(k["lam_s"], 0.0) * n, k["primal"], k
