"""CPU check of the re-orthogonalisation variant used by the v3 device Lanczos kernel.

KrylovKit (and oracle/oracle_eig.c) : local three-term recurrence, then two modified Gram-Schmidt passes.
device kernel v2 (lanczos_cl.cuh)   : two classical Gram-Schmidt passes over the whole basis (CGS2).
device kernel v3 (lanczos_cl3.cuh)  : local three-term recurrence (alpha from a fused dot), then ONE classical
                                      Gram-Schmidt pass over the whole basis (second pass only when the first
                                      removed a visible part of w).
This script runs a numpy restatement of the thick-restart loop with each scheme on a family of test matrices
and on PDHG-like iterates, and compares mat-vec counts, converged counts, Ritz values and basis orthogonality
with the C oracle.  Run: python scripts/lz_variant_check.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import oracle  # noqa: E402


def lanczos(A, x0, howmany, K, scheme, maxiter=100, tol=1e-12):
    n = A.shape[0]
    V = np.zeros((n, K + 1))
    V[:, 0] = x0 / np.linalg.norm(x0)
    Hd = np.zeros(K); He = np.zeros(K); Harr = np.zeros(K)
    arrow_at, arrow_len = -1, 0
    k, numops, numiter, converged = 1, 0, 1, 0
    worst_orth = 0.0
    while True:
        j = k - 1
        w = A @ V[:, j]
        numops += 1
        Vj = V[:, : j + 1]
        if scheme == "cgs2":
            h = Vj.T @ w; w = w - Vj @ h; alpha = h[j]
            h2 = Vj.T @ w; wn2 = w @ w; w = w - Vj @ h2; alpha += h2[j]
            beta2 = wn2 - h2 @ h2
            if not (h2 @ h2 <= 1e-4 * wn2):
                beta2 = w @ w
        elif scheme == "local1":
            alpha = V[:, j] @ w
            w = w - alpha * V[:, j]
            if arrow_at < 0 or j > arrow_at:
                if j > 0:
                    w = w - He[j - 1] * V[:, j - 1]
            elif j == arrow_at:
                # first step after a thick restart: couples to all kept Ritz vectors
                w = w - V[:, :arrow_len] @ Harr[:arrow_len]
            h = Vj.T @ w; wn2 = w @ w; w = w - Vj @ h; alpha += h[j]
            hn2 = h @ h
            beta2 = wn2 - hn2
            if not (hn2 <= 1e-4 * wn2):
                h2 = Vj.T @ w; w = w - Vj @ h2; alpha += h2[j]
                beta2 = w @ w
        else:
            raise ValueError(scheme)
        beta = np.sqrt(max(beta2, 0.0))
        Hd[j] = alpha; He[j] = beta
        V[:, k] = w / beta if beta > 0 else 0.0
        G = V[:, : k + 1].T @ V[:, : k + 1]
        if beta > tol:
            worst_orth = max(worst_orth, np.abs(G - np.eye(k + 1)).max())
        if beta <= tol and k < howmany:
            howmany = k
        finished = False
        if k == K or beta <= tol:
            T = np.diag(Hd[:k])
            for i in range(k - 1):
                lo, hi = i, i + 1
                if not (lo < arrow_len and hi <= arrow_at):
                    T[lo, hi] = T[hi, lo] = He[lo]
            if arrow_at >= 0:
                for lo in range(arrow_len):
                    T[lo, arrow_at] = T[arrow_at, lo] = Harr[lo]
            D, U = np.linalg.eigh(T)
            D = D[::-1]; U = U[:, ::-1]
            f = beta * U[k - 1, :]
            converged = 0
            while converged < k and abs(f[converged]) <= tol:
                converged += 1
            if converged >= howmany:
                finished = True
            elif k == K:
                if numiter == maxiter:
                    finished = True
                else:
                    keep = (3 * K + 2 * converged) // 5
                    Vn = V[:, :K] @ U[:, :keep]
                    V[:, :keep] = Vn
                    V[:, keep] = V[:, K]
                    Hd[:] = 0; He[:] = 0; Harr[:] = 0
                    Hd[:keep] = D[:keep]; Harr[:keep] = f[:keep]
                    arrow_at, arrow_len = keep, keep
                    k = keep + 1
                    numiter += 1
                    continue
        if finished:
            nv = min(max(howmany, converged), k)
            return D[:nv], V[:, :k] @ U[:, :nv], dict(converged=converged, numops=numops, numiter=numiter, orth=worst_orth)
        k += 1


def cases():
    rng = np.random.default_rng(0)
    for n, r, nev in ((150, 4, 2), (300, 6, 4), (600, 3, 3), (1000, 8, 6), (500, 12, 12), (400, 2, 16)):
        B = rng.standard_normal((n, r))
        S = rng.standard_normal((n, n))
        yield f"lowrank+noise n={n} r={r} nev={nev}", B @ B.T - 0.1 * np.eye(n) + 0.01 * (S + S.T), nev
    for n, nev in ((200, 3), (500, 5)):
        S = rng.standard_normal((n, n))
        yield f"wigner n={n} nev={nev}", (S + S.T) / np.sqrt(n), nev          # hard: needs restarts
    for n, nev in ((300, 2), (300, 8)):
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        lam = np.concatenate([[10.0, 10.0 - 1e-6, 9.0, 5.0], -rng.random(n - 4)])      # near-degenerate top pair
        yield f"clustered n={n} nev={nev}", (Q * lam) @ Q.T, nev
    n = 400
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.concatenate([1e6 + np.arange(3.0), 1e-3 * rng.standard_normal(n - 3)])   # huge shift: alpha >> beta
    yield "shifted n=400 nev=3", (Q * lam) @ Q.T, 3
    yield "zero n=200 nev=2", np.zeros((200, 200)), 2
    B = rng.standard_normal((250, 2))
    yield "exact rank-2 n=250 nev=4", B @ B.T, 4


def main():
    bad = 0
    for name, A, nev in cases():
        n = A.shape[0]
        K = max(2 * nev + 1, 25)
        x0 = oracle.eig_resid(n)
        vo, Vo, io = oracle.lanczos(np.triu(A), x0, nev, K)
        line = f"{name:34s} oracle ops={io['numops']:4d} conv={io['converged']:2d}"
        for scheme in ("cgs2", "local1"):
            v, Vv, info = lanczos(A, x0, nev, K, scheme)
            m = min(len(v), len(vo), nev)
            dv = np.abs(v[:m] - vo[:m]).max() / max(1.0, np.abs(vo[:m]).max()) if m else 0.0
            same = info["numops"] == io["numops"] and info["converged"] == io["converged"]
            if scheme == "local1" and (not same or dv > 1e-9 or info["orth"] > 1e-10):
                bad += 1
            line += f" | {scheme}: ops={info['numops']:4d} conv={info['converged']:2d} dval={dv:.1e} orth={info['orth']:.1e}{'' if same else ' DIFF'}"
        print(line, flush=True)
    print("mismatches for local1:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
