"""CPU restatement (numpy / Python floats) of two numerical pieces of the device Ritz analysis that deviate from the oracle's
literal restatement of KrylovKit, checked against LAPACK:

1. ritz_bi.cuh `ritz_value_warp`: an eigenvalue of a symmetric tridiagonal by rounds of 32 Sturm counts — 33-way multisection
   until the eigenvalue is isolated, then 32 points clustered at distances w 8^-i around the secant estimate of the
   characteristic polynomial's root.  Checked: the values agree with `numpy.linalg.eigvalsh` to eps ||T||, and the number of
   rounds stays well below the 11 of plain multisection.
2. lanczos_cl3.cuh `thick_restart_tridiag`: after a thick restart the Rayleigh quotient is diag(theta) bordered by the row f;
   a small Lanczos run on diag(theta) started from f (Gram-Schmidt repeated until a pass no longer shrinks the vector, a
   remainder below 1e-14 ||theta|| treated as a breakdown, continuation with the coordinate vector that sticks out most of the
   span so far) gives an orthogonal Q with Q' diag(theta) Q tridiagonal and Q' f = ||f|| e_0.  Checked on generic, clustered
   and degenerate inputs (repeated theta, zero and negligible components of f — the case where two fixed passes lost
   orthogonality completely): Q orthogonal to 1e-13, the transformed matrix tridiagonal to 1e-13 ||theta||, f mapped onto the
   first basis vector, and the spectrum of [T~, coupling] equal to that of the bordered matrix.

    python scripts/ritz_restart_check.py        # exits 0 when everything agrees
"""
import sys

import numpy as np


# ------------------------------------------------------------------------------------------------ 1. Ritz values
def sturm_count(d, e2, x):
    """# eigenvalues < x and the last minor (scaled matrix, entries <= 1 + |x|): p_j = (d_j - x) p_{j-1} - e2_{j-1} p_{j-2}."""
    p0, p1 = 1.0, d[0] - x
    s1 = p1 <= 0.0
    cnt = 1 if s1 else 0
    ex = 0
    for j in range(1, len(d)):
        p2 = (d[j] - x) * p1 - e2[j - 1] * p0
        s2 = (not s1) if p2 == 0.0 else (p2 < 0.0)
        if s2 != s1:
            cnt += 1
        s1 = s2
        p0, p1 = p1, p2
        if j % 8 == 0:
            a = abs(p1)
            if a > 2.0 ** 400:
                p0 *= 2.0 ** -400; p1 *= 2.0 ** -400; ex += 400
            elif 0.0 < a < 2.0 ** -400:
                p0 *= 2.0 ** 400; p1 *= 2.0 ** 400; ex -= 400
    return cnt, p1, ex


def ritz_value(d, e2, idx, lo, hi, tnorm):
    k = len(d)
    clo, chi, plo, phi, elo, ehi, vlo, vhi = 0, k, 0.0, 0.0, 0, 0, False, False
    cluster, xc, rounds = False, 0.0, 0
    for _ in range(48):
        w = hi - lo
        if cluster:
            xs = [min(max(xc - w * 8.0 ** -(l + 1), lo), hi) for l in range(16)] + [min(max(xc + w * 8.0 ** -(32 - l), lo), hi) for l in range(16, 32)]
        else:
            xs = [lo + w * ((l + 1) / 33.0) for l in range(32)]
        res = [sturm_count(d, e2, x) for x in xs]
        rounds += 1
        L = next((l for l in range(32) if res[l][0] > idx), 32)
        if L > 0:
            lo, (clo, plo, elo), vlo = xs[L - 1], res[L - 1], True
        if L < 32:
            hi, (chi, phi, ehi), vhi = xs[L], res[L], True
        wn = hi - lo
        if wn <= max(2.3e-16 * max(abs(lo), abs(hi)), 2.3e-16 * tnorm):
            break
        iso = (chi - clo == 1) and vlo and vhi and elo == ehi and ((plo < 0.0) != (phi < 0.0))
        fr = 0.5
        if iso:
            a, b = abs(plo), abs(phi)
            fr = a / (a + b) if a + b > 0.0 else 0.5
        cluster = iso and ((not cluster) or wn <= 0.125 * w)
        xc = lo + wn * fr
    return 0.5 * (lo + hi), rounds


def check_values(rng):
    worst, most_rounds, total_rounds, n = 0.0, 0, 0, 0
    for trial in range(60):
        k = int(rng.integers(9, 52))
        kind = trial % 4
        d = rng.standard_normal(k) * (10.0 if kind != 2 else 1e-3) + (50.0 if kind == 1 else 0.0)
        e = np.abs(rng.standard_normal(k - 1)) * (1.0 if kind != 3 else 1e-6) + 1e-12
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        ref = np.linalg.eigvalsh(T)[::-1]
        gl = min(d[j] - (e[j - 1] if j else 0.0) - (e[j] if j < k - 1 else 0.0) for j in range(k))
        gu = max(d[j] + (e[j - 1] if j else 0.0) + (e[j] if j < k - 1 else 0.0) for j in range(k))
        tnorm = max(abs(gl), abs(gu))
        s = 2.0 ** -(int(np.floor(np.log2(tnorm))) + 1)
        ds, e2 = d * s, (e * s) ** 2
        wdt = (gu - gl) * s
        for t in range(min(6, k - 1)):
            lam, rounds = ritz_value(ds, e2, k - 1 - t, gl * s - 1e-3 * wdt, gu * s + 1e-3 * wdt, tnorm * s)
            worst = max(worst, abs(lam / s - ref[t]) / tnorm)
            most_rounds = max(most_rounds, rounds); total_rounds += rounds; n += 1
    print(f"ritz values: max |lam - LAPACK| / ||T|| = {worst:.2e}, rounds mean {total_rounds / n:.1f} max {most_rounds}")
    return worst <= 1e-14 and total_rounds / n <= 8.0      # (LAPACK itself is good to a few eps ||T||)


# ------------------------------------------------------------------------------------------------ 2. restart
def retridiagonalise(theta, f):
    m2 = len(theta)
    thmax = max(np.abs(theta).max(), 1e-300)
    nf = float(np.sqrt(np.sum(f * f)))
    Q = np.zeros((m2, m2))           # row j = Lanczos vector j
    ta, tb = np.zeros(m2), np.zeros(m2)
    q = f / nf if nf > 0.0 else np.eye(m2)[0]
    qprev, bprev = np.zeros(m2), 0.0

    def reorth(w, j):
        nprev = float(np.sqrt(np.sum(w * w)))
        ncur = nprev
        for _ in range(6):
            h = Q[:j + 1] @ w
            w = w - Q[:j + 1].T @ h
            ncur = float(np.sqrt(np.sum(w * w)))
            if not (ncur < 0.7 * nprev):
                break
            nprev = ncur
        return w, ncur

    for j in range(m2):
        Q[j] = q
        w = theta * q
        aj = float(q @ w)
        w = w - aj * q - bprev * qprev
        ta[j] = aj
        if j == m2 - 1:
            break
        w, nb = reorth(w, j)
        bj = nb
        if not (nb > 1e-14 * thmax):
            g = 1.0 - np.sum(Q[:j + 1] ** 2, axis=0)
            w = np.eye(m2)[int(np.argmax(g))]
            w, nb = reorth(w, j)
            bj = 0.0
        tb[j] = bj
        qprev, q, bprev = q, w / nb, bj
    return Q, ta, tb, nf


def check_restart(rng):
    ok = True
    worst = [0.0, 0.0, 0.0, 0.0]
    for trial in range(80):
        m2 = int(rng.integers(2, 32))
        kind = trial % 5
        theta = np.sort(rng.standard_normal(m2) * 5.0 + 30.0)[::-1]
        f = rng.standard_normal(m2)
        if kind == 1:                                    # clustered values
            theta = 37.0 + 1e-9 * rng.standard_normal(m2)
        if kind == 2:                                    # repeated values (multiple eigenvalues)
            theta = np.repeat(theta[: (m2 + 2) // 3], 3)[:m2]
        if kind == 3:                                    # zero and negligible couplings (pairs that all but converged)
            f[rng.random(m2) < 0.4] = 0.0
            f[rng.random(m2) < 0.3] *= 1e-40
        if kind == 4:                                    # graded couplings over 30 orders of magnitude
            f = f * 10.0 ** (-30.0 * rng.random(m2))
        if not np.any(f):
            f[0] = 1.0
        Q, ta, tb, nf = retridiagonalise(theta, f)
        scale = np.abs(theta).max()
        orth = np.abs(Q @ Q.T - np.eye(m2)).max()
        Tt = Q @ np.diag(theta) @ Q.T
        tri = np.diag(ta) + np.diag(tb[:m2 - 1], 1) + np.diag(tb[:m2 - 1], -1)
        tri_err = np.abs(Tt - tri).max() / scale
        qf = Q @ f
        f_err = max(abs(qf[0] - nf), np.abs(qf[1:]).max(initial=0.0)) / max(nf, 1e-300)
        # spectrum of the bordered matrix [diag(theta) f; f' 0] == spectrum of [T~ (nf e_0); (nf e_0)' 0]
        B1 = np.zeros((m2 + 1, m2 + 1)); B1[:m2, :m2] = np.diag(theta); B1[:m2, m2] = f; B1[m2, :m2] = f
        B2 = np.zeros((m2 + 1, m2 + 1)); B2[:m2, :m2] = tri; B2[0, m2] = nf; B2[m2, 0] = nf
        spec_err = np.abs(np.linalg.eigvalsh(B1) - np.linalg.eigvalsh(B2)).max() / scale
        for i, v in enumerate((orth, tri_err, f_err, spec_err)):
            worst[i] = max(worst[i], v)
        ok = ok and orth <= 1e-13 and tri_err <= 1e-13 and f_err <= 1e-12 and spec_err <= 1e-13
    print("restart: max ||QQ' - I|| = %.2e, off-tridiagonal / ||theta|| = %.2e, |Qf - ||f|| e_0| / ||f|| = %.2e, spectrum diff = %.2e" % tuple(worst))
    return ok


def main():
    rng = np.random.default_rng(2024)
    ok1 = check_values(rng)
    ok2 = check_restart(rng)
    print("OK" if ok1 and ok2 else "MISMATCH")
    return 0 if ok1 and ok2 else 1


if __name__ == "__main__":
    sys.exit(main())
