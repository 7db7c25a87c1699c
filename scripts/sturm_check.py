"""CPU check of the division-free, pre-scaled Sturm count used by ritz_bi.cuh (same arithmetic in Python floats):
compares count(x) with the number of eigenvalues below x from LAPACK on random / graded / clustered tridiagonals,
including the rescue of exact zeros and the every-8-steps rescaling."""
import numpy as np


def sturm_count(d, e2s, x, inv_t):
    k = len(d)
    p0, p1 = 1.0, (d[0] - x) * inv_t
    cnt = 1 if p1 <= 0.0 else 0
    t0 = e2s[0] * p0 if k > 1 else 0.0
    for jb in range(1, k, 8):
        for j in range(jb, min(jb + 8, k)):
            dx = (d[j] - x) * inv_t
            p2 = dx * p1 - t0            # fma on the device; rounding differences do not matter for the count
            neg = (p2 == 0.0) or ((p2 < 0.0) != (p1 < 0.0))
            cnt += 1 if neg else 0
            if p2 == 0.0:
                p2 = -p1 * 1e-300 if p1 != 0.0 else -1e-300
            t0 = e2s[j] * p1 if j < k - 1 else 0.0
            p0, p1 = p1, p2
        a = abs(p1)
        if a > 1e100:
            p0 *= 1e-100; p1 *= 1e-100; t0 *= 1e-100
        elif a < 1e-100:
            p0 *= 1e100; p1 *= 1e100; t0 *= 1e100
    return cnt


def sturm_count_v1(d, e, x):
    """the first-generation device recurrence (unscaled, rescaling every step): reference for the convention at
    exact zeros of the minors, where both disagree with LAPACK by one in the same way"""
    k = len(d)
    p0, p1 = 1.0, d[0] - x
    cnt = 1 if p1 <= 0.0 else 0
    for j in range(1, k):
        ej = e[j - 1]
        p2 = (d[j] - x) * p1 - (ej * ej) * p0
        neg = (p2 == 0.0) or ((p2 < 0.0) != (p1 < 0.0))
        cnt += 1 if neg else 0
        if p2 == 0.0:
            p2 = -p1 * 1e-300 if p1 != 0.0 else -1e-300
        a = abs(p2)
        if a > 1e150:
            p1 *= 1e-150; p2 *= 1e-150
        elif a < 1e-150:
            p1 *= 1e150; p2 *= 1e150
        p0, p1 = p1, p2
    return cnt


def main():
    rng = np.random.default_rng(1)
    bad = 0
    total = 0
    for trial in range(400):
        k = int(rng.integers(8, 104))
        kind = trial % 4
        if kind == 0:
            d = rng.standard_normal(k); e = rng.standard_normal(k - 1)
        elif kind == 1:
            d = 10.0 ** rng.uniform(-6, 6, k) * rng.choice([-1, 1], k); e = 10.0 ** rng.uniform(-3, 3, k - 1)
        elif kind == 2:
            d = np.full(k, 3.0) + 1e-9 * rng.standard_normal(k); e = np.full(k - 1, 1.0)
        else:
            d = 1e4 + rng.standard_normal(k); e = 1e-4 * (1.0 + rng.random(k - 1))
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        w = np.linalg.eigvalsh(T)
        tn = max(np.abs(d).max() + 2 * np.abs(e).max(), 1e-300)
        inv_t = 1.0 / tn
        e2s = (e * inv_t) ** 2
        xs = np.concatenate([rng.uniform(w[0] - 0.1 * tn, w[-1] + 0.1 * tn, 40), 0.5 * (w[1:] + w[:-1])])
        for x in d:                               # exact zeros of a minor: same convention as the first generation
            total += 1
            if sturm_count(d, e2s, x, inv_t) != sturm_count_v1(d, e, x):
                bad += 1
        for x in xs:
            gap = np.abs(w - x).min()
            if gap < 1e-10 * tn:
                continue                      # too close to an eigenvalue for the count to be defined in floating point
            total += 1
            c = sturm_count(d, e2s, x, inv_t)
            if c != int((w < x).sum()):
                bad += 1
    print(f"{total} counts checked, {bad} mismatches")
    return bad


if __name__ == "__main__":
    raise SystemExit(1 if main() else 0)
