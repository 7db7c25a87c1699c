"""CPU check of the division-free, pre-scaled Sturm count used by ritz_bi.cuh (same arithmetic in Python floats):
compares count(x) with the number of eigenvalues below x from LAPACK on random / graded / clustered tridiagonals,
including the rescue of exact zeros and the every-8-steps rescaling."""
import numpy as np


def sturm_count(d, e2s, x, inv_t):
    k = len(d)
    p0, p1 = 1.0, (d[0] - x) * inv_t
    s1 = p1 <= 0.0
    cnt = 1 if s1 else 0
    t0 = e2s[0] * p0 if k > 1 else 0.0
    for jb in range(1, k, 8):
        nb = min(8, k - jb)
        for u in range(8):
            j = jb + u
            dx = (d[j] - x) * inv_t if j < k else 1.0
            ee = e2s[j] if j < k - 1 else 0.0
            p2 = dx * p1 - t0            # fma on the device; rounding differences do not matter for the count
            s2 = (not s1) if p2 == 0.0 else (p2 < 0.0)
            if u < nb:
                cnt += 1 if s2 != s1 else 0
                s1 = s2
            t0 = ee * p1
            p0, p1 = p1, p2
        a = abs(p1)
        if a > 1e100:
            p0 *= 1e-100; p1 *= 1e-100; t0 *= 1e-100
        elif 0.0 < a < 1e-100:
            p0 *= 1e100; p1 *= 1e100; t0 *= 1e100
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
        # random points, mid-gaps, and the diagonal entries themselves (x = d[0] makes the first minor exactly zero)
        xs = np.concatenate([rng.uniform(w[0] - 0.1 * tn, w[-1] + 0.1 * tn, 40), 0.5 * (w[1:] + w[:-1]), d])
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
