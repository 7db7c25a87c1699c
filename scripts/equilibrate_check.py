"""CPU restatement of `k_equilibrate` (proxsdp_b200/csrc/runtime.cu) against the literal restatement of the reference's
equilibrate! (oracle/oracle_np.py, reference src/equilibration.jl:1-71).

The device collapses the iteration: the reference re-sets the column scaling to its mean in every step, so D = exp(v) I
with one number v, and M only enters through the row sums r_i = sum_j M_ij^2:
    row_norms_i = (exp(u_i) exp(v))^2 r_i ,   sum_j col_norms_j = sum_i row_norms_i .
This script walks that recurrence in Python floats and compares E and D with the literal form."""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_np                      # noqa: E402
from proxsdp_b200.options import Options          # noqa: E402


def collapsed(M, opt):
    R, n = M.shape
    r = np.asarray(M.multiply(M).sum(axis=1)).ravel()
    alpha2, beta2, gamma = np.sqrt(n / R), np.sqrt(R / n), 0.1
    u, ubar, v, vbar = np.zeros(R), np.zeros(R), 0.0, 0.0
    for it in range(1, int(opt.equilibration_iters) + 1):
        step = 2.0 / (gamma * (it + 1.0))
        e = np.exp(u) * np.exp(v)
        rn = e * e * r
        S = rn.sum()
        u = np.clip(u - step * (rn - alpha2 + gamma * u), opt.equilibration_lb, opt.equilibration_ub)
        ubar = 2.0 / (it + 2.0) * u + it / (it + 2.0) * ubar
        v = min(opt.equilibration_ub, max(v - step * (S / n - beta2 + gamma * v), 0.0))
        vbar = 2.0 / (it + 2.0) * v + it / (it + 2.0) * vbar
    return np.exp(ubar), np.exp(vbar)


def main():
    bad = 0
    rng = np.random.default_rng(0)
    for (R, n, dens, iters) in [(7, 40, 0.3, 50), (60, 30, 0.1, 200), (5, 5, 1.0, 1000), (300, 2000, 0.01, 100)]:
        M = sp.random(R, n, dens, random_state=int(rng.integers(1 << 30)), format="csr")
        M = sp.csr_matrix(sp.diags(np.logspace(-2, 2, R)) @ M)
        opt = Options(equilibration_iters=iters)
        E0, D0 = oracle_np.equilibrate(M, opt)
        E1, d1 = collapsed(M, opt)
        err = max(np.abs(E0 / E1 - 1.0).max(), np.abs(D0 / d1 - 1.0).max())
        spread = np.abs(D0 / D0[0] - 1.0).max()
        ok = err <= 1e-8 and spread <= 1e-14      # the first steps (step size 10, 6.7, 5 ...) amplify rounding
        print(f"R={R} n={n} iters={iters}: max rel diff {err:.2e}, spread of D {spread:.1e} {'ok' if ok else 'FAIL'}")
        bad += not ok
    return bad


if __name__ == "__main__":
    sys.exit(main())
