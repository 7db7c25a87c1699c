import sys, time, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
from oracle import oracle
from proxsdp_b200 import solver
rng = np.random.default_rng(1)
n = 124
Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
x0 = oracle.eig_resid(n)
for case in range(6):
    for nev in (4, 5, 6):
        if case == 0: tail = np.linspace(1e-3, -1e-3, 20)
        elif case == 1: tail = np.linspace(1e-2, -1e-2, 10)
        elif case == 2: tail = -np.abs(rng.standard_normal(20)) * 1e-4
        elif case == 3: tail = np.zeros(20)
        elif case == 4: tail = np.linspace(0.5, -0.5, 20)
        else: tail = np.linspace(1e-6, -1e-6, 30)
        lam = np.concatenate([[100.0, 80.0, 60.0, 50.0], tail, np.linspace(-1.0, -60.0, n - 4 - len(tail))])
        A = (Q * lam) @ Q.T; A = 0.5 * (A + A.T)
        v1, V1, i1 = oracle.lanczos(np.triu(A), x0, nev, 25)
        v2, V2, i2 = solver.lanczos(A, x0, nev, 25)
        print("case", case, "nev", nev, "oracle", i1, "gpu", {k: i2[k] for k in ("converged", "numops", "numiter")},
              "valdiff", np.abs(v1[:nev] - v2[:nev]).max() if len(v1) >= nev and len(v2) >= nev else (len(v1), len(v2)), flush=True)
