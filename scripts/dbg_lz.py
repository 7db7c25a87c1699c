import os, sys
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from oracle import oracle
from proxsdp_b200 import solver
rng = np.random.default_rng(5)
n, K = 300, 25
Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
lam = np.concatenate([[10, 9.99, 9.985, 9.98, 9.9], np.linspace(9.8, -8, n - 5)])
A = (Q * lam) @ Q.T; A = (A + A.T) / 2
x0 = oracle.eig_resid(n)
for nev in (2, 3, 4):
    vo, Vo, io = oracle.lanczos(np.triu(A), x0, nev, K)
    vg, Vg, ig = solver.lanczos(A, x0, nev, K)
    print("nev", nev, "oracle", io, vo[:5], "| gpu", ig, vg[:5], flush=True)
