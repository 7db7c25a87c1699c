import os, sys
import numpy as np
ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from proxsdp_b200 import solver
n, nev = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(n)
Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
lam = np.concatenate([[10.0] * 3, [8.0] * 4, [7.5, 7.5], [6.0] * 3, rng.uniform(-3.0, 3.0, n - 12)])
A = (Q * lam) @ Q.T
A = 0.5 * (A + A.T)
sys.path.insert(0, os.path.join(ROOT))
from oracle import oracle
x0 = oracle.eig_resid(n)
vo, Vo, io = oracle.lanczos(np.triu(A), x0, nev, 25)
print("oracle:", io, np.round(vo[:nev + 2], 9).tolist(), "max residual %.2e" % np.abs(A @ Vo - Vo * vo).max(), flush=True)
os.environ["PROXSDP_B200_LZ_DEBUG"] = "1"
vg, Vg, ig = solver.lanczos(A, x0, nev, 25)
c = min(ig["converged"], len(vg))
print("gpu:", ig, np.round(vg[:nev + 2], 9).tolist(), "max residual of the converged pairs %.2e" % np.abs(A @ Vg[:, :c] - Vg[:, :c] * vg[:c]).max(),
      "orth %.2e" % np.abs(Vg.T @ Vg - np.eye(len(vg))).max(), flush=True)
