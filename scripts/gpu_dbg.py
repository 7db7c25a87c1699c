import sys, time, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
from proxsdp_b200 import solver
rng = np.random.default_rng(0)
for n in (257, 300, 320, 321, 384, 400, 448, 512, 600):
    A = rng.standard_normal((n, n)); A = A + A.T
    w, Z = solver.eigh(A)
    w0 = np.linalg.eigvalsh(A)
    print("eigh", n, "val err", np.abs(w - w0).max(), "recon", np.abs(Z @ np.diag(w) @ Z.T - A).max(), "orth", np.abs(Z.T @ Z - np.eye(n)).max(), flush=True)
