"""Wall time of the device full eigendecomposition (block-Jacobi, proxsdp_b200_eigh) next to LAPACK dsyevr on the host."""
import os, sys, time
import numpy as np
import scipy.linalg as sla
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from proxsdp_b200 import solver
for n in (int(a) for a in (sys.argv[1:] or ["500", "1000", "2000"])):
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, 12))
    S = rng.standard_normal((n, n))
    A = B @ B.T + 0.05 * (S + S.T)
    solver.eigh(A[:64, :64].copy())
    t0 = time.perf_counter(); w, Z = solver.eigh(A); t1 = time.perf_counter()
    t2 = time.perf_counter(); wl = sla.eigh(A, driver="evr", eigvals_only=False)[0]; t3 = time.perf_counter()
    print(f"n={n}: device block-Jacobi {1e3*(t1-t0):.1f} ms (incl. H2D/D2H of the matrices), host LAPACK dsyevr {1e3*(t3-t2):.1f} ms, "
          f"max |dw| {np.abs(w - wl).max():.2e}", flush=True)
