"""Lanczos kernel on cones whose matrix does NOT fit L2 (B200: 126 MB): n = 4000 (128 MB), 5000 (200 MB, the size of
SDPLIB maxG55), 6000 (288 MB).  Checks the eigsolve against the CPU oracle and reports the algorithmic bandwidth
mat-vecs x (8 n^2 + 16 n) / time — here the roofline denominator (HBM copy bandwidth) is the real bound."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from proxsdp_b200 import solver  # noqa: E402

peak = 6535.7
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
rng = np.random.default_rng(5)
SIZES = [int(a) for a in sys.argv[1:]] or [2000, 4000, 5000, 6000]
for n, nev in [(n, 4) for n in SIZES]:
    B = rng.standard_normal((n, 6))
    S = rng.standard_normal((n, n))
    A = B @ B.T - 0.1 * np.eye(n) + 0.01 * (S + S.T)
    del S
    x0 = oracle.eig_resid(n)
    K = max(2 * nev + 1, 25)
    t0 = time.perf_counter()
    vo, Vo, io = oracle.lanczos(np.triu(A), x0, nev, K)
    t_cpu = time.perf_counter() - t0
    vg, Vg, ig = solver.lanczos(A, x0, nev, K, repeat=4)
    same = (ig["converged"], ig["numops"], ig["numiter"]) == (io["converged"], io["numops"], io["numiter"])
    dv = float(np.abs(vg[:nev] - vo[:nev]).max() / np.abs(vo[:nev]).max())
    res = float(np.abs(A @ Vg - Vg * vg).max() / np.abs(vo).max())
    gbs = ig["numops"] * (8.0 * n * n + 16.0 * n) / (ig["ms"] * 1e-3) / 1e9
    print(json.dumps({"kernel_env": os.environ.get("PROXSDP_B200_LZ_KERNEL", "3") + "/" + os.environ.get("PROXSDP_B200_LANCZOS", "cluster"), "n": n, "matrix_MB": round(8e-6 * n * n, 1), "nev": nev, "K": K, "matvecs": ig["numops"], "counts_equal_oracle": same,
                      "ritz_value_rel_diff": dv, "residual_rel": res, "gpu_ms": ig["ms"], "us_per_matvec": 1e3 * ig["ms"] / ig["numops"],
                      "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak, "cpu_oracle_ms": 1e3 * t_cpu}), flush=True)
