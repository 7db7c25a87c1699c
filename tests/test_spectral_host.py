"""CPU test of the host side of `approx_norm = false` (reference src/pdhg.jl:107-118): proxsdp_b200/csrc/spectral.hpp (the
restarted Lanczos run behind `Solver::spectral_norm_device` and its small Jacobi eigensolver) is compiled alone with g++,
the operator v -> M (M' v) supplied by a dense matrix, and compared with LAPACK's largest singular value."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r'''
#include <cstdio>
#include <cstdlib>
#include "spectral.hpp"
int main(int argc, char** argv) {
    long long R = atoll(argv[1]), n = atoll(argv[2]);
    std::vector<double> M((size_t)(R * n)), V0((size_t)R);
    FILE* f = fopen(argv[3], "rb");
    if (fread(M.data(), 8, (size_t)(R * n), f) != (size_t)(R * n)) return 2;
    if (fread(V0.data(), 8, (size_t)R, f) != (size_t)R) return 2;
    fclose(f);
    long long calls = 0;
    auto matvec = [&](const double* vin, double* wout) {
        std::vector<double> t((size_t)n, 0.0);
        for (long long i = 0; i < R; ++i) for (long long j = 0; j < n; ++j) t[j] += M[i * n + j] * vin[i];
        for (long long i = 0; i < R; ++i) { double s = 0; for (long long j = 0; j < n; ++j) s += M[i * n + j] * t[j]; wout[i] = s; }
        ++calls;
    };
    printf("%.17g %lld\n", pb::lanczos_sigma_max(R, V0, matvec), calls);
    // the Jacobi solver alone
    int k = 17;
    std::vector<double> A((size_t)k * k), A0, Z;
    srand(1);
    for (int i = 0; i < k; ++i) for (int j = 0; j <= i; ++j) { double v = rand() / (double)RAND_MAX - 0.5; A[i * k + j] = v; A[j * k + i] = v; }
    A0 = A;
    pb::host_jacobi_eigh(k, A, Z);
    double err = 0, orth = 0;
    for (int c = 0; c < k; ++c) for (int i = 0; i < k; ++i) { double s = 0; for (int j = 0; j < k; ++j) s += A0[i * k + j] * Z[j * k + c]; err = fmax(err, fabs(s - A[c * k + c] * Z[i * k + c])); }
    for (int a = 0; a < k; ++a) for (int b = 0; b < k; ++b) { double s = 0; for (int i = 0; i < k; ++i) s += Z[i * k + a] * Z[i * k + b]; orth = fmax(orth, fabs(s - (a == b))); }
    printf("%.3e %.3e\n", err, orth);
    return 0;
}
'''


def test_restarted_lanczos_and_jacobi_against_lapack(tmp_path):
    src, exe = tmp_path / "driver.cpp", tmp_path / "driver"
    src.write_text(DRIVER)
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "proxsdp_b200", "csrc"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    rng = np.random.default_rng(0)
    cases = [(1, 5, "random"), (2, 3, "random"), (30, 100, "random"), (200, 50, "random"), (300, 400, "clustered"),
             (500, 500, "identity"), (120, 300, "lowrank")]
    for R, n, kind in cases:
        if kind == "random":
            M = rng.standard_normal((R, n))
        elif kind == "identity":
            M = np.eye(R, n)                                   # Max-Cut's constraint matrix: every singular value is 1
        elif kind == "clustered":
            U, _ = np.linalg.qr(rng.standard_normal((R, R)))
            V, _ = np.linalg.qr(rng.standard_normal((n, n)))
            sv = np.concatenate([[1.0, 0.9999, 0.9998, 0.9997], np.linspace(0.99, 0.1, R - 4)])
            M = (U * sv) @ V[:R]
        else:
            M = rng.standard_normal((R, 3)) @ rng.standard_normal((3, n))
        v0 = rng.standard_normal(R)
        v0 /= np.linalg.norm(v0)
        path = tmp_path / "m.bin"
        with open(path, "wb") as f:
            f.write(np.ascontiguousarray(M, dtype=np.float64).tobytes())
            f.write(np.ascontiguousarray(v0, dtype=np.float64).tobytes())
        out = subprocess.run([str(exe), str(R), str(n), str(path)], check=True, capture_output=True, text=True).stdout.split()
        sigma, calls, err, orth = float(out[0]), int(out[1]), float(out[2]), float(out[3])
        ref = np.linalg.svd(M, compute_uv=False).max()
        assert abs(sigma / ref - 1.0) <= 1e-10, (R, n, kind, sigma, ref)
        assert calls <= 40 * 60
        assert err <= 1e-13 and orth <= 1e-13
