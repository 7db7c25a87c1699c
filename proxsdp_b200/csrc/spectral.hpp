// spectral.hpp — host side of the exact operator norm (approx_norm = false, reference src/pdhg.jl:107-118): the largest
// eigenvalue of the symmetric positive semidefinite operator v -> M (M' v) by a restarted Lanczos run with full
// re-orthogonalisation.  Host-only C++ (no CUDA): the operator comes in as a functor, so tests/test_spectral_host.py can
// compile this header alone with g++ and check it against LAPACK; solver.cu plugs in the two device sparse products.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace pb {

// cyclic Jacobi on a small dense symmetric matrix (host; k <= 48): A -> eigenvalues on the diagonal, Z = eigenvectors
inline void host_jacobi_eigh(int k, std::vector<double>& A, std::vector<double>& Z) {
    Z.assign((size_t)k * k, 0.0);
    for (int i = 0; i < k; ++i) Z[(size_t)i * k + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) { const double a = A[(size_t)i * k + j]; if (i == j) dia += a * a; else off += a * a; }
        if (off <= 1e-32 * (dia + off)) break;
        for (int pp = 0; pp < k - 1; ++pp)
            for (int q = pp + 1; q < k; ++q) {
                const double apq = A[(size_t)pp * k + q];
                if (apq == 0.0) continue;
                const double theta = (A[(size_t)q * k + q] - A[(size_t)pp * k + pp]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int r = 0; r < k; ++r) {       // columns pp, q
                    const double arp = A[(size_t)r * k + pp], arq = A[(size_t)r * k + q];
                    A[(size_t)r * k + pp] = c * arp - sn * arq; A[(size_t)r * k + q] = sn * arp + c * arq;
                }
                for (int r = 0; r < k; ++r) {       // rows pp, q
                    const double apr = A[(size_t)pp * k + r], aqr = A[(size_t)q * k + r];
                    A[(size_t)pp * k + r] = c * apr - sn * aqr; A[(size_t)q * k + r] = sn * apr + c * aqr;
                }
                for (int r = 0; r < k; ++r) {
                    const double zrp = Z[(size_t)r * k + pp], zrq = Z[(size_t)r * k + q];
                    Z[(size_t)r * k + pp] = c * zrp - sn * zrq; Z[(size_t)r * k + q] = sn * zrp + c * zrq;
                }
            }
    }
}

// sqrt(lambda_max) of the operator `matvec(v_in, w_out)` on vectors of length R, started from V0 (unit norm, length R):
// cycles of K = min(R, 40) Lanczos steps with two Gram-Schmidt passes against the whole basis, Rayleigh quotient by the
// Jacobi solver above, restart from the Ritz vector; converged when the residual of the leading pair is below
// 1e-11 lambda (or the Krylov space is invariant / complete).  Returns -1 when 60 cycles do not get there.
template <class MatVec>
double lanczos_sigma_max(long long R, const std::vector<double>& V0, MatVec&& matvec) {
    const int K = (int)std::min<long long>(R, 40);
    std::vector<double> V((size_t)R * (size_t)(K + 1)), w((size_t)R), alpha((size_t)K), beta((size_t)K + 1, 0.0);
    std::copy(V0.begin(), V0.begin() + R, V.begin());
    auto dot = [&](const double* a, const double* b) { double s0 = 0.0; for (long long i = 0; i < R; ++i) s0 += a[i] * b[i]; return s0; };
    for (int restart = 0; restart < 60; ++restart) {
        int k = 0;
        bool invariant = false;
        for (int j = 0; j < K; ++j) {
            double* vj = V.data() + (size_t)j * R;
            matvec(vj, w.data());
            alpha[(size_t)j] = dot(w.data(), vj);
            for (int pass = 0; pass < 2; ++pass)
                for (int q = 0; q <= j; ++q) {
                    const double* vq = V.data() + (size_t)q * R;
                    const double h = dot(vq, w.data());
                    for (long long i = 0; i < R; ++i) w[(size_t)i] -= h * vq[i];
                }
            const double bn = std::sqrt(dot(w.data(), w.data()));
            beta[(size_t)j + 1] = bn;
            k = j + 1;
            if (bn <= 1e-14 * std::fabs(alpha[0]) || bn == 0.0) { invariant = true; break; }
            double* vn = V.data() + (size_t)(j + 1) * R;
            for (long long i = 0; i < R; ++i) vn[i] = w[(size_t)i] / bn;
        }
        std::vector<double> T((size_t)k * k, 0.0), Z;
        for (int i = 0; i < k; ++i) {
            T[(size_t)i * k + i] = alpha[(size_t)i];
            if (i + 1 < k) { T[(size_t)i * k + i + 1] = beta[(size_t)i + 1]; T[(size_t)(i + 1) * k + i] = beta[(size_t)i + 1]; }
        }
        host_jacobi_eigh(k, T, Z);
        int best = 0;
        for (int i = 1; i < k; ++i) if (T[(size_t)i * k + i] > T[(size_t)best * k + best]) best = i;
        const double lam = T[(size_t)best * k + best];
        const double resid = std::fabs(beta[(size_t)k] * Z[(size_t)(k - 1) * k + best]);
        if (invariant || k == R || resid <= 1e-11 * std::fabs(lam)) return std::sqrt(std::max(lam, 0.0));
        // restart from the Ritz vector
        std::fill(w.begin(), w.end(), 0.0);
        for (int q = 0; q < k; ++q) {
            const double zq = Z[(size_t)q * k + best];
            const double* vq = V.data() + (size_t)q * R;
            for (long long i = 0; i < R; ++i) w[(size_t)i] += zq * vq[i];
        }
        const double nw = std::sqrt(dot(w.data(), w.data()));
        for (long long i = 0; i < R; ++i) V[(size_t)i] = w[(size_t)i] / nw;
    }
    return -1.0;
}

}  // namespace pb
