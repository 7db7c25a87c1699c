// jacobi.cuh — block-cooperative parallel cyclic Jacobi eigensolver in shared memory.
//
// Replaces the LAPACK dsyevr call sites of the reference for matrices that fit one
// CTA's shared memory: small PSD cones (reference src/prox_operators.jl:111-126,
// `full_eig!`, taken whenever n <= min_size_krylov_eigs = 100) and the K x K Rayleigh
// quotient of the Lanczos process (KrylovKit `tridiageigh!`, reached from
// reference src/eigsolver.jl:802).
//
// Two-sided Jacobi with the round-robin ("chess tournament") parallel ordering:
// m/2 disjoint rotations per round, m-1 rounds per sweep, quadratic convergence.
#pragma once
#include "common.cuh"

namespace pb {

struct JacobiScratch {
    double* c;    // m/2 cosines
    double* s;    // m/2 sines
    double* dii;  // m/2 new a_ii
    double* dkk;  // m/2 new a_kk
    int* pi;      // m/2 first index of the pair
    int* pk;      // m/2 second index of the pair
    double* red;  // >= 34 doubles reduction scratch
};

__host__ __device__ inline size_t jacobi_scratch_bytes(int n) {
    int m2 = (n + 2) / 2;
    return sizeof(double) * (4 * (size_t)m2 + 40) + sizeof(int) * 2 * (size_t)m2 + 16;
}

__device__ inline JacobiScratch jacobi_carve(void* base, int n) {
    int m2 = (n + 2) / 2;
    JacobiScratch js;
    double* d = reinterpret_cast<double*>(base);
    js.c = d; d += m2;
    js.s = d; d += m2;
    js.dii = d; d += m2;
    js.dkk = d; d += m2;
    js.red = d; d += 40;
    js.pi = reinterpret_cast<int*>(d);
    js.pk = js.pi + m2;
    return js;
}

// A: n x n symmetric, column-major with leading dimension lda (full storage, both
// triangles valid).  V: n x n (ldv), overwritten with the eigenvectors (columns).
// On exit diag(A) holds the eigenvalues (unsorted).  Every thread of the block must
// call this.  Returns the number of sweeps performed.
__device__ inline int jacobi_eigh_smem(int n, double* A, int lda, double* V, int ldv, JacobiScratch js) {
    const int tid = threadIdx.x, nt = blockDim.x;
    // V = I, fro^2
    double fro2 = 0.0;
    for (int idx = tid; idx < n * n; idx += nt) {
        int r = idx % n, c = idx / n;
        V[r + c * ldv] = (r == c) ? 1.0 : 0.0;
        double a = A[r + c * lda];
        fro2 += a * a;
    }
    fro2 = block_sum(fro2, js.red);
    if (n <= 1 || fro2 == 0.0) return 0;
    const double skip = sqrt(fro2) * 1.0e-17;
    const int m = (n + 1) & ~1;       // even number of players (one dummy if n is odd)
    const int npairs = m / 2;
    int sweep = 0;
    for (; sweep < 40; ++sweep) {
        int rotated_in_sweep = 0;
        for (int round = 0; round < m - 1; ++round) {
            // 1. rotation parameters for this round's disjoint pairs
            int any = 0;
            if (tid < npairs) {
                int a_, b_;
                if (tid == 0) { a_ = m - 1; b_ = round; }
                else { a_ = (round + tid) % (m - 1); b_ = (round - tid + (m - 1)) % (m - 1); }
                int i = a_ < b_ ? a_ : b_, k = a_ < b_ ? b_ : a_;
                double c = 1.0, s = 0.0, dii = 0.0, dkk = 0.0;
                if (k < n) {
                    double aik = A[i + k * lda], aii = A[i + i * lda], akk = A[k + k * lda];
                    dii = aii; dkk = akk;
                    if (fabs(aik) > skip) {
                        double theta = (akk - aii) / (2.0 * aik);
                        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
                        if (theta < 0.0) t = -t;
                        c = 1.0 / sqrt(t * t + 1.0);
                        s = t * c;
                        dii = aii - t * aik;
                        dkk = akk + t * aik;
                        any = 1;
                    }
                } else {
                    i = -1;   // pair with the dummy player: nothing to do
                }
                js.c[tid] = c; js.s[tid] = s; js.dii[tid] = dii; js.dkk[tid] = dkk;
                js.pi[tid] = i; js.pk[tid] = k;
            }
            any = __syncthreads_or(any);
            if (!any) continue;
            rotated_in_sweep = 1;
            // 2. column pass: A <- A J, V <- V J   (items = pair x row)
            for (int item = tid; item < npairs * n; item += nt) {
                int p = item / n, r = item - p * n;
                int i = js.pi[p];
                double s = js.s[p];
                if (i < 0 || s == 0.0) continue;
                int k = js.pk[p];
                double c = js.c[p];
                double ai = A[r + i * lda], ak = A[r + k * lda];
                A[r + i * lda] = c * ai - s * ak;
                A[r + k * lda] = s * ai + c * ak;
                double vi = V[r + i * ldv], vk = V[r + k * ldv];
                V[r + i * ldv] = c * vi - s * vk;
                V[r + k * ldv] = s * vi + c * vk;
            }
            __syncthreads();
            // 3. row pass: A <- J' A, with the 2x2 pivot blocks set analytically
            for (int item = tid; item < npairs * n; item += nt) {
                int p = item / n, q = item - p * n;
                int i = js.pi[p];
                double s = js.s[p];
                if (i < 0 || s == 0.0) continue;
                int k = js.pk[p];
                double c = js.c[p];
                if (q == i) { A[i + q * lda] = js.dii[p]; A[k + q * lda] = 0.0; }
                else if (q == k) { A[i + q * lda] = 0.0; A[k + q * lda] = js.dkk[p]; }
                else {
                    double ai = A[i + q * lda], ak = A[k + q * lda];
                    A[i + q * lda] = c * ai - s * ak;
                    A[k + q * lda] = s * ai + c * ak;
                }
            }
            __syncthreads();
        }
        if (!rotated_in_sweep) break;
    }
    return sweep;
}

// ---------------------------------------------------------------------------
// Low-latency variant used on the Lanczos critical path (K x K Rayleigh quotient):
//   * the matrix is padded to an even dimension m (zero row/column), so there is no dummy player;
//   * A is double-buffered: one fused pass computes  A_new = J' A_old J  per 2x2 block
//     (rotation pair p for the rows, pair q for the columns) together with V <- V J, so a round
//     costs two block barriers instead of three;
//   * rotation parameters use rsqrt + a Newton-refined reciprocal (no IEEE divide/sqrt sequences
//     on the dependent chain).
// A0/A1: m x m (lda >= m), A0 holds the input (both triangles, padding zeroed by the caller);
// V: m x m (ldv).  Returns the buffer (A0 or A1) whose diagonal holds the eigenvalues.
// scratch js must be carved for dimension m.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp_pos(double x) {
    // x > 0.  rcp.approx (>= 20 bits) + 2 Newton steps; falls back to the IEEE divide outside the safe range.
    if (!(x > 1e-290 && x < 1e290)) return 1.0 / x;
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

// rotation sweeps shared by the cold and warm entry points (V already initialised)
__device__ inline double* jacobi_sweeps_fast(int m, double* A0, double* A1, int lda, double* V, int ldv,
                                             JacobiScratch js, double fro2, int max_sweeps = 40) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (m < 2 || fro2 == 0.0) return A0;
    const double skip = sqrt(fro2) * 1.0e-17;
    const int npairs = m / 2;
    const int nblk = npairs * npairs;
    const int nitems = nblk + m * npairs;
    // the item -> (pair, pair) / (pair, row) decode does not change between rounds: do it once when every
    // thread owns at most one item (K = 25: 507 items for 512 threads)
    const bool single = nitems <= nt;
    int my_a = 0, my_b = 0, my_kind = -1;      // kind 0: block (p = a, q = b); kind 1: V item (q = a, r = b)
    if (single && tid < nitems) {
        if (tid < nblk) { my_kind = 0; my_a = tid / npairs; my_b = tid - my_a * npairs; }
        else { my_kind = 1; int it = tid - nblk; my_a = it / m; my_b = it - my_a * m; }
    }
    double* A = A0;
    double* B = A1;
    int nrounds = 0;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        int rotated_in_sweep = 0;
        for (int round = 0; round < m - 1; ++round) {
            int any = 0;
            if (tid < npairs) {
                int a_, b_;
                if (tid == 0) { a_ = m - 1; b_ = round; }
                else { a_ = (round + tid) % (m - 1); b_ = (round - tid + (m - 1)) % (m - 1); }
                const int i = a_ < b_ ? a_ : b_, k = a_ < b_ ? b_ : a_;
                const double aik = A[i + k * lda], aii = A[i + i * lda], akk = A[k + k * lda];
                double c = 1.0, s = 0.0, dii = aii, dkk = akk;
                if (fabs(aik) > skip) {
                    const double del = akk - aii;
                    const double w = fma(del, del, 4.0 * aik * aik);
                    const double sq = w * rsqrt(w);
                    double t = (2.0 * aik) * fast_rcp_pos(fabs(del) + sq);
                    if (del < 0.0) t = -t;
                    c = rsqrt(fma(t, t, 1.0));
                    s = t * c;
                    dii = aii - t * aik;
                    dkk = akk + t * aik;
                    any = 1;
                }
                js.c[tid] = c; js.s[tid] = s; js.dii[tid] = dii; js.dkk[tid] = dkk;
                js.pi[tid] = i; js.pk[tid] = k;
            }
            any = __syncthreads_or(any);
            if (!any) continue;
            rotated_in_sweep = 1;
            ++nrounds;
            for (int item = tid; item < nitems; item += nt) {
                int kind, ia, ib;
                if (single) { kind = my_kind; ia = my_a; ib = my_b; }
                else if (item < nblk) { kind = 0; ia = item / npairs; ib = item - ia * npairs; }
                else { kind = 1; int it = item - nblk; ia = it / m; ib = it - ia * m; }
                if (kind == 0) {
                    const int p = ia, q = ib;
                    const int ip = js.pi[p], kp = js.pk[p], iq = js.pi[q], kq = js.pk[q];
                    if (p == q) {
                        B[ip + ip * lda] = js.dii[p]; B[kp + kp * lda] = js.dkk[p];
                        B[ip + kp * lda] = 0.0; B[kp + ip * lda] = 0.0;
                    } else {
                        const double cp = js.c[p], sp = js.s[p], cq = js.c[q], sq_ = js.s[q];
                        const double b00 = A[ip + iq * lda], b01 = A[ip + kq * lda];
                        const double b10 = A[kp + iq * lda], b11 = A[kp + kq * lda];
                        // columns: (col_i, col_k) <- (c col_i - s col_k, s col_i + c col_k)
                        const double c00 = cq * b00 - sq_ * b01, c01 = sq_ * b00 + cq * b01;
                        const double c10 = cq * b10 - sq_ * b11, c11 = sq_ * b10 + cq * b11;
                        // rows: (row_i, row_k) <- (c row_i - s row_k, s row_i + c row_k)
                        B[ip + iq * lda] = cp * c00 - sp * c10; B[ip + kq * lda] = cp * c01 - sp * c11;
                        B[kp + iq * lda] = sp * c00 + cp * c10; B[kp + kq * lda] = sp * c01 + cp * c11;
                    }
                } else {
                    const int q = ia, r = ib;
                    const double sq_ = js.s[q];
                    if (sq_ != 0.0) {
                        const int iq = js.pi[q], kq = js.pk[q];
                        const double cq = js.c[q];
                        const double vi = V[r + iq * ldv], vk = V[r + kq * ldv];
                        V[r + iq * ldv] = cq * vi - sq_ * vk;
                        V[r + kq * ldv] = sq_ * vi + cq * vk;
                    }
                }
            }
            __syncthreads();
            double* t = A; A = B; B = t;
        }
        if (!rotated_in_sweep) break;
    }
    if (tid == 0) js.red[38] = (double)nrounds;
    return A;
}

__device__ inline double* jacobi_eigh_smem_fast(int m, double* A0, double* A1, int lda, double* V, int ldv,
                                                JacobiScratch js, int max_sweeps = 40) {
    const int tid = threadIdx.x, nt = blockDim.x;
    double fro2 = 0.0;
    for (int idx = tid; idx < m * m; idx += nt) {
        int r = idx % m, c = idx / m;
        V[r + c * ldv] = (r == c) ? 1.0 : 0.0;
        double a = A0[r + c * lda];
        fro2 += a * a;
    }
    fro2 = block_sum(fro2, js.red);
    return jacobi_sweeps_fast(m, A0, A1, lda, V, ldv, js, fro2, max_sweeps);
}

// Warm start: W (m x m, leading dimension ldv, global memory) is an orthogonal basis that nearly diagonalises
// A0 — the eigenvectors of the previous PDHG iteration's Rayleigh quotient.  Forms W' A0 W, then sweeps
// (2-3 instead of ~8).  Jacobi converges from any orthogonal start, so the result is the eigendecomposition
// of A0 to working precision either way; only the rounding path differs from the cold start.
__device__ inline double* jacobi_eigh_smem_warm(int m, int k, double* A0, double* A1, int lda, double* V, int ldv,
                                                const double* W, JacobiScratch js) {
    const int tid = threadIdx.x, nt = blockDim.x;
    (void)k;
    for (int idx = tid; idx < m * m; idx += nt) { int r = idx % m, c = idx / m; V[r + c * ldv] = W[r + c * ldv]; }
    __syncthreads();
    // A1 = A0 * V
    for (int idx = tid; idx < m * m; idx += nt) {
        int r = idx % m, c = idx / m;
        double s = 0.0;
        for (int i = 0; i < m; ++i) s = fma(A0[r + i * lda], V[i + c * ldv], s);
        A1[r + c * lda] = s;
    }
    __syncthreads();
    // A0 = V' * A1 (upper triangle computed, mirrored: exactly symmetric)
    double fro2 = 0.0;
    for (int idx = tid; idx < m * m; idx += nt) {
        int r = idx % m, c = idx / m;
        if (r <= c) {
            double s = 0.0;
            for (int i = 0; i < m; ++i) s = fma(V[i + r * ldv], A1[i + c * lda], s);
            A0[r + c * lda] = s;
            if (r != c) A0[c + r * lda] = s;
            fro2 += (r == c) ? s * s : 2.0 * s * s;
        }
    }
    fro2 = block_sum(fro2, js.red);
    return jacobi_sweeps_fast(m, A0, A1, lda, V, ldv, js, fro2);
}

// order[r] = index of the r-th largest diagonal entry of A (descending, ties by index).
// n threads participate; caller syncs afterwards.
__device__ inline void rank_sort_desc(int n, const double* A, int lda, int* order) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double di = A[i + i * lda];
        int pos = 0;
        for (int j = 0; j < n; ++j) {
            double dj = A[j + j * lda];
            pos += (dj > di) || (dj == di && j < i);
        }
        order[pos] = i;
    }
}

}  // namespace pb
