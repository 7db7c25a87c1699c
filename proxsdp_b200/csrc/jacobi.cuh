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
