// fulleig.cuh — exact (full spectrum) PSD projections.
//
//   k_small_cone_proj : one CTA per small PSD cone (side <= SMALL_CONE_MAX): fused
//       primal update + svec->matrix + Jacobi eigendecomposition + positive clip +
//       V diag(lambda+) V' + matrix->svec, all in shared memory.  Replaces, per cone,
//       pdhg.jl:622 + prox_operators.jl:1-16 + full_eig! (prox_operators.jl:111-126,
//       LAPACK dsyevr + rank-1 dgemm loop) + prox_operators.jl:17-31.  Batched over all
//       small cones of the problem in one launch (the reference walks them serially).
//
//   block-Jacobi path for large cones (fallback / full_eig_decomp / target_rank > 16):
//       k_bj_pair_eig + k_bj_apply — see below.
#pragma once
#include "common.cuh"
#include "jacobi.cuh"
#include "kernels_vec.cuh"

namespace pb {

constexpr int SMALL_CONE_MAX = 100;   // = default min_size_krylov_eigs: such cones never take the Krylov path

// fast = 1: three matrices of the even-padded side m (double-buffered Jacobi, jacobi_sweeps_fast); fast = 0: two
// matrices, in-place three-pass Jacobi (sides 94 .. 100 do not fit three matrices in 227 KB)
__host__ __device__ inline int small_cone_m(int n) { return (n + 1) & ~1; }
__host__ __device__ inline size_t small_cone_smem_bytes(int n, int fast = 0) {
    if (fast) {
        const int m = small_cone_m(n), lda = m | 1;
        return sizeof(double) * (3 * (size_t)m * lda + 2 * (size_t)n) + sizeof(int) * (size_t)((n + 2 + 3) & ~3) +
               jacobi_scratch_bytes(m) + 64;
    }
    int lda = n | 1;
    return sizeof(double) * (2 * (size_t)n * lda + 2 * (size_t)n) + sizeof(int) * (size_t)((n + 2 + 3) & ~3) +
           jacobi_scratch_bytes(n) + 64;
}

// mode 0: projection (writes x_out, rank, min_eig).  mode 1: eigenvalues only — writes the
// minimum eigenvalue of mat(v*scale) to out_min[cone] (cone_feas, pdhg.jl:678-699).
struct SmallConeArgs {
    const int* cone_ids;         // small cones handled by this launch
    const int* cone_side;        // per cone
    const long long* cone_off;   // per cone svec offset
    const double* x; const double* Mty; const double* c;
    double tau, tol_psd;
    double* x_out;
    double* scal;
    int mode;
    double scale;
    double* out_min;
    int fast;                    // 1: double-buffered Jacobi on three shared-memory matrices (see small_cone_smem_bytes)
    // warm start (fast path, mode 0): the eigenvector basis of each cone's previous projection, m x (m | 1) doubles per
    // cone at warm + blockIdx.x * warm_stride.  Consecutive PDHG iterates have nearly the same eigenvectors, so
    // W' A W is nearly diagonal and 2-3 sweeps replace ~9; Jacobi converges from any orthogonal start, only the
    // rounding path differs.  warm_read = 0 on the first projection of a solve and every 32nd one (so that rounding
    // in the accumulated rotations cannot build up).
    double* warm;
    long long warm_stride;
    int warm_read;
};

__global__ void __launch_bounds__(512) k_small_cone_proj(SmallConeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cone = a.cone_ids[blockIdx.x];
    const int n = a.cone_side[cone];
    const long long off = a.cone_off[cone];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lda = n | 1;
    const int tri = n * (n + 1) / 2;
    const double sqrt2 = 1.41421356237309504880;

    if (n == 1) {   // prox_operators.jl:43-45
        if (tid == 0) {
            if (a.mode == 0) {
                double u = sub_rn(a.x[off], mul_rn(a.tau, add_rn(a.Mty[off], a.c[off])));
                u = fmax(0.0, u);
                a.x_out[off] = u;
                a.scal[S_HEADER + 3 * cone + 0] = 0.0;
                a.scal[S_HEADER + 3 * cone + 1] = u;
                a.scal[S_HEADER + 3 * cone + 2] = -1.0;
            } else {
                a.out_min[cone] = a.x[off] * a.scale;
            }
        }
        return;
    }
    // fast path: side padded to the even m with a zero row/column (an exactly decoupled zero eigenvalue that no
    // rotation ever touches); ld is the leading dimension of every matrix of this cone from here on
    const int m = a.fast ? small_cone_m(n) : n;
    const int ld = a.fast ? (m | 1) : lda;
    double* A = reinterpret_cast<double*>(smem_raw);
    double* V = A + (size_t)m * ld;
    double* A1 = V + (size_t)m * ld;                       // fast path only
    double* lam = a.fast ? A1 + (size_t)m * ld : A1;       // n
    double* lamp = lam + n;                // n: positive eigenvalues (compacted)
    int* pos = reinterpret_cast<int*>(lamp + n);   // n + 2
    JacobiScratch js = jacobi_carve(reinterpret_cast<void*>(pos + ((n + 2 + 3) & ~3)), m);

    // svec -> full symmetric matrix (fused primal update)
    if (m > n) for (int i = tid; i < m; i += nt) { A[i + n * ld] = 0.0; A[n + i * ld] = 0.0; }
    for (int k = tid; k < tri; k += nt) {
        int j = (int)((sqrt(8.0 * (double)k + 1.0) - 1.0) * 0.5);
        while ((j + 1) * (j + 2) / 2 <= k) ++j;
        while (j * (j + 1) / 2 > k) --j;
        int i = k - j * (j + 1) / 2;
        double u;
        if (a.mode == 0) u = sub_rn(a.x[off + k], mul_rn(a.tau, add_rn(a.Mty[off + k], a.c[off + k])));
        else u = a.x[off + k] * a.scale;
        double v = (i != j) ? u / sqrt2 : u;
        A[i + j * ld] = v;
        A[j + i * ld] = v;
    }
    __syncthreads();
    if (a.fast) {                                                       // eigenvalues on the diagonal of the returned buffer
        double* W = (a.mode == 0 && a.warm) ? a.warm + (size_t)blockIdx.x * (size_t)a.warm_stride : nullptr;
        if (W && a.warm_read) A = jacobi_eigh_smem_warm(m, m, A, A1, ld, V, ld, W, js);
        else A = jacobi_eigh_smem_fast(m, A, A1, ld, V, ld, js);
        __syncthreads();
        if (W) for (int idx = tid; idx < m * ld; idx += nt) W[idx] = V[idx];
    } else {
        jacobi_eigh_smem(n, A, ld, V, ld, js);
    }
    __syncthreads();
    if (a.mode == 1) {
        double mn = 1.0e300;
        for (int i = tid; i < n; i += nt) mn = fmin(mn, A[i + i * ld]);
        mn = -block_max_id(-mn, js.red, -1.0e300);
        if (tid == 0) a.out_min[cone] = mn;
        return;
    }
    // clip: keep lambda > 0; rank counts lambda > tol_psd (prox_operators.jl:116-124)
    if (tid == 0) {
        int np = 0, rk = 0;
        for (int i = 0; i < n; ++i) {
            double l = A[i + i * ld];
            if (l > 0.0) { pos[np] = i; lamp[np] = l; np++; if (l > a.tol_psd) rk++; }
        }
        pos[n] = np;
        a.scal[S_HEADER + 3 * cone + 0] = (double)rk;
        a.scal[S_HEADER + 3 * cone + 1] = 0.0;      // full_eig! sets min_eig = 0.0
        a.scal[S_HEADER + 3 * cone + 2] = -1.0;
    }
    __syncthreads();
    const int np = pos[n];
    // X+ = sum lambda v v' written straight to svec (off-diagonals * sqrt(2))
    for (int k = tid; k < tri; k += nt) {
        int j = (int)((sqrt(8.0 * (double)k + 1.0) - 1.0) * 0.5);
        while ((j + 1) * (j + 2) / 2 <= k) ++j;
        while (j * (j + 1) / 2 > k) --j;
        int i = k - j * (j + 1) / 2;
        double s = 0.0;
        for (int q = 0; q < np; ++q) {
            int col = pos[q];
            s = fma(lamp[q] * V[i + col * ld], V[j + col * ld], s);
        }
        a.x_out[off + k] = (i != j) ? s * sqrt2 : s;
    }
}


// ===========================================================================
// Block-Jacobi full eigendecomposition for large cones (any n), device resident.
//
// The matrix (padded to NP = 32*nb rows/cols with zeros) is cut into nb blocks of 32.
// A sweep is a round-robin tournament over block pairs (I, J); per round
//   1. k_bj_pair_eig : one CTA per pair rotates the 64x64 pivot sub-matrix [A_II A_IJ; A_JI A_JJ] in shared
//                      memory: `inner_sweeps` cyclic Jacobi sweeps (double-buffered two-sided rounds), NOT to
//                      convergence — every rotation lowers the off-diagonal norm of the whole matrix by exactly
//                      2 a_pq^2, so the outer iteration converges whatever the inner effort, and one or two inner
//                      sweeps cost 60-130 dependent rounds instead of ~500 -> Q and the rotated pivot block P
//   2. k_bj_apply<COLS>: A[:, I∪J] <- A[:, I∪J] Q ;  V[:, I∪J] <- V[:, I∪J] Q
//   3. k_bj_apply<ROWS>: A[I∪J, :] <- Q' A[I∪J, :]   (pivot block set to P, which is symmetric to the last bit)
// Pairs of one round touch disjoint block rows/columns, so each pass is one launch; the pair of a CTA is computed
// from (round, blockIdx) on the device — the host only launches.
// Used for full_eig! on cones with side > SMALL_CONE_MAX (Krylov fallback,
// full_eig_decomp = true, target_rank > max_target_rank_krylov_eigs) and for cone_feas.
// ===========================================================================
constexpr int BJ_B = 32;          // block size
constexpr int BJ_P = 2 * BJ_B;    // pivot size
constexpr int BJ_LD = BJ_P + 1;

struct BjRound {       // the tournament round: CTA t handles the pair below
    int round, mplayers, nb;
};
__device__ __forceinline__ void bj_pair_of(const BjRound& r, int t, int& I, int& J) {
    int a_, b_;
    if (r.mplayers == 2) { a_ = 0; b_ = 1; }
    else if (t == 0) { a_ = r.mplayers - 1; b_ = r.round; }
    else { a_ = (r.round + t) % (r.mplayers - 1); b_ = (r.round - t + (r.mplayers - 1)) % (r.mplayers - 1); }
    I = min(a_, b_); J = max(a_, b_);
    if (J >= r.nb) { I = -1; J = -1; }        // pair with the dummy player of an odd tournament
}

__host__ __device__ inline size_t bj_pair_smem_bytes() {
    return sizeof(double) * (3 * (size_t)BJ_P * BJ_LD) + jacobi_scratch_bytes(BJ_P) + 64;
}

// Q layout: per pair, Q (BJ_P x BJ_P column-major, ld = BJ_P) followed by the rotated pivot block P (same layout).
constexpr int BJ_QSTRIDE = 2 * BJ_P * BJ_P;
constexpr size_t BJ_APPLY_SMEM = sizeof(double) * 2 * BJ_P * (BJ_P + 1);

__global__ void __launch_bounds__(512)
k_bj_pair_eig(double* __restrict__ A, int ld, BjRound pr, int inner_sweeps, double* __restrict__ Qbuf, int* __restrict__ rotated) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int I, J;
    bj_pair_of(pr, blockIdx.x, I, J);
    if (I < 0) return;
    double* S0 = reinterpret_cast<double*>(smem_raw);
    double* S1 = S0 + (size_t)BJ_P * BJ_LD;
    double* U = S1 + (size_t)BJ_P * BJ_LD;
    JacobiScratch js = jacobi_carve(reinterpret_cast<void*>(U + (size_t)BJ_P * BJ_LD), BJ_P);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int idx = tid; idx < BJ_P * BJ_P; idx += nt) {
        int r = idx % BJ_P, c = idx / BJ_P;
        int gr = (r < BJ_B ? I * BJ_B + r : J * BJ_B + (r - BJ_B));
        int gc = (c < BJ_B ? I * BJ_B + c : J * BJ_B + (c - BJ_B));
        S0[r + c * BJ_LD] = A[(size_t)gr + (size_t)gc * ld];
    }
    __syncthreads();
    // symmetrise (the two triangles may differ by rounding after the GEMM passes)
    for (int idx = tid; idx < BJ_P * BJ_P; idx += nt) {
        int r = idx % BJ_P, c = idx / BJ_P;
        if (r < c) {
            double v = 0.5 * (S0[r + c * BJ_LD] + S0[c + r * BJ_LD]);
            S0[r + c * BJ_LD] = v; S0[c + r * BJ_LD] = v;
        }
    }
    __syncthreads();
    const double* P = jacobi_eigh_smem_fast(BJ_P, S0, S1, BJ_LD, U, BJ_LD, js, inner_sweeps);
    __syncthreads();
    const int nrounds = (int)js.red[38];
    double* Q = Qbuf + (size_t)blockIdx.x * BJ_QSTRIDE;
    for (int idx = tid; idx < BJ_P * BJ_P; idx += nt) {
        int r = idx % BJ_P, c = idx / BJ_P;
        Q[idx] = U[r + c * BJ_LD];
        // rotated pivot block, symmetrised (both triangles carry the same rotations up to rounding)
        Q[BJ_P * BJ_P + idx] = (r == c) ? P[r + c * BJ_LD] : 0.5 * (P[r + c * BJ_LD] + P[c + r * BJ_LD]);
    }
    if (tid == 0 && nrounds > 0) atomicAdd(rotated, 1);
}

// MODE 0: columns of A and V;  MODE 1: rows of A.  grid = (pairs, NP/64 chunks, MODE==0 ? 2 : 1)
template <int MODE>
__global__ void __launch_bounds__(256)
k_bj_apply(double* __restrict__ A, double* __restrict__ V, int ld, int NP, BjRound pr,
           const double* __restrict__ Qbuf) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double (*Qs)[BJ_P + 1] = reinterpret_cast<double (*)[BJ_P + 1]>(smem_raw);
    double (*Ts)[BJ_P + 1] = Qs + BJ_P;
    int I, J;
    bj_pair_of(pr, blockIdx.x, I, J);
    if (I < 0) return;
    const double* Q = Qbuf + (size_t)blockIdx.x * BJ_QSTRIDE;
    const int tid = threadIdx.x;
    double* M = (MODE == 0 && blockIdx.z == 1) ? V : A;
    const int chunk0 = blockIdx.y * BJ_P;
    for (int idx = tid; idx < BJ_P * BJ_P; idx += 256) {
        int r = idx % BJ_P, c = idx / BJ_P;
        Qs[r][c] = Q[idx];      // Qs[r][c] = Q(r, c)
    }
    // load tile: MODE 0 -> Ts[rr][cc] = M(chunk0 + rr, pivot col cc); MODE 1 -> Ts[rr][cc] = M(pivot row rr, chunk0 + cc)
    for (int idx = tid; idx < BJ_P * BJ_P; idx += 256) {
        int a_ = idx % BJ_P, b_ = idx / BJ_P;
        if (MODE == 0) {
            int gr = chunk0 + a_;
            int gc = (b_ < BJ_B ? I * BJ_B + b_ : J * BJ_B + (b_ - BJ_B));
            Ts[a_][b_] = (gr < NP) ? M[(size_t)gr + (size_t)gc * ld] : 0.0;
        } else {
            int gr = (a_ < BJ_B ? I * BJ_B + a_ : J * BJ_B + (a_ - BJ_B));
            int gc = chunk0 + b_;
            Ts[a_][b_] = (gc < NP) ? M[(size_t)gr + (size_t)gc * ld] : 0.0;
        }
    }
    __syncthreads();
    // each thread computes a 4x4 patch of the 64x64 product
    const int tr = (tid % 16) * 4, tc = (tid / 16) * 4;
    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
    for (int k = 0; k < BJ_P; ++k) {
        double av[4], bv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0) { av[u] = Ts[tr + u][k]; bv[u] = Qs[k][tc + u]; }      // T * Q
            else           { av[u] = Qs[k][tr + u]; bv[u] = Ts[k][tc + u]; }      // Q' * T
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] = fma(av[u], bv[v], acc[u][v]);
    }
    const double* Pm = Q + BJ_P * BJ_P;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            int a_ = tr + u, b_ = tc + v;
            if (MODE == 0) {
                int gr = chunk0 + a_;
                int gc = (b_ < BJ_B ? I * BJ_B + b_ : J * BJ_B + (b_ - BJ_B));
                if (gr < NP) M[(size_t)gr + (size_t)gc * ld] = acc[u][v];
            } else {
                int gr = (a_ < BJ_B ? I * BJ_B + a_ : J * BJ_B + (a_ - BJ_B));
                int gc = chunk0 + b_;
                if (gc < NP) {
                    double val = acc[u][v];
                    // pivot block: the block the pair kernel rotated in shared memory (exactly symmetric)
                    int pc = -1;
                    if (gc >= I * BJ_B && gc < (I + 1) * BJ_B) pc = gc - I * BJ_B;
                    else if (gc >= J * BJ_B && gc < (J + 1) * BJ_B) pc = BJ_B + gc - J * BJ_B;
                    if (pc >= 0) val = Pm[a_ + pc * BJ_P];
                    M[(size_t)gr + (size_t)gc * ld] = val;
                }
            }
        }
}

// C = op(A) B for NP x NP matrices (NP a multiple of 64, column-major, leading dimension ld), op = transpose when TA.
// Plain tiled FP64 product (64 x 64 tile per CTA, 4 x 4 patch per thread, k in slices of 16): used twice per warm-started
// eigendecomposition (W' (A W)), i.e. a few % of it; FP64 has no tcgen05 kind, so this is CUDA-core FMA by construction.
template <bool TA>
__global__ void __launch_bounds__(256)
k_gemm64(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int NP, int ld) {
    __shared__ double As[16][64 + 4];     // As[k][i] = op(A)(i0 + i, k0 + k)
    __shared__ double Bs[16][64 + 4];     // Bs[k][j] = B(k0 + k, j0 + j)
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
    const int tr = (tid % 16) * 4, tc = (tid / 16) * 4;
    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
    for (int k0 = 0; k0 < NP; k0 += 16) {
        for (int idx = tid; idx < 16 * 64; idx += 256) {
            if (TA) { const int k = idx % 16, i = idx / 16; As[k][i] = A[(size_t)(k0 + k) + (size_t)(i0 + i) * ld]; }      // A'(i, k) = A(k, i): k contiguous
            else    { const int i = idx % 64, k = idx / 64; As[k][i] = A[(size_t)(i0 + i) + (size_t)(k0 + k) * ld]; }
            { const int k = idx % 16, j = idx / 16; Bs[k][j] = B[(size_t)(k0 + k) + (size_t)(j0 + j) * ld]; }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            double av[4], bv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { av[u] = As[k][tr + u]; bv[u] = Bs[k][tc + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fma(av[u], bv[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) C[(size_t)(i0 + tr + u) + (size_t)(j0 + tc + v) * ld] = acc[u][v];
}

// zero padding of A beyond n (warm start: V keeps the previous eigenvectors)
__global__ void k_bj_pad(double* __restrict__ A, int ld, int n, int NP) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)NP * NP;
    if (idx >= tot) return;
    int r = (int)(idx % NP), c = (int)(idx / NP);
    if (r >= n || c >= n) A[(size_t)r + (size_t)c * ld] = 0.0;
}

// V = I (NP x NP), zero padding of A beyond n
__global__ void k_bj_init(double* __restrict__ A, double* __restrict__ V, int ld, int n, int NP) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)NP * NP;
    if (idx >= tot) return;
    int r = (int)(idx % NP), c = (int)(idx / NP);
    V[(size_t)r + (size_t)c * ld] = (r == c) ? 1.0 : 0.0;
    if (r >= n || c >= n) A[(size_t)r + (size_t)c * ld] = 0.0;
}

// off-diagonal and total squared Frobenius norms of A (n x n) -> out[0], out[1]
__global__ void __launch_bounds__(256)
k_bj_offnorm(const double* __restrict__ A, int ld, int n, double* __restrict__ out, ReduceWs ws) {
    __shared__ double red[40];
    __shared__ int s_last;
    double off = 0.0, tot = 0.0;
    long long total = (long long)n * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int r = (int)(idx % n), c = (int)(idx / n);
        double v = A[(size_t)r + (size_t)c * ld];
        tot += v * v;
        if (r != c) off += v * v;
    }
    off = block_sum(off, red);
    tot = block_sum(tot, red);
    int nb = gridDim.x;
    if (threadIdx.x == 0) { ws.partials[blockIdx.x] = off; ws.partials[nb + blockIdx.x] = tot; }
    if (last_block_arrive(ws.counters + 4, &s_last)) {
        double a = 0.0, b = 0.0;
        for (int k = threadIdx.x; k < nb; k += blockDim.x) { a += __ldcg(ws.partials + k); b += __ldcg(ws.partials + nb + k); }
        a = block_sum(a, red);
        b = block_sum(b, red);
        if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
    }
}

// diag(A) -> out (n)
__global__ void k_bj_diag(const double* __restrict__ A, int ld, int n, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = A[(size_t)i + (size_t)i * ld];
}

}  // namespace pb
