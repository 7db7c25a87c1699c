// lanczos.cuh — persistent cooperative thick-restart Lanczos for the top eigenpairs
// of a dense symmetric FP64 matrix, plus the fused reconstruction kernel.
//
// Replaces KrylovKit.eigsolve(A, resid, nev, :LR, Lanczos(orth, K, maxiter, tol))
// as called from reference src/eigsolver.jl:802-812 and its post-processing in
// reference src/prox_operators.jl:89-109 (clip lambda > 0, X = sum lambda v v').
//
// One launch = one eigsolve.  The grid (<= one CTA per SM, co-resident via a
// cooperative launch) owns contiguous row slabs of X.  Per Lanczos step:
//     symv slab x v  ->  CGS pass 1 partial dots  -> [grid barrier]
//     w -= V h ; pass 2 partial dots              -> [grid barrier]
//     w -= V h2 ; beta ; publish v_{j+1}          -> [grid barrier]
// i.e. three grid barriers per mat-vec, everything else CTA-local:
//   * the basis V lives only as per-CTA row slabs in shared memory — the single
//     global vector is the newest Lanczos vector, which every CTA stages for the
//     next mat-vec;
//   * the K x K Rayleigh quotient (tridiagonal, or arrow + tridiagonal after a thick
//     restart) is held redundantly by every CTA and diagonalised redundantly with the
//     shared-memory Jacobi solver, so no broadcast/extra barrier is needed and all
//     CTAs take identical control-flow decisions (bitwise identical inputs).
// Thick restart (Krylov-Schur): keep = (3K + 2*converged)/5 Ritz vectors + residual,
// the restarted Rayleigh quotient is diag(D[1:keep]) bordered by f[1:keep]; KrylovKit
// re-tridiagonalises that arrow by Householder reflections, which is only a change of
// basis inside the kept subspace (same Ritz values/vectors).
#pragma once
#include "common.cuh"
#include "jacobi.cuh"
#include "kernels_vec.cuh"

namespace pb {

struct LanczosArgs {
    const double* X;     // n x n symmetric, column-major, leading dimension ld (multiple of 64, finite padding)
    int n, ld;
    const double* x0;    // start vector (n)
    double* Y;           // out: Ritz vectors, n x K, leading dimension ld
    uint4* xbuf;         // exchange buffer for partial dot products: [2][K + 2][G] flagged doubles
    uint4* vx;           // exchange buffer for the newest Lanczos vector: [2][ld] flagged doubles
    unsigned int epoch_base;   // flags used by this launch are epoch_base + 1, +2, ...
    int nev, K, maxiter;
    double tol;
    int rows_max;        // max rows owned by a CTA = ceil(n / grid)
    double* vals;        // out: K Ritz values (descending)
    int* info;           // out: [0] nvals, [1] converged, [2] numops, [3] numiter
    double* scal;        // iteration scalar record (S_POISON, S_NUMOPS, per-cone slots)
    int cone;            // cone index for the per-cone slots
    int n_cones_total;
    int jac_inplace;     // 1: the three K x K Jacobi matrices do not fit shared memory -> two-matrix in-place solver
    long long* prof;     // optional (debug): per-phase clock64 totals of CTA 0, 8 slots
};

// ---------------------------------------------------------------------------
// Flagged ("LL") exchange: a double travels as two 8-byte words {lo32, flag}, {hi32, flag}.
// 8-byte accesses are single-copy atomic, so a reader that sees the expected flag in both
// words has the whole value — data and "ready" signal arrive in the same L2 round trip and no
// fence, atomic or separate barrier is needed.  Every CTA reads every other CTA's words, so
// completing the read IS the grid-wide barrier.  Buffers alternate with the exchange parity;
// a buffer is reused only two exchanges later, when every CTA has provably finished reading it.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned int flag) {
    unsigned int lo = (unsigned int)__double2loint(v), hi = (unsigned int)__double2hiint(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(flag), "r"(hi), "r"(flag)
                 : "memory");
}
__device__ __forceinline__ uint4 ll_peek(const uint4* p) {
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ double ll_value(const uint4& r) { return __hiloint2double((int)r.z, (int)r.x); }

// sum over the G per-CTA partials of one slot, lanes strided over CTAs (fixed order => deterministic)
__device__ __forceinline__ double ll_reduce_slot(const uint4* slot, int G, unsigned int flag, int lane) {
    double s = 0.0;
    for (int c0 = 0; c0 < G; c0 += 128) {
        const int i0 = c0 + lane, i1 = i0 + 32, i2 = i0 + 64, i3 = i0 + 96;
        uint4 r0, r1, r2, r3;
        bool ok;
        do {
            ok = true;
            if (i0 < G) { r0 = ll_peek(slot + i0); }
            if (i1 < G) { r1 = ll_peek(slot + i1); }
            if (i2 < G) { r2 = ll_peek(slot + i2); }
            if (i3 < G) { r3 = ll_peek(slot + i3); }
            if (i0 < G) ok = ok && r0.y == flag && r0.w == flag;
            if (i1 < G) ok = ok && r1.y == flag && r1.w == flag;
            if (i2 < G) ok = ok && r2.y == flag && r2.w == flag;
            if (i3 < G) ok = ok && r3.y == flag && r3.w == flag;
        } while (!ok);
        if (i0 < G) s += ll_value(r0);
        if (i1 < G) s += ll_value(r1);
        if (i2 < G) s += ll_value(r2);
        if (i3 < G) s += ll_value(r3);
    }
    return warp_sum(s);
}

// shared-memory carve-up (host mirrors this in lanczos_smem_bytes)
struct LanczosSmem {
    double* vbuf;    // cpr * 64: the newest Lanczos vector, zero padded
    double* slabA;   // (K+1) * RLp
    double* slabB;   // (K+1) * RLp
    double* wloc;    // RLp
    double* wpart;   // LZ_NW * LZ_TMAX
    double* hred;    // K + 2
    double* Hd;      // K      diag of the Rayleigh quotient
    double* He;      // K      sub-diagonal (tridiagonal part)
    double* Harr;    // K      arrow row (entries coupling column `arrow_at` to the kept Ritz values)
    double* D;       // K      sorted Ritz values
    double* f;       // K      Ritz residuals beta * U[k-1, :]
    double* JA;      // Kp * Kp   (Kp = K + 2: even padding + odd leading dimension)
    double* JB;      // Kp * Kp
    double* JU;      // Kp * Kp
    int* order;      // K
    int* wgs;        // LZ_NW + 1 chunk boundaries of the warps
    void* jscratch;
};

constexpr int LZ_THREADS = 512;
constexpr int LZ_NW = LZ_THREADS / 32;
constexpr int LZ_TMAX = 10;    // rows a warp's chunk range may touch at most (checked on the host)

__host__ __device__ inline int lanczos_rlp(int rows_max) { return rows_max | 1; }
__host__ __device__ inline int lanczos_kp(int K) { return (K + 2) | 1; }
__host__ __device__ inline int lanczos_cpr(int n) { return (n + 63) / 64; }

__host__ __device__ inline size_t lanczos_smem_bytes(int K, int rows_max, int n, int jac_inplace = 0) {
    size_t RLp = (size_t)lanczos_rlp(rows_max), Kp = (size_t)lanczos_kp(K);
    size_t d = (size_t)lanczos_cpr(n) * 64 + 2 * (size_t)(K + 1) * RLp + RLp + (size_t)LZ_NW * LZ_TMAX + (size_t)(K + 2) +
               5 * (size_t)K + (jac_inplace ? 2 : 3) * Kp * Kp;
    return d * sizeof(double) + sizeof(int) * (size_t)(((K + 2 + 3) & ~3) + 32) + jacobi_scratch_bytes((int)Kp) + 64;
}

__device__ inline LanczosSmem lanczos_carve(unsigned char* base, int K, int rows_max, int n, int jac_inplace) {
    LanczosSmem s;
    size_t RLp = (size_t)lanczos_rlp(rows_max), Kp = (size_t)lanczos_kp(K);
    double* d = reinterpret_cast<double*>(base);
    s.vbuf = d; d += (size_t)lanczos_cpr(n) * 64;
    s.slabA = d; d += (size_t)(K + 1) * RLp;
    s.slabB = d; d += (size_t)(K + 1) * RLp;
    s.wloc = d; d += RLp;
    s.wpart = d; d += LZ_NW * LZ_TMAX;
    s.hred = d; d += K + 2;
    s.Hd = d; d += K;
    s.He = d; d += K;
    s.Harr = d; d += K;
    s.D = d; d += K;
    s.f = d; d += K;
    s.JA = d; d += Kp * Kp;
    s.JB = d; if (!jac_inplace) d += Kp * Kp;      // in-place Jacobi (huge K): JB is not used
    s.JU = d; d += Kp * Kp;
    s.order = reinterpret_cast<int*>(d);
    s.wgs = s.order + ((K + 2 + 3) & ~3);
    s.jscratch = reinterpret_cast<void*>(s.wgs + 32);
    return s;
}

__global__ void __launch_bounds__(LZ_THREADS, 1) k_lanczos(LanczosArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int n = a.n, ld = a.ld, K = a.K;
    const int r0 = (int)((long long)cta * n / G), r1 = (int)((long long)(cta + 1) * n / G);
    const int rl = r1 - r0;
    const int RLp = lanczos_rlp(a.rows_max);
    const int cpr = lanczos_cpr(n);            // 64-double chunks per row
    const int Kp2 = K + 2;
    LanczosSmem sm = lanczos_carve(smem_raw, K, a.rows_max, n, a.jac_inplace);
    JacobiScratch js = jacobi_carve(sm.jscratch, lanczos_kp(K));
    double* slab = sm.slabA;       // current basis slab: slab[q * RLp + r]
    double* slab_alt = sm.slabB;
    unsigned int epoch = a.epoch_base;

    // this warp's share of the slab symv: chunks [g0, g1) of the rl x cpr chunk grid (row-major)
    const int nchunks = rl * cpr;
    const int g0 = (int)((long long)warp * nchunks / LZ_NW), g1 = (int)((long long)(warp + 1) * nchunks / LZ_NW);
    const int wrow0 = g0 / cpr;
    if (tid <= LZ_NW) sm.wgs[tid] = (int)((long long)tid * nchunks / LZ_NW);

    // ---- ||x0|| (every CTA redundantly; deterministic) ----
    double nrm = 0.0;
    for (int i = tid; i < n; i += LZ_THREADS) { double t = a.x0[i]; nrm += t * t; }
    nrm = block_sum(nrm, js.red);
    const double inv_beta0 = 1.0 / sqrt(nrm);

    for (int i = tid; i < K; i += LZ_THREADS) { sm.Hd[i] = 0.0; sm.He[i] = 0.0; sm.Harr[i] = 0.0; }
    for (int r = tid; r < rl; r += LZ_THREADS) slab[r] = a.x0[r0 + r] * inv_beta0;          // v_0 slab
    for (int c = tid; c < cpr * 64; c += LZ_THREADS) sm.vbuf[c] = (c < n) ? a.x0[c] * inv_beta0 : 0.0;
    __syncthreads();

    int howmany = a.nev;
    int k = 1;              // number of basis vectors currently in the slab
    int arrow_at = -1;      // index of the column bordered by the arrow row (after a restart), else -1
    int arrow_len = 0;
    int numops = 0, numiter = 1, converged = 0;
    double beta = 0.0;
    int finished = 0;
    int kfin = 1;

    long long tprev = clock64();
#define LZ_STAMP(slot) do { if (a.prof && tid == 0 && numops == 10) a.prof[8 + cta * 8 + slot] = gtimer(); } while (0)
#define LZ_TICK(slot) do { if (a.prof && cta == 0 && tid == 0) { long long tn = clock64(); a.prof[slot] += tn - tprev; tprev = tn; } } while (0)
    while (!finished) {
        const int j = k - 1;   // newest basis vector index
        LZ_TICK(7);
        LZ_STAMP(0);
        // ================= symv: wloc = X[rows, :] * v_j (v_j is in vbuf) =================
        {
            int g = g0;
            while (g < g1) {
                const int row = g / cpr;
                const int gend = min(g1, (row + 1) * cpr);
                const double* xr = a.X + (size_t)(r0 + row) * ld + 2 * lane;
                const double* vb = sm.vbuf + 2 * lane;
                int cc = (g - row * cpr) * 64;
                const int ce = (gend - row * cpr) * 64;
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
                for (; cc + 448 < ce; cc += 512) {
                    double2 x0v = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc));
                    double2 x1v = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc + 64));
                    double2 x2v = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc + 128));
                    double2 x3v = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc + 192));
                    double2 x4v = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc + 256));
                    double2 x5v = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc + 320));
                    double2 x6v = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc + 384));
                    double2 x7v = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc + 448));
                    double2 v0 = *reinterpret_cast<const double2*>(vb + cc);
                    double2 v1 = *reinterpret_cast<const double2*>(vb + cc + 64);
                    double2 v2 = *reinterpret_cast<const double2*>(vb + cc + 128);
                    double2 v3 = *reinterpret_cast<const double2*>(vb + cc + 192);
                    double2 v4 = *reinterpret_cast<const double2*>(vb + cc + 256);
                    double2 v5 = *reinterpret_cast<const double2*>(vb + cc + 320);
                    double2 v6 = *reinterpret_cast<const double2*>(vb + cc + 384);
                    double2 v7 = *reinterpret_cast<const double2*>(vb + cc + 448);
                    acc0 = fma(x0v.x, v0.x, acc0); acc0 = fma(x0v.y, v0.y, acc0);
                    acc1 = fma(x1v.x, v1.x, acc1); acc1 = fma(x1v.y, v1.y, acc1);
                    acc2 = fma(x2v.x, v2.x, acc2); acc2 = fma(x2v.y, v2.y, acc2);
                    acc3 = fma(x3v.x, v3.x, acc3); acc3 = fma(x3v.y, v3.y, acc3);
                    acc0 = fma(x4v.x, v4.x, acc0); acc0 = fma(x4v.y, v4.y, acc0);
                    acc1 = fma(x5v.x, v5.x, acc1); acc1 = fma(x5v.y, v5.y, acc1);
                    acc2 = fma(x6v.x, v6.x, acc2); acc2 = fma(x6v.y, v6.y, acc2);
                    acc3 = fma(x7v.x, v7.x, acc3); acc3 = fma(x7v.y, v7.y, acc3);
                }
                for (; cc + 192 < ce; cc += 256) {
                    double2 x0v = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc));
                    double2 x1v = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc + 64));
                    double2 x2v = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc + 128));
                    double2 x3v = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc + 192));
                    double2 v0 = *reinterpret_cast<const double2*>(vb + cc);
                    double2 v1 = *reinterpret_cast<const double2*>(vb + cc + 64);
                    double2 v2 = *reinterpret_cast<const double2*>(vb + cc + 128);
                    double2 v3 = *reinterpret_cast<const double2*>(vb + cc + 192);
                    acc0 = fma(x0v.x, v0.x, acc0); acc0 = fma(x0v.y, v0.y, acc0);
                    acc1 = fma(x1v.x, v1.x, acc1); acc1 = fma(x1v.y, v1.y, acc1);
                    acc2 = fma(x2v.x, v2.x, acc2); acc2 = fma(x2v.y, v2.y, acc2);
                    acc3 = fma(x3v.x, v3.x, acc3); acc3 = fma(x3v.y, v3.y, acc3);
                }
                for (; cc < ce; cc += 64) {
                    double2 xv = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc));
                    double2 vv = *reinterpret_cast<const double2*>(vb + cc);
                    acc0 = fma(xv.x, vv.x, acc0); acc0 = fma(xv.y, vv.y, acc0);
                }
                double acc = warp_sum((acc0 + acc1) + (acc2 + acc3));
                if (lane == 0) sm.wpart[warp * LZ_TMAX + (row - wrow0)] = acc;
                g = gend;
            }
        }
        LZ_STAMP(1);
        __syncthreads();
        LZ_STAMP(2);
        LZ_TICK(0);
        // fold the per-warp row partials in warp order (deterministic)
        for (int r = tid; r < rl; r += LZ_THREADS) {
            const int ga = r * cpr, gb = ga + cpr;      // chunk range of row r
            double s = 0.0;
            for (int w = 0; w < LZ_NW; ++w) {
                const int wg0 = sm.wgs[w], wg1 = sm.wgs[w + 1];
                if (wg0 < gb && wg1 > ga && wg1 > wg0) s += sm.wpart[w * LZ_TMAX + (r - wg0 / cpr)];
            }
            sm.wloc[r] = s;
        }
        __syncthreads();
        LZ_STAMP(3);
        numops++;
        LZ_TICK(1);

        // ================= CGS pass 1: partial dots -> exchange -> h =================
        ++epoch;
        {
            uint4* xb = a.xbuf + (size_t)(epoch & 1u) * Kp2 * G;
            // quad per dot: q = tid / 4 in [0, j+1]; q == j+1 is ||w||^2
            for (int qb = warp * 32; qb < 4 * (j + 2); qb += LZ_THREADS) {     // warp-uniform trip count
                const int q4 = qb + lane;
                const int q = q4 >> 2, sub = q4 & 3;
                double s = 0.0;
                if (q <= j) { for (int r = sub; r < rl; r += 4) s = fma(slab[q * RLp + r], sm.wloc[r], s); }
                else if (q == j + 1) { for (int r = sub; r < rl; r += 4) s = fma(sm.wloc[r], sm.wloc[r], s); }
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (sub == 0 && q <= j + 1) ll_store(xb + (size_t)q * G + cta, s, epoch);
            }
            for (int q = warp; q <= j + 1; q += LZ_NW) {
                double s = ll_reduce_slot(xb + (size_t)q * G, G, epoch, lane);
                if (lane == 0) sm.hred[q] = s;
            }
        }
        __syncthreads();
        if (a.prof && tid == 0 && numops == 11) a.prof[8 + cta * 8 + 4] = gtimer();
        LZ_TICK(2);
        double alpha = sm.hred[j];
        // w' = w - V h   (quad per row)
        for (int rb = warp * 32; rb < 4 * rl; rb += LZ_THREADS) {              // warp-uniform trip count
            const int r4 = rb + lane;
            const int r = r4 >> 2, sub = r4 & 3;
            double s = 0.0;
            if (r < rl) { for (int q = sub; q <= j; q += 4) s = fma(sm.hred[q], slab[q * RLp + r], s); }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (sub == 0 && r < rl) sm.wloc[r] -= s;
        }
        __syncthreads();
        // ================= CGS pass 2 =================
        ++epoch;
        {
            uint4* xb = a.xbuf + (size_t)(epoch & 1u) * Kp2 * G;
            for (int qb = warp * 32; qb < 4 * (j + 2); qb += LZ_THREADS) {     // warp-uniform trip count
                const int q4 = qb + lane;
                const int q = q4 >> 2, sub = q4 & 3;
                double s = 0.0;
                if (q <= j) { for (int r = sub; r < rl; r += 4) s = fma(slab[q * RLp + r], sm.wloc[r], s); }
                else if (q == j + 1) { for (int r = sub; r < rl; r += 4) s = fma(sm.wloc[r], sm.wloc[r], s); }
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (sub == 0 && q <= j + 1) ll_store(xb + (size_t)q * G + cta, s, epoch);
            }
            for (int q = warp; q <= j + 1; q += LZ_NW) {
                double s = ll_reduce_slot(xb + (size_t)q * G, G, epoch, lane);
                if (lane == 0) sm.hred[q] = s;
            }
        }
        __syncthreads();
        if (a.prof && tid == 0 && numops == 11) a.prof[8 + cta * 8 + 5] = gtimer();
        LZ_TICK(3);
        alpha += sm.hred[j];
        const double wn2 = sm.hred[j + 1];
        double h2n2 = 0.0;
        for (int q = 0; q <= j; ++q) h2n2 = fma(sm.hred[q], sm.hred[q], h2n2);
        for (int rb = warp * 32; rb < 4 * rl; rb += LZ_THREADS) {              // warp-uniform trip count
            const int r4 = rb + lane;
            const int r = r4 >> 2, sub = r4 & 3;
            double s = 0.0;
            if (r < rl) { for (int q = sub; q <= j; q += 4) s = fma(sm.hred[q], slab[q * RLp + r], s); }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (sub == 0 && r < rl) sm.wloc[r] -= s;
        }
        __syncthreads();
        double beta2 = wn2 - h2n2;
        if (!(h2n2 <= 1e-4 * wn2)) {
            // the second pass removed a visible fraction of w: recompute ||w|| exactly (grid-uniform branch)
            ++epoch;
            uint4* xb = a.xbuf + (size_t)(epoch & 1u) * Kp2 * G;
            double s = 0.0;
            for (int r = tid; r < rl; r += LZ_THREADS) s += sm.wloc[r] * sm.wloc[r];
            s = block_sum(s, js.red);
            if (tid == 0) ll_store(xb + cta, s, epoch);
            if (warp == 0) {
                double t = ll_reduce_slot(xb, G, epoch, lane);
                if (lane == 0) sm.hred[j + 1] = t;
            }
            __syncthreads();
            beta2 = sm.hred[j + 1];
            __syncthreads();
        }
        beta = sqrt(fmax(beta2, 0.0));
        LZ_TICK(4);
        if (tid == 0) {
            sm.Hd[j] = alpha;
            sm.He[j] = beta;        // couples j and j+1
        }
        // ================= publish v_{j+1} = w / beta and stage it for the next mat-vec =================
        // (skipped on breakdown and after the last vector of a Krylov cycle when the solve is over — decided below —
        //  but the exchange itself is unconditional so that epochs stay grid-uniform)
        ++epoch;
        {
            uint4* vx = a.vx + (size_t)(epoch & 1u) * ld;
            const double ib = (beta > 0.0) ? 1.0 / beta : 0.0;
            for (int r = tid; r < rl; r += LZ_THREADS) {
                const double v = sm.wloc[r] * ib;
                slab[k * RLp + r] = v;          // slab has K+1 columns and k <= K
                ll_store(vx + r0 + r, v, epoch);
            }
            for (int c0 = tid; c0 < n; c0 += 4 * LZ_THREADS) {
                const int i0 = c0, i1 = c0 + LZ_THREADS, i2 = c0 + 2 * LZ_THREADS, i3 = c0 + 3 * LZ_THREADS;
                uint4 q0, q1, q2, q3;
                bool ok;
                do {
                    ok = true;
                    q0 = ll_peek(vx + i0);
                    if (i1 < n) q1 = ll_peek(vx + i1);
                    if (i2 < n) q2 = ll_peek(vx + i2);
                    if (i3 < n) q3 = ll_peek(vx + i3);
                    ok = q0.y == epoch && q0.w == epoch;
                    if (i1 < n) ok = ok && q1.y == epoch && q1.w == epoch;
                    if (i2 < n) ok = ok && q2.y == epoch && q2.w == epoch;
                    if (i3 < n) ok = ok && q3.y == epoch && q3.w == epoch;
                } while (!ok);
                sm.vbuf[i0] = ll_value(q0);
                if (i1 < n) sm.vbuf[i1] = ll_value(q1);
                if (i2 < n) sm.vbuf[i2] = ll_value(q2);
                if (i3 < n) sm.vbuf[i3] = ll_value(q3);
            }
        }
        __syncthreads();
        if (a.prof && tid == 0 && numops == 11) a.prof[8 + cta * 8 + 6] = gtimer();
        LZ_TICK(5);

        // ================= Ritz analysis (redundant in every CTA) =================
        if (beta <= a.tol && k < howmany) howmany = k;
        if (k == K || beta <= a.tol) {
            // dense Rayleigh quotient from its compact form, padded to an even dimension m
            const int lda = lanczos_kp(K);
            const int m = (k + 1) & ~1;
            for (int idx = tid; idx < m * m; idx += LZ_THREADS) {
                int r = idx % m, c = idx / m;
                double v = 0.0;
                if (r < k && c < k) {
                    if (r == c) v = sm.Hd[r];
                    else {
                        int lo = min(r, c), hi = max(r, c);
                        if (hi == arrow_at && lo < arrow_len) v = sm.Harr[lo];
                        else if (hi == lo + 1 && !(lo < arrow_len && hi <= arrow_at)) v = sm.He[lo];
                    }
                }
                sm.JA[r + c * lda] = v;
            }
            __syncthreads();
            const double* Jd = sm.JA;
            if (a.jac_inplace) jacobi_eigh_smem(m, sm.JA, lda, sm.JU, lda, js);
            else Jd = jacobi_eigh_smem_fast(m, sm.JA, sm.JB, lda, sm.JU, lda, js);
            __syncthreads();
            rank_sort_desc(k, Jd, lda, sm.order);
            __syncthreads();
            for (int i = tid; i < k; i += LZ_THREADS) {
                int o = sm.order[i];
                sm.D[i] = Jd[o + o * lda];
                sm.f[i] = beta * sm.JU[(k - 1) + o * lda];
            }
            __syncthreads();
            converged = 0;
            while (converged < k && fabs(sm.f[converged]) <= a.tol) converged++;
            kfin = k;
            if (converged >= howmany) {
                finished = 1;
            } else if (k == K) {
                if (numiter == a.maxiter) {
                    finished = 1;
                } else {
                    // ---- thick restart ----
                    const int keep = (3 * K + 2 * converged) / 5;
                    for (int idx = tid; idx < keep * rl; idx += LZ_THREADS) {
                        int q = idx / rl, r = idx - q * rl;
                        int o = sm.order[q];
                        double s = 0.0;
                        for (int i = 0; i < K; ++i) s = fma(slab[i * RLp + r], sm.JU[i + o * lda], s);
                        slab_alt[q * RLp + r] = s;
                    }
                    for (int r = tid; r < rl; r += LZ_THREADS) slab_alt[keep * RLp + r] = slab[K * RLp + r];
                    __syncthreads();
                    for (int i = tid; i < K; i += LZ_THREADS) {
                        double d = (i < keep) ? sm.D[i] : 0.0;
                        double fa = (i < keep) ? sm.f[i] : 0.0;
                        sm.Hd[i] = d; sm.Harr[i] = fa; sm.He[i] = 0.0;
                    }
                    __syncthreads();
                    double* t = slab; slab = slab_alt; slab_alt = t;
                    arrow_at = keep; arrow_len = keep;
                    k = keep + 1;
                    numiter++;
                    continue;
                }
            }
        }
        LZ_TICK(6);
        if (!finished) k++;
    }

    // ================= outputs =================
    k = kfin;
    int nvals = howmany > converged ? howmany : converged;
    if (nvals > k) nvals = k;
    {
        const int lda = lanczos_kp(K);
        for (int idx = tid; idx < nvals * rl; idx += LZ_THREADS) {
            int q = idx / rl, r = idx - q * rl;
            int o = sm.order[q];
            double s = 0.0;
            for (int i = 0; i < k; ++i) s = fma(slab[i * RLp + r], sm.JU[i + o * lda], s);
            a.Y[(size_t)q * ld + r0 + r] = s;
        }
    }
    if (cta == 0) {
        for (int i = tid; i < nvals; i += LZ_THREADS) a.vals[i] = sm.D[i];
        if (tid == 0) {
            a.info[0] = nvals; a.info[1] = converged; a.info[2] = numops; a.info[3] = numiter;
            a.scal[S_NUMOPS] += (double)numops;
            a.scal[S_HEADER + 3 * a.cone + 2] = (double)converged;
            if (converged == 0) a.scal[S_POISON] = 1.0;
        }
    }
}

// ---------------------------------------------------------------------------
// post-processing of one eigsolve (prox_operators.jl:91-107): which eigenpairs are kept
//   kept = { r < min(nev, converged) : lambda_r > 0 },  current_rank = |kept|,
//   min_eig = minimum over ALL returned values (length max(nev', converged)).
// ---------------------------------------------------------------------------
__global__ void k_lanczos_select(const double* __restrict__ vals, const int* __restrict__ info, int nev,
                                 int* __restrict__ kept_idx, double* __restrict__ kept_lam,
                                 int* __restrict__ nkept, double* __restrict__ scal, int cone) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (scal[S_POISON] != 0.0) { *nkept = 0; return; }
    const int nvals = info[0], conv = info[1];
    const int lim = min(nev, conv);
    double mn = vals[0];
    for (int i = 1; i < nvals; ++i) mn = fmin(mn, vals[i]);
    int cnt = 0;
    for (int i = 0; i < lim; ++i)
        if (vals[i] > 0.0) { kept_idx[cnt] = i; kept_lam[cnt] = vals[i]; cnt++; }
    *nkept = cnt;
    scal[S_HEADER + 3 * cone + 0] = (double)cnt;
    scal[S_HEADER + 3 * cone + 1] = mn;
}

// ---------------------------------------------------------------------------
// K7+K8 fused: x_out[k(i,j)] = s_ij * sum_q lam[q] Y[i, idx[q]] Y[j, idx[q]]
// s_ij = 1 on the diagonal, sqrt(2) off it (prox_operators.jl:17-31).  Written once,
// directly in svec form — the reference's r+1 read-modify-write passes over the dense
// n x n matrix (fill! + one dgemm per kept pair, prox_operators.jl:92,104) never happen.
// grid = tile pairs (bi <= bj) of 32x32, block = (32, 8).
// ---------------------------------------------------------------------------
constexpr int RC_CH = 32;   // kept eigenpairs staged per pass

__global__ void __launch_bounds__(256)
k_reconstruct_svec(const double* __restrict__ Y, int ld, int n, const int* __restrict__ kept_idx,
                   const double* __restrict__ kept_lam, const int* __restrict__ nkept_ptr,
                   double* __restrict__ x_out, const double* __restrict__ poison) {
    __shared__ double Yi[32][RC_CH + 1];
    __shared__ double Yj[32][RC_CH + 1];
    if (poison && *poison != 0.0) return;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    int t = blockIdx.x;
    int bj = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((bj + 1) * (bj + 2) / 2 <= t) ++bj;
    while (bj * (bj + 1) / 2 > t) --bj;
    int bi = t - bj * (bj + 1) / 2;
    const int nkept = *nkept_ptr;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};   // this thread: i = bi*32+tx, j = bj*32 + ty + 8*u
    for (int base = 0; base < nkept; base += RC_CH) {
        const int nk = min(RC_CH, nkept - base);
        __syncthreads();
        for (int idx = tid; idx < 32 * nk; idx += 256) {
            int q = idx / 32, r = idx - q * 32;
            int gi = bi * 32 + r, gj = bj * 32 + r;
            int col = kept_idx[base + q];
            double l = kept_lam[base + q];
            Yi[r][q] = (gi < n) ? Y[(size_t)col * ld + gi] * l : 0.0;
            Yj[r][q] = (gj < n) ? Y[(size_t)col * ld + gj] : 0.0;
        }
        __syncthreads();
        for (int q = 0; q < nk; ++q) {
            double yi = Yi[tx][q];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fma(yi, Yj[ty + 8 * u][q], acc[u]);
        }
    }
    const double sqrt2 = 1.41421356237309504880;
    int i = bi * 32 + tx;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        int j = bj * 32 + ty + 8 * u;
        if (i < n && j < n && i <= j) {
            size_t k = (size_t)j * (size_t)(j + 1) / 2 + (size_t)i;
            x_out[k] = (i != j) ? acc[u] * sqrt2 : acc[u];
        }
    }
}

}  // namespace pb
