// lanczos_cl3.cuh — third generation of the cluster-replicated thick-restart Lanczos kernel.
//
// Same contract as k_lanczos_cl (lanczos_cl.cuh): KrylovKit.eigsolve(A, resid, nev, :LR, Lanczos(orth, K,
// maxiter, tol)) as called from reference src/eigsolver.jl:802-812.  The basis still lives as one replica
// per thread-block cluster in distributed shared memory.  What changed, each item measured on B200
// (profiles/r1f_*):
//
//  1. Grid exchange without atomics.  120 CTAs incrementing ONE counter serialise in the L2 atomic unit
//     (~27 cycles each => 1.7 us per barrier).  Here every CTA release-stores its own epoch word and one warp
//     per CTA polls the whole flag array with four independent relaxed loads per lane: arrival costs one
//     store latency + one load round trip.  Epochs are monotone across launches, so no memset node either.
//  2. One re-orthogonalisation exchange per step instead of two.  KrylovKit's recurrence is the local
//     three-term step (w -= alpha v_j + beta_{j-1} v_{j-1}) followed by two Gram-Schmidt passes.  The local
//     step needs alpha = v_j . w, a grid-wide dot: its per-CTA partials ride the w all-gather that the grid
//     exchange performs anyway.  After the local step the components of w along the basis are O(eps ||X||)
//     (or the known arrow row right after a thick restart), so ONE classical Gram-Schmidt pass over the whole
//     basis restores orthogonality to machine precision; a second pass runs only when the first one removed a
//     visible part of w (||h||^2 > 1e-4 ||w||^2: breakdown / invariant subspace).  scripts/lz_variant_check.py
//     shows identical mat-vec counts, converged counts and 1e-15 orthogonality versus the two-pass schemes.
//  3. X stays on chip.  The 8 n^2 bytes of X are split over the CTAs' row slabs; the part of a slab that fits
//     the shared memory left over by the basis replica is kept there by the first mat-vec of the launch, so
//     the following ~24 mat-vecs read it at shared-memory instead of L2 bandwidth.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "jacobi.cuh"
#include "kernels_vec.cuh"
#include "lanczos.cuh"
#include "lanczos_cl.cuh"
#include "ritz_bi.cuh"

namespace pb {

constexpr int LZ3_GMAX = 256;    // largest grid the flag array / partial-alpha rows are sized for

struct LanczosCl3Args {
    const double* X; int n, ld;
    const double* x0;
    double* Y;                 // out: Ritz vectors, ld x K
    double* wg;                // [2][ld] gathered mat-vec result (global)
    double* apart;             // [2][LZ3_GMAX] per-CTA partial alpha = v_j[slab] . w[slab]
    unsigned int* flags;       // [LZ3_GMAX] per-CTA arrival epochs (monotone over launches, never reset)
    unsigned int epoch_base;   // the g-th exchange of this launch publishes epoch_base + g
    const double* ritz_rd;     // optional warm start of the Ritz eigenproblem (see LanczosClArgs)
    double* ritz_wr;
    int nev, K, maxiter;
    double tol;
    int rows_max;              // ceil(n / grid): symv rows per CTA
    int vn_max;                // ceil(n / C): basis rows per CTA
    int xres_chunks;           // 64-double chunks of X each WARP keeps in shared memory (0: none)
    double* vals; int* info; double* scal; int cone;
    int use_bi;
    long long* prof;
};

struct LanczosCl3Smem {
    double* xs;      // LZ_NW * xres_chunks * 64   resident part of my X slab, one contiguous run per warp
    LanczosClSmem b; // everything the second-generation kernel keeps (lanczos_cl.cuh)
};

__host__ __device__ inline size_t lanczos_cl3_smem_bytes(int K, int rows_max, int vn_max, int n, int C, int xres_chunks) {
    return (size_t)LZ_NW * (size_t)xres_chunks * 64 * sizeof(double) + lanczos_cl_smem_bytes(K, rows_max, vn_max, n, C);
}

// one row span of the slab symv.  MODE 0: X from global ; 1: X from global, copy kept in shared memory ;
// 2: X from shared memory.  xr / xsr / vb already carry the lane offset; cc, ce are column offsets (multiples of 64).
template <int MODE>
__device__ __forceinline__ void symv_span(const double* __restrict__ xr, double* xsr, const double* vb, int cc, const int ce,
                                          double& acc0, double& acc1, double& acc2, double& acc3) {
    for (; cc + 448 < ce; cc += 512) {
        double2 x[8], w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 2) x[u] = *reinterpret_cast<const double2*>(xsr + cc + 64 * u);
            else x[u] = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc + 64 * u));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = *reinterpret_cast<const double2*>(vb + cc + 64 * u);
        if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < 8; ++u) *reinterpret_cast<double2*>(xsr + cc + 64 * u) = x[u];
        }
        acc0 = fma(x[0].x, w[0].x, acc0); acc0 = fma(x[0].y, w[0].y, acc0);
        acc1 = fma(x[1].x, w[1].x, acc1); acc1 = fma(x[1].y, w[1].y, acc1);
        acc2 = fma(x[2].x, w[2].x, acc2); acc2 = fma(x[2].y, w[2].y, acc2);
        acc3 = fma(x[3].x, w[3].x, acc3); acc3 = fma(x[3].y, w[3].y, acc3);
        acc0 = fma(x[4].x, w[4].x, acc0); acc0 = fma(x[4].y, w[4].y, acc0);
        acc1 = fma(x[5].x, w[5].x, acc1); acc1 = fma(x[5].y, w[5].y, acc1);
        acc2 = fma(x[6].x, w[6].x, acc2); acc2 = fma(x[6].y, w[6].y, acc2);
        acc3 = fma(x[7].x, w[7].x, acc3); acc3 = fma(x[7].y, w[7].y, acc3);
    }
    for (; cc < ce; cc += 64) {
        double2 xv;
        if (MODE == 2) xv = *reinterpret_cast<const double2*>(xsr + cc);
        else xv = ld_stream_d2(reinterpret_cast<const double2*>(xr + cc));
        if (MODE == 1) *reinterpret_cast<double2*>(xsr + cc) = xv;
        const double2 vv = *reinterpret_cast<const double2*>(vb + cc);
        acc0 = fma(xv.x, vv.x, acc0); acc0 = fma(xv.y, vv.y, acc0);
    }
}

__global__ void __launch_bounds__(LZ_THREADS, 1) k_lanczos_cl3(LanczosCl3Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int n = a.n, ld = a.ld, K = a.K;
    // symv rows of this CTA (grid-wide split) and basis rows of this CTA (cluster-wide split)
    const int r0 = (int)((long long)cta * n / G), r1 = (int)((long long)(cta + 1) * n / G);
    const int rl = r1 - r0;
    const int v0 = (int)((long long)crank * n / C), v1 = (int)((long long)(crank + 1) * n / C);
    const int vn = v1 - v0;
    const int VNp = lanczos_cl_vnp(a.vn_max);
    const int cpr = lanczos_cpr(n);
    const int Kp2 = K + 2;
    double* const xs = reinterpret_cast<double*>(smem_raw);
    LanczosClSmem sm = lanczos_cl_carve(smem_raw + (size_t)LZ_NW * (size_t)a.xres_chunks * 64 * sizeof(double), K, a.rows_max,
                                        a.vn_max, n, C);
    JacobiScratch js = jacobi_carve(sm.jscratch, lanczos_kp(K));
    double* Vs = sm.Vs;

    // peers' views of my staging buffer (compile-time indexed: stays in registers)
    double* peer_vbuf[LZC_MAXC];
#pragma unroll
    for (int c = 0; c < LZC_MAXC; ++c) peer_vbuf[c] = (c < C) ? cluster.map_shared_rank(sm.vbuf, c) : nullptr;

    const int nchunks = rl * cpr;
    if (tid <= LZ_NW) sm.wgs[tid] = (int)((long long)tid * nchunks / LZ_NW);
    const int g0 = (int)((long long)warp * nchunks / LZ_NW), g1 = (int)((long long)(warp + 1) * nchunks / LZ_NW);
    const int wrow0 = g0 / cpr;
    const int gres = g0 + min(a.xres_chunks, g1 - g0);        // chunks [g0, gres) of this warp stay in shared memory
    double* const xsw = xs + (size_t)warp * (size_t)a.xres_chunks * 64;
    // which (warp, slot) partials make up row `lane` of my slab (at most 3 when a row straddles warps)
    int fold_n = 0, fold_s0 = 0, fold_s1 = 0, fold_s2 = 0;
    if (warp == 0 && lane < rl) {
        const int ga = lane * cpr, gb = ga + cpr;
        for (int w = 0; w < LZ_NW; ++w) {
            const int wa = (int)((long long)w * nchunks / LZ_NW), wb = (int)((long long)(w + 1) * nchunks / LZ_NW);
            if (wa < gb && wb > ga && wb > wa) {
                const int slot = w * LZ_TMAX + (lane - wa / cpr);
                if (fold_n == 0) fold_s0 = slot; else if (fold_n == 1) fold_s1 = slot; else fold_s2 = slot;
                ++fold_n;
            }
        }
    }

    double nrm = 0.0;
    for (int i = tid; i < n; i += LZ_THREADS) { double t = a.x0[i]; nrm += t * t; }
    nrm = block_sum(nrm, js.red);
    const double inv_beta0 = 1.0 / sqrt(nrm);
    for (int i = tid; i < K; i += LZ_THREADS) { sm.Hd[i] = 0.0; sm.He[i] = 0.0; sm.Harr[i] = 0.0; }
    for (int t = tid; t < vn; t += LZ_THREADS) Vs[t] = a.x0[v0 + t] * inv_beta0;
    for (int c = tid; c < cpr * 64; c += LZ_THREADS) sm.vbuf[c] = (c < n) ? a.x0[c] * inv_beta0 : 0.0;
    cluster.sync();       // everybody's shared memory is initialised before any peer writes into it

    int howmany = a.nev;
    int k = 1, arrow_at = -1, arrow_len = 0;
    int numops = 0, numiter = 1, converged = 0;
    double beta = 0.0;
    int finished = 0, failed = 0;
    unsigned int gsync = 0;
    bool first_analysis = true;

    long long tprev = clock64();
#define LZ3_TICK(slot) do { if (a.prof && cta == 0 && tid == 0) { long long tn = clock64(); a.prof[slot] += tn - tprev; tprev = tn; } } while (0)

    while (!finished) {
        const int j = k - 1;
        LZ3_TICK(7);
        // ================= symv on my slab of rows: wpart = X[r0:r1, :] v_j =================
        {
            int g = g0;
            while (g < g1) {
                const int row = g / cpr;
                const int gend = min(g1, (row + 1) * cpr);
                const double* xr = a.X + (size_t)(r0 + row) * ld + 2 * lane;
                const double* vb = sm.vbuf + 2 * lane;
                // the shared-memory copy of chunk gg sits at xsw + (gg - g0) * 64; as a function of the column offset
                double* xsr = xsw + ((long long)row * cpr - g0) * 64 + 2 * lane;
                const int c_lo = (g - row * cpr) * 64, c_hi = (gend - row * cpr) * 64;
                const int c_res = (min(gend, max(g, gres)) - row * cpr) * 64;      // [c_lo, c_res) resident, [c_res, c_hi) global
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
                if (numops == 0) symv_span<1>(xr, xsr, vb, c_lo, c_res, acc0, acc1, acc2, acc3);
                else symv_span<2>(xr, xsr, vb, c_lo, c_res, acc0, acc1, acc2, acc3);
                symv_span<0>(xr, xsr, vb, c_res, c_hi, acc0, acc1, acc2, acc3);
                double acc = warp_sum((acc0 + acc1) + (acc2 + acc3));
                if (lane == 0) sm.wpart[warp * LZ_TMAX + (row - wrow0)] = acc;
                g = gend;
            }
        }
        __syncthreads();
        LZ3_TICK(0);
        // ======== fold the per-warp row partials, publish my w slab + partial alpha, grid exchange (warp 0) ========
        ++gsync;
        if (warp == 0) {
            double* wgp = a.wg + (size_t)(numops & 1) * ld;
            double ap = 0.0;
            if (rl <= 32) {
                if (lane < rl) {
                    double s = 0.0;
                    if (fold_n > 0) s += sm.wpart[fold_s0];
                    if (fold_n > 1) s += sm.wpart[fold_s1];
                    if (fold_n > 2) s += sm.wpart[fold_s2];
                    __stcg(wgp + r0 + lane, s);
                    ap = s * sm.vbuf[r0 + lane];
                }
            } else {
                for (int r = lane; r < rl; r += 32) {       // large cones: generic fold
                    const int ga = r * cpr, gb = ga + cpr;
                    double s = 0.0;
                    for (int w = 0; w < LZ_NW; ++w) {
                        const int wa = sm.wgs[w], wb = sm.wgs[w + 1];
                        if (wa < gb && wb > ga && wb > wa) s += sm.wpart[w * LZ_TMAX + (r - wa / cpr)];
                    }
                    __stcg(wgp + r0 + r, s);
                    ap = fma(s, sm.vbuf[r0 + r], ap);
                }
            }
            ap = warp_sum(ap);
            if (lane == 0) __stcg(a.apart + (size_t)(numops & 1) * LZ3_GMAX + cta, ap);
            __syncwarp();
            const unsigned int target = a.epoch_base + gsync;
            if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.flags + cta), "r"(target) : "memory");
            int ok = 1;
            const long long t0 = clock64();
            while (true) {
                bool all = true;
                for (int c = lane; c < G; c += 32) {
                    unsigned int v;
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.flags + c) : "memory");
                    all = all && ((int)(v - target) >= 0);
                }
                if (__all_sync(0xffffffffu, all)) break;
                if (clock64() - t0 > 4000000000LL) ok = 0;             // ~2 s: a peer died; give up instead of hanging
                ok = __shfl_sync(0xffffffffu, ok, 0);
                if (!ok) break;
            }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            if (lane == 0) sm.wgs[LZ_NW + 1] = ok;
        }
        __syncthreads();
        if (!sm.wgs[LZ_NW + 1]) { failed = 1; break; }
        // ======== gather the w entries of my basis rows; alpha = sum of the per-CTA partials (fixed order) ========
        {
            const double* wgp = a.wg + (size_t)(numops & 1) * ld;
            for (int t = tid; t < vn; t += LZ_THREADS) sm.wv[t] = __ldcg(wgp + v0 + t);
            if (warp == LZ_NW - 1) {
                const double* app = a.apart + (size_t)(numops & 1) * LZ3_GMAX;
                double s = 0.0;
                for (int c = lane; c < G; c += 32) s += __ldcg(app + c);
                s = warp_sum(s);
                if (lane == 0) sm.hred[K + 1] = s;
            }
        }
        __syncthreads();
        numops++;
        double alpha = sm.hred[K + 1];
        // ======== local three-term step on my rows (arrow row right after a thick restart) ========
        {
            const double bprev = (j > 0 && j != arrow_at) ? sm.He[j - 1] : 0.0;
            for (int t = tid; t < vn; t += LZ_THREADS) {
                double w = sm.wv[t];
                w = fma(-alpha, Vs[(size_t)j * VNp + t], w);
                if (j == arrow_at) {
                    for (int i = 0; i < arrow_len; ++i) w = fma(-sm.Harr[i], Vs[(size_t)i * VNp + t], w);
                } else if (j > 0) {
                    w = fma(-bprev, Vs[(size_t)(j - 1) * VNp + t], w);
                }
                sm.wv[t] = w;
            }
        }
        __syncthreads();
        LZ3_TICK(1);

        // ================= one Gram-Schmidt pass inside the cluster (a second one only on breakdown) =================
        double wn2 = 0.0, hn2 = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            // partial dots over my basis rows: half-warp per q (q == j+1: ||w||^2), four accumulators per lane
            const int hw = tid >> 4, hl = tid & 15;
            for (int qb = 0; qb <= j + 1; qb += LZ_THREADS / 16) {
                const int q = qb + hw;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                if (q <= j + 1) {
                    const double* vq = (q <= j) ? Vs + (size_t)q * VNp : sm.wv;
                    int t = hl;
                    for (; t + 48 < vn; t += 64) {          // four independent chains: FP64 FMA latency is the bound here
                        s0 = fma(vq[t], sm.wv[t], s0); s1 = fma(vq[t + 16], sm.wv[t + 16], s1);
                        s2 = fma(vq[t + 32], sm.wv[t + 32], s2); s3 = fma(vq[t + 48], sm.wv[t + 48], s3);
                    }
                    for (; t < vn; t += 16) s0 = fma(vq[t], sm.wv[t], s0);
                }
                double s = (s0 + s1) + (s2 + s3);
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                if (q <= j + 1 && hl < C) cluster.map_shared_rank(sm.hpart, hl)[(size_t)pass * C * Kp2 + (size_t)crank * Kp2 + q] = s;   // lane c -> peer c
            }
            LZ3_TICK(2);
            cluster.sync();
            LZ3_TICK(3);
            // h[q] = sum over the C peers in rank order; ||h||^2 on the fly (32 q's per warp)
            for (int qb = warp * 32; qb <= j + 1; qb += LZ_THREADS) {
                const int q = qb + lane;
                double s = 0.0;
                if (q <= j + 1) {
                    const double* hp = sm.hpart + (size_t)pass * C * Kp2 + q;
                    if (C == 8) {        // fixed pairwise tree (3 dependent adds instead of 8)
                        s = ((hp[0] + hp[(size_t)Kp2]) + (hp[(size_t)2 * Kp2] + hp[(size_t)3 * Kp2])) +
                            ((hp[(size_t)4 * Kp2] + hp[(size_t)5 * Kp2]) + (hp[(size_t)6 * Kp2] + hp[(size_t)7 * Kp2]));
                    } else {
                        for (int c = 0; c < C; ++c) s += hp[(size_t)c * Kp2];
                    }
                    sm.hred[q] = s;
                }
                double sq = (q <= j) ? s * s : 0.0;
                sq = warp_sum(sq);
                if (lane == 0) sm.wpart[warp] = sq;         // wpart is free between symv phases
            }
            __syncthreads();
            alpha += sm.hred[j];
            wn2 = sm.hred[j + 1];
            hn2 = 0.0;
            for (int w = 0; w * 32 <= j + 1; ++w) hn2 += sm.wpart[w];
            // w <- w - V h on my rows (two threads per row when there are enough threads)
            if (2 * vn <= LZ_THREADS) {
                const int t = tid >> 1, sub = tid & 1;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                if (t < vn) {
                    int q = sub;
                    for (; q + 6 <= j; q += 8) {
                        s0 = fma(sm.hred[q], Vs[(size_t)q * VNp + t], s0);
                        s1 = fma(sm.hred[q + 2], Vs[(size_t)(q + 2) * VNp + t], s1);
                        s2 = fma(sm.hred[q + 4], Vs[(size_t)(q + 4) * VNp + t], s2);
                        s3 = fma(sm.hred[q + 6], Vs[(size_t)(q + 6) * VNp + t], s3);
                    }
                    for (; q <= j; q += 2) s0 = fma(sm.hred[q], Vs[(size_t)q * VNp + t], s0);
                }
                double sacc = (s0 + s1) + (s2 + s3);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
                if (t < vn && sub == 0) sm.wv[t] -= sacc;
            } else {
                for (int t = tid; t < vn; t += LZ_THREADS) {
                    double s0 = 0.0, s1 = 0.0;
                    int q = 0;
                    for (; q + 1 <= j; q += 2) {
                        s0 = fma(sm.hred[q], Vs[(size_t)q * VNp + t], s0);
                        s1 = fma(sm.hred[q + 1], Vs[(size_t)(q + 1) * VNp + t], s1);
                    }
                    if (q <= j) s0 = fma(sm.hred[q], Vs[(size_t)q * VNp + t], s0);
                    sm.wv[t] -= (s0 + s1);
                }
            }
            __syncthreads();
            LZ3_TICK(6);
            if (hn2 <= 1e-4 * wn2) break;      // cluster- and grid-uniform: every CTA holds bitwise identical h
        }
        // after the last pass executed: ||w_new||^2 = ||w||^2 - ||h||^2 (h is tiny unless w collapsed, and then beta <= tol anyway)
        beta = sqrt(fmax(wn2 - hn2, 0.0));
        if (tid == 0) { sm.Hd[j] = alpha; sm.He[j] = beta; }
        // ================= v_{j+1} = w / beta: keep my rows, push them into every peer's staging buffer =================
        {
            const double ib = (beta > 0.0) ? 1.0 / beta : 0.0;
            for (int t = tid; t < vn; t += LZ_THREADS) {
                const double v = sm.wv[t] * ib;
                Vs[(size_t)k * VNp + t] = v;
#pragma unroll
                for (int c = 0; c < LZC_MAXC; ++c) if (c < C) peer_vbuf[c][v0 + t] = v;
            }
        }
        cluster.sync();
        LZ3_TICK(4);

        // ================= Ritz analysis (redundant in every CTA) =================
        if (beta <= a.tol && k < howmany) howmany = k;
        if (k == K || beta <= a.tol) {
            const int lda = lanczos_kp(K);
            const int m = (k + 1) & ~1;
            // ---- fast path: leading pairs of the plain tridiagonal by bisection + twisted vectors ----
            bool done_bi = false;
            if (a.use_bi && arrow_at < 0 && 2 * (size_t)lda * lda >= ritz_bi_scratch_doubles(K)) {
                RitzBiScratch bs = ritz_bi_carve(sm.JA, K);          // JA and JB are contiguous and unused here
                const int mb = ritz_top_bi(k, sm.Hd, sm.He, howmany + 4, sm.D, sm.JU, lda, bs);
                if (mb > 0) {
                    for (int i = tid; i < mb; i += LZ_THREADS) { sm.order[i] = i; sm.f[i] = beta * sm.JU[(k - 1) + (size_t)i * lda]; }
                    __syncthreads();
                    int cv = 0;
                    while (cv < mb && fabs(sm.f[cv]) <= a.tol) cv++;
                    if (cv >= howmany && cv < mb) { converged = cv; finished = 1; done_bi = true; }
                    __syncthreads();
                }
            }
            if (a.prof && cta == 0 && tid == 0) { a.prof[8 + (done_bi ? 0 : 1)] += 1; a.prof[10] += clock64() - tprev; }
            if (done_bi) { LZ3_TICK(5); continue; }
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
            for (int idx = tid; idx < lda * lda; idx += LZ_THREADS) sm.JB[idx] = 0.0;
            __syncthreads();
            const double* Jd;
            const bool warm = first_analysis && a.ritz_rd && (int)a.ritz_rd[0] == k && k == K;
            if (warm) Jd = jacobi_eigh_smem_warm(m, k, sm.JA, sm.JB, lda, sm.JU, lda, a.ritz_rd + 1, js);
            else Jd = jacobi_eigh_smem_fast(m, sm.JA, sm.JB, lda, sm.JU, lda, js);
            __syncthreads();
            if (first_analysis && a.ritz_wr && cta == 0) {
                for (int idx = tid; idx < lda * lda; idx += LZ_THREADS) a.ritz_wr[1 + idx] = sm.JU[idx];
                if (tid == 0) a.ritz_wr[0] = (k == K) ? (double)k : -1.0;
            }
            first_analysis = false;
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
            if (converged >= howmany) {
                finished = 1;
            } else if (k == K) {
                if (numiter == a.maxiter) {
                    finished = 1;
                } else {
                    // ---- thick restart: V[:, 0:keep] <- V U[:, order[0:keep]], in place row by row ----
                    const int keep = (3 * K + 2 * converged) / 5;
                    for (int t = tid; t < vn; t += LZ_THREADS) {
                        double row[LZC_KMAX];
                        for (int i = 0; i < K; ++i) row[i] = Vs[(size_t)i * VNp + t];
                        for (int q = 0; q < keep; ++q) {
                            const double* u = sm.JU + (size_t)sm.order[q] * lda;
                            double s = 0.0;
                            for (int i = 0; i < K; ++i) s = fma(row[i], u[i], s);
                            Vs[(size_t)q * VNp + t] = s;
                        }
                        Vs[(size_t)keep * VNp + t] = Vs[(size_t)K * VNp + t];
                    }
                    __syncthreads();
                    for (int i = tid; i < K; i += LZ_THREADS) {
                        double d = (i < keep) ? sm.D[i] : 0.0;
                        double fa = (i < keep) ? sm.f[i] : 0.0;
                        sm.Hd[i] = d; sm.Harr[i] = fa; sm.He[i] = 0.0;
                    }
                    __syncthreads();
                    arrow_at = keep; arrow_len = keep;
                    k = keep + 1;
                    numiter++;
                    LZ3_TICK(5);
                    continue;
                }
            }
        }
        LZ3_TICK(5);
        if (!finished) k++;
    }

    // ================= outputs (cluster 0 holds a full replica) =================
    int nvals = howmany > converged ? howmany : converged;
    if (nvals > k) nvals = k;
    if (!failed && cta < C) {
        const int lda = lanczos_kp(K);
        for (int idx = tid; idx < nvals * vn; idx += LZ_THREADS) {
            int q = idx / vn, t = idx - q * vn;
            const double* u = sm.JU + (size_t)sm.order[q] * lda;
            double s = 0.0;
            for (int i = 0; i < k; ++i) s = fma(Vs[(size_t)i * VNp + t], u[i], s);
            a.Y[(size_t)q * ld + v0 + t] = s;
        }
    }
    if (cta == 0) {
        if (!failed) for (int i = tid; i < nvals; i += LZ_THREADS) a.vals[i] = sm.D[i];
        if (tid == 0) {
            a.info[0] = failed ? 0 : nvals; a.info[1] = failed ? 0 : converged; a.info[2] = numops; a.info[3] = numiter;
            a.scal[S_NUMOPS] += (double)numops;
            a.scal[S_HEADER + 3 * a.cone + 2] = failed ? 0.0 : (double)converged;
            if (failed || converged == 0) a.scal[S_POISON] = 1.0;
        }
    }
    cluster.sync();      // no CTA leaves while a peer may still address its shared memory
}

}  // namespace pb
