// lanczos_cl.cuh — cluster-replicated thick-restart Lanczos: ONE grid-wide synchronisation per mat-vec.
//
// Same contract as k_lanczos (lanczos.cuh): KrylovKit.eigsolve(A, resid, nev, :LR, Lanczos(orth, K, maxiter,
// tol)) as called from reference src/eigsolver.jl:802-812.  What differs is where the Krylov basis lives.
//
// Measured on B200 (scripts/xchg_bench.cu): any grid-wide exchange through L2 costs >= 1.7-2.2 us (two dies),
// so the three exchanges per step of the row-distributed kernel (CGS pass 1, pass 2, publish v) bound a step
// at ~3 x 2.5 us no matter how fast the symv is.  Here the grid is made of thread-block clusters of C CTAs and
// EVERY cluster keeps a full replica of the basis V in its distributed shared memory (CTA rank c owns rows
// [c n/C, (c+1) n/C) of every basis vector).  Per Lanczos step:
//     symv on this CTA's slab of X rows (all CTAs of the grid share the 8 n^2 bytes)  -> w slab to global
//     ONE grid barrier (counter in L2) ; every CTA loads the w entries of its V rows  (all-gather)
//     CGS pass 1 / pass 2 / normalisation inside the cluster: partial dots are pushed into the peers'
//     shared memory (DSMEM) and a hardware cluster barrier (~0.2 us) replaces each grid barrier;
//     the new Lanczos vector is pushed into every peer's staging buffer for the next symv.
// All clusters execute the same arithmetic on bitwise identical data, so every CTA of the grid takes the
// same control decisions with no broadcast.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "jacobi.cuh"
#include "kernels_vec.cuh"
#include "lanczos.cuh"
#include "ritz_bi.cuh"

namespace pb {
namespace cg = cooperative_groups;

constexpr int LZC_MAXC = 16;     // largest supported cluster size
constexpr int LZC_KMAX = 104;    // largest Krylov dimension (thick-restart scratch row lives in local memory)

struct LanczosClArgs {
    const double* X; int n, ld;
    const double* x0;
    double* Y;                 // out: Ritz vectors, ld x K
    double* wg;                // [2][ld] gathered mat-vec result (global)
    unsigned int* bar;         // grid barrier counter (zeroed by the host before the launch)
    const double* ritz_rd;     // optional warm start of the Ritz eigenproblem: [0] = k of the stored basis, then Kp*Kp
    double* ritz_wr;           // where this launch stores its Ritz basis for the next one (never the same buffer as ritz_rd)
    int nev, K, maxiter;
    double tol;
    int rows_max;              // ceil(n / grid): symv rows per CTA
    int vn_max;                // ceil(n / C): basis rows per CTA
    double* vals; int* info; double* scal; int cone;
    int use_bi;                // 1: try bisection + twisted vectors for the leading Ritz pairs before the dense Jacobi solve
    long long* prof;
};

struct LanczosClSmem {
    double* vbuf;    // cpr * 64  newest Lanczos vector (written by the cluster peers)
    double* Vs;      // (K+1) * VNp  basis rows owned by this CTA
    double* wv;      // VNp       w entries of my basis rows
    double* hpart;   // 2 * C * (K+2)  per-peer partial dots (written by the peers), double buffered
    double* hred;    // K + 2
    double* wloc;    // RLp
    double* wpart;   // LZ_NW * LZ_TMAX
    double* Hd; double* He; double* Harr; double* D; double* f;   // K each
    double* JA; double* JB; double* JU;                              // Kp * Kp each
    int* order;      // K
    int* wgs;        // LZ_NW + 1 chunk boundaries of the warps
    void* jscratch;
};

__host__ __device__ inline int lanczos_cl_vnp(int vn_max) { return vn_max | 1; }

__host__ __device__ inline size_t lanczos_cl_smem_bytes(int K, int rows_max, int vn_max, int n, int C) {
    size_t Kp = (size_t)lanczos_kp(K), VNp = (size_t)lanczos_cl_vnp(vn_max), RLp = (size_t)lanczos_rlp(rows_max);
    size_t d = (size_t)lanczos_cpr(n) * 64 + (size_t)(K + 1) * VNp + VNp + 2 * (size_t)C * (size_t)(K + 2) + (size_t)(K + 2) + RLp +
               (size_t)LZ_NW * LZ_TMAX + 5 * (size_t)K + 3 * Kp * Kp;
    return d * sizeof(double) + sizeof(int) * (size_t)(((K + 3) & ~3) + 32) + jacobi_scratch_bytes((int)Kp) + 64;
}

__device__ inline LanczosClSmem lanczos_cl_carve(unsigned char* base, int K, int rows_max, int vn_max, int n, int C) {
    LanczosClSmem s;
    size_t Kp = (size_t)lanczos_kp(K), VNp = (size_t)lanczos_cl_vnp(vn_max), RLp = (size_t)lanczos_rlp(rows_max);
    double* d = reinterpret_cast<double*>(base);
    s.vbuf = d; d += (size_t)lanczos_cpr(n) * 64;
    s.Vs = d; d += (size_t)(K + 1) * VNp;
    s.wv = d; d += VNp;
    s.hpart = d; d += 2 * (size_t)C * (size_t)(K + 2);
    s.hred = d; d += K + 2;
    s.wloc = d; d += RLp;
    s.wpart = d; d += LZ_NW * LZ_TMAX;
    s.Hd = d; d += K; s.He = d; d += K; s.Harr = d; d += K; s.D = d; d += K; s.f = d; d += K;
    s.JA = d; d += Kp * Kp; s.JB = d; d += Kp * Kp; s.JU = d; d += Kp * Kp;
    s.order = reinterpret_cast<int*>(d);
    s.wgs = s.order + ((K + 3) & ~3);
    s.jscratch = reinterpret_cast<void*>(s.wgs + 32);
    return s;
}

// grid barrier: one release-increment per CTA, one acquire-poll by thread 0, bounded spin
__device__ __forceinline__ bool grid_arrive_wait(unsigned int* bar, unsigned int target) {
    __syncthreads();
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned int v;
        long long t0 = clock64();
        int ok = 1;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (v >= target) break;
            if (clock64() - t0 > 4000000000LL) { ok = 0; break; }      // ~2 s: a peer died; give up instead of hanging
        } while (true);
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}

__global__ void __launch_bounds__(LZ_THREADS, 1) k_lanczos_cl(LanczosClArgs a) {
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
    LanczosClSmem sm = lanczos_cl_carve(smem_raw, K, a.rows_max, a.vn_max, n, C);
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
#define LZC_TICK(slot) do { if (a.prof && cta == 0 && tid == 0) { long long tn = clock64(); a.prof[slot] += tn - tprev; tprev = tn; } } while (0)

    while (!finished) {
        const int j = k - 1;
        LZC_TICK(7);
        // ================= symv on my slab of rows: wloc = X[r0:r1, :] v_j =================
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
                    double2 w0 = *reinterpret_cast<const double2*>(vb + cc);
                    double2 w1 = *reinterpret_cast<const double2*>(vb + cc + 64);
                    double2 w2 = *reinterpret_cast<const double2*>(vb + cc + 128);
                    double2 w3 = *reinterpret_cast<const double2*>(vb + cc + 192);
                    double2 w4 = *reinterpret_cast<const double2*>(vb + cc + 256);
                    double2 w5 = *reinterpret_cast<const double2*>(vb + cc + 320);
                    double2 w6 = *reinterpret_cast<const double2*>(vb + cc + 384);
                    double2 w7 = *reinterpret_cast<const double2*>(vb + cc + 448);
                    acc0 = fma(x0v.x, w0.x, acc0); acc0 = fma(x0v.y, w0.y, acc0);
                    acc1 = fma(x1v.x, w1.x, acc1); acc1 = fma(x1v.y, w1.y, acc1);
                    acc2 = fma(x2v.x, w2.x, acc2); acc2 = fma(x2v.y, w2.y, acc2);
                    acc3 = fma(x3v.x, w3.x, acc3); acc3 = fma(x3v.y, w3.y, acc3);
                    acc0 = fma(x4v.x, w4.x, acc0); acc0 = fma(x4v.y, w4.y, acc0);
                    acc1 = fma(x5v.x, w5.x, acc1); acc1 = fma(x5v.y, w5.y, acc1);
                    acc2 = fma(x6v.x, w6.x, acc2); acc2 = fma(x6v.y, w6.y, acc2);
                    acc3 = fma(x7v.x, w7.x, acc3); acc3 = fma(x7v.y, w7.y, acc3);
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
        __syncthreads();
        LZC_TICK(0);
        // fold the per-warp row partials in warp order, publish my w slab and arrive at the grid barrier:
        // all done by warp 0 (rows_max <= 32 is checked on the host), so no block barrier sits between the
        // last store and the release
        ++gsync;
        if (warp == 0) {
            if (rl <= 32) {
                if (lane < rl) {
                    double s = 0.0;
                    if (fold_n > 0) s += sm.wpart[fold_s0];
                    if (fold_n > 1) s += sm.wpart[fold_s1];
                    if (fold_n > 2) s += sm.wpart[fold_s2];
                    __stcg(a.wg + (size_t)(numops & 1) * ld + r0 + lane, s);
                }
            } else {
                for (int r = lane; r < rl; r += 32) {       // large cones: generic fold
                    const int ga = r * cpr, gb = ga + cpr;
                    double s = 0.0;
                    for (int w = 0; w < LZ_NW; ++w) {
                        const int wa = sm.wgs[w], wb = sm.wgs[w + 1];
                        if (wa < gb && wb > ga && wb > wa) s += sm.wpart[w * LZ_TMAX + (r - wa / cpr)];
                    }
                    __stcg(a.wg + (size_t)(numops & 1) * ld + r0 + r, s);
                }
            }
            __syncwarp();
            if (lane == 0) {
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar) : "memory");
                const unsigned int target = gsync * (unsigned int)G;
                unsigned int v;
                long long t0 = clock64();
                int ok = 1;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.bar) : "memory");
                    if (v >= target) break;
                    if (clock64() - t0 > 4000000000LL) { ok = 0; break; }   // ~2 s: give up instead of hanging the GPU
                } while (true);
                sm.wgs[LZ_NW + 1] = ok;
            }
        }
        __syncthreads();
        if (!sm.wgs[LZ_NW + 1]) { failed = 1; break; }
        {
            const double* wgp = a.wg + (size_t)(numops & 1) * ld;
            for (int t = tid; t < vn; t += LZ_THREADS) sm.wv[t] = __ldcg(wgp + v0 + t);
        }
        __syncthreads();
        numops++;
        LZC_TICK(1);

        // ================= CGS2 inside the cluster =================
        double alpha = 0.0, wn2 = 0.0, h2n2 = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            // partial dots over my basis rows: half-warp per q (q == j+1: ||w||^2), two accumulators per lane
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
            LZC_TICK(2);
            cluster.sync();
            LZC_TICK(3);
            // h[q] = sum over the C peers in rank order; ||h||^2 on the fly (warps 0..: 32 q's per warp)
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
            if (pass == 0) alpha = sm.hred[j];
            else {
                alpha += sm.hred[j];
                wn2 = sm.hred[j + 1];
                h2n2 = 0.0;
                for (int w = 0; w * 32 <= j + 1; ++w) h2n2 += sm.wpart[w];
            }
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
            LZC_TICK(6);
        }
        double beta2 = wn2 - h2n2;
        if (!(h2n2 <= 1e-4 * wn2)) {
            // the second pass removed a visible fraction of w: recompute ||w|| exactly (cluster-uniform branch)
            double s = 0.0;
            for (int t = tid; t < vn; t += LZ_THREADS) s += sm.wv[t] * sm.wv[t];
            s = block_sum(s, js.red);
            if (tid < C) cluster.map_shared_rank(sm.hpart, tid)[(size_t)crank * Kp2] = s;          // pass-0 buffer, slot 0 (free again)
            cluster.sync();
            if (tid == 0) { double t = 0.0; for (int c = 0; c < C; ++c) t += sm.hpart[(size_t)c * Kp2]; sm.hred[j + 1] = t; }
            __syncthreads();
            beta2 = sm.hred[j + 1];
            cluster.sync();      // the slot is rewritten by the next step's pass 0
        }
        beta = sqrt(fmax(beta2, 0.0));
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
        LZC_TICK(4);

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
            if (done_bi) { LZC_TICK(5); continue; }
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
                    LZC_TICK(5);
                    continue;
                }
            }
        }
        LZC_TICK(5);
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
