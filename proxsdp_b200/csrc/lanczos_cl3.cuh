// lanczos_cl3.cuh — third generation of the cluster-replicated thick-restart Lanczos kernel.
//
// Same contract as k_lanczos_cl (lanczos_cl.cuh): KrylovKit.eigsolve(A, resid, nev, :LR, Lanczos(orth, K,
// maxiter, tol)) as called from reference src/eigsolver.jl:802-812.  The basis still lives as one replica
// per thread-block cluster in distributed shared memory.  What changed, each item measured on B200
// (profiles/r1g_*):
//
//  1. One re-orthogonalisation exchange per step instead of two.  KrylovKit's recurrence is the local
//     three-term step (w -= alpha v_j + beta_{j-1} v_{j-1}) followed by two Gram-Schmidt passes.  The local
//     step needs alpha = v_j . w, a grid-wide dot: its per-CTA partials ride the w all-gather that the grid
//     exchange performs anyway.  After the local step the components of w along the basis are O(eps ||X||)
//     (or the known arrow row right after a thick restart), so ONE classical Gram-Schmidt pass over the whole
//     basis restores orthogonality to machine precision; a second pass runs only when the first one removed a
//     visible part of w (||h||^2 > 1e-4 ||w||^2: breakdown / invariant subspace).  scripts/lz_variant_check.py
//     shows identical mat-vec counts, converged counts and 1e-15 orthogonality versus the two-pass schemes.
//  2. No register spills on the per-step path.  Every acquire at gpu/cluster scope makes ptxas emit CCTL.IVALL
//     (L1 invalidate), so a spilled value re-read after a barrier costs a full L2 round trip; the second
//     generation kept ~25 shared-memory pointers live (50 registers) and spilled loop state, which showed up as
//     ~1.7 us of "fold" per step.  Here the shared-memory layout travels as integer offsets in the kernel
//     parameters (constant bank) and addresses are rebuilt where they are used.
//  3. Strip symv with X partly on chip.  The second-generation symv gave every warp a run of 64-double chunks
//     of the slab: ragged row tails made up to 13 dependent L2 round trips per mat-vec and the v operand was
//     re-read from shared memory for every row.  Here warp w owns a column strip of CPW chunks for ALL rows of
//     the slab: its piece of v sits in registers, the loads of a batch of RB rows are issued back to back,
//     the per-row partial sums of the 32 lanes are combined by a transposing butterfly (16 shuffles per 16
//     rows), and the first xres_rows rows of the slab are kept in the shared memory left over by the basis
//     replica, filled by the first mat-vec of the launch.
//  4. Grid exchange: relaxed polling + one acquire fence (an acquire load per poll costs a CCTL.IVALL each);
//     bar_mode 2 adds a two-level counter (group of C CTAs, then global) against the serialisation of 120
//     atomics on one address.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "jacobi.cuh"
#include "kernels_vec.cuh"
#include "lanczos.cuh"
#include "lanczos_cl.cuh"
#include "ritz_bi.cuh"

namespace pb {

constexpr int LZ3_GMAX = 256;    // largest grid the partial-alpha rows / group counters are sized for

// shared-memory layout, in doubles from the start of dynamic shared memory (host-computed, read from the constant bank)
struct Lz3Layout {
    int xs;      // xres_rows * cpr * 64   resident rows of my X slab (laid out like X)
    int rls;     // row stride of wrow (odd, > rows_max)
    int wrow;    // LZ_NW * rls            strip partials of the slab symv, [strip][row]
    int walpha;  // LZ_NW                  per-warp partial of alpha = v_j[slab] . (X v_j)[slab]
    int dpart;   // LZ_NW * (K+2)          per-warp partial Gram-Schmidt dots
    int vbuf;    // cpr * 64               newest Lanczos vector (written by the cluster peers)
    int Vs;      // (K+1) * VNp            basis rows owned by this CTA
    int wv;      // VNp                    w entries of my basis rows
    int hpart;   // 2 * C * (K+2)          per-peer partial dots (written by the peers), one buffer per pass
    int hred;    // K + 2                  reduced dots; [K+1] = alpha of the local step
    int red;     // 40                     block reduction scratch
    int Hd, He, Harr, D, f;   // K each
    int JA, JB, JU;           // Kp * Kp each (JA and JB contiguous: bisection scratch)
    int order;   // K ints
    int jscratch;
    int total;   // doubles
};

__host__ __device__ inline int lanczos_cl3_rls(int rows_max) { return (rows_max + 1) | 1; }

__host__ inline Lz3Layout lanczos_cl3_layout(int K, int rows_max, int vn_max, int n, int C, int xres_rows) {
    Lz3Layout L{};
    const int Kp = lanczos_kp(K), VNp = lanczos_cl_vnp(vn_max), cpr = lanczos_cpr(n);
    int d = 0;
    auto take = [&](int cnt) { int o = d; d += (cnt + 1) & ~1; return o; };     // keep everything 16-byte aligned
    L.xs = take(xres_rows * cpr * 64);
    L.rls = lanczos_cl3_rls(rows_max);
    L.wrow = take(L.rls * LZ_NW);
    L.walpha = take(LZ_NW);
    L.dpart = take(LZ_NW * (K + 2));
    L.vbuf = take(cpr * 64);
    L.Vs = take((K + 1) * VNp);
    L.wv = take(VNp);
    L.hpart = take(2 * C * (K + 2));
    L.hred = take(K + 2);
    L.red = take(40);
    L.Hd = take(K); L.He = take(K); L.Harr = take(K); L.D = take(K); L.f = take(K);
    L.JA = take(Kp * Kp); L.JB = take(Kp * Kp); L.JU = take(Kp * Kp);
    L.order = take((K + 3) / 2 + 2);
    L.jscratch = take((int)((jacobi_scratch_bytes(Kp) + 7) / 8));
    L.total = d;
    return L;
}

struct LanczosCl3Args {
    const double* X; int n, ld;
    const double* x0;
    double* Y;                 // out: Ritz vectors, ld x K
    double* wg;                // [2][ld] gathered mat-vec result (global)
    double* apart;             // [2][LZ3_GMAX] per-CTA partial alpha = v_j[slab] . w[slab]
    unsigned int* bar;         // [(1 + LZ3_GMAX / 2) * 32] counters, one per 128-byte line, zeroed per launch
    int bar_mode;              // 0: one counter ; 2: group counters + global counter
    int cpw;                   // 64-double chunks per symv strip (1..8), strips = ceil(cpr / cpw) <= LZ_NW
    int xres_rows;             // slab rows kept in shared memory
    const double* ritz_rd;     // optional warm start of the Ritz eigenproblem (see LanczosClArgs)
    double* ritz_wr;
    int nev, K, maxiter;
    double tol;
    int vn_max;                // ceil(n / C): basis rows per CTA
    int rbase, rrem;           // symv rows of CTA c: rbase + (c < rrem), starting at c * rbase + min(c, rrem)   (n = G * rbase + rrem)
    int vbase, vrem;           // basis rows of cluster rank c, same formula with C
    double* vals; int* info; double* scal; int cone;
    int use_bi;
    long long* prof;
    Lz3Layout L;
};

// sums NP per-lane values over the 32 lanes of a warp with a transposing butterfly: after the call the lanes
// whose low (5 - log2 NP) bits are zero hold the warp total of value number `sel` (returned per lane).
template <int NP>
__device__ __forceinline__ double warp_multi_sum(double (&a)[NP], int lane, int& sel) {
    int o = 16, row = 0;
#pragma unroll
    for (int cnt = NP; cnt > 1; cnt >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < cnt / 2; ++i) {
            const double send = up ? a[i] : a[i + cnt / 2];
            const double keep = up ? a[i + cnt / 2] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
        if (up) row += cnt / 2;
        o >>= 1;
    }
    double s = a[0];
#pragma unroll
    for (int oo = 32 / NP / 2; oo > 0; oo >>= 1) s += __shfl_xor_sync(0xffffffffu, s, oo);
    sel = row;
    return s;
}

// One batch of the strip symv: RB rows x CPW chunks of X (16 bytes per lane each) against the warp's piece of v.
// SRC 0: X from global ; 1: from global, copy kept in shared memory ; 2: from shared memory.  A full batch
// (nrows == RB, nch == CPW) is branch-free: all loads are issued back to back, then the FMAs.
template <int CPW, int RB, int NP, int SRC>
__device__ __forceinline__ void strip_batch(const double* __restrict__ gp, const size_t gstride, double* sp, const int sstride,
                                            const int nrows, const int nch, const double2 (&vr)[CPW], double (&acc)[NP]) {
    double2 x[RB][CPW];
    if (nrows == RB && nch == CPW) {
#pragma unroll
        for (int i = 0; i < RB; ++i) {
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                if (SRC == 2) x[i][c] = *reinterpret_cast<const double2*>(sp + i * sstride + c * 64);
                else x[i][c] = ld_stream_d2(reinterpret_cast<const double2*>(gp + i * gstride + c * 64));
            }
        }
        if (SRC == 1) {
#pragma unroll
            for (int i = 0; i < RB; ++i) {
#pragma unroll
                for (int c = 0; c < CPW; ++c) *reinterpret_cast<double2*>(sp + i * sstride + c * 64) = x[i][c];
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < RB; ++i) {
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                x[i][c] = make_double2(0.0, 0.0);
                if (i < nrows && c < nch) {
                    if (SRC == 2) x[i][c] = *reinterpret_cast<const double2*>(sp + i * sstride + c * 64);
                    else x[i][c] = ld_stream_d2(reinterpret_cast<const double2*>(gp + i * gstride + c * 64));
                    if (SRC == 1) *reinterpret_cast<double2*>(sp + i * sstride + c * 64) = x[i][c];
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double t = 0.0;
        if (i < RB) {
#pragma unroll
            for (int c = 0; c < CPW; ++c) { t = fma(x[i][c].x, vr[c].x, t); t = fma(x[i][c].y, vr[c].y, t); }
        }
        acc[i] = t;
    }
}

// strip symv: warp (strip s, row group rg) owns the columns of CPW chunks for the rows rg, rg + nrg, ... of the slab;
// the per-row partial dot products go to wrow[s * RLs + row].  The first xres_rows rows of the slab live in shared
// memory (xs, laid out like X: row * cpr chunks) once the first mat-vec of the launch has put them there.
// The warp's share of alpha = v[slab] . (X v)[slab] rides the last butterfly in the spare slot NP - 1 (RB < NP) and
// lands in walpha[warp]: no extra shuffle.
template <int CPW, int RB>
__device__ __forceinline__ void symv_strips(const double* __restrict__ X, const int ld, const int r0, const int rl, const int cpr,
                                            const double* vbuf, double* xs, const int xres_rows, const bool first, double* wrow,
                                            const int RLs, double* walpha, const int warp, const int lane) {
    constexpr int NP = RB < 2 ? 2 : RB < 4 ? 4 : RB < 8 ? 8 : 16;
    static_assert(RB < NP, "one butterfly slot must stay free for alpha");
    const int NS = (cpr + CPW - 1) / CPW;
    const int nrg = LZ_NW / NS;
    const int s = warp % NS, rg = warp / NS;
    const int n_my = (warp < NS * nrg && rl > rg) ? (rl - rg + nrg - 1) / nrg : 0;      // my rows: rg + i * nrg, i < n_my
    if (n_my == 0) { if (lane == 0) walpha[warp] = 0.0; return; }
    const int nch = min(CPW, cpr - s * CPW);
    double2 vr[CPW];
#pragma unroll
    for (int c = 0; c < CPW; ++c) vr[c] = (c < nch) ? *reinterpret_cast<const double2*>(vbuf + (s * CPW + c) * 64 + 2 * lane) : make_double2(0.0, 0.0);
    const int n_res = min(n_my, (xres_rows > rg) ? (xres_rows - rg + nrg - 1) / nrg : 0);   // the first n_res of my rows are resident
    const double* gp = X + (size_t)(r0 + rg) * ld + (size_t)s * CPW * 64 + 2 * lane;
    double* sp = xs + (rg * cpr + s * CPW) * 64 + 2 * lane;
    const size_t gstride = (size_t)nrg * ld;
    const int sstride = nrg * cpr * 64;
    double pa = 0.0;
    for (int ib = 0; ib < n_my; ) {
        double acc[NP];
        int nrows;
        if (ib < n_res) {
            nrows = min(RB, n_res - ib);
            if (first) strip_batch<CPW, RB, NP, 1>(gp + ib * gstride, gstride, sp + ib * sstride, sstride, nrows, nch, vr, acc);
            else strip_batch<CPW, RB, NP, 2>(gp + ib * gstride, gstride, sp + ib * sstride, sstride, nrows, nch, vr, acc);
        } else {
            nrows = min(RB, n_my - ib);
            strip_batch<CPW, RB, NP, 0>(gp + ib * gstride, gstride, sp, sstride, nrows, nch, vr, acc);
        }
        // alpha share of this batch: sum_i acc[i] * v[row_i] (three chains), kept per lane until the last batch
        {
            double p0 = 0.0, p1 = 0.0, p2 = 0.0;
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                const double vi = (i < nrows) ? vbuf[r0 + rg + (ib + i) * nrg] : 0.0;
                if (i % 3 == 0) p0 = fma(acc[i], vi, p0); else if (i % 3 == 1) p1 = fma(acc[i], vi, p1); else p2 = fma(acc[i], vi, p2);
            }
            pa += (p0 + p1) + p2;
        }
        const bool last = (ib + nrows >= n_my);
        if (last) acc[NP - 1] = pa;
        int sel;
        const double tot = warp_multi_sum<NP>(acc, lane, sel);
        if ((lane & (32 / NP - 1)) == 0) {
            if (sel < nrows) wrow[s * RLs + rg + (ib + sel) * nrg] = tot;
            else if (last && sel == NP - 1) walpha[warp] = tot;
        }
        ib += nrows;
    }
}

__global__ void __launch_bounds__(LZ_THREADS, 1) k_lanczos_cl3(const __grid_constant__ LanczosCl3Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int n = a.n, K = a.K;
    // symv rows of this CTA (grid-wide split) and basis rows of this CTA (cluster-wide split)
    // (cheap integer forms on purpose: ptxas rematerialises them instead of spilling them across the symv)
    const int r0 = cta * a.rbase + min(cta, a.rrem);
    const int rl = a.rbase + (cta < a.rrem ? 1 : 0);
    const int v0 = crank * a.vbase + min(crank, a.vrem);
    const int vn = a.vbase + (crank < a.vrem ? 1 : 0);
    const int VNp = lanczos_cl_vnp(a.vn_max);
    const int cpr = lanczos_cpr(n);
    const int Kp2 = K + 2;
    const int RLs = a.L.rls;
    // shared-memory arrays: rebuilt from the constant-bank offsets at the point of use (no long-lived pointers)
    double* const sbase = reinterpret_cast<double*>(smem_raw);
#define SMD(field) (sbase + a.L.field)
    // state that every thread reads but only the Ritz analysis changes lives in registers; all of it is uniform
    int howmany = a.nev;
    int k = 1, arrow = -1;          // arrow >= 0 after a thick restart: column `arrow` couples to the kept Ritz values 0..arrow-1
    int numiter = 1, converged = 0;
    double beta = 0.0;
    int finished = 0, failed = 0;
    unsigned int gsync = 0;         // grid exchanges so far == mat-vecs so far
    bool first_analysis = true;

    // the per-phase clocks of the profiled CTAs (0, C - 1, G / 2, G - 1) accumulate in shared memory (a global
    // read-modify-write per tick would stall warp 0, the warp on the critical path) and are flushed at the end
    __shared__ long long s_prof[32];
    __shared__ int s_ok;
    __shared__ int s_state[8];
    __shared__ double s_beta;
    const int prow = (cta == 0) ? 0 : (cta == C - 1) ? 1 : (cta == G / 2) ? 2 : (cta == G - 1) ? 3 : -1;
    const bool profiling = (a.prof != nullptr) && prow >= 0;
    if (tid < 32) s_prof[tid] = 0;
    // (the running clock sits in s_prof[31], not in a register that would have to live across the symv)
#define tprev s_prof[31]
#define LZ3_TICK(slot) do { if (profiling && tid == 0) { long long tn = clock64(); s_prof[slot] += tn - tprev; tprev = tn; } } while (0)

    {
        double nrm = 0.0;
        for (int i = tid; i < n; i += LZ_THREADS) { double t = a.x0[i]; nrm += t * t; }
        nrm = block_sum(nrm, SMD(red));
        const double inv_beta0 = 1.0 / sqrt(nrm);
        for (int i = tid; i < K; i += LZ_THREADS) { SMD(Hd)[i] = 0.0; SMD(He)[i] = 0.0; SMD(Harr)[i] = 0.0; }
        for (int t = tid; t < vn; t += LZ_THREADS) SMD(Vs)[t] = a.x0[v0 + t] * inv_beta0;
        for (int c = tid; c < cpr * 64; c += LZ_THREADS) SMD(vbuf)[c] = (c < n) ? a.x0[c] * inv_beta0 : 0.0;
    }
    cluster.sync();       // everybody's shared memory is initialised before any peer writes into it
    if (profiling && tid == 0) tprev = clock64();

    while (!finished) {
        LZ3_TICK(7);
        // The loop state is parked in shared memory across the symv (the register-hungry part of the step) and
        // re-read behind the barrier that follows it: a value kept live across the symv gets spilled, and a
        // spill re-read after a fence is an L2 round trip because the fence invalidates L1.
        if (tid == 0) {
            s_state[0] = k; s_state[1] = (int)gsync; s_state[2] = arrow; s_state[3] = howmany; s_state[4] = numiter;
            s_state[5] = first_analysis ? 1 : 0; s_beta = beta;
        }
        // ================= symv on my slab of rows: strip partials of X[r0:r1, :] v_j =================
        {
            const bool first = (gsync == 0);
            switch (a.cpw) {
#define LZ3_SYMV(CPW, RB) symv_strips<CPW, RB>(a.X, a.ld, r0, rl, cpr, SMD(vbuf), SMD(xs), a.xres_rows, first, SMD(wrow), RLs, SMD(walpha), warp, lane)
                case 1: LZ3_SYMV(1, 15); break;
                case 2: LZ3_SYMV(2, 9); break;
                case 3: LZ3_SYMV(3, 6); break;
                case 4: LZ3_SYMV(4, 3); break;
                case 5: LZ3_SYMV(5, 3); break;
                case 6: LZ3_SYMV(6, 3); break;
                default: LZ3_SYMV(8, 2); break;
#undef LZ3_SYMV
            }
        }
        __syncthreads();
        k = s_state[0]; gsync = (unsigned int)s_state[1]; arrow = s_state[2]; howmany = s_state[3]; numiter = s_state[4];
        first_analysis = s_state[5] != 0; beta = s_beta;
        const int j = k - 1;
        LZ3_TICK(0);
        // ======== fold the strip partials, publish my w slab + partial alpha, grid exchange (warp 0) ========
        // Every dependent step of a single warp costs 25-110 cycles on B200 (DFMA 23, LDS ~35, one 64-bit shuffle
        // step ~110: scripts/lat_bench.cu), so the phases below are laid out for short chains: conflict-free
        // shared-memory columns and serial adds in 4 chains instead of shuffle trees.
        ++gsync;
        if (warp == 0) {
            double* wgp = a.wg + (size_t)(gsync & 1) * a.ld + r0;
            for (int r = lane; r <= rl; r += 32) {           // lane r: row r of my slab; lane rl: my share of alpha
                const double* wr = (r < rl) ? SMD(wrow) + r : SMD(walpha);
                const int stride = (r < rl) ? RLs : 1;
                const int cnt = (r < rl) ? (cpr + a.cpw - 1) / a.cpw : LZ_NW;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int q = 0;
                for (; q + 3 < cnt; q += 4) { s0 += wr[q * stride]; s1 += wr[(q + 1) * stride]; s2 += wr[(q + 2) * stride]; s3 += wr[(q + 3) * stride]; }
                for (; q < cnt; ++q) s0 += wr[q * stride];
                const double sres = (s0 + s1) + (s2 + s3);
                if (r < rl) __stcg(wgp + r, sres);
                else __stcg(a.apart + (size_t)(gsync & 1) * LZ3_GMAX + cta, sres);
            }
            __syncwarp();
            long long t0 = 0;
            if (profiling && lane == 0) { t0 = clock64(); s_prof[15] += t0 - tprev; }      // fold + publish
            if (lane == 0) {
                int ok = 1;
                unsigned int target;
                if (a.bar_mode == 2) {
                    // group counter first; the last arriver of the group bumps the global counter
                    unsigned int old;
                    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(a.bar + (size_t)(1 + cta / C) * 32) : "memory");
                    if (old + 1 == gsync * (unsigned int)C) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar) : "memory");
                    target = gsync * (unsigned int)(G / C);
                } else {
                    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar) : "memory");
                    target = gsync * (unsigned int)G;
                }
                unsigned int v;
                const long long tw = clock64();
                do {
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.bar) : "memory");
                    if (v >= target) break;
                    if (clock64() - tw > 4000000000LL) { ok = 0; break; }   // ~2 s: give up instead of hanging the GPU
                } while (true);
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                s_ok = ok;
                if (profiling) s_prof[16] += clock64() - t0;      // arrive + wait
            }
        }
        __syncthreads();
        if (!s_ok) { failed = 1; break; }
        LZ3_TICK(17);
        // ======== gather + local three-term step + partial Gram-Schmidt dots, warp by warp ========
        // Warp w owns the basis rows [w RW, (w+1) RW) of this CTA for the whole phase, so no block barrier is needed
        // between gathering w, the local step and the dots; the only block-wide value is alpha, which warp LZ_NW-1
        // sums from the per-CTA partials.  alpha is only needed to about 1e-7: the Gram-Schmidt pass removes
        // whatever is left along v_j exactly (alpha_j = alpha~ + h_j), so its shuffle tree runs in FP32 (35 instead
        // of 110 cycles per step).
        const int RW = (a.vn_max + LZ_NW - 1) / LZ_NW;
        const int t_lo = min(vn, warp * RW), t_hi = min(vn, (warp + 1) * RW);
        {
            const double* wgp = a.wg + (size_t)(gsync & 1) * a.ld + v0;
            const double* Vs = SMD(Vs);
            double* wv = SMD(wv);
            double wreg[2];                                  // RW <= 64 rows per warp is checked on the host
#pragma unroll
            for (int u = 0; u < 2; ++u) { const int t = t_lo + lane + 32 * u; wreg[u] = (t < t_hi) ? __ldcg(wgp + t) : 0.0; }
            if (warp == LZ_NW - 1) {
                const double* app = a.apart + (size_t)(gsync & 1) * LZ3_GMAX;
                double sd = 0.0;
                for (int c = lane; c < G; c += 32) sd += __ldcg(app + c);
                float sf = (float)sd;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sf += __shfl_xor_sync(0xffffffffu, sf, o);
                if (lane == 0) SMD(hred)[K + 1] = (double)sf;
            }
            if (j == arrow) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int t = t_lo + lane + 32 * u;
                    if (t < t_hi) { double w = wreg[u]; for (int i = 0; i < arrow; ++i) w = fma(-SMD(Harr)[i], Vs[i * VNp + t], w); wreg[u] = w; }
                }
            } else if (j > 0) {
                const double bprev = SMD(He)[j - 1];
#pragma unroll
                for (int u = 0; u < 2; ++u) { const int t = t_lo + lane + 32 * u; if (t < t_hi) wreg[u] = fma(-bprev, Vs[(j - 1) * VNp + t], wreg[u]); }
            }
            __syncthreads();
            const double alpha0 = SMD(hred)[K + 1];
#pragma unroll
            for (int u = 0; u < 2; ++u) { const int t = t_lo + lane + 32 * u; if (t < t_hi) wv[t] = fma(-alpha0, Vs[j * VNp + t], wreg[u]); }
            __syncwarp();
        }
        double alpha = SMD(hred)[K + 1];
        LZ3_TICK(1);

        // ================= one Gram-Schmidt pass inside the cluster (a second one only on breakdown) =================
        double wn2 = 0.0, hn2 = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            // partial dots: lane <-> q (q == j+1: ||w||^2), warp <-> my RW rows; no cross-lane reduction at all
            {
                const double* Vs = SMD(Vs);
                const double* wv = SMD(wv);
                for (int qb = 0; qb <= j + 1; qb += 32) {
                    const int q = qb + lane;
                    if (q <= j + 1) {
                        const double* vq = (q <= j) ? Vs + q * VNp : wv;
                        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                        int t = t_lo;
                        for (; t + 3 < t_hi; t += 4) {
                            s0 = fma(vq[t], wv[t], s0); s1 = fma(vq[t + 1], wv[t + 1], s1);
                            s2 = fma(vq[t + 2], wv[t + 2], s2); s3 = fma(vq[t + 3], wv[t + 3], s3);
                        }
                        for (; t < t_hi; ++t) s0 = fma(vq[t], wv[t], s0);
                        SMD(dpart)[warp * Kp2 + q] = (s0 + s1) + (s2 + s3);
                    }
                }
            }
            __syncthreads();
            // my CTA's dots = sum over the warps (4 chains), pushed to every peer (thread q <-> dot q)
            for (int q = tid; q <= j + 1; q += LZ_THREADS) {
                const double* dp = SMD(dpart) + q;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int w = 0; w < LZ_NW; w += 4) { s0 += dp[w * Kp2]; s1 += dp[(w + 1) * Kp2]; s2 += dp[(w + 2) * Kp2]; s3 += dp[(w + 3) * Kp2]; }
                const double sres = (s0 + s1) + (s2 + s3);
                for (int c = 0; c < C; ++c) cluster.map_shared_rank(SMD(hpart), c)[pass * C * Kp2 + crank * Kp2 + q] = sres;
            }
            LZ3_TICK(2);
            cluster.sync();
            LZ3_TICK(3);
            // h[q] = sum over the C peers in rank order (fixed tree)
            for (int q = tid; q <= j + 1; q += LZ_THREADS) {
                const double* hp = SMD(hpart) + pass * C * Kp2 + q;
                double sres;
                if (C == 8) {
                    sres = ((hp[0] + hp[Kp2]) + (hp[2 * Kp2] + hp[3 * Kp2])) + ((hp[4 * Kp2] + hp[5 * Kp2]) + (hp[6 * Kp2] + hp[7 * Kp2]));
                } else {
                    sres = 0.0;
                    for (int c = 0; c < C; ++c) sres += hp[c * Kp2];
                }
                SMD(hred)[q] = sres;
            }
            __syncthreads();
            alpha += SMD(hred)[j];
            wn2 = SMD(hred)[j + 1];
            // w <- w - V h on my rows, ||h||^2 on the fly in every thread (two threads per row when there are enough)
            {
                const double* Vs = SMD(Vs);
                const double* hred = SMD(hred);
                if (2 * vn <= LZ_THREADS) {
                    const int t = min(tid >> 1, vn - 1), sub = tid & 1;
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
                    int q = sub;
                    for (; q + 6 <= j; q += 8) {
                        const double h0 = hred[q], h1 = hred[q + 2], h2 = hred[q + 4], h3 = hred[q + 6];
                        s0 = fma(h0, Vs[q * VNp + t], s0); g0 = fma(h0, h0, g0);
                        s1 = fma(h1, Vs[(q + 2) * VNp + t], s1); g1 = fma(h1, h1, g1);
                        s2 = fma(h2, Vs[(q + 4) * VNp + t], s2); g2 = fma(h2, h2, g2);
                        s3 = fma(h3, Vs[(q + 6) * VNp + t], s3); g3 = fma(h3, h3, g3);
                    }
                    for (; q <= j; q += 2) { const double h0 = hred[q]; s0 = fma(h0, Vs[q * VNp + t], s0); g0 = fma(h0, h0, g0); }
                    double sacc = (s0 + s1) + (s2 + s3), gacc = (g0 + g1) + (g2 + g3);
                    sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
                    // the two halves of ||h||^2 are added in a fixed order (even + odd) so every thread gets the same bits
                    const double gother = __shfl_xor_sync(0xffffffffu, gacc, 1);
                    hn2 = sub ? gother + gacc : gacc + gother;
                    if ((tid >> 1) < vn && sub == 0) SMD(wv)[t] -= sacc;
                } else {
                    double g0 = 0.0, g1 = 0.0;
                    for (int q = 0; q <= j; ++q) { const double h0 = hred[q]; if (q & 1) g1 = fma(h0, h0, g1); else g0 = fma(h0, h0, g0); }
                    hn2 = g0 + g1;
                    for (int t = tid; t < vn; t += LZ_THREADS) {
                        double s0 = 0.0, s1 = 0.0;
                        int q = 0;
                        for (; q + 1 <= j; q += 2) {
                            s0 = fma(hred[q], Vs[q * VNp + t], s0);
                            s1 = fma(hred[q + 1], Vs[(q + 1) * VNp + t], s1);
                        }
                        if (q <= j) s0 = fma(hred[q], Vs[q * VNp + t], s0);
                        SMD(wv)[t] -= (s0 + s1);
                    }
                }
            }
            if (hn2 <= 0.25 * wn2) break;      // DGKS: the pass shrank w by less than 1/sqrt(2)... (margin 2x): one pass is enough
            __syncthreads();                   // second pass (breakdown only): its dots read the updated wv of every warp
        }
        // ||w_new||^2 = ||w||^2 - ||h||^2 ; v_{j+1} = w / beta: keep my rows, push them into every peer's staging buffer
        {
            const double beta2 = fmax(wn2 - hn2, 0.0);
            beta = sqrt(beta2);
            const double ib = (beta > 0.0) ? 1.0 / beta : 0.0;
            if (tid == 0) { SMD(Hd)[j] = alpha; SMD(He)[j] = beta; }
            if (2 * vn <= LZ_THREADS) {
                // the thread pair of row t shares the pushes: sub 0 -> peers 0, 2, ..., sub 1 -> peers 1, 3, ...
                const int t = tid >> 1, sub = tid & 1;
                double wnew = (t < vn && sub == 0) ? SMD(wv)[t] : 0.0;
                wnew = __shfl_sync(0xffffffffu, wnew, (tid & 31) & ~1);
                if (t < vn) {
                    const double v = wnew * ib;
                    if (sub == 0) SMD(Vs)[k * VNp + t] = v;
                    for (int c = sub; c < C; c += 2) cluster.map_shared_rank(SMD(vbuf), c)[v0 + t] = v;
                }
            } else {
                __syncthreads();
                for (int t = tid; t < vn; t += LZ_THREADS) {
                    const double v = SMD(wv)[t] * ib;
                    SMD(Vs)[k * VNp + t] = v;
                    for (int c = 0; c < C; ++c) cluster.map_shared_rank(SMD(vbuf), c)[v0 + t] = v;
                }
            }
        }
        LZ3_TICK(6);
        cluster.sync();
        LZ3_TICK(4);

        // ================= Ritz analysis (redundant in every CTA) =================
        if (beta <= a.tol && k < howmany) howmany = k;
        if (k == K || beta <= a.tol) {
            const int lda = lanczos_kp(K);
            const int m = (k + 1) & ~1;
            // ---- fast path: leading pairs of the plain tridiagonal by bisection + twisted vectors ----
            bool done_bi = false;
            if (a.use_bi && arrow < 0 && 2 * (size_t)lda * lda >= ritz_bi_scratch_doubles(K)) {
                RitzBiScratch bs = ritz_bi_carve(SMD(JA), K);          // JA and JB are contiguous and unused here
                const int mb = ritz_top_bi(k, SMD(Hd), SMD(He), howmany + 4, SMD(D), SMD(JU), lda, bs, profiling ? s_prof : nullptr);
                if (mb > 0) {
                    int* order = reinterpret_cast<int*>(SMD(order));
                    for (int i = tid; i < mb; i += LZ_THREADS) { order[i] = i; SMD(f)[i] = beta * SMD(JU)[(k - 1) + i * lda]; }
                    __syncthreads();
                    int cv = 0;
                    while (cv < mb && fabs(SMD(f)[cv]) <= a.tol) cv++;
                    if (cv >= howmany && cv < mb) { converged = cv; finished = 1; done_bi = true; }
                    __syncthreads();
                }
            }
            if (profiling && tid == 0) { s_prof[8 + (done_bi ? 0 : 1)] += 1; s_prof[10] += clock64() - tprev; }
            if (done_bi) { LZ3_TICK(5); continue; }
            {
                double* JA = SMD(JA);
                for (int idx = tid; idx < m * m; idx += LZ_THREADS) {
                    int r = idx % m, c = idx / m;
                    double v = 0.0;
                    if (r < k && c < k) {
                        if (r == c) v = SMD(Hd)[r];
                        else {
                            int lo = min(r, c), hi = max(r, c);
                            if (hi == arrow && lo < arrow) v = SMD(Harr)[lo];
                            else if (hi == lo + 1 && !(lo < arrow && hi <= arrow)) v = SMD(He)[lo];
                        }
                    }
                    JA[r + c * lda] = v;
                }
                for (int idx = tid; idx < lda * lda; idx += LZ_THREADS) SMD(JB)[idx] = 0.0;
            }
            __syncthreads();
            JacobiScratch js = jacobi_carve(SMD(jscratch), lda);
            const double* Jd;
            const bool warm = first_analysis && a.ritz_rd && (int)a.ritz_rd[0] == k && k == K;
            if (warm) Jd = jacobi_eigh_smem_warm(m, k, SMD(JA), SMD(JB), lda, SMD(JU), lda, a.ritz_rd + 1, js);
            else Jd = jacobi_eigh_smem_fast(m, SMD(JA), SMD(JB), lda, SMD(JU), lda, js);
            __syncthreads();
            if (first_analysis && a.ritz_wr && cta == 0) {
                for (int idx = tid; idx < lda * lda; idx += LZ_THREADS) a.ritz_wr[1 + idx] = SMD(JU)[idx];
                if (tid == 0) a.ritz_wr[0] = (k == K) ? (double)k : -1.0;
            }
            first_analysis = false;
            int* order = reinterpret_cast<int*>(SMD(order));
            rank_sort_desc(k, Jd, lda, order);
            __syncthreads();
            for (int i = tid; i < k; i += LZ_THREADS) {
                int o = order[i];
                SMD(D)[i] = Jd[o + o * lda];
                SMD(f)[i] = beta * SMD(JU)[(k - 1) + o * lda];
            }
            __syncthreads();
            converged = 0;
            while (converged < k && fabs(SMD(f)[converged]) <= a.tol) converged++;
            if (converged >= howmany) {
                finished = 1;
            } else if (k == K) {
                if (numiter == a.maxiter) {
                    finished = 1;
                } else {
                    // ---- thick restart: V[:, 0:keep] <- V U[:, order[0:keep]], in place row by row ----
                    const int keep = (3 * K + 2 * converged) / 5;
                    double* Vs = SMD(Vs);
                    for (int t = tid; t < vn; t += LZ_THREADS) {
                        double row[LZC_KMAX];
                        for (int i = 0; i < K; ++i) row[i] = Vs[i * VNp + t];
                        for (int q = 0; q < keep; ++q) {
                            const double* u = SMD(JU) + order[q] * lda;
                            double s = 0.0;
                            for (int i = 0; i < K; ++i) s = fma(row[i], u[i], s);
                            Vs[q * VNp + t] = s;
                        }
                        Vs[keep * VNp + t] = Vs[K * VNp + t];
                    }
                    __syncthreads();
                    for (int i = tid; i < K; i += LZ_THREADS) {
                        double d = (i < keep) ? SMD(D)[i] : 0.0;
                        double fa = (i < keep) ? SMD(f)[i] : 0.0;
                        SMD(Hd)[i] = d; SMD(Harr)[i] = fa; SMD(He)[i] = 0.0;
                    }
                    __syncthreads();
                    arrow = keep;
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
        const int* order = reinterpret_cast<const int*>(SMD(order));
        const double* Vs = SMD(Vs);
        for (int idx = tid; idx < nvals * vn; idx += LZ_THREADS) {
            int q = idx / vn, t = idx - q * vn;
            const double* u = SMD(JU) + order[q] * lda;
            double s = 0.0;
            for (int i = 0; i < k; ++i) s = fma(Vs[i * VNp + t], u[i], s);
            a.Y[(size_t)q * a.ld + v0 + t] = s;
        }
    }
    if (cta == 0) {
        if (!failed) for (int i = tid; i < nvals; i += LZ_THREADS) a.vals[i] = SMD(D)[i];
        if (tid == 0) {
            a.info[0] = failed ? 0 : nvals; a.info[1] = failed ? 0 : converged; a.info[2] = (int)gsync; a.info[3] = numiter;
            a.scal[S_NUMOPS] += (double)gsync;
            a.scal[S_HEADER + 3 * a.cone + 2] = failed ? 0.0 : (double)converged;
            if (failed || converged == 0) a.scal[S_POISON] = 1.0;
        }
    }
    if (profiling && tid < 32) a.prof[32 * prow + tid] += s_prof[tid];
    cluster.sync();      // no CTA leaves while a peer may still address its shared memory
#undef SMD
#undef LZ3_TICK
#undef tprev
}

}  // namespace pb
