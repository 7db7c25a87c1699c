// lanczos_cl3.cuh — third generation of the cluster-replicated thick-restart Lanczos kernel.
//
// Same contract as k_lanczos_cl (lanczos_cl.cuh): KrylovKit.eigsolve(A, resid, nev, :LR, Lanczos(orth, K,
// maxiter, tol)) as called from reference src/eigsolver.jl:802-812.  The basis still lives as one replica
// per thread-block cluster in distributed shared memory.  What changed, each item measured on B200
// (profiles/r1h_* ... r1k_*, scripts/lat_bench.cu, scripts/symv_bench.cu):
//
//  1. One re-orthogonalisation exchange per step instead of two.  KrylovKit's recurrence is the local
//     three-term step (w -= alpha v_j + beta_{j-1} v_{j-1}) followed by two Gram-Schmidt passes.  The local
//     step needs alpha = v_j . w, a grid-wide dot: its per-CTA partials ride the w all-gather that the grid
//     exchange performs anyway.  After the local step the components of w along the basis are O(eps ||X||)
//     (or the known arrow row right after a thick restart), so ONE classical Gram-Schmidt pass over the whole
//     basis restores orthogonality to machine precision; a second pass runs only when the first one removed a
//     visible part of w (DGKS criterion, ||h||^2 > ||w||^2 / 4: breakdown / invariant subspace).  alpha for the
//     local step is only needed to ~1e-7 (the pass removes the rest along v_j exactly), so its sum over the CTAs
//     runs through an FP32 shuffle tree.  scripts/lz_variant_check.py
//     shows identical mat-vec counts, converged counts and 1e-15 orthogonality versus the two-pass schemes.
//  2. No register spills on the per-step path.  Every acquire at gpu/cluster scope makes ptxas emit CCTL.IVALL
//     (L1 invalidate), so a spilled value re-read after a barrier costs a full L2 round trip; the second
//     generation kept ~25 shared-memory pointers live (50 registers) and spilled loop state, which showed up as
//     ~1.7 us of "fold" per step.  Here the shared-memory layout travels as integer offsets in the kernel
//     parameters (constant bank) and addresses are rebuilt where they are used.
//  3. Strip symv without shuffles.  The second-generation symv gave every warp a run of 64-double chunks of the
//     slab: ragged row tails made up to 13 dependent L2 round trips per mat-vec, the v operand was re-read from
//     shared memory for every row, and every row cost a 5-step shuffle tree (~110 cycles per 64-bit step).  Here
//     warp w owns a column strip of CPW chunks for ALL rows of the slab: its piece of v sits in registers, the
//     loads of a batch of RB rows are issued back to back, every lane parks its per-row partial sums in shared
//     memory ([row][thread], conflict-free) and after one block barrier half-warp h adds up the 512 partials of
//     row h in eight chains and stores the finished w entry straight to global memory — the separate "fold" phase
//     and all shuffles are gone.  alpha's partial is row number rl of the same table.  (Keeping half of the slab
//     resident in shared memory was tried and measured: no gain at n = 2000.)
//  4. No grid barrier.  A counter barrier costs release fence (~500 cycles with stores in flight) + atomic + poll
//     + acquire fence, ~0.9 us on top of the slowest CTA, and then the data still has to be fetched.  Here the w
//     entries and the alpha partials travel as flagged 16-byte words (two 8-byte halves {lo32, tag}, {hi32, tag}:
//     a reader that sees the tag in both halves has the value — the "LL" exchange of lanczos.cuh): the lanes that
//     need a value poll that value, no fence on either side.  Completing the alpha all-gather (every CTA needs
//     every CTA's partial) is what orders a step after the previous one, so two buffers suffice.
//  5. No cluster barrier either: the Gram-Schmidt dots and the new Lanczos vector are delivered into the peers'
//     shared memory with st.async ... mbarrier::complete_tx and consumed behind an mbarrier phase (helpers below).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "jacobi.cuh"
#include "kernels_vec.cuh"
#include "lanczos.cuh"
#include "lanczos_cl.cuh"
#include "ritz_bi.cuh"

namespace pb {

constexpr int LZ3_GMAX = 256;    // largest grid the per-CTA alpha partials are sized for

// shared-memory layout, in doubles from the start of dynamic shared memory (host-computed, read from the constant bank)
struct Lz3Layout {
    int wpart;   // 2 * 32 * LZ_NW         per-warp totals of the slab symv, [buffer][row of the round][warp]
    int aprod;   // 32                     w[r] v[r] of my slab rows (this CTA's share of alpha)
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
    int xres;    // nres * cpr * 64        the LAST nres rows of my slab of X, staged once per launch by the TMA engine
    int nres;
    // implicit-operator mode (k_lanczos_cl3<..., true>): instead of slab rows the region holds
    int imp_Ys;     // imp_r * VNp         my rows of the previous iterate's kept eigenvectors
    int imp_lam;    // LZ3_RMAX            their eigenvalues
    int imp_tq;     // LZ3_RMAX            lambda_q (Y_q . v)
    int imp_tpart;  // C * LZ3_RMAX        per-peer partial Y' v (written by the peers)
    int imp_twarp;  // LZ_NW * LZ3_RMAX    per-warp partials
    int imp_apart;  // C + LZ_NW           per-peer partial alpha (written by the peers), per-warp partials
    int imp_sval;   // imp_cap             values of my rows of the sparse part, tau folded in
    int imp_scol;   // imp_cap ints
    int imp_sptr;   // VNp + 1 ints
    int imp_cap;
    int imp_r;      // kept eigenpairs the layout has room for
    int total;   // doubles
};
constexpr int LZ3_RMAX = 32;     // kept eigenpairs of the previous iterate the implicit operator can carry

// everything except the resident slab rows
__host__ inline Lz3Layout lanczos_cl3_layout(int K, int nres, int vn_max, int n, int C, int imp_cap = -1, int imp_r = 0) {
    Lz3Layout L{};
    const int Kp = lanczos_kp(K), VNp = lanczos_cl_vnp(vn_max), cpr = lanczos_cpr(n);
    int d = 0;
    auto take = [&](int cnt) { int o = d; d += (cnt + 1) & ~1; return o; };     // keep everything 16-byte aligned
    L.wpart = take(2 * 32 * LZ_NW);
    L.aprod = take(32);
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
    L.nres = nres;
    L.xres = take(nres * cpr * 64);
    if (imp_cap >= 0) {
        L.imp_Ys = take((imp_r > 0 ? imp_r : LZ3_RMAX) * VNp);      // imp_r: kept pairs the launch can meet (<= target rank)
        L.imp_lam = take(LZ3_RMAX);
        L.imp_tq = take(LZ3_RMAX);
        L.imp_tpart = take(C * LZ3_RMAX);
        L.imp_twarp = take(LZ_NW * LZ3_RMAX);
        L.imp_apart = take(C + LZ_NW);
        L.imp_sval = take(imp_cap);
        L.imp_scol = take((imp_cap + 1) / 2 + 1);
        L.imp_sptr = take((VNp + 2) / 2 + 1);
        L.imp_cap = imp_cap;
        L.imp_r = imp_r > 0 ? imp_r : LZ3_RMAX;
    }
    L.total = d;
    return L;
}

struct LanczosCl3Args {
    const double* X; int n, ld;
    const double* x0;
    double* Y;                 // out: Ritz vectors, ld x K
    uint4* wg;                 // [2][ld] mat-vec result, flagged ("LL") words: data and ready-signal in one 16-byte store
    uint4* apart;              // [2][LZ3_GMAX] per-CTA partial alpha = v_j[slab] . w[slab], flagged words
    unsigned int epoch_base;   // the g-th exchange of this launch is tagged epoch_base + g (unique over launches: no memset)
    int res_begin_off;         // rl - res_begin_off is the first RESIDENT row of a slab of rl rows (the last L.nres rows of the
                               // shortest slab live in shared memory; a slab that is one row longer streams one row more)
    const double* ritz_rd;     // optional warm start of the Ritz eigenproblem (see LanczosClArgs)
    double* ritz_wr;
    int nev, K, maxiter;
    double tol;
    int vn_max;                // ceil(n / C): basis rows per CTA
    int rbase, rrem;           // symv rows of CTA c: rbase + (c < rrem), starting at c * rbase + min(c, rrem)   (n = G * rbase + rrem)
    int vbase, vrem;           // basis rows of cluster rank c, same formula with C
    double* vals; int* info; double* scal; int cone;
    int use_bi;
    double stop_above;         // finish as soon as the largest Ritz value exceeds this (it is a lower bound of lambda_max);
                               // 1e300 = never: used by cone_feas, which only needs to know whether lambda_min < -tol
    // implicit operator: the matrix is NOT stored; it is  Y diag(lam) Y' - tau S  with Y, lam the kept
    // eigenpairs of the previous projection (x_k = svec(Y lam Y')) and S = mat(M'y + c), sparse
    const double* imp_Y; const int* imp_kept_idx; const double* imp_kept_lam; const int* imp_nkept;
    const int* imp_rowptr; const int* imp_col; const int* imp_pos; const double* imp_coef;   // CSR by row of the pattern of S
    const double* imp_Mty; const double* imp_c;      // this cone's svec blocks of M'y and of the (scaled) objective
    double imp_tau;
    int debug;                 // PROXSDP_B200_LZ_DEBUG=1: CTA 0 prints every Ritz analysis / restart (device printf)
    int arrow_restart;         // 1: keep the arrowhead form after a thick restart (dense Jacobi Ritz solves; PROXSDP_B200_LZ_ARROW=1)
    long long spin_limit;      // cycles a spin loop waits for a peer before the launch gives up (~2 s; PROXSDP_B200_LZ_SPIN_S
                               // stretches it for runs under compute-sanitizer, where a peer can be 100x slower)
    int bi_memory;             // 1: a declined bisection Ritz solve sends the later analyses of the launch straight to the dense solver
    int pf_rows;               // > 0: slab rows prefetched into L2 ahead of the register loads (matrices larger than L2;
                               // PROXSDP_B200_LZ_PF overrides the host's choice)
    int poll_ns;               // back-off between two polls of the flagged exchange words (PROXSDP_B200_LZ_POLL_NS)
    int eager;                 // 1: KrylovKit's `eager` schedule (opt.krylovkit_eager, reference src/eigsolver.jl:809): the Ritz
                               // analysis also runs after every expansion step once k >= howmany
    int strict;                // 1: KrylovKit's arithmetic to the letter — alpha of the local step summed in FP64 and two
                               // Gram-Schmidt passes on every step (PROXSDP_B200_LZ_STRICT=1); 0: FP32 tree for the provisional
                               // alpha and a second pass only when the DGKS test asks for it (same counts)
    long long* prof;
    Lz3Layout L;
};

// ---------------------------------------------------------------------------
// Cluster exchange without fences: st.async delivers an 8-byte value into a peer's shared memory and credits the
// peer's mbarrier with the bytes; the consumer arms the phase with the byte count it expects and waits on the
// phase parity.  A cluster barrier costs MEMBAR.ALL.GPU + barrier + L1 invalidate on every thread (~12 % of the
// kernel's stall samples in profiles/r1h); here only the data travels.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned int mapa_u32(unsigned int addr, int cta_rank) {
    unsigned int r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void st_async_f64(unsigned int remote_addr, double v, unsigned int remote_mbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr), "l"(__double_as_longlong(v)),
                 "r"(remote_mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// all threads call it; returns false after `limit` cycles (~2 s by default: a peer died)
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned int parity, const long long limit) {
    const unsigned int a = smem_u32(bar);
    unsigned int done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return true;
        if (clock64() - t0 > limit) return false;
    }
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// L2 prefetch of `bytes` contiguous bytes (multiple of 16) by the TMA engine: no registers, no shared memory, no completion
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned int bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// One step of the transposing butterfly that sums per-lane partials of N rows over the warp: lanes with the mask bit
// clear keep the lower half [0, H) of the rows, lanes with the bit set the upper half [H, N); after the five steps
// (masks 16 ... 1) every lane holds the warp total of at most one row.
template <int N, int NMAX>
__device__ __forceinline__ void bfly_step(double (&v)[NMAX], const bool hi, const int mask) {
    constexpr int H = (N + 1) / 2;
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const double upper = (H + i < N) ? v[H + i] : 0.0;
        const double send = hi ? v[i] : upper;
        const double recv = __shfl_xor_sync(0xffffffffu, send, mask);
        v[i] = (hi ? upper : v[i]) + recv;
    }
}
// acc[i]: this lane's partial of row i (i < nrows <= NMAX <= 32).  On return acc[0] is the warp total of row `off` when
// valid != 0.  H1 + ... + H5 (20 for NMAX = 18) 64-bit shuffles instead of 5 per row.
template <int NMAX>
__device__ __forceinline__ void bfly_reduce(double (&acc)[NMAX], const int lane, const int nrows, int& off, int& valid) {
    constexpr int N1 = NMAX, H1 = (N1 + 1) / 2, H2 = (H1 + 1) / 2, H3 = (H2 + 1) / 2, H4 = (H3 + 1) / 2, H5 = (H4 + 1) / 2;
    static_assert(NMAX <= 32 && H5 == 1, "at most 32 rows per round");
    int cnt = nrows;
    off = 0;
    { const bool hi = (lane & 16) != 0; bfly_step<N1, NMAX>(acc, hi, 16); off += hi ? H1 : 0; cnt = hi ? max(cnt - H1, 0) : min(cnt, H1); }
    { const bool hi = (lane & 8) != 0;  bfly_step<H1, NMAX>(acc, hi, 8);  off += hi ? H2 : 0; cnt = hi ? max(cnt - H2, 0) : min(cnt, H2); }
    { const bool hi = (lane & 4) != 0;  bfly_step<H2, NMAX>(acc, hi, 4);  off += hi ? H3 : 0; cnt = hi ? max(cnt - H3, 0) : min(cnt, H3); }
    { const bool hi = (lane & 2) != 0;  bfly_step<H3, NMAX>(acc, hi, 2);  off += hi ? H4 : 0; cnt = hi ? max(cnt - H4, 0) : min(cnt, H4); }
    { const bool hi = (lane & 1) != 0;  bfly_step<H4, NMAX>(acc, hi, 1);  off += hi ? H5 : 0; cnt = hi ? max(cnt - H5, 0) : min(cnt, H5); }
    valid = cnt >= 1;
}

// loads of one chunk of RB streamed rows (rows first, first + 1, ... below `stream_end`): CPW 16-byte loads per row and lane
template <int CPW, int RB>
__device__ __forceinline__ void strip_load(double2 (&x)[RB][CPW], const double* __restrict__ gcol, const int ld, const int first,
                                           const int stream_end, const int nch) {
#pragma unroll
    for (int i = 0; i < RB; ++i) {
        if (first + i < stream_end) {
            const double* rp = gcol + (size_t)(first + i) * ld;
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                x[i][c] = make_double2(0.0, 0.0);
                if (c < nch) x[i][c] = ld_stream_d2(reinterpret_cast<const double2*>(rp + c * 64));
            }
        }
    }
}

// up to NMAX STREAMED rows first .. first + count - 1 of the slab: strip loads, RB rows per chunk, the next chunk in flight
// while the current one is multiplied; warp totals of the rows -> wrow[row - first][warp]
template <int CPW, int RB, int NMAX>
__device__ __forceinline__ void symv_sub_streamed(const double* __restrict__ gcol, const int ld, const int first, const int count,
                                                  const int nch, const double2 (&vr)[CPW], double* wrow, const int lane, const int warp) {
    double acc[NMAX];
    const int stream_end = first + count;
    double2 x[RB][CPW];
    strip_load<CPW, RB>(x, gcol, ld, first, stream_end, nch);
#pragma unroll
    for (int c0 = 0; c0 < NMAX; c0 += RB) {
        double2 xn[RB][CPW];
        if (c0 + RB < NMAX) strip_load<CPW, RB>(xn, gcol, ld, first + c0 + RB, stream_end, nch);
#pragma unroll
        for (int i = 0; i < RB; ++i) {
            if (c0 + i < NMAX) {
                double t = 0.0;
                if (c0 + i < count) {
#pragma unroll
                    for (int c = 0; c < CPW; ++c) { t = fma(x[i][c].x, vr[c].x, t); t = fma(x[i][c].y, vr[c].y, t); }
                }
                acc[c0 + i] = t;
            }
        }
        if (c0 + RB < NMAX) {
#pragma unroll
            for (int i = 0; i < RB; ++i)
#pragma unroll
                for (int c = 0; c < CPW; ++c) x[i][c] = xn[i][c];
        }
    }
    int off, valid;
    bfly_reduce<NMAX>(acc, lane, count, off, valid);
    if (valid) wrow[off * LZ_NW + warp] = acc[0];
}

// up to NMAX RESIDENT rows: srow points at this lane's piece of the first of them in shared memory
template <int CPW, int NMAX>
__device__ __forceinline__ void symv_sub_resident(const double* srow, const int rstride, const int count, const int nch,
                                                  const double2 (&vr)[CPW], double* wrow, const int lane, const int warp) {
    double acc[NMAX];
#pragma unroll
    for (int i = 0; i < NMAX; ++i) {
        double t = 0.0;
        if (i < count) {
            const double* rp = srow + (size_t)i * rstride;
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                if (c < nch) { const double2 q = *reinterpret_cast<const double2*>(rp + c * 64); t = fma(q.x, vr[c].x, t); t = fma(q.y, vr[c].y, t); }
            }
        }
        acc[i] = t;
    }
    int off, valid;
    bfly_reduce<NMAX>(acc, lane, count, off, valid);
    if (valid) wrow[off * LZ_NW + warp] = acc[0];
}

// Slab symv w[r0 : r0 + rl] = X[r0 : r0 + rl, :] v plus this CTA's share of alpha = v[slab] . w[slab], written straight
// to global memory as flagged words (wg_slab[row], *alpha_out).
//   * Warp s owns the column strip of CPW 64-double chunks for ALL rows of the slab; its piece of v sits in registers.
//   * The last rows of the slab (from res_begin on) are RESIDENT in shared memory: the TMA engine staged them once when
//     the kernel started (cp.async.bulk + mbarrier, see the kernel prologue); at n = 2000 that takes ~40 % of the matrix
//     off the L2 -> SM path of every mat-vec.  The rows before them are streamed with 128-bit loads, RB rows per chunk,
//     the next chunk in flight while the current one is multiplied.
//   * Every lane keeps one partial sum per row of a sub-round (NMAX rows) in registers; a transposing butterfly (12
//     shuffles for 9 rows instead of 5 per row) leaves one warp total per row in one lane, the NS <= 16 warp totals of
//     a row go through a 4 KB table, and after one block barrier per 32 rows thread r of warp 0 publishes row r.
//     (The first generation parked 512 partials per row in a 70 KB shared-memory table that half-warps summed after a
//     block barrier: 1.2 us per mat-vec, and the table took the room the resident rows now use.)
// All threads of the block must call it.
template <int CPW, int RB, int NMAX, bool PFL2>
__device__ __forceinline__ void symv_slab(const double* __restrict__ X, const int ld, const int r0, const int rl, const int cpr,
                                          const double* vbuf, const double* xres, const int res_begin, double* wpart, double* aprod,
                                          uint4* wg_slab, uint4* alpha_out, const unsigned int tag, const int pf_rows, long long* prof) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long tp = (prof && tid == 0) ? clock64() : 0;
    const int NS = (cpr + CPW - 1) / CPW;
    const bool active = warp < NS;
    const int nch = min(CPW, cpr - warp * CPW);
    double2 vr[CPW];
#pragma unroll
    for (int c = 0; c < CPW; ++c) vr[c] = (active && c < nch) ? *reinterpret_cast<const double2*>(vbuf + (warp * CPW + c) * 64 + 2 * lane) : make_double2(0.0, 0.0);
    if (tid < 32) aprod[tid] = 0.0;      // (the same warp reads and updates it below: no barrier needed)
    for (int rbeg = 0; rbeg < rl; rbeg += 32) {          // one block barrier per 32 rows
        const int rend = min(rl, rbeg + 32);
        const int buf = (rbeg >> 5) & 1;                 // two tables: the writers of round r + 2 are behind the barrier of
        double* wtab = wpart + buf * 32 * LZ_NW;         // round r + 1, which warp 0 passes only after reading round r
        if (active) {
            const int se = min(rend, res_begin);
            const double* gcol = X + (size_t)r0 * ld + (size_t)warp * CPW * 64 + 2 * lane;
            for (int f = rbeg; f < se; f += NMAX) {
                // matrices beyond L2: the TMA engine pulls the rows pf_rows ahead of the ones being multiplied into L2, so the
                // register loads below see L2 latency instead of HBM latency (same bytes in flight -> more bandwidth)
                // (a warp-uniform instruction: the whole warp issues it once per row)
                if (PFL2 && pf_rows > 0 && warp == 0) {
                    for (int q = 0; q < NMAX; ++q)
                        if (f + pf_rows + q < min(rl, res_begin))
                            bulk_prefetch_l2(X + (size_t)(r0 + f + pf_rows + q) * ld, (unsigned int)(cpr * 64 * sizeof(double)));
                }
                symv_sub_streamed<CPW, RB, NMAX>(gcol, ld, f, min(NMAX, se - f), nch, vr, wtab + (f - rbeg) * LZ_NW, lane, warp);
            }
            const int rstride = cpr * 64;
            const double* scol = xres + (size_t)warp * CPW * 64 + 2 * lane;
            for (int f = max(rbeg, res_begin); f < rend; f += NMAX)
                symv_sub_resident<CPW, NMAX>(scol + (size_t)(f - res_begin) * rstride, rstride, min(NMAX, rend - f), nch, vr,
                                             wtab + (f - rbeg) * LZ_NW, lane, warp);
        }
        if (prof && tid == 0) { const long long tn = clock64(); prof[15] += tn - tp; tp = tn; }          // warp 0: loads + FMAs + butterflies
        __syncthreads();
        if (prof && tid == 0) { const long long tn = clock64(); prof[16] += tn - tp; tp = tn; }          // waiting for the slowest warp
        if (tid < rend - rbeg) {
            const double* q = wtab + tid * LZ_NW;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int w = 0;
            for (; w + 3 < NS; w += 4) { s0 += q[w]; s1 += q[w + 1]; s2 += q[w + 2]; s3 += q[w + 3]; }
            for (; w < NS; ++w) s0 += q[w];
            const double sres = (s0 + s1) + (s2 + s3);
            ll_store(wg_slab + rbeg + tid, sres, tag);
            aprod[tid] = fma(sres, vbuf[r0 + rbeg + tid], aprod[tid]);
        }
    }
    // the first pf_rows rows of the slab for the NEXT mat-vec: HBM is idle during the exchange and Gram-Schmidt phases
    if (PFL2 && pf_rows > 0 && warp == 1) {
        for (int q = 0; q < min(pf_rows, min(rl, res_begin)); ++q)
            bulk_prefetch_l2(X + (size_t)(r0 + q) * ld, (unsigned int)(cpr * 64 * sizeof(double)));
    }
    // alpha: the products of my rows (threads 0 .. 31, all in warp 0), summed in four chains by lane 0
    if (tid < 32) {
        __syncwarp();
        if (tid == 0) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) { s0 += aprod[i]; s1 += aprod[i + 1]; s2 += aprod[i + 2]; s3 += aprod[i + 3]; }
            ll_store(alpha_out, (s0 + s1) + (s2 + s3), tag);
        }
    }
    if (prof && tid == 0) { const long long tn = clock64(); prof[17] += tn - tp; tp = tn; }              // row totals + publish
}

// scratch of the re-tridiagonalising restart inside JB (lda x lda doubles): m2 Lanczos vectors of pitch m2 | 1, T~, two work vectors
__device__ __forceinline__ bool lz3_restart_fits(const int m2, const int lda) {
    return m2 >= 1 && m2 <= 31 && m2 * (m2 | 1) + 2 * m2 + 64 <= lda * lda;
}

// Ritz analysis + thick restart, out of line: it runs once or twice per launch, and keeping its ~100 KB of code
// out of the per-step loop keeps the loop inside the instruction cache.  State in/out through st[] (shared memory):
//   st[0] k, st[2] arrow, st[3] howmany, st[4] numiter, st[5] first_analysis, st[6] converged, st[7] finished,
//   st[9] bisection verdicts of this launch: bit 0 = the leading pairs were declined once, bit 1 = the `keep` pairs were
//   (a near-multiple eigenvalue stays one for the rest of the eigsolve: later analyses go straight to the dense solver).
// All threads of the CTA call it; every CTA of the grid computes the same result from bitwise identical data.
__device__ __noinline__ void lz3_ritz(const LanczosCl3Args& a, double* sbase, int* st, const double beta, const int vn, const int cta,
                                      long long* prof) {
#define SMD(field) (sbase + a.L.field)
    const int tid = threadIdx.x;
    const int K = a.K, VNp = lanczos_cl_vnp(a.vn_max);
    int k = st[0], arrow = st[2], howmany = st[3], numiter = st[4];
    const bool first_analysis = st[5] != 0;
    int bi_bad = a.bi_memory ? st[9] : 0;
    int converged = 0, finished = 0;
    __syncthreads();                              // everybody has read the state before thread 0 rewrites it
    const int lda = lanczos_kp(K);
    const int m = (k + 1) & ~1;
    long long t_in = prof ? clock64() : 0;
    bool first_out = first_analysis;
    // Thick restart as KrylovKit does it: keep the `keep` leading Ritz pairs (theta_t, y_t = V u_t) and the residual
    // vector r.  In the basis [y_0 .. y_keep-1, r] the Rayleigh quotient is diag(theta) bordered by the row
    // f_t = beta u_t[K-1]; an orthogonal change of basis inside span(y) turns it back into a TRIDIAGONAL matrix:
    // Lanczos on diag(theta) started from f / ||f|| (full re-orthogonalisation, at most 31 steps on vectors of length
    // keep: one warp, lane <-> component) gives Q~ with Q~' diag(theta) Q~ = T~ and Q~' f = ||f|| e_0; taken in
    // reverse order the coupling to r sits in the last row, as the three-term recurrence expects.  The basis is
    // rotated once by U Q~, the recurrence carries on with a plain tridiagonal (so every later Ritz analysis takes the
    // bisection path), and nothing downstream knows an arrow ever existed.
    auto thick_restart_tridiag = [&](const int keep, const int nlock) {
        // nlock: leading Ritz pairs that have converged (|f_t| <= tol).  They are LOCKED: kept as decoupled 1 x 1 blocks
        // at the front of the new tridiagonal, their residual coupling (<= tol) dropped.  Feeding them to the small
        // Lanczos run below would be numerically fatal: their components of f are ~1e-15 of the others and shrink by
        // that factor at every restart, so the run exhausts its Krylov space early and rounding noise along the dominant
        // locked eigenvector is amplified ~100x per step until it pollutes the rotated basis (seen on Max-Cut n = 2000
        // once two of three wanted pairs had converged: the locked Ritz value moved from 1283.66 to 1275.47).
        double* Vs = SMD(Vs);
        const int* order = reinterpret_cast<const int*>(SMD(order));
        double* Wc = SMD(JA);                    // K x keep coefficients of the new basis vectors in the old Lanczos basis
        double* Qs = SMD(JB);                    // m2 x MP: Qs[j * MP + c] = component c of Lanczos vector j
        const int m2 = keep - nlock;             // pairs that take part in the re-tridiagonalisation (1 <= m2 <= 31)
        const int MP = m2 | 1;                   // odd row pitch: conflict-free columns
        double* ta = Qs + m2 * MP;               // diagonal of T~
        double* tb = ta + m2;                    // off-diagonal of T~
        double* ws = tb + m2;                    // work vector
        double* hs = ws + 32;                    // re-orthogonalisation coefficients
        __syncthreads();                         // D, f, order, JU are final; the bisection / Jacobi scratch in JA, JB is dead
        if (tid < 32) {
            const int c = tid;                   // component c <-> kept pair nlock + c
            const double th = (c < m2) ? SMD(D)[nlock + c] : 0.0;
            const double fc = (c < m2) ? SMD(f)[nlock + c] : 0.0;
            const double nf = sqrt(warp_sum(fc * fc));
            double thmax = fabs(th);
            for (int o = 16; o > 0; o >>= 1) thmax = fmax(thmax, __shfl_xor_sync(0xffffffffu, thmax, o));
            thmax = fmax(thmax, 1e-300);
            double qprev = 0.0, qcur = (nf > 0.0) ? fc / nf : ((c == 0) ? 1.0 : 0.0);
            double bprev = 0.0;
            for (int j = 0; j < m2; ++j) {
                if (c < m2) Qs[j * MP + c] = qcur;
                double w = th * qcur;
                const double aj = warp_sum(qcur * w);
                w -= aj * qcur + bprev * qprev;
                if (c == 0) ta[j] = aj;
                if (j == m2 - 1) break;
                // Gram-Schmidt passes of w against q_0 .. q_j (lane i <-> coefficient of q_i), repeated until a pass no longer
                // shrinks the vector (DGKS): with components of f that are tiny or exactly zero — multiple eigenvalues, pairs
                // that have all but converged — w is annihilated down to rounding noise, and a fixed number of passes then
                // leaves a "unit" vector that is not orthogonal to its predecessors at all (seen: a rotated basis with
                // || V'V - I || = 1 and Ritz residuals of 1e4 on matrices with 3-fold eigenvalues)
                auto reorth = [&]() -> double {
                    double nprev = sqrt(warp_sum((c < m2) ? w * w : 0.0)), ncur = nprev;
                    for (int pass = 0; pass < 6; ++pass) {
                        __syncwarp();
                        ws[c] = (c < m2) ? w : 0.0;
                        __syncwarp();
                        double h = 0.0;
                        if (c <= j) for (int t = 0; t < m2; ++t) h = fma(Qs[c * MP + t], ws[t], h);
                        hs[c] = (c <= j) ? h : 0.0;
                        __syncwarp();
                        if (c < m2) for (int i = 0; i <= j; ++i) w = fma(-hs[i], Qs[i * MP + c], w);
                        ncur = sqrt(warp_sum((c < m2) ? w * w : 0.0));
                        if (!(ncur < 0.7 * nprev)) break;
                        nprev = ncur;
                    }
                    return ncur;
                };
                double nb = reorth();
                double bj = nb;
                if (!(nb > 1e-14 * thmax)) {
                    // breakdown (an f_t that is zero or negligible, a repeated theta): what is left of w is rounding noise.
                    // T~ decouples here (off-diagonal 0: the coupling dropped is below 1e-14 ||Theta||) and the run carries on
                    // with the coordinate vector that has the largest component outside span(q_0 .. q_j), which is >= 1/sqrt(m2)
                    double g = (c < m2) ? 1.0 : -1.0;
                    if (c < m2) for (int i = 0; i <= j; ++i) { const double qi = Qs[i * MP + c]; g = fma(-qi, qi, g); }
                    int t0 = c;
                    for (int o = 16; o > 0; o >>= 1) {
                        const double og = __shfl_xor_sync(0xffffffffu, g, o);
                        const int ot = __shfl_xor_sync(0xffffffffu, t0, o);
                        if (og > g || (og == g && ot < t0)) { g = og; t0 = ot; }
                    }
                    w = (c == t0) ? 1.0 : 0.0;
                    nb = reorth();
                    bj = 0.0;
                }
                if (c == 0) tb[j] = bj;
                qprev = qcur;
                qcur = w / nb;
                bprev = bj;
            }
            if (c == 0) tb[m2 - 1] = nf;           // coupling of the last rotated vector to the residual vector
        }
        __syncthreads();
        // new basis: the locked Ritz vectors first, then vector nlock + i = sum_t y_{nlock + t} Q~[m2 - 1 - i][t]
        for (int idx = tid; idx < K * keep; idx += LZ_THREADS) {
            const int r = idx % K, i = idx / K;
            double sacc;
            if (i < nlock) {
                sacc = SMD(JU)[r + order[i] * lda];
            } else {
                const double* qv = Qs + (m2 - 1 - (i - nlock)) * MP;
                sacc = 0.0;
                for (int t = 0; t < m2; ++t) sacc = fma(SMD(JU)[r + order[nlock + t] * lda], qv[t], sacc);
            }
            Wc[r + i * K] = sacc;
        }
        __syncthreads();
        for (int t = tid; t < vn; t += LZ_THREADS) {
            double row[LZC_KMAX];
            for (int i = 0; i < K; ++i) row[i] = Vs[i * VNp + t];
            for (int q = 0; q < keep; ++q) {
                const double* u = Wc + q * K;
                double sacc = 0.0;
                for (int i = 0; i < K; ++i) sacc = fma(row[i], u[i], sacc);
                Vs[q * VNp + t] = sacc;
            }
            Vs[keep * VNp + t] = Vs[K * VNp + t];
        }
        __syncthreads();
#ifdef PROXSDP_B200_LZ_DEBUG_PRINTF
        if (a.debug && cta == 0 && tid == 0)
            printf("[lz] restart keep %d locked %d  ||f|| %.3e  T~ diag %.6g ... %.6g  off %.3e ... %.3e\n", keep, nlock, tb[m2 - 1], ta[0], ta[m2 - 1],
                   tb[0], m2 > 1 ? tb[m2 - 2] : 0.0);
#endif
        // (the locked values are read before anything is overwritten: D is a separate array)
        for (int i = tid; i < K; i += LZ_THREADS) {
            double hd = 0.0, he = 0.0;
            if (i < nlock) { hd = SMD(D)[i]; he = 0.0; }
            else if (i < keep) {
                const int q = i - nlock;
                hd = ta[m2 - 1 - q];
                he = (q < m2 - 1) ? tb[m2 - 2 - q] : tb[m2 - 1];
            }
            SMD(Hd)[i] = hd; SMD(He)[i] = he; SMD(Harr)[i] = 0.0;
        }
        arrow = -1;
        k = keep;                       // the caller's k++ makes it keep + 1
        numiter++;
    };
    // restart that keeps the arrowhead form (large Krylov dimensions: more than 31 unlocked pairs do not fit the one-warp reduction above;
    // the dense Jacobi solver diagonalises arrow + tridiagonal tail directly, which is the same Krylov space)
    auto thick_restart_arrow = [&](const int keep) {
        double* Vs = SMD(Vs);
        const int* order = reinterpret_cast<const int*>(SMD(order));
        for (int t = tid; t < vn; t += LZ_THREADS) {
            double row[LZC_KMAX];
            for (int i = 0; i < K; ++i) row[i] = Vs[i * VNp + t];
            for (int q = 0; q < keep; ++q) {
                const double* u = SMD(JU) + order[q] * lda;
                double sacc = 0.0;
                for (int i = 0; i < K; ++i) sacc = fma(row[i], u[i], sacc);
                Vs[q * VNp + t] = sacc;
            }
            Vs[keep * VNp + t] = Vs[K * VNp + t];
        }
        __syncthreads();
        for (int i = tid; i < K; i += LZ_THREADS) {
            double d = (i < keep) ? SMD(D)[i] : 0.0;
            double fa = (i < keep) ? SMD(f)[i] : 0.0;
            SMD(Hd)[i] = d; SMD(Harr)[i] = fa; SMD(He)[i] = 0.0;
        }
        arrow = keep;
        k = keep;
        numiter++;
    };
    // ---- fast path: the leading pairs of the tridiagonal Rayleigh quotient by bisection + twisted vectors ----
    bool done_bi = false;
    if (a.use_bi && arrow < 0 && !(bi_bad & 1) && 2 * (size_t)lda * lda >= ritz_bi_scratch_doubles(K)) {
        RitzBiScratch bs = ritz_bi_carve(SMD(JA), K);          // JA and JB are contiguous and unused here
        auto solve = [&](const int want, const int have) -> int {
            const int got = ritz_top_bi(k, SMD(Hd), SMD(He), want, SMD(D), SMD(JU), lda, bs, prof, have);
            if (got > 0) {
                int* order = reinterpret_cast<int*>(SMD(order));
                for (int i = tid; i < got; i += LZ_THREADS) { order[i] = i; SMD(f)[i] = beta * SMD(JU)[(k - 1) + i * lda]; }
            }
            __syncthreads();
            return got;
        };
        int mb = solve(howmany + 4, 0);
        if (mb == 0 && !(a.eager && k < 8)) bi_bad |= 1;     // (eager analyses of a basis younger than 8 vectors go to the dense solver without prejudice)
        if (mb > 0) {
            int cv = 0;
            while (cv < mb && fabs(SMD(f)[cv]) <= a.tol) cv++;
#ifdef PROXSDP_B200_LZ_DEBUG_PRINTF
            if (a.debug && cta == 0 && tid == 0)
                printf("[lz] analysis numiter %d k %d beta %.3e bisection mb %d cv %d  D %.9g %.9g %.9g %.9g  f %.2e %.2e %.2e %.2e\n", numiter, k, beta,
                       mb, cv, SMD(D)[0], SMD(D)[1], SMD(D)[2], SMD(D)[3], SMD(f)[0], SMD(f)[1], SMD(f)[2], SMD(f)[3]);
#endif
            if (cv >= howmany && cv < mb) { converged = cv; finished = 1; done_bi = true; }
            else if (SMD(D)[0] > a.stop_above) { converged = max(cv, 1); finished = 1; done_bi = true; }      // bound certified
            else if (a.eager && beta > a.tol && cv < howmany && k < K) { converged = cv; done_bi = true; }      // eager analysis, not there yet: expand further
            else if (cv < howmany && k == K && numiter < a.maxiter) {
                // not converged at the end of a Krylov cycle: the restart needs the `keep` leading pairs — one warp each
                const int keep = (3 * K + 2 * cv) / 5;
                if (keep <= RITZ_BI_MAXM && keep < k && !a.arrow_restart && !(bi_bad & 2) && lz3_restart_fits(keep - cv, lda)) {
                    __syncthreads();
                    if (mb < keep) mb = solve(keep, mb);      // the pairs of the first call stay
                    if (mb < keep) bi_bad |= 2;
                    int cv2 = 0;
                    while (cv2 < mb && fabs(SMD(f)[cv2]) <= a.tol) cv2++;
                    if (mb >= keep && cv2 == cv) {
                        converged = cv;
                        thick_restart_tridiag(keep, cv);
                        done_bi = true;
                        first_out = false;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (prof && tid == 0) { prof[8 + (done_bi ? 0 : 1)] += 1; prof[10] += clock64() - t_in; }
    if (!done_bi) {
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
        first_out = false;
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
#ifdef PROXSDP_B200_LZ_DEBUG_PRINTF
        if (a.debug && cta == 0 && tid == 0)
            printf("[lz] analysis numiter %d k %d arrow %d beta %.3e dense cv %d  D %.9g %.9g %.9g %.9g  f %.2e %.2e %.2e %.2e\n", numiter, k, arrow, beta,
                   converged, SMD(D)[0], SMD(D)[1], SMD(D)[2], SMD(D)[3], SMD(f)[0], SMD(f)[1], SMD(f)[2], SMD(f)[3]);
#endif
        if (converged >= howmany) {
            finished = 1;
        } else if (SMD(D)[0] > a.stop_above) {      // the largest Ritz value never exceeds lambda_max: bound certified
            converged = max(converged, 1);
            finished = 1;
        } else if (k == K) {
            if (numiter == a.maxiter) {
                finished = 1;
            } else {
                const int keep = (3 * K + 2 * converged) / 5;
                const int nlk = min(converged, keep - 1);
                if (keep - nlk <= 31 && arrow < 0 && !a.arrow_restart && lz3_restart_fits(keep - nlk, lda)) thick_restart_tridiag(keep, nlk);
                else thick_restart_arrow(keep);
            }
        }
    }
    if (tid == 0) {
        st[0] = k; st[2] = arrow; st[3] = howmany; st[4] = numiter; st[5] = first_out ? 1 : 0; st[6] = converged; st[7] = finished;
        st[9] = bi_bad;
    }
    __syncthreads();
#undef SMD
}

// OP selects the operator: 0 = dense matrix in global memory, row slabs over the whole grid (the default); 1 = implicit
// low-rank + sparse operator on ONE cluster; 2 = dense matrix RESIDENT in the distributed shared memory of ONE cluster
// (mid-size cones: CTA rank c holds rows [c n/C, (c+1) n/C) of X for the whole eigsolve, no grid exchange at all).
template <int CPW, int RB, int NMAX, int OP = 0, bool PFL2 = false>
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
    constexpr bool IMP = (OP == 1);
    const int VNp = lanczos_cl_vnp(a.vn_max);
    const int cpr = lanczos_cpr(n);
    const int Kp2 = K + 2;
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
    __shared__ int s_state[10];
    __shared__ double s_beta;
    __shared__ __align__(8) unsigned long long s_mbar[5];      // [0] Gram-Schmidt dots exchange, [1] publication of v_{j+1},
                                                               // [2] TMA staging of the resident slab rows,
                                                               // implicit operator: [3] Y'v exchange, [4] alpha exchange
    unsigned int ph_dots = 0;                                  // dots exchanges completed so far (parity = phase & 1); the
                                                               // publication phase is the step number: parity (gsync - 1) & 1
    const int prow = (cta == 0) ? 0 : (cta == C - 1) ? 1 : (cta == G / 2) ? 2 : (cta == G - 1) ? 3 : -1;
    const bool profiling = (a.prof != nullptr) && prow >= 0;
    if (tid < 32) s_prof[tid] = 0;
    if (tid == 0) { s_ok = 1; s_state[9] = 0; }
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
    if (tid == 0) {
        mbar_init(&s_mbar[0], 1);
        mbar_init(&s_mbar[1], 1);
        mbar_init(&s_mbar[2], 1);
        mbar_init(&s_mbar[3], 1);
        mbar_init(&s_mbar[4], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // stage the resident rows of my slab (the last a.res_begin_off of its rl rows) into shared memory: one bulk
        // copy per row on the TMA engine (row = cpr * 64 doubles = ld doubles: X is zero-padded to ld columns),
        // completion counted in bytes on s_mbar[2]; the copies run while the cluster barrier below is crossed
        // (OP == 2: ALL my basis rows of X, rows v0 .. v0 + vn - 1 — the matrix lives in the cluster's shared memory)
        const int nres = (OP == 1) ? 0 : (OP == 2) ? vn : min(a.res_begin_off, rl);
        const int first_res = (OP == 2) ? v0 : r0 + rl - nres;
        if (nres > 0) {
            const unsigned int rowbytes = (unsigned int)(cpr * 64 * sizeof(double));
            mbar_expect_tx(&s_mbar[2], (unsigned int)nres * rowbytes);
            for (int q = 0; q < nres; ++q)
                bulk_g2s(SMD(xres) + (size_t)q * cpr * 64, a.X + (size_t)(first_res + q) * a.ld, rowbytes, &s_mbar[2]);
        }
    }
    int imp_nk = 0;
    if constexpr (IMP) {
        // my rows of the kept eigenvectors of the previous projection, their eigenvalues, and my rows of the sparse part
        // with (M'y + c) and -tau folded into the values: all of it stays in shared memory for the whole eigsolve
        imp_nk = min(*a.imp_nkept, a.L.imp_r);
        double* Ys = SMD(imp_Ys);
        for (int idx = tid; idx < imp_nk * vn; idx += LZ_THREADS) {
            const int q = idx / vn, t = idx - q * vn;
            Ys[q * VNp + t] = a.imp_Y[(size_t)(v0 + t) + (size_t)a.imp_kept_idx[q] * a.ld];
        }
        for (int q = tid; q < LZ3_RMAX; q += LZ_THREADS) SMD(imp_lam)[q] = (q < imp_nk) ? a.imp_kept_lam[q] : 0.0;
        int* sptr = reinterpret_cast<int*>(SMD(imp_sptr));
        int* scol = reinterpret_cast<int*>(SMD(imp_scol));
        const int base = a.imp_rowptr[v0];
        for (int t = tid; t <= vn; t += LZ_THREADS) sptr[t] = a.imp_rowptr[v0 + t] - base;
        const int cnt = min(a.imp_rowptr[v0 + vn] - base, a.L.imp_cap);
        for (int e = tid; e < cnt; e += LZ_THREADS) {
            const int pos = a.imp_pos[base + e];
            scol[e] = a.imp_col[base + e];
            SMD(imp_sval)[e] = -a.imp_tau * (a.imp_coef[base + e] * (a.imp_Mty[pos] + a.imp_c[pos]));
        }
    }
    cluster.sync();       // everybody's shared memory and barriers are initialised before any peer writes into them
    if ((OP == 2 ? vn : (OP == 0 ? min(a.res_begin_off, rl) : 0)) > 0 && !mbar_wait(&s_mbar[2], 0, a.spin_limit)) s_ok = 0;
    if (profiling && tid == 0) tprev = clock64();

    while (!finished) {
        LZ3_TICK(7);
        // The loop state is parked in shared memory across the symv (the register-hungry part of the step) and
        // re-read behind the barrier that follows it: a value kept live across the symv gets spilled, and a
        // spill re-read after a fence is an L2 round trip because the fence invalidates L1.
        if (tid == 0) {
            s_state[0] = k; s_state[1] = (int)gsync; s_state[2] = arrow; s_state[3] = howmany; s_state[4] = numiter;
            s_state[5] = first_analysis ? 1 : 0; s_state[8] = (int)ph_dots; s_beta = beta;
        }
        // ================= symv on my slab of rows: w slab and my share of alpha straight to global =================
        // (gsync still counts the exchanges done so far: this step's buffers have parity (gsync + 1) & 1)
        if constexpr (OP == 0)
            symv_slab<CPW, RB, NMAX, PFL2>(a.X, a.ld, r0, rl, cpr, SMD(vbuf), SMD(xres), max(rl - a.res_begin_off, 0), SMD(wpart), SMD(aprod),
                                     a.wg + (size_t)((gsync + 1) & 1) * a.ld + r0, a.apart + (size_t)((gsync + 1) & 1) * LZ3_GMAX + cta,
                                     a.epoch_base + gsync + 1, a.pf_rows, profiling ? s_prof : nullptr);
        else
            __syncthreads();      // the state words just parked are re-read below
        // (no block barrier here: the state words were written a whole step ago)
        k = s_state[0]; gsync = (unsigned int)s_state[1]; arrow = s_state[2]; howmany = s_state[3]; numiter = s_state[4];
        first_analysis = s_state[5] != 0; ph_dots = (unsigned int)s_state[8]; beta = s_beta;
        const int j = k - 1;
        LZ3_TICK(0);
        ++gsync;
        // ======== gather + local three-term step + partial Gram-Schmidt dots, warp by warp ========
        // Every dependent step of a single warp costs 25-110 cycles on B200 (DFMA 23, LDS ~35-70, one 64-bit shuffle
        // step ~110: scripts/lat_bench.cu), so the phases below are laid out for short chains: conflict-free
        // shared-memory columns and serial adds in 4 chains instead of shuffle trees.
        // Warp w owns the basis rows [w RW, (w+1) RW) of this CTA for the whole phase: its lanes poll the flagged w
        // entries of exactly those rows, so no block barrier is needed between the exchange, the local step and the
        // dots; the only block-wide value is alpha, which warp LZ_NW-1 sums from the per-CTA partials.  alpha is only
        // needed to about 1e-7: the Gram-Schmidt pass removes whatever is left along v_j exactly (alpha_j = alpha~ +
        // h_j), so its shuffle tree runs in FP32 (35 instead of 110 cycles per step).
        const int RW = (a.vn_max + LZ_NW - 1) / LZ_NW;
        const int t_lo = min(vn, warp * RW), t_hi = min(vn, (warp + 1) * RW);
        {
            const unsigned int tag = a.epoch_base + gsync;
            const uint4* wgp = a.wg + (size_t)(gsync & 1) * a.ld + v0;
            const double* Vs = SMD(Vs);
            double* wv = SMD(wv);
            double wreg[2];                                  // RW <= 64 rows per warp is checked on the host
            if constexpr (OP != 0) {
                // ---- single-cluster operators: w on my basis rows, no grid exchange (the cluster is the whole grid) ----
                const double* vb = SMD(vbuf);
                double ap = 0.0;
                if constexpr (OP == 2) {
                    // dense rows resident in shared memory: warp <-> its RW <= 4 basis rows, lane <-> a double2 of every
                    // 64-column chunk; the row totals end up in lanes 0 .. RW-1 (lane u <-> row t_lo + u)
                    const double* xr = SMD(xres) + 2 * lane;
                    const int rstride = cpr * 64;
                    double acc[4] = {0.0, 0.0, 0.0, 0.0};
                    for (int c = 0; c < cpr; ++c) {
                        const double2 vv = *reinterpret_cast<const double2*>(vb + c * 64 + 2 * lane);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (t_lo + u < t_hi) {
                                const double2 q = *reinterpret_cast<const double2*>(xr + (size_t)(t_lo + u) * rstride + c * 64);
                                acc[u] = fma(q.x, vv.x, acc[u]);
                                acc[u] = fma(q.y, vv.y, acc[u]);
                            }
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
                    }
                    double w = 0.0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (lane == u && t_lo + u < t_hi) w = acc[u];
                    wreg[0] = w; wreg[1] = 0.0;
                    if (t_lo + lane < t_hi) ap = w * vb[v0 + t_lo + lane];
                } else {
                // ---- w = Y (lam (Y' v)) - tau S v ----
                const double* Ys = SMD(imp_Ys);
                {   // (1) Y' v: lane <-> eigenvector q, warp <-> its rows, then across warps and across the cluster
                    double s0 = 0.0, s1 = 0.0;
                    if (lane < imp_nk) {
                        const double* yq = Ys + lane * VNp;
                        int t = t_lo;
                        for (; t + 1 < t_hi; t += 2) { s0 = fma(yq[t], vb[v0 + t], s0); s1 = fma(yq[t + 1], vb[v0 + t + 1], s1); }
                        if (t < t_hi) s0 = fma(yq[t], vb[v0 + t], s0);
                    }
                    SMD(imp_twarp)[warp * LZ3_RMAX + lane] = s0 + s1;
                }
                __syncthreads();
                if (tid < imp_nk) {
                    const double* tw = SMD(imp_twarp) + tid;
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                    for (int w = 0; w < LZ_NW; w += 4) { s0 += tw[w * LZ3_RMAX]; s1 += tw[(w + 1) * LZ3_RMAX]; s2 += tw[(w + 2) * LZ3_RMAX]; s3 += tw[(w + 3) * LZ3_RMAX]; }
                    const double sres = (s0 + s1) + (s2 + s3);
                    const unsigned int dst = smem_u32(SMD(imp_tpart) + crank * LZ3_RMAX + tid), mb = smem_u32(&s_mbar[3]);
                    for (int c = 0; c < C; ++c) st_async_f64(mapa_u32(dst, c), sres, mapa_u32(mb, c));
                }
                if (tid == 0) mbar_expect_tx(&s_mbar[3], (unsigned int)(C * imp_nk * 8));
                if (imp_nk > 0 && !mbar_wait(&s_mbar[3], (gsync - 1) & 1, a.spin_limit)) s_ok = 0;
                if (tid < imp_nk) {
                    double sres = 0.0;
                    for (int c = 0; c < C; ++c) sres += SMD(imp_tpart)[c * LZ3_RMAX + tid];      // rank order: same bits in every CTA
                    SMD(imp_tq)[tid] = SMD(imp_lam)[tid] * sres;
                }
                __syncthreads();
                // (2) my rows of w, and my share of alpha = v . w
                const int* sptr = reinterpret_cast<const int*>(SMD(imp_sptr));
                const int* scol = reinterpret_cast<const int*>(SMD(imp_scol));
                const double* sval = SMD(imp_sval);
                const double* tq = SMD(imp_tq);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int t = t_lo + lane + 32 * u;
                    double w = 0.0;
                    if (t < t_hi) {
                        double w0 = 0.0, w1 = 0.0;
                        int q = 0;
                        for (; q + 1 < imp_nk; q += 2) { w0 = fma(Ys[q * VNp + t], tq[q], w0); w1 = fma(Ys[(q + 1) * VNp + t], tq[q + 1], w1); }
                        if (q < imp_nk) w0 = fma(Ys[q * VNp + t], tq[q], w0);
                        double z0 = 0.0, z1 = 0.0;
                        int e = sptr[t];
                        const int ee = min(sptr[t + 1], a.L.imp_cap);
                        for (; e + 1 < ee; e += 2) { z0 = fma(sval[e], vb[scol[e]], z0); z1 = fma(sval[e + 1], vb[scol[e + 1]], z1); }
                        if (e < ee) z0 = fma(sval[e], vb[scol[e]], z0);
                        w = (w0 + w1) + (z0 + z1);
                        ap = fma(w, vb[v0 + t], ap);
                    }
                    wreg[u] = w;
                }
                }      // OP == 1
                ap = warp_sum(ap);
                if (lane == 0) SMD(imp_apart)[C + warp] = ap;
                __syncthreads();
                if (tid == 0) {
                    const double* aw = SMD(imp_apart) + C;
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                    for (int w = 0; w < LZ_NW; w += 4) { s0 += aw[w]; s1 += aw[w + 1]; s2 += aw[w + 2]; s3 += aw[w + 3]; }
                    const double sres = (s0 + s1) + (s2 + s3);
                    const unsigned int dst = smem_u32(SMD(imp_apart) + crank), mb = smem_u32(&s_mbar[4]);
                    for (int c = 0; c < C; ++c) st_async_f64(mapa_u32(dst, c), sres, mapa_u32(mb, c));
                    mbar_expect_tx(&s_mbar[4], (unsigned int)(C * 8));
                }
                if (!mbar_wait(&s_mbar[4], (gsync - 1) & 1, a.spin_limit)) s_ok = 0;
                if (tid == 0) {
                    const double* ap_ = SMD(imp_apart);
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                    int c = 0;
                    for (; c + 3 < C; c += 4) { s0 += ap_[c]; s1 += ap_[c + 1]; s2 += ap_[c + 2]; s3 += ap_[c + 3]; }
                    for (; c < C; ++c) s0 += ap_[c];
                    SMD(hred)[K + 1] = (s0 + s1) + (s2 + s3);
                }
            } else {
            {
                uint4 rr[2];
                bool have[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) have[u] = !(t_lo + lane + 32 * u < t_hi);
                const long long tw = clock64();
                while (true) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (!have[u]) { rr[u] = ll_peek(wgp + t_lo + lane + 32 * u); have[u] = (rr[u].y == tag && rr[u].w == tag); }
                    }
                    if (__all_sync(0xffffffffu, have[0] && have[1])) break;
                    if (clock64() - tw > a.spin_limit) { s_ok = 0; break; }      // ~2 s: a peer died; give up instead of hanging
                    // back off between polls: 120 CTAs x 16 warps spinning on L2 take request slots from the CTAs that are
                    // still streaming their slab (measured: profiles/r2_poll_sweep.txt)
                    if (a.poll_ns > 0) __nanosleep((unsigned int)a.poll_ns);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) wreg[u] = (t_lo + lane + 32 * u < t_hi) ? ll_value(rr[u]) : 0.0;
            }
            if (warp == LZ_NW - 1) {
                const uint4* app = a.apart + (size_t)(gsync & 1) * LZ3_GMAX;
                // four partials per lane polled concurrently (one L2 round trip when they are there), 128 CTAs per sweep
                double sd = 0.0;
                const long long tw = clock64();
                for (int cb = 0; cb < G; cb += 128) {
                    uint4 pr[4];
                    bool got[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) got[u] = !(cb + lane + 32 * u < G);
                    while (true) {
                        bool all = true;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (!got[u]) { pr[u] = ll_peek(app + cb + lane + 32 * u); got[u] = (pr[u].y == tag && pr[u].w == tag); }
                            all = all && got[u];
                        }
                        if (__all_sync(0xffffffffu, all)) break;
                        if (clock64() - tw > a.spin_limit) { s_ok = 0; break; }
                        if (a.poll_ns > 0) __nanosleep((unsigned int)a.poll_ns);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (cb + lane + 32 * u < G) sd += ll_value(pr[u]);
                }
                if (a.strict) {
                    sd = warp_sum(sd);
                    if (lane == 0) SMD(hred)[K + 1] = sd;
                } else {
                    float sf = (float)sd;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sf += __shfl_xor_sync(0xffffffffu, sf, o);
                    if (lane == 0) SMD(hred)[K + 1] = (double)sf;
                }
            }
            }      // dense operator
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
            if (!s_ok) { failed = 1; break; }
            const double alpha0 = SMD(hred)[K + 1];
#pragma unroll
            for (int u = 0; u < 2; ++u) { const int t = t_lo + lane + 32 * u; if (t < t_hi) wv[t] = fma(-alpha0, Vs[j * VNp + t], wreg[u]); }
            __syncwarp();
        }
        double alpha = SMD(hred)[K + 1];
        LZ3_TICK(1);

        // ================= one Gram-Schmidt pass inside the cluster (a second one only on breakdown) =================
        double wn2 = 0.0, hn2 = 0.0;
        for (int pass = 0; pass < 3; ++pass) {
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
                const unsigned int dst = smem_u32(SMD(hpart) + (pass & 1) * C * Kp2 + crank * Kp2 + q), mb = smem_u32(&s_mbar[0]);
                for (int c = 0; c < C; ++c) st_async_f64(mapa_u32(dst, c), sres, mapa_u32(mb, c));
            }
            if (tid == 0) mbar_expect_tx(&s_mbar[0], (unsigned int)(C * (j + 2) * 8));      // j + 2 dots from each of the C peers (me included)
            LZ3_TICK(2);
            if (!mbar_wait(&s_mbar[0], ph_dots & 1, a.spin_limit)) s_ok = 0;
            ++ph_dots;
            LZ3_TICK(3);
            // h[q] = sum over the C peers in rank order (fixed tree)
            for (int q = tid; q <= j + 1; q += LZ_THREADS) {
                const double* hp = SMD(hpart) + (pass & 1) * C * Kp2 + q;
                double sres;
                if (C == 8) {
                    sres = ((hp[0] + hp[Kp2]) + (hp[2 * Kp2] + hp[3 * Kp2])) + ((hp[4 * Kp2] + hp[5 * Kp2]) + (hp[6 * Kp2] + hp[7 * Kp2]));
                } else {
                    // any other cluster size: four chains in a fixed order (same bits in every CTA)
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                    int c = 0;
                    for (; c + 3 < C; c += 4) { s0 += hp[c * Kp2]; s1 += hp[(c + 1) * Kp2]; s2 += hp[(c + 2) * Kp2]; s3 += hp[(c + 3) * Kp2]; }
                    for (; c < C; ++c) s0 += hp[c * Kp2];
                    sres = (s0 + s1) + (s2 + s3);
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
                if (2 * a.vn_max <= LZ_THREADS) {      // cluster-uniform choice: every CTA sums ||h||^2 in the same order
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
            // DGKS: the pass shrank w by less than 1/sqrt(2) (margin 2x): one pass is enough.  After a second pass the norm
            // below, ||w||^2 - ||h||^2, is only trusted when that pass removed next to nothing; otherwise (w lies in the span
            // of the basis to working precision: an invariant subspace, e.g. a multiple eigenvalue whose copies are exhausted)
            // a third pass measures the norm of what is left directly — its ||h||^2 is then pure rounding.  Without it the
            // difference cancels, beta comes out wrong by a large factor, the next "unit" vector has norm 60 and its alpha
            // pollutes the Rayleigh quotient (seen on mcp500-1 at a 3-fold eigenvalue: a Ritz value of 142 933 for ||A|| = 38).
            if (pass == 2) break;
            if (pass == 1 && hn2 <= 1e-4 * wn2) break;
            if (pass == 0 && !a.strict && hn2 <= 0.25 * wn2) break;
            // second pass (breakdown only).  Its dots read the updated wv of every warp (block barrier), and they travel
            // through the SAME mbarrier as the first pass: a peer that is already in its second pass must not credit
            // bytes to my barrier while my first-pass phase is still open (I may be waiting for a third CTA), hence a
            // cluster barrier on this rare path — every CTA of the cluster has closed phase one before anyone sends.
            cluster.sync();
        }
        // ||w_new||^2 = ||w||^2 - ||h||^2 ; v_{j+1} = w / beta: keep my rows, push them into every peer's staging buffer
        {
            // 1/beta by rsqrt (one MUFU + two Newton steps) instead of sqrt followed by a division: both sit on the
            // critical path of every thread; beta itself is beta2 * (1/beta), good to an ulp
            const double beta2 = fmax(wn2 - hn2, 0.0);
            const double ib = (beta2 > 0.0) ? rsqrt(beta2) : 0.0;
            beta = beta2 * ib;
            if (tid == 0) { SMD(Hd)[j] = alpha; SMD(He)[j] = beta; }
            if (2 * a.vn_max <= LZ_THREADS) {      // cluster-uniform choice: every CTA sums ||h||^2 in the same order
                // the thread pair of row t shares the pushes: sub 0 -> peers 0, 2, ..., sub 1 -> peers 1, 3, ...
                const int t = tid >> 1, sub = tid & 1;
                double wnew = (t < vn && sub == 0) ? SMD(wv)[t] : 0.0;
                wnew = __shfl_sync(0xffffffffu, wnew, (tid & 31) & ~1);
                if (t < vn) {
                    const double v = wnew * ib;
                    if (sub == 0) SMD(Vs)[k * VNp + t] = v;
                    const unsigned int dst = smem_u32(SMD(vbuf) + v0 + t), mb = smem_u32(&s_mbar[1]);
                    for (int c = sub; c < C; c += 2) st_async_f64(mapa_u32(dst, c), v, mapa_u32(mb, c));
                }
            } else {
                __syncthreads();
                for (int t = tid; t < vn; t += LZ_THREADS) {
                    const double v = SMD(wv)[t] * ib;
                    SMD(Vs)[k * VNp + t] = v;
                    const unsigned int dst = smem_u32(SMD(vbuf) + v0 + t), mb = smem_u32(&s_mbar[1]);
                    for (int c = 0; c < C; ++c) st_async_f64(mapa_u32(dst, c), v, mapa_u32(mb, c));
                }
            }
        }
        if (tid == 0) mbar_expect_tx(&s_mbar[1], (unsigned int)(n * 8));      // every row of v_{j+1}, from whichever peer owns it
        LZ3_TICK(6);
        if (!mbar_wait(&s_mbar[1], (gsync + 1) & 1, a.spin_limit)) s_ok = 0;
        LZ3_TICK(4);

        // ================= Ritz analysis (redundant in every CTA; out of line) =================
        if (beta <= a.tol && k < howmany) howmany = k;
        if (k == K || beta <= a.tol || (a.eager && k >= howmany)) {
            if (tid == 0) {
                s_state[0] = k; s_state[2] = arrow; s_state[3] = howmany; s_state[4] = numiter; s_state[5] = first_analysis ? 1 : 0;
            }
            __syncthreads();
            lz3_ritz(a, sbase, s_state, beta, vn, cta, profiling ? s_prof : nullptr);
            k = s_state[0]; arrow = s_state[2]; howmany = s_state[3]; numiter = s_state[4]; first_analysis = s_state[5] != 0;
            converged = s_state[6]; finished = s_state[7];
            __syncthreads();                    // the state words are rewritten at the top of the next step
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
