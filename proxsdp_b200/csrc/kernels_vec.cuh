// kernels_vec.cuh — fused HBM-bound vector kernels of the PDHG iteration.
//
//   k_svec_to_mat<1>   K1+K2  x - tau*(Mty + c) -> full symmetric matrix  (pdhg.jl:622, prox_operators.jl:1-16)
//   k_primal_tail      K1     same update for the SOC / free tail of x     (pdhg.jl:622)
//   k_soc_project      K11    second-order-cone projection                 (prox_operators.jl:138-158)
//   k_spmv_csr         K12    Mx = M x                                     (pdhg.jl:634)
//   k_dual_trial       K14-16 over-relaxed dual step + box projection      (pdhg.jl:544-553, prox_operators.jl:160-170)
//   k_spmv_mt_norm     K13+17 Mty = M' y fused with the linesearch norms   (pdhg.jl:556-566)
//   k_residual_primal  K18+21 primal fixed-point residual + c'x            (residuals.jl:41-48,22)
//   k_residual_dual    K18+20+21 dual residual, feasibility, b'y, h'y      (residuals.jl:52-59,5-29)
//
// All reductions are two-phase (per-block partials folded by the last block in a
// fixed order), hence bit-reproducible run to run for a fixed grid.
#pragma once
#include "common.cuh"

namespace pb {

// ---------------------------------------------------------------------------
// per-iteration scalar record shared between kernels and the host control loop
// ---------------------------------------------------------------------------
enum ScalarSlot {
    S_LS_ACCEPTED = 0,  // 1.0 once a linesearch trial has been accepted
    S_LS_TRIAL,         // index of the accepted trial
    S_TAU,              // accepted primal step
    S_YNORM2,           // ||y_new - y||^2 of the last evaluated trial
    S_MTYNORM2,         // ||Mty_new - Mty||^2 of the last evaluated trial
    S_LS_EVALS,         // number of trials actually evaluated
    S_RES_P_NUM,        // max |(x - tau Mty) - (x_old - tau Mty_old)|
    S_RES_P_DEN,        // max |x_old - tau Mty_old|
    S_RES_D_NUM,
    S_RES_D_DEN,
    S_EQ_MAX,           // max |Mx - b|
    S_IN_MAX,           // max(0, max(Mx - h))
    S_PRIM_OBJ,         // c'x
    S_BY,               // b'y_eq
    S_HY,               // h'y_in
    S_SOC_GAP,          // max over SOC cones of ||v|| - t
    S_POISON,           // != 0: a Lanczos call did not converge; the rest of the iteration was skipped
    S_NUMOPS,           // Lanczos mat-vecs performed this iteration
    S_ELAPSED,          // host wall-clock seconds since the start of the solve (sharded runs: max over ranks decides time limits)
    S_HEADER            // per-cone records follow: current_rank, min_eig, converged (3 doubles each)
};

struct ReduceWs {
    double* partials;        // [slots][max_blocks]
    unsigned int* counters;  // one per kernel kind
    int max_blocks;
};

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

// ---------------------------------------------------------------------------
// Structural-nonzero bitmap of (M'y, c): bit k set iff M' has a non-empty row k or c[k] != 0.  On Max-Cut n = 2000 that is
// 24 000 of 2 001 000 entries: the streaming kernels below skip the loads of Mty / Mty_old / c where the bit is clear
// (x - tau (0 + 0) == x to the last bit), i.e. 32 - 48 MB of zeros per iteration are not read.
// ---------------------------------------------------------------------------
__global__ void k_mask_from_c(const double* __restrict__ c, long long N, unsigned int* __restrict__ mask) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < N; i += stride)
        if (c[i] != 0.0) atomicOr(mask + (i >> 5), 1u << (i & 31));
}
__global__ void k_mask_from_rows(const int* __restrict__ rows, int n_rows, unsigned int* __restrict__ mask) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_rows) { const int r = rows[q]; atomicOr(mask + (r >> 5), 1u << (r & 31)); }
}
__device__ __forceinline__ bool mask_bit(const unsigned int* __restrict__ mask, const long long k) {
    return mask == nullptr || ((mask[k >> 5] >> (k & 31)) & 1u) != 0u;
}

// ---------------------------------------------------------------------------
// K1+K2: X = mat(x - tau*(Mty + c)) for one PSD cone, full n x n (both triangles).
// 32x32 tiles of the upper triangle; the mirrored tile goes through shared memory so
// that both global writes are coalesced.  svec order: k(i,j) = j(j+1)/2 + i, i <= j.
// grid = (#tile pairs with bi <= bj), block = (32, 8)
// ---------------------------------------------------------------------------
template <bool PRIMAL>
__global__ void __launch_bounds__(256)
k_svec_to_mat(const double* __restrict__ x, const double* __restrict__ Mty, const double* __restrict__ c,
              double tau, double scale, int n, int ld, double* __restrict__ X,
              const unsigned int* __restrict__ mask = nullptr, long long mask_off = 0) {
    __shared__ double tile[32][33];
    // decode tile pair index t -> (bi, bj), bi <= bj, column-major over the upper triangle of tiles
    int t = blockIdx.x;
    int bj = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((bj + 1) * (bj + 2) / 2 <= t) ++bj;
    while (bj * (bj + 1) / 2 > t) --bj;
    int bi = t - bj * (bj + 1) / 2;
    const double sqrt2 = 1.41421356237309504880;
    int i = bi * 32 + threadIdx.x;
#pragma unroll
    for (int jj = threadIdx.y; jj < 32; jj += 8) {
        int j = bj * 32 + jj;
        double v = 0.0;
        if (i < n && j < n && i <= j) {
            size_t k = (size_t)j * (size_t)(j + 1) / 2 + (size_t)i;
            double u;
            if (PRIMAL) {                                                      // x .-= tau .* (Mty .+ c)
                const double mc = mask_bit(mask, mask_off + (long long)k) ? add_rn(Mty[k], c[k]) : 0.0;
                u = sub_rn(x[k], mul_rn(tau, mc));
            }
            else u = x[k] * scale;
            v = (i != j) ? u / sqrt2 : u;                                      // off-diagonals / sqrt(2)
            X[(size_t)i + (size_t)j * ld] = v;
        }
        tile[jj][threadIdx.x] = v;   // tile[j_local][i_local]
    }
    __syncthreads();
    // mirrored element (j, i): write X[j + i*ld]; thread x runs over j (contiguous)
    int j2 = bj * 32 + threadIdx.x;
#pragma unroll
    for (int ii = threadIdx.y; ii < 32; ii += 8) {
        int i2 = bi * 32 + ii;
        if (i2 < n && j2 < n && i2 < j2) X[(size_t)j2 + (size_t)i2 * ld] = tile[threadIdx.x][ii];
    }
}

// K1 for the non-PSD tail (SOC blocks and free variables): x_new = x - tau*(Mty + c)
__global__ void k_primal_tail(const double* __restrict__ x, const double* __restrict__ Mty,
                              const double* __restrict__ c, double tau, long long begin, long long end,
                              double* __restrict__ x_new) {
    long long i = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < end; i += stride) x_new[i] = sub_rn(x[i], mul_rn(tau, add_rn(Mty[i], c[i])));
}

// ---------------------------------------------------------------------------
// K11: SOC projection, one block per cone.  x[off] = t, x[off+1 .. off+len) = v.
// Also records max(||v|| - t) AFTER projection (soc_convergence, residuals.jl:73-86).
// ---------------------------------------------------------------------------
__global__ void k_soc_project(double* __restrict__ x, const long long* __restrict__ soc_off,
                              const int* __restrict__ soc_len, double* __restrict__ soc_gap,
                              const double* __restrict__ poison) {
    __shared__ double red[40];
    if (poison && *poison != 0.0) return;
    int k = blockIdx.x;
    double* t = x + soc_off[k];
    double* v = t + 1;
    int len = soc_len[k] - 1;
    double s = 0.0;
    for (int i = threadIdx.x; i < len; i += blockDim.x) s += v[i] * v[i];
    double nv = sqrt(block_sum(s, red));
    double t0 = t[0];
    __syncthreads();
    double nv_after, t_after;
    if (nv <= -t0) {
        for (int i = threadIdx.x; i < len; i += blockDim.x) v[i] = 0.0;
        if (threadIdx.x == 0) t[0] = 0.0;
        nv_after = 0.0; t_after = 0.0;
    } else if (nv <= t0) {
        nv_after = nv; t_after = t0;
    } else {
        double val = 0.5 * (1.0 + t0 / nv);
        for (int i = threadIdx.x; i < len; i += blockDim.x) v[i] *= val;
        if (threadIdx.x == 0) t[0] = val * nv;
        // the reference re-evaluates norm(v) - s on the projected data; recompute exactly
        __syncthreads();
        double s2 = 0.0;
        for (int i = threadIdx.x; i < len; i += blockDim.x) s2 += v[i] * v[i];
        nv_after = sqrt(block_sum(s2, red));
        t_after = val * nv;
    }
    if (threadIdx.x == 0) soc_gap[k] = nv_after - t_after;
}

// ---------------------------------------------------------------------------
// K12: y = A x for a CSR matrix, GROUP lanes per row (GROUP in {1,2,4,8,16,32}).
// Rows flagged long (handled by k_spmv_long) are skipped through row_skip.
// ---------------------------------------------------------------------------
template <int GROUP>
__global__ void k_spmv_csr(int nrows, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                           const double* __restrict__ val, const double* __restrict__ x,
                           double* __restrict__ y, int long_threshold, const double* __restrict__ poison) {
    if (poison && *poison != 0.0) return;
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int row = (int)(gid / GROUP);
    int lane = (int)(gid % GROUP);
    bool active = row < nrows;
    int b = 0, e = 0;
    if (active) { b = rowptr[row]; e = rowptr[row + 1]; }
    if (e - b > long_threshold) { active = false; e = b; }
    double s = 0.0;
    for (int k = b + lane; k < e; k += GROUP) s += val[k] * x[colidx[k]];
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, GROUP);
    if (active && lane == 0) y[row] = s;
}

// one block per long row (deterministic block reduction)
__global__ void k_spmv_long(const int* __restrict__ long_rows, const int* __restrict__ rowptr,
                            const int* __restrict__ colidx, const double* __restrict__ val,
                            const double* __restrict__ x, double* __restrict__ y,
                            const double* __restrict__ poison) {
    __shared__ double red[40];
    if (poison && *poison != 0.0) return;
    int row = long_rows[blockIdx.x];
    int b = rowptr[row], e = rowptr[row + 1];
    double s = 0.0;
    for (int k = b + threadIdx.x; k < e; k += blockDim.x) s += val[k] * x[colidx[k]];
    s = block_sum(s, red);
    if (threadIdx.x == 0) y[row] = s;
}

// ---------------------------------------------------------------------------
// K14-K16: one linesearch trial of the dual update (also the fixed-step dual_step!).
//   tau_t   = tau0 * decay^trial (sequential products, as pdhg.jl:569 does)
//   theta   = tau_t / tau_old ; sigma = beta * tau_t
//   y_half  = y + sigma*((1+theta)*Mx - theta*Mx_old)
//   proj    = [ b ; min(y_half/sigma, h) ] ;  y_new = y_half - sigma*proj
// fixed-step variant (use_theta=0): y_half = y + sigma*(2 Mx - Mx_old), sigma given.
// Accumulates ||y_new - y||^2.  Skipped when an earlier trial was accepted.
// ---------------------------------------------------------------------------
struct DualArgs {
    const double* y; const double* Mx; const double* Mx_old; const double* b; const double* h;
    double* y_new;
    int p, m;
    double tau0, decay, tau_old, beta, sigma_fixed;
    int trial, use_theta;
};

__global__ void __launch_bounds__(256)
k_dual_trial(DualArgs a, double* __restrict__ scal, ReduceWs ws) {
    __shared__ double red[40];
    __shared__ int s_last;
    if (scal[S_POISON] != 0.0 || scal[S_LS_ACCEPTED] != 0.0) return;
    double tau = a.tau0;
    for (int t = 0; t < a.trial; ++t) tau = mul_rn(tau, a.decay);
    double sigma, w1, w2;
    if (a.use_theta) {
        double theta = tau / a.tau_old;
        sigma = mul_rn(a.beta, tau);
        w1 = add_rn(1.0, theta);
        w2 = theta;
    } else {
        sigma = a.sigma_fixed;
        w1 = 2.0; w2 = 1.0;
    }
    int R = a.p + a.m;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R; i += gridDim.x * blockDim.x) {
        double yh = add_rn(a.y[i], mul_rn(sigma, sub_rn(mul_rn(w1, a.Mx[i]), mul_rn(w2, a.Mx_old[i]))));
        double proj;
        if (i < a.p) proj = a.b[i];
        else proj = fmin(yh / sigma, a.h[i - a.p]);
        double yn = sub_rn(yh, mul_rn(sigma, proj));
        a.y_new[i] = yn;
        double d = sub_rn(yn, a.y[i]);
        acc += d * d;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) ws.partials[blockIdx.x] = acc;
    if (last_block_arrive(ws.counters + 0, &s_last)) {
        double s = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += __ldcg(ws.partials + b);
        s = block_sum(s, red);
        if (threadIdx.x == 0) { scal[S_YNORM2] = s; scal[S_TAU] = tau; }
    }
}

// long rows of M' (a variable that appears in thousands of constraints, e.g. the anchor block of the
// sensor-localisation SDP): one block per row, same gating as the line-search trial it belongs to
// (M' is stored as DCSR: compact row q is variable nz_rows[q], entries rowptr[q] .. rowptr[q + 1]; long_rows holds
//  compact indices)
__global__ void k_spmv_mt_long(const int* __restrict__ long_rows, const int* __restrict__ nz_rows, const int* __restrict__ rowptr,
                               const int* __restrict__ colidx, const double* __restrict__ val,
                               const double* __restrict__ y, double* __restrict__ out, const double* __restrict__ scal) {
    __shared__ double red[40];
    if (scal && (scal[S_POISON] != 0.0 || scal[S_LS_ACCEPTED] != 0.0)) return;
    const int q = long_rows[blockIdx.x];
    int b = rowptr[q], e = rowptr[q + 1];
    double s = 0.0;
    for (int k = b + threadIdx.x; k < e; k += blockDim.x) s += val[k] * y[colidx[k]];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[nz_rows[q]] = s;
}

// ---------------------------------------------------------------------------
// K13+K17: Mty_new = M' y_new (CSR of M' == CSC of M, one thread per row of M'),
// fused with ||Mty_new - Mty||^2 and the accept test
//     sqrt(beta) * tau * ||Mty_new - Mty|| <= delta * ||y_new - y||     (pdhg.jl:566)
// evaluated by the last block.
// ---------------------------------------------------------------------------
struct MtArgs {
    int N; const int* rowptr; const int* colidx; const double* val;   // DCSR: rowptr is compact (n_nz + 1 entries)
    int long_threshold;      // rows of M' longer than this were computed by k_spmv_mt_long into Mty_new already
    const int* nz_rows; int n_nz;   // non-empty rows of M' (DCSR): Mty is identically zero on all other rows, so only these are walked
    const double* y_new; const double* Mty; double* Mty_new;
    double beta, delta; int trial, do_test;
};

__global__ void __launch_bounds__(256)
k_spmv_mt_norm(MtArgs a, double* __restrict__ scal, ReduceWs ws) {
    __shared__ double red[40];
    __shared__ int s_last;
    if (scal[S_POISON] != 0.0 || scal[S_LS_ACCEPTED] != 0.0) return;
    double acc = 0.0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < a.n_nz; q += gridDim.x * blockDim.x) {
        const int i = a.nz_rows[q];
        int b = a.rowptr[q], e = a.rowptr[q + 1];
        double s = 0.0;
        if (e - b > a.long_threshold) {
            s = a.Mty_new[i];
        } else {
            for (int k = b; k < e; ++k) s += a.val[k] * a.y_new[a.colidx[k]];
            a.Mty_new[i] = s;
        }
        double d = sub_rn(s, a.Mty[i]);
        acc += d * d;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) ws.partials[blockIdx.x] = acc;
    if (last_block_arrive(ws.counters + 1, &s_last)) {
        double s = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += __ldcg(ws.partials + b);
        s = block_sum(s, red);
        if (threadIdx.x == 0) {
            scal[S_MTYNORM2] = s;
            scal[S_LS_EVALS] = (double)(a.trial + 1);
            double tau = scal[S_TAU];
            double lhs = mul_rn(mul_rn(sqrt(a.beta), tau), sqrt(s));
            double rhs = mul_rn(a.delta, sqrt(scal[S_YNORM2]));
            if (a.do_test != 2 && (!a.do_test || lhs <= rhs)) {      // do_test == 2: sharded run, k_ls_decide decides
                scal[S_LS_ACCEPTED] = 1.0;
                scal[S_LS_TRIAL] = (double)a.trial;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K13-K17, whole speculative ladder in two launches (single-GPU runs).
//
// The line search (pdhg.jl:543-571) tries tau_t = tau0 decay^t, t = 0, 1, ...; trial t is accepted when
// sqrt(beta) tau_t ||M'(y_t - y)|| <= delta ||y_t - y||.  Mean number of trials on Max-Cut: 2.2, so the first T <= 4
// trials are evaluated side by side:
//   k_ls_ladder  computes, for every trial of the ladder, ||y_t - y||^2 and ||M'y_t - M'y||^2 WITHOUT writing any
//                vector: y_t is an elementwise function of (y, Mx, Mx_old, b, h), so the M' product recomputes the few
//                entries of y_t it needs (M' is walked through its non-empty rows, DCSR).  The last block applies the
//                accept test to trial 0, 1, ... in order and records the first accepted one.
//   k_ls_apply   writes y_new and Mty_new for the recorded trial (or for the last one when none was accepted: the
//                reference keeps the last trial's vectors when its loop runs out, and the host carries on with
//                further single trials otherwise).
// Rows of M' longer than the threshold are summed by one CTA each (k_ls_long_rows) before k_ls_ladder.
// ---------------------------------------------------------------------------
constexpr int LS_MAXT = 4;

struct LadderArgs {
    const double* y; const double* Mx; const double* Mx_old; const double* b; const double* h;
    int p, m;
    double tau0, decay, tau_old, beta, sigma_fixed, delta;
    int use_theta;           // 1: line search (theta from the trial step); 0: fixed-step dual_step! (pdhg.jl:584-609)
    int ntrials;             // T: trials evaluated by this launch (1 .. LS_MAXT)
    int do_test;             // 0: accept trial 0 unconditionally (fixed step); 2: sharded run — only publish this rank's
                             // partial sums (partial_out), k_ls_decide_gathered applies the test to the sums over all ranks
    double* partial_out;     // [2 * LS_MAXT]: ||y_t - y||^2, ||M'y_t - M'y||^2 per trial (do_test == 2)
    // M' in DCSR
    const int* nz_rows; const int* nz_ptr; const int* colidx; const double* val; int n_nz;
    int long_threshold; const int* long_rows; int n_long;
    double* long_sums;       // [n_long][LS_MAXT]
    const double* Mty;       // current M'y
    double* y_new; double* Mty_new;
};

struct TrialCoef { double tau, sigma, w1, w2; };

__device__ __forceinline__ TrialCoef ls_trial_coef(const LadderArgs& a, int trial) {
    TrialCoef c;
    c.tau = a.tau0;
    for (int t = 0; t < trial; ++t) c.tau = mul_rn(c.tau, a.decay);      // sequential products, as pdhg.jl:569 does
    if (a.use_theta) {
        const double theta = c.tau / a.tau_old;
        c.sigma = mul_rn(a.beta, c.tau);
        c.w1 = add_rn(1.0, theta);
        c.w2 = theta;
    } else {
        c.sigma = a.sigma_fixed; c.w1 = 2.0; c.w2 = 1.0;
    }
    return c;
}
// y_new[i] of one trial (pdhg.jl:544-553 with box_projection!, prox_operators.jl:160-170)
__device__ __forceinline__ double ls_y_new(const LadderArgs& a, const TrialCoef& c, int i, double yi, double mx, double mxo) {
    const double yh = add_rn(yi, mul_rn(c.sigma, sub_rn(mul_rn(c.w1, mx), mul_rn(c.w2, mxo))));
    const double proj = (i < a.p) ? a.b[i] : fmin(yh / c.sigma, a.h[i - a.p]);
    return sub_rn(yh, mul_rn(c.sigma, proj));
}

// one CTA per long row of M': long_sums[row][t] = sum_k val[k] y_t[col[k]] for every trial of the ladder
__global__ void __launch_bounds__(512) k_ls_long_rows(LadderArgs a, const double* __restrict__ scal, int trial0) {
    __shared__ double red[40];
    if (scal[S_POISON] != 0.0) return;
    const int q = a.long_rows[blockIdx.x];
    const int kb = a.nz_ptr[q], ke = a.nz_ptr[q + 1];
    TrialCoef c[LS_MAXT];
#pragma unroll
    for (int t = 0; t < LS_MAXT; ++t) c[t] = ls_trial_coef(a, trial0 + min(t, a.ntrials - 1));
    double s[LS_MAXT] = {0.0, 0.0, 0.0, 0.0};
    for (int k = kb + threadIdx.x; k < ke; k += blockDim.x) {
        const int col = a.colidx[k];
        const double v = a.val[k], yi = a.y[col], mx = a.Mx[col], mxo = a.Mx_old[col];
#pragma unroll
        for (int t = 0; t < LS_MAXT; ++t) if (t < a.ntrials) s[t] = fma(v, ls_y_new(a, c[t], col, yi, mx, mxo), s[t]);
    }
#pragma unroll
    for (int t = 0; t < LS_MAXT; ++t) {
        const double r = block_sum(s[t], red);
        if (threadIdx.x == 0) a.long_sums[(size_t)blockIdx.x * LS_MAXT + t] = r;
    }
}

__global__ void __launch_bounds__(256) k_ls_ladder(LadderArgs a, double* __restrict__ scal, ReduceWs ws, int trial0) {
    __shared__ double red[40];
    __shared__ int s_last;
    if (scal[S_POISON] != 0.0 || scal[S_LS_ACCEPTED] != 0.0) return;
    TrialCoef c[LS_MAXT];
#pragma unroll
    for (int t = 0; t < LS_MAXT; ++t) c[t] = ls_trial_coef(a, trial0 + min(t, a.ntrials - 1));
    double yn[LS_MAXT] = {0.0, 0.0, 0.0, 0.0}, mn[LS_MAXT] = {0.0, 0.0, 0.0, 0.0};
    const int R = a.p + a.m;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = t0; i < R; i += stride) {
        const double yi = a.y[i], mx = a.Mx[i], mxo = a.Mx_old[i];
#pragma unroll
        for (int t = 0; t < LS_MAXT; ++t)
            if (t < a.ntrials) { const double d = sub_rn(ls_y_new(a, c[t], i, yi, mx, mxo), yi); yn[t] += d * d; }
    }
    for (int q = t0; q < a.n_nz; q += stride) {
        const int kb = a.nz_ptr[q], ke = a.nz_ptr[q + 1];
        double s[LS_MAXT] = {0.0, 0.0, 0.0, 0.0};
        if (ke - kb > a.long_threshold) {
            // compact index of this row among the long rows: long_rows is ascending
            int lo = 0, hi = a.n_long - 1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.long_rows[mid] < q) lo = mid + 1; else hi = mid; }
#pragma unroll
            for (int t = 0; t < LS_MAXT; ++t) s[t] = a.long_sums[(size_t)lo * LS_MAXT + t];
        } else {
            for (int k = kb; k < ke; ++k) {
                const int col = a.colidx[k];
                const double v = a.val[k], yi = a.y[col], mx = a.Mx[col], mxo = a.Mx_old[col];
#pragma unroll
                for (int t = 0; t < LS_MAXT; ++t) if (t < a.ntrials) s[t] += v * ls_y_new(a, c[t], col, yi, mx, mxo);
            }
        }
        const double old = a.Mty[a.nz_rows[q]];
#pragma unroll
        for (int t = 0; t < LS_MAXT; ++t) if (t < a.ntrials) { const double d = sub_rn(s[t], old); mn[t] += d * d; }
    }
    const int nb = gridDim.x;
#pragma unroll
    for (int t = 0; t < LS_MAXT; ++t) {
        const double u = block_sum(yn[t], red), w = block_sum(mn[t], red);
        if (threadIdx.x == 0) { ws.partials[(2 * t) * nb + blockIdx.x] = u; ws.partials[(2 * t + 1) * nb + blockIdx.x] = w; }
    }
    if (last_block_arrive(ws.counters + 0, &s_last)) {
        double tot[2 * LS_MAXT];
#pragma unroll
        for (int t = 0; t < 2 * LS_MAXT; ++t) {
            double v = 0.0;
            for (int b = threadIdx.x; b < nb; b += blockDim.x) v += __ldcg(ws.partials + t * nb + b);
            tot[t] = block_sum(v, red);
        }
        if (threadIdx.x == 0 && a.do_test == 2) {
            for (int t = 0; t < 2 * LS_MAXT; ++t) a.partial_out[t] = tot[t];
        } else if (threadIdx.x == 0) {
            int chosen = -1;
            for (int t = 0; t < a.ntrials && chosen < 0; ++t) {
                const double lhs = mul_rn(mul_rn(sqrt(a.beta), c[t].tau), sqrt(tot[2 * t + 1]));
                const double rhs = mul_rn(a.delta, sqrt(tot[2 * t]));
                if (!a.do_test || lhs <= rhs) chosen = t;
            }
            const int last = chosen >= 0 ? chosen : a.ntrials - 1;
            scal[S_YNORM2] = tot[2 * last]; scal[S_MTYNORM2] = tot[2 * last + 1];
            scal[S_TAU] = c[last].tau;
            scal[S_LS_EVALS] = (double)(trial0 + last + 1);
            scal[S_LS_TRIAL] = (double)(trial0 + last);
            if (chosen >= 0) scal[S_LS_ACCEPTED] = 1.0;
        }
    }
}

// sharded line search: the per-rank partial sums of every trial of the ladder arrive in ONE all-gather
// ([nranks][2 * LS_MAXT]); every rank adds them up in rank order and applies the accept test of pdhg.jl:566 to trial
// 0, 1, ... exactly as the last block of k_ls_ladder does on a single GPU, so all ranks record the same trial.
__global__ void k_ls_decide_gathered(LadderArgs a, const double* __restrict__ gathered, int nranks, double* __restrict__ scal, int trial0) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (scal[S_POISON] != 0.0 || scal[S_LS_ACCEPTED] != 0.0) return;
    int chosen = -1, last = a.ntrials - 1;
    double yn_l = 0.0, mn_l = 0.0, tau_l = 0.0;
    for (int t = 0; t < a.ntrials; ++t) {
        double yn = 0.0, mn = 0.0;
        for (int r = 0; r < nranks; ++r) { yn += gathered[(size_t)r * 2 * LS_MAXT + 2 * t]; mn += gathered[(size_t)r * 2 * LS_MAXT + 2 * t + 1]; }
        const TrialCoef c = ls_trial_coef(a, trial0 + t);
        const double lhs = mul_rn(mul_rn(sqrt(a.beta), c.tau), sqrt(mn));
        const double rhs = mul_rn(a.delta, sqrt(yn));
        const bool ok = lhs <= rhs;
        if (ok || t == last) { yn_l = yn; mn_l = mn; tau_l = c.tau; if (ok) chosen = t; last = t; break; }
    }
    scal[S_YNORM2] = yn_l; scal[S_MTYNORM2] = mn_l;
    scal[S_TAU] = tau_l;
    scal[S_LS_EVALS] = (double)(trial0 + last + 1);
    scal[S_LS_TRIAL] = (double)(trial0 + last);
    if (chosen >= 0) scal[S_LS_ACCEPTED] = 1.0;
}

// writes y_new / Mty_new of trial scal[S_LS_TRIAL] (set by k_ls_ladder of the same ladder)
__global__ void __launch_bounds__(256) k_ls_apply(LadderArgs a, const double* __restrict__ scal, int trial0) {
    if (scal[S_POISON] != 0.0) return;
    const int trial = (int)scal[S_LS_TRIAL];
    if (trial < trial0 || trial >= trial0 + a.ntrials) return;      // an earlier ladder was accepted
    const TrialCoef c = ls_trial_coef(a, trial);
    const int R = a.p + a.m;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = t0; i < R; i += stride) a.y_new[i] = ls_y_new(a, c, i, a.y[i], a.Mx[i], a.Mx_old[i]);
    for (int q = t0; q < a.n_nz; q += stride) {
        const int kb = a.nz_ptr[q], ke = a.nz_ptr[q + 1];
        double s = 0.0;
        if (ke - kb > a.long_threshold) {
            int lo = 0, hi = a.n_long - 1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.long_rows[mid] < q) lo = mid + 1; else hi = mid; }
            s = a.long_sums[(size_t)lo * LS_MAXT + (trial - trial0)];
        } else {
            for (int k = kb; k < ke; ++k) {
                const int col = a.colidx[k];
                s += a.val[k] * ls_y_new(a, c, col, a.y[col], a.Mx[col], a.Mx_old[col]);
            }
        }
        a.Mty_new[a.nz_rows[q]] = s;
    }
}

// ---------------------------------------------------------------------------
// K18 (primal half) + K21 (c'x):  one pass over x, x_old, Mty, Mty_old, c.
//   num = max |(x - tau*Mty) - (x_old - tau*Mty_old)| ; den = max |x_old - tau*Mty_old|
// tau is the accepted step (scal[S_TAU]).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4)      // 4 CTAs per SM = the whole grid in one resident wave (solver.cu: blocksN)
k_residual_primal(long long N, const double* __restrict__ x, const double* __restrict__ x_old,
                  const double* __restrict__ Mty, const double* __restrict__ Mty_old,
                  const double* __restrict__ c, double* __restrict__ scal, ReduceWs ws,
                  const unsigned int* __restrict__ mask = nullptr) {
    __shared__ double red[40];
    __shared__ int s_last;
    if (scal[S_POISON] != 0.0 || scal[S_LS_ACCEPTED] == 0.0) return;
    const double tau = scal[S_TAU];
    double num = 0.0, den = 0.0, obj = 0.0;
    long long stride = (long long)gridDim.x * blockDim.x;
    // 128-bit loads, two independent pairs per thread per trip: ten 16-byte loads in flight per thread
    const long long npair = N >> 1;
    const double2* x2 = reinterpret_cast<const double2*>(x);
    const double2* xo2 = reinterpret_cast<const double2*>(x_old);
    const double2* m2 = reinterpret_cast<const double2*>(Mty);
    const double2* mo2 = reinterpret_cast<const double2*>(Mty_old);
    const double2* c2 = reinterpret_cast<const double2*>(c);
    double obj1 = 0.0;
    auto one = [&](double xn, double xo, double mt, double mo, double cc, double& ob) {
        double pold = sub_rn(xo, mul_rn(tau, mo));
        double pnew = sub_rn(xn, mul_rn(tau, mt));
        num = nanmax(num, fabs(sub_rn(pnew, pold)));
        den = nanmax(den, fabs(pold));
        ob = fma(cc, xn, ob);
    };
    // pair i = entries 2i, 2i + 1: both bits sit in one mask word; where they are clear Mty, Mty_old and c are exact zeros
    auto live = [&](const long long i) -> bool { return mask == nullptr || ((mask[i >> 4] >> ((2 * i) & 31)) & 3u) != 0u; };
    const double2 z2 = make_double2(0.0, 0.0);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < npair; i += 2 * stride) {
        const bool l0 = live(i), l1 = live(i + stride);
        const double2 a0 = x2[i], b0 = xo2[i];
        const double2 a1 = x2[i + stride], b1 = xo2[i + stride];
        const double2 d0 = l0 ? m2[i] : z2, e0 = l0 ? mo2[i] : z2, f0 = l0 ? c2[i] : z2;
        const double2 d1 = l1 ? m2[i + stride] : z2, e1 = l1 ? mo2[i + stride] : z2, f1 = l1 ? c2[i + stride] : z2;
        one(a0.x, b0.x, d0.x, e0.x, f0.x, obj); one(a0.y, b0.y, d0.y, e0.y, f0.y, obj1);
        one(a1.x, b1.x, d1.x, e1.x, f1.x, obj); one(a1.y, b1.y, d1.y, e1.y, f1.y, obj1);
    }
    for (; i < npair; i += stride) {
        const bool l0 = live(i);
        const double2 a0 = x2[i], b0 = xo2[i];
        const double2 d0 = l0 ? m2[i] : z2, e0 = l0 ? mo2[i] : z2, f0 = l0 ? c2[i] : z2;
        one(a0.x, b0.x, d0.x, e0.x, f0.x, obj); one(a0.y, b0.y, d0.y, e0.y, f0.y, obj1);
    }
    if ((N & 1) && blockIdx.x == 0 && threadIdx.x == 0) one(x[N - 1], x_old[N - 1], Mty[N - 1], Mty_old[N - 1], c[N - 1], obj);
    obj += obj1;
    num = block_nanmax(num, red);
    den = block_nanmax(den, red);
    obj = block_sum(obj, red);
    int nb = gridDim.x;
    if (threadIdx.x == 0) {
        ws.partials[blockIdx.x] = num;
        ws.partials[nb + blockIdx.x] = den;
        ws.partials[2 * nb + blockIdx.x] = obj;
    }
    if (last_block_arrive(ws.counters + 2, &s_last)) {
        double a = 0.0, b = 0.0, o = 0.0;
        for (int k = threadIdx.x; k < nb; k += blockDim.x) {
            a = nanmax(a, __ldcg(ws.partials + k));
            b = nanmax(b, __ldcg(ws.partials + nb + k));
            o += __ldcg(ws.partials + 2 * nb + k);
        }
        a = block_nanmax(a, red);
        b = block_nanmax(b, red);
        o = block_sum(o, red);
        if (threadIdx.x == 0) { scal[S_RES_P_NUM] = a; scal[S_RES_P_DEN] = b; scal[S_PRIM_OBJ] = o; }
    }
}

// K18 (dual half) + K20 + K21: one pass over the R rows.
__global__ void __launch_bounds__(256)
k_residual_dual(int p, int m, double beta, int use_beta, double sigma_fixed,
                const double* __restrict__ y, const double* __restrict__ y_old,
                const double* __restrict__ Mx, const double* __restrict__ Mx_old,
                const double* __restrict__ b, const double* __restrict__ h,
                double* __restrict__ scal, ReduceWs ws) {
    __shared__ double red[40];
    __shared__ int s_last;
    if (scal[S_POISON] != 0.0 || scal[S_LS_ACCEPTED] == 0.0) return;
    const double sigma = use_beta ? mul_rn(beta, scal[S_TAU]) : sigma_fixed;   // p.dual_step (pdhg.jl:579)
    int R = p + m;
    double num = 0.0, den = 0.0, eq = 0.0, in = 0.0, by = 0.0, hy = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R; i += gridDim.x * blockDim.x) {
        double pold = sub_rn(y_old[i], mul_rn(sigma, Mx_old[i]));
        double pnew = sub_rn(y[i], mul_rn(sigma, Mx[i]));
        num = nanmax(num, fabs(sub_rn(pnew, pold)));
        den = nanmax(den, fabs(pold));
        if (i < p) {
            eq = fmax(eq, fabs(Mx[i] - b[i]));
            by += b[i] * y[i];
        } else {
            in = fmax(in, Mx[i] - h[i - p]);
            hy += h[i - p] * y[i];
        }
    }
    double v[6] = {num, den, eq, in, by, hy};
    v[0] = block_nanmax(v[0], red);
    v[1] = block_nanmax(v[1], red);
    v[2] = block_nanmax(v[2], red);
    v[3] = block_nanmax(v[3], red);
    v[4] = block_sum(v[4], red);
    v[5] = block_sum(v[5], red);
    int nb = gridDim.x;
    if (threadIdx.x == 0)
        for (int q = 0; q < 6; ++q) ws.partials[q * nb + blockIdx.x] = v[q];
    if (last_block_arrive(ws.counters + 3, &s_last)) {
        double r[6] = {0, 0, 0, 0, 0, 0};
        for (int k = threadIdx.x; k < nb; k += blockDim.x) {
            r[0] = nanmax(r[0], __ldcg(ws.partials + 0 * nb + k));
            r[1] = nanmax(r[1], __ldcg(ws.partials + 1 * nb + k));
            r[2] = nanmax(r[2], __ldcg(ws.partials + 2 * nb + k));
            r[3] = nanmax(r[3], __ldcg(ws.partials + 3 * nb + k));
            r[4] += __ldcg(ws.partials + 4 * nb + k);
            r[5] += __ldcg(ws.partials + 5 * nb + k);
        }
        r[0] = block_nanmax(r[0], red);
        r[1] = block_nanmax(r[1], red);
        r[2] = block_nanmax(r[2], red);
        r[3] = block_nanmax(r[3], red);
        r[4] = block_sum(r[4], red);
        r[5] = block_sum(r[5], red);
        if (threadIdx.x == 0) {
            scal[S_RES_D_NUM] = r[0]; scal[S_RES_D_DEN] = r[1];
            scal[S_EQ_MAX] = r[2]; scal[S_IN_MAX] = r[3];
            scal[S_BY] = r[4]; scal[S_HY] = r[5];
        }
    }
}

// reset the per-iteration scalar record (keeps nothing from the previous iteration)
__global__ void k_scal_reset(double* __restrict__ scal, int n, double soc_gap_init, double elapsed) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) scal[i] = (i == S_SOC_GAP) ? soc_gap_init : (i == S_ELAPSED ? elapsed : 0.0);
}

// End of an iteration on a single GPU: hand the scalar record to the host through mapped page-locked memory (the host
// spins on the sequence word: no cudaMemcpy, no stream synchronisation) and clear it for the next iteration.
__global__ void k_publish_record(double* __restrict__ scal, int n, int reset_mode, double* __restrict__ host_rec,
                                 volatile unsigned long long* __restrict__ host_seq, unsigned long long seq) {
    // reset_mode 0: publish only (the host is in the middle of an iteration: more trials / a fallback follow);
    // 1 / 2: the iteration is complete unless the record says otherwise (poisoned eigsolve; 1: line search without an
    // accepted trial) — exactly the conditions under which the host keeps working on this record
    __shared__ int s_reset;
    if (threadIdx.x == 0)
        s_reset = reset_mode != 0 && scal[S_POISON] == 0.0 && (reset_mode == 2 || scal[S_LS_ACCEPTED] != 0.0);
    __syncthreads();
    const bool reset = s_reset != 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { host_rec[i] = scal[i]; if (reset && i < S_HEADER) scal[i] = 0.0; }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) { *host_seq = seq; __threadfence_system(); }
}

// ---------------------------------------------------------------------------
// sharded runs: combine the per-rank iteration records (gathered in rank order) in place.
// sums: objective dot products, mat-vec count; maxima (NaN-propagating for the residual norms): the rest.
// The line-search slots are already identical on every rank (k_ls_decide).
// ---------------------------------------------------------------------------
__global__ void k_fold_header(const double* __restrict__ gathered, int nranks, double* __restrict__ scal) {
    int slot = threadIdx.x;
    if (slot >= S_HEADER) return;
    double v = gathered[slot];
    for (int r = 1; r < nranks; ++r) {
        double o = gathered[(size_t)r * S_HEADER + slot];
        switch (slot) {
            case S_PRIM_OBJ: case S_BY: case S_HY: case S_NUMOPS: v += o; break;
            case S_RES_P_NUM: case S_RES_P_DEN: case S_RES_D_NUM: case S_RES_D_DEN: v = nanmax(v, o); break;
            case S_EQ_MAX: case S_IN_MAX: case S_SOC_GAP: case S_POISON: case S_ELAPSED: v = fmax(v, o); break;
            default: break;     // line-search slots: rank 0's copy (identical everywhere)
        }
    }
    scal[slot] = v;
}

// sharded line search: sums of the two squared norms over ranks (rank order), then the accept test of
// pdhg.jl:566 exactly as k_spmv_mt_norm evaluates it in the single-GPU case.  gathered: [nranks][2].
__global__ void k_ls_decide(const double* __restrict__ gathered, int nranks, double* __restrict__ scal,
                            double beta, double delta, int trial) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (scal[S_LS_ACCEPTED] != 0.0) return;
    double yn = 0.0, mn = 0.0;
    for (int r = 0; r < nranks; ++r) { yn += gathered[2 * r]; mn += gathered[2 * r + 1]; }
    scal[S_YNORM2] = yn; scal[S_MTYNORM2] = mn;
    scal[S_LS_EVALS] = (double)(trial + 1);
    double tau = scal[S_TAU];
    double lhs = mul_rn(mul_rn(sqrt(beta), tau), sqrt(mn));
    double rhs = mul_rn(delta, sqrt(yn));
    if (lhs <= rhs) { scal[S_LS_ACCEPTED] = 1.0; scal[S_LS_TRIAL] = (double)trial; }
}

// max over SOC cones of the stored gaps -> scal[S_SOC_GAP]  (single small block)
__global__ void k_soc_gap_max(const double* __restrict__ soc_gap, int n_soc, double* __restrict__ scal) {
    __shared__ double red[40];
    if (scal[S_POISON] != 0.0) return;
    double v = -1.0e300;
    for (int i = threadIdx.x; i < n_soc; i += blockDim.x) v = fmax(v, soc_gap[i]);
    // block max via nanmax tree on shifted values is unnecessary here: plain max
    v = warp_max(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < ((int)blockDim.x + 31) / 32) ? red[lane] : -1.0e300;
        t = warp_max(t);
        if (lane == 0) scal[S_SOC_GAP] = t;
    }
}

}  // namespace pb
