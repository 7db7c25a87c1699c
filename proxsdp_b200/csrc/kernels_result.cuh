// kernels_result.cuh — result assembly on the device (reference src/pdhg.jl:678-787):
//
//   k_slack            slack = M_orig x - [b; h]                          (cache_solution, pdhg.jl:757-758)
//   k_dual_cone_rows   dual_cone += M_orig' y on the non-empty rows        (get_duals, pdhg.jl:701-710)
//   k_feas_ineq_tail   ineq_viol = -min(0, min y_in), zero_viol = max |dual_cone[free]|   (dual_feas, pdhg.jl:716-732)
//   k_feas_soc         SOC part of cone_feas                               (pdhg.jl:691-697)
//   k_feas_small_max   PSD part of cone_feas over the small cones (their minimum eigenvalues come from k_small_cone_proj)
//
// x and y never leave the GPU for these: the host downloads the finished vectors once.
#pragma once
#include "common.cuh"
#include "kernels_vec.cuh"

namespace pb {

enum FeasSlot { FS_INEQ = 0, FS_ZERO, FS_SOC, FS_SMALL_PSD, FS_COUNT };

// one warp per row of a CSR matrix: out[row] = sum val[k] x[colidx[k]] - rhs[row]  (rows of any length)
__global__ void __launch_bounds__(256)
k_slack(int nrows, const int* __restrict__ rowptr, const int* __restrict__ colidx, const double* __restrict__ val,
        const double* __restrict__ x, const double* __restrict__ rhs_eq, int p, const double* __restrict__ rhs_in,
        double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < nrows; row += warps) {
        const int b = rowptr[row], e = rowptr[row + 1];
        double s = 0.0;
        for (int k = b + lane; k < e; k += 32) s = fma(val[k], x[colidx[k]], s);
        s = warp_sum(s);
        if (lane == 0) out[row] = s - (row < p ? rhs_eq[row] : rhs_in[row - p]);
    }
}

// one warp per compact row q of M' (DCSR): dc[nz_rows[q]] += sum val[k] y[colidx[k]]
__global__ void __launch_bounds__(256)
k_dual_cone_rows(int n_nz, const int* __restrict__ nz_rows, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                 const double* __restrict__ val, const double* __restrict__ y, double* __restrict__ dc) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n_nz; q += warps) {
        const int b = rowptr[q], e = rowptr[q + 1];
        double s = 0.0;
        for (int k = b + lane; k < e; k += 32) s = fma(val[k], y[colidx[k]], s);
        s = warp_sum(s);
        if (lane == 0) dc[nz_rows[q]] += s;
    }
}

// feas[FS_INEQ] = -min(0, min y_in) ; feas[FS_ZERO] = max |dc[tail_begin .. n)|   (two-phase, deterministic)
__global__ void __launch_bounds__(256)
k_feas_ineq_tail(const double* __restrict__ y_in, long long m, const double* __restrict__ dc, long long tail_begin,
                 long long n, double* __restrict__ feas, ReduceWs ws) {
    __shared__ double red[40];
    __shared__ int s_last;
    double mn = 0.0, mx = 0.0;      // min(0, ...) and max(0, ...) fold the identities in
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = t0; i < m; i += stride) mn = fmin(mn, y_in[i]);
    for (long long i = tail_begin + t0; i < n; i += stride) mx = fmax(mx, fabs(dc[i]));
    mn = -block_max_id(-mn, red, 0.0);
    mx = block_max_id(mx, red, 0.0);
    const int nb = gridDim.x;
    if (threadIdx.x == 0) { ws.partials[blockIdx.x] = mn; ws.partials[nb + blockIdx.x] = mx; }
    if (last_block_arrive(ws.counters + 5, &s_last)) {
        double a = 0.0, b = 0.0;
        for (int k = threadIdx.x; k < nb; k += blockDim.x) { a = fmin(a, __ldcg(ws.partials + k)); b = fmax(b, __ldcg(ws.partials + nb + k)); }
        a = -block_max_id(-a, red, 0.0);
        b = block_max_id(b, red, 0.0);
        if (threadIdx.x == 0) { feas[FS_INEQ] = -a; feas[FS_ZERO] = b; }
    }
}

// one block per SOC cone: viol[k] = -min(0, t - ||v||)
__global__ void k_feas_soc(const double* __restrict__ dc, const long long* __restrict__ soc_off, const int* __restrict__ soc_len,
                           double* __restrict__ viol) {
    __shared__ double red[40];
    const int k = blockIdx.x;
    const double* t = dc + soc_off[k];
    const int len = soc_len[k] - 1;
    double s = 0.0;
    for (int i = threadIdx.x; i < len; i += blockDim.x) s += t[1 + i] * t[1 + i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) viol[k] = -fmin(0.0, t[0] - sqrt(s));
}

// feas[slot] = max(0, max over ids of (negate ? -min(0, v[id]) : v[id]))   (single block)
__global__ void k_feas_max(const double* __restrict__ v, const int* __restrict__ ids, int count, int negate_min,
                           double* __restrict__ feas, int slot) {
    __shared__ double red[40];
    double mx = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const double x = v[ids ? ids[i] : i];
        mx = fmax(mx, negate_min ? -fmin(0.0, x) : x);
    }
    mx = block_max_id(mx, red, 0.0);
    if (threadIdx.x == 0) feas[slot] = mx;
}

}  // namespace pb
