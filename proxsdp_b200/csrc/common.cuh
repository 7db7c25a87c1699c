// common.cuh — shared device/host helpers for the ProxSDP B200 hot path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

namespace pb {

// ---------------------------------------------------------------------------
// error handling: every CUDA failure becomes a C++ exception that the C-ABI
// layer converts into a negative return code + last-error string.
// ---------------------------------------------------------------------------
struct CudaError : public std::runtime_error {
    int code;
    CudaError(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

#define PB_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            char _buf[512];                                                                   \
            snprintf(_buf, sizeof(_buf), "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, \
                     cudaGetErrorString(_e));                                                 \
            throw pb::CudaError(-100 - (int)_e, _buf);                                        \
        }                                                                                     \
    } while (0)

constexpr int WARP = 32;

// ---------------------------------------------------------------------------
// warp / block reductions (deterministic: fixed shuffle tree, fixed warp order)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// NaN-propagating max for the infinity norms (Julia's norm(x, Inf) returns NaN if any NaN)
__device__ __forceinline__ double nanmax(double a, double b) {
    return (a != a) ? a : ((b != b) ? b : fmax(a, b));
}
__device__ __forceinline__ double warp_nanmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum; result valid in every thread.  scratch: >= 33 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < nw) ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}
__device__ __forceinline__ double block_nanmax(double v, double* scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_nanmax(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < nw) ? scratch[lane] : 0.0;
        t = warp_nanmax(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// plain block max with an explicit identity (no NaN propagation)
__device__ __forceinline__ double block_max_id(double v, double* scratch, double identity) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < nw) ? scratch[lane] : identity;
        t = warp_max(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// ---------------------------------------------------------------------------
// memory helpers
// ---------------------------------------------------------------------------
// streaming 128-bit read of read-only data that is used once per pass
__device__ __forceinline__ double2 ld_stream_d2(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
// loads that must observe other CTAs' writes made before a grid barrier (bypass L1)
__device__ __forceinline__ double ld_cg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ double2 ld_cg_d2(const double2* p) { return __ldcg(p); }

// "last block done" pattern: returns true in every thread of the block that
// arrives last on `counter`; that block may then fold the per-block partials in
// a fixed order (deterministic).  The counter is reset for the next launch.
__device__ __forceinline__ bool last_block_arrive(unsigned int* counter, int* s_flag) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        unsigned int prev = atomicAdd(counter, 1u);
        *s_flag = (prev == total - 1);
        if (*s_flag) *counter = 0;
    }
    __syncthreads();
    bool last = (*s_flag != 0);
    if (last) __threadfence();
    return last;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Is position pos (working order: [PSD svec blocks | SOC blocks | free]) an off-diagonal entry of a PSD block?
// svec order inside a block: k = j (j + 1) / 2 + i, i <= j (reference src/scaling.jl:28-58, prox_operators.jl:1-16).
__device__ __forceinline__ bool offdiag_position(long long pos, long long psd_end, const long long* __restrict__ cone_off, int n_sdp) {
    if (pos >= psd_end) return false;
    int lo = 0, hi = n_sdp - 1;               // last cone with off <= pos
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (cone_off[mid] <= pos) lo = mid; else hi = mid - 1; }
    const long long k = pos - cone_off[lo];
    long long j = (long long)((sqrt(8.0 * (double)k + 1.0) - 1.0) * 0.5);
    while ((j + 1) * (j + 2) / 2 <= k) ++j;
    while (j * (j + 1) / 2 > k) --j;
    return k != j * (j + 1) / 2 + j;
}

}  // namespace pb
