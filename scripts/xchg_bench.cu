// xchg_bench.cu — micro-benchmark of grid-wide "all-reduce of a few doubles" mechanisms on one GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/xchg_bench scripts/xchg_bench.cu
// Each variant: G CTAs x 512 threads (cooperative), ROUNDS rounds; every round each CTA contributes
// NS doubles and every CTA needs the G-way sum of each.  Reports ns per round (CTA 0 clock).
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

template <int SCOPE>   // 0 = volatile (sys), 1 = relaxed.gpu
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned flag) {
    unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    if (SCOPE == 0) asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
    else asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
}
template <int SCOPE>
__device__ __forceinline__ uint4 ll_peek(const uint4* p) {
    uint4 r;
    if (SCOPE == 0) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ double ll_value(const uint4& r) { return __hiloint2double((int)r.z, (int)r.x); }

template <int SCOPE>
__device__ __forceinline__ double ll_reduce_slot(const uint4* slot, int G, unsigned flag, int lane) {
    double s = 0.0;
    for (int c0 = 0; c0 < G; c0 += 160) {
        uint4 r[5];
        bool ok;
        do {
            ok = true;
#pragma unroll
            for (int u = 0; u < 5; ++u) { int i = c0 + lane + 32 * u; if (i < G) r[u] = ll_peek<SCOPE>(slot + i); }
#pragma unroll
            for (int u = 0; u < 5; ++u) { int i = c0 + lane + 32 * u; if (i < G) ok = ok && r[u].y == flag && r[u].w == flag; }
        } while (!ok);
#pragma unroll
        for (int u = 0; u < 5; ++u) { int i = c0 + lane + 32 * u; if (i < G) s += ll_value(r[u]); }
    }
    return warp_sum(s);
}

// ---- variant A: LL all-to-all ----
template <int SCOPE>
__global__ void __launch_bounds__(512, 1) k_ll(uint4* xbuf, int NS, int rounds, double* out, long long* tns) {
    __shared__ double hred[128];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, G = gridDim.x, cta = blockIdx.x;
    unsigned epoch = 0;
    double acc = 0.0;
    long long t0 = gtimer();
    for (int r = 0; r < rounds; ++r) {
        ++epoch;
        uint4* xb = xbuf + (size_t)(epoch & 1u) * NS * G;
        if (tid < NS) ll_store<SCOPE>(xb + (size_t)tid * G + cta, (double)(cta + tid + r), epoch);
        for (int q = warp; q < NS; q += 16) {
            double s = ll_reduce_slot<SCOPE>(xb + (size_t)q * G, G, epoch, lane);
            if (lane == 0) hred[q] = s;
        }
        __syncthreads();
        acc += hred[r % NS];
        __syncthreads();
    }
    long long t1 = gtimer();
    if (tid == 0) { out[cta] = acc; if (cta == 0) *tns = t1 - t0; }
}

// ---- variant B: counter barrier (red.release + ld.acquire poll by one thread) + plain partial arrays ----
__global__ void __launch_bounds__(512, 1) k_bar(double* part, unsigned* bar, int NS, int rounds, double* out, long long* tns) {
    __shared__ double hred[128];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, G = gridDim.x, cta = blockIdx.x;
    double acc = 0.0;
    long long t0 = gtimer();
    for (int r = 0; r < rounds; ++r) {
        double* pb = part + (size_t)(r & 1) * NS * G;
        if (tid < NS) __stcg(pb + (size_t)tid * G + cta, (double)(cta + tid + r));
        __syncthreads();
        if (tid == 0) {
            unsigned target = (unsigned)(r + 1) * (unsigned)G;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            unsigned v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
        }
        __syncthreads();
        for (int q = warp; q < NS; q += 16) {
            double s = 0.0;
            for (int c = lane; c < G; c += 32) s += __ldcg(pb + (size_t)q * G + c);
            s = warp_sum(s);
            if (lane == 0) hred[q] = s;
        }
        __syncthreads();
        acc += hred[r % NS];
        __syncthreads();
    }
    long long t1 = gtimer();
    if (tid == 0) { out[cta] = acc; if (cta == 0) *tns = t1 - t0; }
}

// ---- variant C: cooperative groups grid.sync + plain partial arrays ----
__global__ void __launch_bounds__(512, 1) k_cg(double* part, int NS, int rounds, double* out, long long* tns) {
    __shared__ double hred[128];
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, G = gridDim.x, cta = blockIdx.x;
    double acc = 0.0;
    long long t0 = gtimer();
    for (int r = 0; r < rounds; ++r) {
        double* pb = part + (size_t)(r & 1) * NS * G;
        if (tid < NS) __stcg(pb + (size_t)tid * G + cta, (double)(cta + tid + r));
        grid.sync();
        for (int q = warp; q < NS; q += 16) {
            double s = 0.0;
            for (int c = lane; c < G; c += 32) s += __ldcg(pb + (size_t)q * G + c);
            s = warp_sum(s);
            if (lane == 0) hred[q] = s;
        }
        __syncthreads();
        acc += hred[r % NS];
        __syncthreads();
    }
    long long t1 = gtimer();
    if (tid == 0) { out[cta] = acc; if (cta == 0) *tns = t1 - t0; }
}

// ---- variant D: atomic accumulate (red.add.f64 into NS slots) + counter; nondeterministic order ----
__global__ void __launch_bounds__(512, 1) k_red(double* slots, unsigned* bar, int NS, int rounds, double* out, long long* tns) {
    __shared__ double hred[128];
    const int tid = threadIdx.x, G = gridDim.x, cta = blockIdx.x;
    double acc = 0.0;
    long long t0 = gtimer();
    for (int r = 0; r < rounds; ++r) {
        double* sb = slots + (size_t)(r % 3) * 128;
        if (tid < NS) atomicAdd(sb + tid, (double)(cta + tid + r));
        if (tid < NS && cta == 0) slots[(size_t)((r + 1) % 3) * 128 + tid] = 0.0;   // clear the buffer used next round (unused this round and last)
        __syncthreads();
        if (tid == 0) {
            unsigned target = (unsigned)(r + 1) * (unsigned)G;
            __threadfence();
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            unsigned v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
        }
        __syncthreads();
        if (tid < NS) hred[tid] = __ldcg(sb + tid);
        __syncthreads();
        acc += hred[r % NS];
        __syncthreads();
    }
    long long t1 = gtimer();
    if (tid == 0) { out[cta] = acc; if (cta == 0) *tns = t1 - t0; }
}

// ---- variant E: LL all-to-all where only ONE warp polls a per-CTA hint word first ----
__global__ void __launch_bounds__(512, 1) k_hint(uint4* xbuf, unsigned* hint, int NS, int rounds, double* out, long long* tns) {
    __shared__ double hred[128];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, G = gridDim.x, cta = blockIdx.x;
    unsigned epoch = 0;
    double acc = 0.0;
    long long t0 = gtimer();
    for (int r = 0; r < rounds; ++r) {
        ++epoch;
        uint4* xb = xbuf + (size_t)(epoch & 1u) * NS * G;
        if (tid < NS) ll_store<1>(xb + (size_t)tid * G + cta, (double)(cta + tid + r), epoch);
        if (tid == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(hint + cta), "r"(epoch) : "memory");
        if (warp == 0) {
            bool ok;
            do {
                ok = true;
                for (int c = lane; c < G; c += 32) {
                    unsigned v;
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(hint + c) : "memory");
                    ok = ok && (v == epoch);
                }
                ok = __all_sync(0xffffffffu, ok);
            } while (!ok);
        }
        __syncthreads();
        for (int q = warp; q < NS; q += 16) {
            double s = ll_reduce_slot<1>(xb + (size_t)q * G, G, epoch, lane);
            if (lane == 0) hred[q] = s;
        }
        __syncthreads();
        acc += hred[r % NS];
        __syncthreads();
    }
    long long t1 = gtimer();
    if (tid == 0) { out[cta] = acc; if (cta == 0) *tns = t1 - t0; }
}

int main(int argc, char** argv) {
    int NS = argc > 1 ? atoi(argv[1]) : 27, rounds = argc > 2 ? atoi(argv[2]) : 2000;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int G = prop.multiProcessorCount;
    uint4* xbuf; double* part; double* slots; unsigned* bar; unsigned* hint; double* out; long long* tns;
    CK(cudaMalloc(&xbuf, sizeof(uint4) * 2 * 128 * G)); CK(cudaMalloc(&part, sizeof(double) * 2 * 128 * G));
    CK(cudaMalloc(&slots, sizeof(double) * 3 * 128)); CK(cudaMalloc(&bar, 64)); CK(cudaMalloc(&hint, 4 * G));
    CK(cudaMalloc(&out, sizeof(double) * G)); CK(cudaMallocManaged(&tns, 8));
    auto reset = [&]() { CK(cudaMemset(xbuf, 0, sizeof(uint4) * 2 * 128 * G)); CK(cudaMemset(part, 0, sizeof(double) * 2 * 128 * G));
        CK(cudaMemset(slots, 0, sizeof(double) * 3 * 128)); CK(cudaMemset(bar, 0, 64)); CK(cudaMemset(hint, 0, 4 * G)); };
    for (int rep = 0; rep < 2; ++rep) {
        {   reset(); void* args[] = {&xbuf, &NS, &rounds, &out, &tns};
            CK(cudaLaunchCooperativeKernel((void*)k_ll<0>, dim3(G), dim3(512), args, 0, 0)); CK(cudaDeviceSynchronize());
            printf("LL all-to-all (volatile/sys)   NS=%d: %.0f ns/round\n", NS, (double)*tns / rounds); }
        {   reset(); void* args[] = {&xbuf, &NS, &rounds, &out, &tns};
            CK(cudaLaunchCooperativeKernel((void*)k_ll<1>, dim3(G), dim3(512), args, 0, 0)); CK(cudaDeviceSynchronize());
            printf("LL all-to-all (relaxed.gpu)    NS=%d: %.0f ns/round\n", NS, (double)*tns / rounds); }
        {   reset(); void* args[] = {&part, &bar, &NS, &rounds, &out, &tns};
            CK(cudaLaunchCooperativeKernel((void*)k_bar, dim3(G), dim3(512), args, 0, 0)); CK(cudaDeviceSynchronize());
            printf("counter barrier + ld.cg        NS=%d: %.0f ns/round\n", NS, (double)*tns / rounds); }
        {   reset(); void* args[] = {&part, &NS, &rounds, &out, &tns};
            CK(cudaLaunchCooperativeKernel((void*)k_cg, dim3(G), dim3(512), args, 0, 0)); CK(cudaDeviceSynchronize());
            printf("cg grid.sync + ld.cg           NS=%d: %.0f ns/round\n", NS, (double)*tns / rounds); }
        {   reset(); void* args[] = {&slots, &bar, &NS, &rounds, &out, &tns};
            CK(cudaLaunchCooperativeKernel((void*)k_red, dim3(G), dim3(512), args, 0, 0)); CK(cudaDeviceSynchronize());
            printf("red.add.f64 + counter          NS=%d: %.0f ns/round\n", NS, (double)*tns / rounds); }
        {   reset(); void* args[] = {&xbuf, &hint, &NS, &rounds, &out, &tns};
            CK(cudaLaunchCooperativeKernel((void*)k_hint, dim3(G), dim3(512), args, 0, 0)); CK(cudaDeviceSynchronize());
            printf("hint poll + LL                 NS=%d: %.0f ns/round\n", NS, (double)*tns / rounds); }
    }
    return 0;
}
