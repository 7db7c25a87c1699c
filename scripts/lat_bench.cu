// lat_bench.cu — single-warp latency probes on B200 for the instruction mixes on the Lanczos step's critical path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lat_bench scripts/lat_bench.cu ; run: /tmp/lat_bench
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* g, long long* out, int nbusy) {
    __shared__ double sm[4096];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    if (warp != 0) {           // other warps: optional FP64 background load
        double a = lane;
        for (int i = 0; i < nbusy; ++i) a = fma(a, 1.0000001, 1e-9);
        if (a == 123.456) g[1000 + threadIdx.x] = a;
        return;
    }
    long long t0, t1;
    double a = lane * 1e-3, b = 1.0000001;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) a = fma(a, b, 1e-9);
    t1 = clock64(); if (lane == 0) out[0] = t1 - t0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) a = a + b;
    t1 = clock64(); if (lane == 0) out[1] = t1 - t0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) a += __shfl_xor_sync(0xffffffffu, a, 1 << (i % 5));
    t1 = clock64(); if (lane == 0) out[2] = t1 - t0;
    t0 = clock64();
    int idx = lane;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { double v = sm[idx]; idx = ((int)v + idx * 7 + 1) & 4095; a += v; }
    t1 = clock64(); if (lane == 0) out[3] = t1 - t0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { a += sm[(lane * 16 + (i & 15)) & 4095]; }       // 32-way bank conflict (stride 128 B)
    t1 = clock64(); if (lane == 0) out[4] = t1 - t0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) { __stcg(g + lane + 32 * (i & 7), a); a += 1.0; }
    t1 = clock64(); if (lane == 0) out[5] = t1 - t0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) { g[lane + 32 * (i & 7)] = a; a += 1.0; }
    t1 = clock64(); if (lane == 0) out[6] = t1 - t0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) { a += __ldcg(g + ((lane + (int)a) & 255)); }      // dependent L2 loads
    t1 = clock64(); if (lane == 0) out[7] = t1 - t0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) { __stcg(g + 512 + lane, a); __threadfence(); a += 1.0; }
    t1 = clock64(); if (lane == 0) out[8] = t1 - t0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) { asm volatile("fence.acq_rel.gpu;" ::: "memory"); a += 1.0; }
    t1 = clock64(); if (lane == 0) out[9] = t1 - t0;
    unsigned int* gu = reinterpret_cast<unsigned int*>(g + 2048);
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) { if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gu) : "memory"); __syncwarp(); }
    t1 = clock64(); if (lane == 0) out[10] = t1 - t0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gu) : "memory"); a += v; }
    t1 = clock64(); if (lane == 0) out[11] = t1 - t0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gu) : "memory"); a += v; }
    t1 = clock64(); if (lane == 0) out[12] = t1 - t0;
    if (a == 123.456) g[lane] = a;
}
int main() {
    double* g; long long* out;
    cudaMalloc(&g, 1 << 20); cudaMemset(g, 0, 1 << 20); cudaMallocManaged(&out, 256);
    const char* nm[13] = {"dependent DFMA", "dependent DADD", "SHFL+DADD", "dependent LDS.64", "LDS.64 32-way conflict", "st.cg (issue)", "st (issue)",
                          "dependent ld.cg (L2)", "st.cg + threadfence", "fence.acq_rel.gpu alone", "red.release.gpu", "ld.acquire.gpu", "ld.relaxed.gpu"};
    const int cnt[13] = {256, 256, 256, 256, 256, 64, 64, 64, 16, 16, 16, 16, 16};
    for (int busy = 0; busy < 2; ++busy) {
        for (int rep = 0; rep < 2; ++rep) { k<<<1, 512>>>(g, out, busy ? 200000 : 0); cudaDeviceSynchronize(); }
        printf("other 15 warps %s:\n", busy ? "running FP64 FMAs" : "idle");
        for (int i = 0; i < 13; ++i) printf("  %-26s %8.1f cycles each\n", nm[i], (double)out[i] / cnt[i]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
