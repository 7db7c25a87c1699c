// symv_bench.cu — how fast can 120 CTAs x 512 threads pull an L2-resident 2000 x 2000 FP64 matrix through the SMs?
// Variants of the slab mat-vec of the Lanczos kernels, WITHOUT any cross-CTA exchange: each CTA repeats its slab
// REP times (block barrier in between), CTA 0 reports cycles per pass.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/symv_bench.bin scripts/symv_bench.cu
// (modes 0-3 measured in profiles/r1j_symv_bench.txt; mode 4, the bulk-async ring, was added after the round's GPU
//  budget was spent: it compiles for sm_100a but has not been run yet)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double2 ld_stream_d2(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ld_plain_d2(const double2* p) { return __ldg(p); }
constexpr int T = 512, NW = 16;
// ---- bulk-async (TMA engine, 1-D) staging helpers for mode 4 ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
constexpr int STAGES = 6;     // mode 4: ring of STAGES slab rows (16 KB each) in shared memory
// mode 0: strip layout (warp = 128 columns, all rows), batches of RB rows, per-lane accumulation only
// mode 1: same + per-row partials parked in shared memory and reduced by half-warps (the cl3 symv)
// mode 2: chunk layout (warp = contiguous run of 64-double chunks, 8 loads in flight, shuffle tree per row)
// mode 3: strip layout, plain ld.global (L1-allocating) instead of the streaming load
// mode 4: slab rows staged through a shared-memory ring by 1-D bulk-async copies (cp.async.bulk + mbarrier, one
//         elected thread issues them), consumed from shared memory by the strip layout + smem row reduction
template <int RB, int MODE>
__global__ void __launch_bounds__(T, 1) k(const double* X, int n, int ld, int rep, double* out, long long* cyc) {
    extern __shared__ double sm[];
    double* vbuf = sm;            // 2048
    double* part = sm + 2048;     // 18 * 512
    double* red2 = part + 18 * 512;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, G = gridDim.x, cta = blockIdx.x;
    const int rbase = n / G, rrem = n % G;
    const int r0 = cta * rbase + min(cta, rrem), rl = rbase + (cta < rrem);
    for (int i = tid; i < 2048; i += T) vbuf[i] = (i < n) ? 1.0 + 1e-3 * i : 0.0;
    __syncthreads();
    double sink = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < rep; ++it) {
        if (MODE == 4) {
            // ring[STAGES][2048] doubles after red2; full[s] (1 arrival + tx bytes), empty[s] (NW arrivals)
            double* ring = red2 + 512;
            unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + STAGES * 2048);
            unsigned long long* empty = full + STAGES;
            if (it == 0) {
                if (tid == 0) { for (int q = 0; q < STAGES; ++q) { mbar_init(&full[q], 1); mbar_init(&empty[q], NW); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
                __syncthreads();
            }
            const double2 v0 = *(const double2*)(vbuf + warp * 128 + 2 * lane), v1 = *(const double2*)(vbuf + warp * 128 + 64 + 2 * lane);
            const unsigned rowbytes = (unsigned)(2048 * sizeof(double));
            const long long base = (long long)it * rl;            // rows issued before this pass (ring position continues over passes)
            if (tid == 0) {                                        // prologue: fill the ring
                for (int r = 0; r < min(STAGES, rl); ++r) {
                    const long long g = base + r; const int st = (int)(g % STAGES);
                    if (g >= STAGES) mbar_wait(&empty[st], (unsigned)(((g / STAGES) - 1) & 1));
                    mbar_expect_tx(&full[st], rowbytes);
                    bulk_g2s(ring + st * 2048, X + (size_t)(r0 + r) * ld, rowbytes, &full[st]);
                }
            }
            for (int r = 0; r < rl; ++r) {
                const long long g = base + r; const int st = (int)(g % STAGES);
                mbar_wait(&full[st], (unsigned)((g / STAGES) & 1));
                const double* rp = ring + st * 2048 + warp * 128 + 2 * lane;
                const double2 x0 = *(const double2*)rp, x1 = *(const double2*)(rp + 64);
                double t = 0.0;
                t = fma(x0.x, v0.x, t); t = fma(x0.y, v0.y, t); t = fma(x1.x, v1.x, t); t = fma(x1.y, v1.y, t);
                part[r * T + tid] = t;
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
                if (tid == 0 && r + STAGES < rl) {                  // refill the slot just freed (by everybody) with row r + STAGES
                    mbar_wait(&empty[st], (unsigned)((g / STAGES) & 1));
                    mbar_expect_tx(&full[st], rowbytes);
                    bulk_g2s(ring + st * 2048, X + (size_t)(r0 + r + STAGES) * ld, rowbytes, &full[st]);
                }
            }
            __syncthreads();
            const int h = tid >> 4, l = tid & 15;
            if (h < rl) {
                const double* p = part + h * T;
                double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                for (int i = l; i < T; i += 64) { s0 += p[i]; s1 += p[i + 16]; s2 += p[i + 32]; s3 += p[i + 48]; }
                red2[h * 16 + l] = (s0 + s1) + (s2 + s3);
            }
            __syncwarp();
            if (h < rl && l == 0) { const double* q = red2 + h * 16; double sacc = 0; for (int i = 0; i < 16; ++i) sacc += q[i]; sink += sacc; }
        } else if (MODE == 2) {
            const int cpr = 32, nchunks = rl * cpr;
            const int g0 = (int)((long long)warp * nchunks / NW), g1 = (int)((long long)(warp + 1) * nchunks / NW);
            int g = g0;
            while (g < g1) {
                const int row = g / cpr, gend = min(g1, (row + 1) * cpr);
                const double* xr = X + (size_t)(r0 + row) * ld + 2 * lane;
                const double* vb = vbuf + 2 * lane;
                int cc = (g - row * cpr) * 64; const int ce = (gend - row * cpr) * 64;
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                for (; cc + 448 < ce; cc += 512) {
                    double2 x[8], w[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) x[u] = ld_stream_d2((const double2*)(xr + cc + 64 * u));
#pragma unroll
                    for (int u = 0; u < 8; ++u) w[u] = *(const double2*)(vb + cc + 64 * u);
#pragma unroll
                    for (int u = 0; u < 8; u += 4) {
                        a0 = fma(x[u].x, w[u].x, a0); a0 = fma(x[u].y, w[u].y, a0);
                        a1 = fma(x[u + 1].x, w[u + 1].x, a1); a1 = fma(x[u + 1].y, w[u + 1].y, a1);
                        a2 = fma(x[u + 2].x, w[u + 2].x, a2); a2 = fma(x[u + 2].y, w[u + 2].y, a2);
                        a3 = fma(x[u + 3].x, w[u + 3].x, a3); a3 = fma(x[u + 3].y, w[u + 3].y, a3);
                    }
                }
                for (; cc < ce; cc += 64) { double2 xv = ld_stream_d2((const double2*)(xr + cc)); double2 vv = *(const double2*)(vb + cc); a0 = fma(xv.x, vv.x, a0); a0 = fma(xv.y, vv.y, a0); }
                double acc = (a0 + a1) + (a2 + a3);
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) part[warp * 10 + (row - g0 / cpr)] = acc;
                g = gend;
            }
            __syncthreads();
            if (tid < rl) sink += part[tid % 160];
        } else {
            const double2 v0 = *(const double2*)(vbuf + warp * 128 + 2 * lane), v1 = *(const double2*)(vbuf + warp * 128 + 64 + 2 * lane);
            const double* gp = X + (size_t)r0 * ld + warp * 128 + 2 * lane;
            for (int row = 0; row < rl; row += RB) {
                double2 x[RB][2];
#pragma unroll
                for (int i = 0; i < RB; ++i) {
                    const int r = min(row + i, rl - 1);
                    if (MODE == 3) { x[i][0] = ld_plain_d2((const double2*)(gp + (size_t)r * ld)); x[i][1] = ld_plain_d2((const double2*)(gp + (size_t)r * ld + 64)); }
                    else { x[i][0] = ld_stream_d2((const double2*)(gp + (size_t)r * ld)); x[i][1] = ld_stream_d2((const double2*)(gp + (size_t)r * ld + 64)); }
                }
#pragma unroll
                for (int i = 0; i < RB; ++i) {
                    double t = 0.0;
                    t = fma(x[i][0].x, v0.x, t); t = fma(x[i][0].y, v0.y, t); t = fma(x[i][1].x, v1.x, t); t = fma(x[i][1].y, v1.y, t);
                    if (MODE == 1) { if (row + i < rl) part[(row + i) * T + tid] = t; } else sink += t;
                }
            }
            if (MODE == 1) {
                __syncthreads();
                const int h = tid >> 4, l = tid & 15;
                if (h < rl) {
                    const double* p = part + h * T;
                    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                    for (int i = l; i < T; i += 64) { s0 += p[i]; s1 += p[i + 16]; s2 += p[i + 32]; s3 += p[i + 48]; }
                    red2[h * 16 + l] = (s0 + s1) + (s2 + s3);
                }
                __syncwarp();
                if (h < rl && l == 0) { const double* q = red2 + h * 16; double s = 0; for (int i = 0; i < 16; ++i) s += q[i]; sink += s; }
            }
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (cta == 0 && tid == 0) cyc[0] = (t1 - t0) / rep;
    if (sink == 123.456) out[cta * T + tid] = sink;
}
template <int RB, int MODE>
void run(const char* name, const double* X, int n, int ld, double* out, long long* cyc, size_t pad = 0) {
    size_t smem = (2048 + 18 * 512 + 512) * sizeof(double) + pad;
    if (MODE == 4) smem += STAGES * 2048 * sizeof(double) + 2 * STAGES * sizeof(unsigned long long);
    cudaFuncSetAttribute(k<RB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int G : {120, 148}) {
        k<RB, MODE><<<G, T, smem>>>(X, n, ld, 5, out, cyc); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<RB, MODE><<<G, T, smem>>>(X, n, ld, 200, out, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s smem %3zu KB G=%3d  %7.2f us/pass (events)  CTA0 %6lld cycles/pass  => %5.2f TB/s  [%s]\n", name, smem >> 10, G, ms * 1e3 / 200, cyc[0], 8.0 * n * n / (ms * 1e-3 / 200) * 1e-12, cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    const int n = 2000, ld = 2048;
    double *X, *out; long long* cyc;
    cudaMalloc(&X, (size_t)ld * ld * 8); cudaMemset(X, 0, (size_t)ld * ld * 8); cudaMalloc(&out, 148 * 512 * 8); cudaMallocManaged(&cyc, 64);
    run<9, 0>("strip, RB=9, accumulate only", X, n, ld, out, cyc);
    run<17, 0>("strip, RB=17, accumulate only", X, n, ld, out, cyc);
    run<6, 0>("strip, RB=6, accumulate only", X, n, ld, out, cyc);
    run<9, 3>("strip, RB=9, plain ld.global (L1)", X, n, ld, out, cyc);
    run<9, 1>("strip, RB=9, smem row reduction (cl3)", X, n, ld, out, cyc);
    run<9, 2>("chunk runs + shuffle tree (cl2)", X, n, ld, out, cyc);
    run<9, 4>("bulk-async ring of 6 rows + smem reduction", X, n, ld, out, cyc);
    // the Lanczos kernels run with ~170-220 KB of shared memory per CTA, i.e. with a small L1: does that matter?
    run<9, 0>("strip, RB=9, accumulate only", X, n, ld, out, cyc, 80 << 10);
    run<9, 0>("strip, RB=9, accumulate only", X, n, ld, out, cyc, 130 << 10);
    run<9, 3>("strip, RB=9, plain ld.global (L1)", X, n, ld, out, cyc, 130 << 10);
    run<9, 1>("strip, RB=9, smem row reduction (cl3)", X, n, ld, out, cyc, 130 << 10);
    run<9, 2>("chunk runs + shuffle tree (cl2)", X, n, ld, out, cyc, 130 << 10);
    run<9, 4>("bulk-async ring of 6 rows + smem reduction", X, n, ld, out, cyc, 30 << 10);
    return 0;
}
