// symv_bench2.cu — the slab mat-vec of the Lanczos kernel with part of the slab RESIDENT in shared memory.
//
// 120 CTAs x 512 threads, 2000 x 2000 FP64 matrix (L2 resident), every pass ends with the same grid-wide exchange the
// real kernel has (each CTA publishes a flagged word, one warp per CTA polls all of them), so the CTAs hit L2 in one
// burst exactly as they do in k_lanczos_cl3.  Variants:
//   mode 0  the r1 symv: strip loads from global + [row][thread] partial table + half-warp sums (lanczos_cl3.cuh)
//   mode 1  RES rows of the slab staged ONCE into shared memory by cp.async.bulk (TMA engine, mbarrier completion),
//           the other rows streamed by strip loads; per-lane partials of all rows reduced by a transposing butterfly
//           (20 shuffles for 18 rows), 16 warp totals per row folded through a 2 KB table
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scripts/symv_bench2.bin scripts/symv_bench2.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int T = 512, NW = 16, NMAX = 18, CPW = 2;
__device__ __forceinline__ double2 ld_stream_d2(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned int flag) {
    unsigned int lo = (unsigned int)__double2loint(v), hi = (unsigned int)__double2hiint(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
}
__device__ __forceinline__ uint4 ll_peek(const uint4* p) {
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
// one step of the transposing butterfly: N values per lane -> ceil(N/2); lanes with the mask bit clear keep the lower
// half [0, H) of the rows, lanes with the bit set the upper half [H, N)
template <int N>
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
// grid-wide exchange at the end of a pass: publish one flagged word, warp NW-1 polls everybody's
__device__ __forceinline__ void exchange(uint4* flags, int G, int cta, unsigned tag, double val) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) ll_store(flags + (size_t)(tag & 1) * 256 + cta, val, tag);
    if (warp == NW - 1) {
        for (int cb = 0; cb < G; cb += 32) {
            const int c = cb + lane;
            bool got = c >= G;
            while (!__all_sync(0xffffffffu, got)) { if (!got) { uint4 r = ll_peek(flags + (size_t)(tag & 1) * 256 + c); got = (r.y == tag && r.w == tag); } }
        }
    }
    __syncthreads();
}

template <int MODE, int RES, int RB>
__global__ void __launch_bounds__(T, 1) k(const double* __restrict__ X, int n, int ld, int rep, uint4* wg, uint4* flags, unsigned tag0,
                                          double* out, long long* cyc) {
    extern __shared__ __align__(16) double sm[];
    double* vbuf = sm;                    // 2048
    double* wpart = sm + 2048;            // 32 * 16
    double* aprod = wpart + 512;          // 32
    double* big = aprod + 32;             // mode 0: part[18][512] + red2[512]; mode 1: xres[RES][2048]
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, G = gridDim.x, cta = blockIdx.x;
    const int rbase = n / G, rrem = n % G;
    const int r0 = cta * rbase + min(cta, rrem), rl = rbase + (cta < rrem);
    for (int i = tid; i < 2048; i += T) vbuf[i] = (i < n) ? 1.0 + 1e-3 * i : 0.0;
    if (MODE == 1) {
        if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncthreads();
        if (tid == 0) {
            const int nres = min(RES, rl);
            mbar_expect_tx(&bar, (unsigned)(nres * 2048 * sizeof(double)));
            for (int r = 0; r < nres; ++r) bulk_g2s(big + (size_t)r * 2048, X + (size_t)(r0 + r) * ld, (unsigned)(2048 * sizeof(double)), &bar);
        }
        mbar_wait(&bar, 0);
    }
    __syncthreads();
    double sink = 0.0;
    long long t_sym = 0, t_red = 0;
    long long t0 = clock64();
    for (int it = 0; it < rep; ++it) {
        const unsigned tag = tag0 + it + 1;
        long long ta = clock64();
        const double2 v0 = *(const double2*)(vbuf + warp * 128 + 2 * lane), v1 = *(const double2*)(vbuf + warp * 128 + 64 + 2 * lane);
        const double* gp = X + (size_t)r0 * ld + warp * 128 + 2 * lane;
        if (MODE == 0) {
            double* part = big; double* red2 = big + 18 * 512;
            double pa = 0.0;
            for (int row = 0; row < rl; row += RB) {
                double2 x[RB][2];
#pragma unroll
                for (int i = 0; i < RB; ++i) {
                    const int r = min(row + i, rl - 1);
                    x[i][0] = ld_stream_d2((const double2*)(gp + (size_t)r * ld)); x[i][1] = ld_stream_d2((const double2*)(gp + (size_t)r * ld + 64));
                }
#pragma unroll
                for (int i = 0; i < RB; ++i) {
                    double t = 0.0;
                    t = fma(x[i][0].x, v0.x, t); t = fma(x[i][0].y, v0.y, t); t = fma(x[i][1].x, v1.x, t); t = fma(x[i][1].y, v1.y, t);
                    if (row + i < rl) { part[(row + i) * T + tid] = t; pa = fma(t, vbuf[r0 + row + i], pa); }
                }
            }
            part[rl * T + tid] = pa;
            long long tb = clock64(); t_sym += tb - ta;
            __syncthreads();
            const int h = tid >> 4, l = tid & 15;
            if (h <= rl) {
                const double* p = part + h * T;
                double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
                for (int i = l; i < T; i += 128) { s0 += p[i]; s1 += p[i + 16]; s2 += p[i + 32]; s3 += p[i + 48]; s4 += p[i + 64]; s5 += p[i + 80]; s6 += p[i + 96]; s7 += p[i + 112]; }
                red2[h * 16 + l] = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
            }
            __syncwarp();
            double alpha = 0.0;
            if (h <= rl && l == 0) {
                const double* q = red2 + h * 16; double s = 0; for (int i = 0; i < 16; ++i) s += q[i];
                if (h < rl) ll_store(wg + (size_t)(tag & 1) * ld + r0 + h, s, tag); else alpha = s;
            }
            t_red += clock64() - tb;
            if (tid == rl * 16) aprod[0] = alpha;
            __syncthreads();
            exchange(flags, G, cta, tag, aprod[0]);
        } else {
            const double* xres = big;
            double acc[NMAX];
            constexpr int NS = NMAX - RES;          // streamed rows (at most)
            // first batch of streamed rows: loads in flight while the resident rows come from shared memory
            double2 x[RB][CPW];
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                const int r = min(RES + i, rl - 1);
                x[i][0] = ld_stream_d2((const double2*)(gp + (size_t)r * ld)); x[i][1] = ld_stream_d2((const double2*)(gp + (size_t)r * ld + 64));
            }
#pragma unroll
            for (int r = 0; r < RES; ++r) {
                const double* rp = xres + (size_t)r * 2048 + warp * 128 + 2 * lane;
                const double2 a = *(const double2*)rp, b = *(const double2*)(rp + 64);
                double t = 0.0;
                t = fma(a.x, v0.x, t); t = fma(a.y, v0.y, t); t = fma(b.x, v1.x, t); t = fma(b.y, v1.y, t);
                acc[r] = t;
            }
#pragma unroll
            for (int b0 = 0; b0 < NS; b0 += RB) {
                double2 xn[RB][CPW];
                if (b0 + RB < NS) {
#pragma unroll
                    for (int i = 0; i < RB; ++i) {
                        const int r = min(RES + b0 + RB + i, rl - 1);
                        xn[i][0] = ld_stream_d2((const double2*)(gp + (size_t)r * ld)); xn[i][1] = ld_stream_d2((const double2*)(gp + (size_t)r * ld + 64));
                    }
                }
#pragma unroll
                for (int i = 0; i < RB; ++i) {
                    if (b0 + i < NS) {
                        double t = 0.0;
                        t = fma(x[i][0].x, v0.x, t); t = fma(x[i][0].y, v0.y, t); t = fma(x[i][1].x, v1.x, t); t = fma(x[i][1].y, v1.y, t);
                        acc[RES + b0 + i] = t;
                    }
                }
                if (b0 + RB < NS) {
#pragma unroll
                    for (int i = 0; i < RB; ++i) { x[i][0] = xn[i][0]; x[i][1] = xn[i][1]; }
                }
            }
            long long tb = clock64(); t_sym += tb - ta;
            // transposing butterfly: 18 -> 9 -> 5 -> 3 -> 2 -> 1 values per lane
            int off = 0, cnt = rl;
            { const bool hi = lane & 16; bfly_step<18>(acc, hi, 16); off += hi ? 9 : 0; cnt = hi ? max(cnt - 9, 0) : min(cnt, 9); }
            { const bool hi = lane & 8;  bfly_step<9>(acc, hi, 8);   off += hi ? 5 : 0; cnt = hi ? max(cnt - 5, 0) : min(cnt, 5); }
            { const bool hi = lane & 4;  bfly_step<5>(acc, hi, 4);   off += hi ? 3 : 0; cnt = hi ? max(cnt - 3, 0) : min(cnt, 3); }
            { const bool hi = lane & 2;  bfly_step<3>(acc, hi, 2);   off += hi ? 2 : 0; cnt = hi ? max(cnt - 2, 0) : min(cnt, 2); }
            { const bool hi = lane & 1;  bfly_step<2>(acc, hi, 1);   off += hi ? 1 : 0; cnt = hi ? max(cnt - 1, 0) : min(cnt, 1); }
            if (cnt >= 1) wpart[off * 16 + warp] = acc[0];
            __syncthreads();
            if (tid < 32) {
                double w = 0.0;
                if (tid < rl) {
                    const double* q = wpart + tid * 16;
                    w = ((q[0] + q[4]) + (q[8] + q[12])) + ((q[1] + q[5]) + (q[9] + q[13])) + (((q[2] + q[6]) + (q[10] + q[14])) + ((q[3] + q[7]) + (q[11] + q[15])));
                    ll_store(wg + (size_t)(tag & 1) * ld + r0 + tid, w, tag);
                }
                aprod[tid] = (tid < rl) ? w * vbuf[r0 + tid] : 0.0;
                __syncwarp();
                if (tid == 0) {
                    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                    for (int i = 0; i < 32; i += 4) { s0 += aprod[i]; s1 += aprod[i + 1]; s2 += aprod[i + 2]; s3 += aprod[i + 3]; }
                    aprod[0] = (s0 + s1) + (s2 + s3);
                }
            }
            t_red += clock64() - tb;
            __syncthreads();
            exchange(flags, G, cta, tag, aprod[0]);
        }
    }
    long long t1 = clock64();
    if (cta == 0 && tid == 0) { cyc[0] = (t1 - t0) / rep; cyc[1] = t_sym / rep; cyc[2] = t_red / rep; }
    if (sink == 123.456) out[cta * T + tid] = sink;
}

__global__ void k_touch(double* X, size_t n) {      // rewrite every element: all lines of X become DIRTY in L2
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) X[i] = X[i] * 1.0;
}
static bool g_dirty = false;
template <int MODE, int RES, int RB>
void run(const char* name, const double* X, int n, int ld, uint4* wg, uint4* flags, double* out, long long* cyc, double* host_w) {
    static unsigned tag0 = 16;
    if (g_dirty) { k_touch<<<148 * 8, 256>>>(const_cast<double*>(X), (size_t)ld * ld); cudaDeviceSynchronize(); }
    size_t smem = (2048 + 512 + 32) * sizeof(double) + (MODE == 0 ? (size_t)(18 * 512 + 512) * sizeof(double) : (size_t)RES * 2048 * sizeof(double));
    cudaFuncSetAttribute(k<MODE, RES, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int G = 120;
    k<MODE, RES, RB><<<G, T, smem>>>(X, n, ld, 5, wg, flags, tag0, out, cyc); tag0 += 5;
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (g_dirty) { k_touch<<<148 * 8, 256>>>(const_cast<double*>(X), (size_t)ld * ld); }
    cudaEventRecord(e0); k<MODE, RES, RB><<<G, T, smem>>>(X, n, ld, 25, wg, flags, tag0, out, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    tag0 += 25;
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // check the published mat-vec result of the last pass against the host's
    uint4* h = new uint4[2 * ld];
    cudaMemcpy(h, wg, sizeof(uint4) * 2 * ld, cudaMemcpyDeviceToHost);
    const unsigned last = tag0; double err = 0.0;
    for (int r = 0; r < n; ++r) {
        const uint4 q = h[(size_t)(last & 1) * ld + r];
        unsigned long long bits = ((unsigned long long)q.z << 32) | q.x; double v; memcpy(&v, &bits, 8);
        if (q.y != last || q.w != last) { err = 1e300; break; }
        err = fmax(err, fabs(v - host_w[r]) / fmax(1.0, fabs(host_w[r])));
    }
    delete[] h;
    printf("%-58s smem %3zu KB  %6.2f us/pass (events)  CTA0 %5lld cyc/pass: symv %5lld reduce+publish %5lld  max rel err %.1e [%s]\n", name, smem >> 10,
           ms * 1e3 / 25, cyc[0], cyc[1], cyc[2], err, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int n = 2000, ld = 2048;
    double *X, *out; long long* cyc; uint4 *wg, *flags;
    cudaMalloc(&X, (size_t)ld * ld * 8); cudaMalloc(&out, 148 * 512 * 8); cudaMallocManaged(&cyc, 64);
    cudaMalloc(&wg, sizeof(uint4) * 2 * ld); cudaMemset(wg, 0, sizeof(uint4) * 2 * ld);
    cudaMalloc(&flags, sizeof(uint4) * 512); cudaMemset(flags, 0, sizeof(uint4) * 512);
    double* hX = new double[(size_t)ld * ld]; double* hw = new double[n];
    for (size_t i = 0; i < (size_t)ld * ld; ++i) hX[i] = 0.0;
    for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) hX[(size_t)r * ld + c] = ((r * 131 + c * 71) % 97 - 48) * 0.01;
    for (int r = 0; r < n; ++r) { double s = 0; for (int c = 0; c < n; ++c) s += hX[(size_t)r * ld + c] * (1.0 + 1e-3 * c); hw[r] = s; }
    cudaMemcpy(X, hX, (size_t)ld * ld * 8, cudaMemcpyHostToDevice);
    for (int dirty = 0; dirty < 2; ++dirty) {
    g_dirty = dirty != 0;
    printf("---- X %s before the timed launch (25 passes per launch, like one eigsolve)\n", dirty ? "REWRITTEN by a kernel (dirty in L2)" : "clean");
    run<0, 0, 9>("r1 symv: strip loads, [row][thread] table", X, n, ld, wg, flags, out, cyc, hw);
    run<1, 0, 9>("butterfly reduce, nothing resident (RB 9)", X, n, ld, wg, flags, out, cyc, hw);
    run<1, 0, 6>("butterfly reduce, nothing resident (RB 6)", X, n, ld, wg, flags, out, cyc, hw);
    run<1, 4, 7>("TMA-staged: 4 rows resident, 13 streamed (RB 7)", X, n, ld, wg, flags, out, cyc, hw);
    run<1, 6, 6>("TMA-staged: 6 rows resident, 11 streamed (RB 6)", X, n, ld, wg, flags, out, cyc, hw);
    run<1, 8, 5>("TMA-staged: 8 rows resident, 9 streamed (RB 5)", X, n, ld, wg, flags, out, cyc, hw);
    run<1, 9, 9>("TMA-staged: 9 rows resident, 8 streamed (RB 9: one batch)", X, n, ld, wg, flags, out, cyc, hw);
    run<1, 9, 4>("TMA-staged: 9 rows resident, 8 streamed (RB 4)", X, n, ld, wg, flags, out, cyc, hw);
    run<1, 10, 4>("TMA-staged: 10 rows resident, 7 streamed (RB 4)", X, n, ld, wg, flags, out, cyc, hw);
    run<1, 12, 3>("TMA-staged: 12 rows resident, 5 streamed (RB 3)", X, n, ld, wg, flags, out, cyc, hw);
    }
    return 0;
}
