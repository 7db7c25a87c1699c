// ritz_bench.cu — cycles of ritz_top_bi (bisection + twisted vectors on the K x K Lanczos tridiagonal) in isolation,
// one 512-thread CTA like the eigsolve kernel's, on a tridiagonal produced by a host Lanczos run.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. -o /tmp/ritz_bench scripts/ritz_bench.cu
// Run:   /tmp/ritz_bench [K]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "proxsdp_b200/csrc/ritz_bi.cuh"

__global__ void __launch_bounds__(512, 1) k_bench(int k, const double* d_in, const double* e_in, int want, int reps, double* lam_out, double* U_out,
                                                  long long* prof, int* got_out) {
    extern __shared__ __align__(16) double sm[];
    const int lda = (k + 1) | 1;
    double* d = sm;
    double* e = d + 128;
    double* lam = e + 128;
    double* U = lam + 128;
    double* scratch = U + (size_t)lda * 128;
    for (int i = threadIdx.x; i < k; i += blockDim.x) { d[i] = d_in[i]; e[i] = (i < k - 1) ? e_in[i] : 0.0; }
    __shared__ long long sprof[32];
    if (threadIdx.x < 32) sprof[threadIdx.x] = 0;
    __syncthreads();
    pb::RitzBiScratch sc = pb::ritz_bi_carve(scratch, k);
    int got = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        got = pb::ritz_top_bi(k, d, e, want, lam, U, lda, sc, sprof);
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { prof[0] = t1 - t0; *got_out = got; }
    if (threadIdx.x < 32 && threadIdx.x > 0) prof[threadIdx.x] = sprof[threadIdx.x];
    for (int i = threadIdx.x; i < got; i += blockDim.x) lam_out[i] = lam[i];
    for (int i = threadIdx.x; i < got * k; i += blockDim.x) U_out[i] = U[(i / k) * lda + (i % k)];
}

static double rnd() { return (double)rand() / RAND_MAX - 0.5; }

int main(int argc, char** argv) {
    const int K = argc > 1 ? atoi(argv[1]) : 25;
    const int n = 300;
    srand(7);
    std::vector<double> A((size_t)n * n);
    for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double v = rnd(); A[(size_t)i * n + j] = v; A[(size_t)j * n + i] = v; }
    for (int s = 0; s < 3; ++s) {      // a few outliers, as the projected matrices of the PDHG loop have
        std::vector<double> u(n);
        double nn = 0; for (auto& x : u) { x = rnd(); nn += x * x; }
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) A[(size_t)i * n + j] += (30.0 - 8.0 * s) * u[i] * u[j] / nn;
    }
    std::vector<std::vector<double>> V;
    std::vector<double> d(K), e(K, 0.0), v(n), w(n);
    double nn = 0; for (auto& x : v) { x = rnd(); nn += x * x; }
    for (auto& x : v) x /= sqrt(nn);
    for (int j = 0; j < K; ++j) {
        V.push_back(v);
        for (int i = 0; i < n; ++i) { double s = 0; for (int c = 0; c < n; ++c) s += A[(size_t)i * n + c] * v[c]; w[i] = s; }
        double a = 0; for (int i = 0; i < n; ++i) a += w[i] * v[i];
        d[j] = a;
        for (int pass = 0; pass < 2; ++pass)
            for (auto& q : V) { double h = 0; for (int i = 0; i < n; ++i) h += q[i] * w[i]; for (int i = 0; i < n; ++i) w[i] -= h * q[i]; }
        double b = 0; for (int i = 0; i < n; ++i) b += w[i] * w[i];
        b = sqrt(b);
        e[j] = b;
        for (int i = 0; i < n; ++i) v[i] = w[i] / b;
    }
    double *dd, *de, *dl, *dU; long long* dp; int* dg;
    cudaMalloc(&dd, K * 8); cudaMalloc(&de, K * 8); cudaMalloc(&dl, 128 * 8); cudaMalloc(&dU, 128 * 128 * 8);
    cudaMallocManaged(&dp, 32 * 8); cudaMallocManaged(&dg, 4);
    cudaMemcpy(dd, d.data(), K * 8, cudaMemcpyHostToDevice); cudaMemcpy(de, e.data(), K * 8, cudaMemcpyHostToDevice);
    const size_t smem = (3 * 128 + (size_t)((K + 1) | 1) * 128 + pb::ritz_bi_scratch_doubles(K) + 64) * 8;
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int reps : {20, 1})
    for (int want : {3, 6, 10, 16, 20, 24}) {
        for (int rep = 0; rep < 1; ++rep) {
            k_bench<<<1, 512, smem>>>(K, dd, de, want, reps, dl, dU, dp, dg);
            cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
        }
        std::vector<double> lam(128), U(128 * 128);
        cudaMemcpy(lam.data(), dl, 128 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(U.data(), dU, 128 * 128 * 8, cudaMemcpyDeviceToHost);
        const int got = *dg;
        double worst = 0;
        for (int q = 0; q < got; ++q) {
            const double* u = U.data() + (size_t)q * K;
            for (int j = 0; j < K; ++j) {
                double t = (d[j] - lam[q]) * u[j];
                if (j > 0) t += e[j - 1] * u[j - 1];
                if (j < K - 1) t += e[j] * u[j + 1];
                worst = fmax(worst, fabs(t));
            }
        }
        printf("reps %2d K %d want %2d got %2d: %7lld cycles per call  (setup %6lld, values %6lld, vectors %6lld, checks %6lld; %4.1f rounds for the top value)  max residual %.2e  lam0 %.12f lam1 %.12f\n",
               reps, K, want, got, dp[0] / reps, dp[11] / reps, dp[12] / reps, dp[13] / reps, dp[14] / reps, (double)dp[18] / reps, worst, lam[0], lam[1]);
    }
    return 0;
}
