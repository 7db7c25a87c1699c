// jacobi_bench.cu — latency of the K x K Ritz eigen-solve variants (one CTA, 512 threads), B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I proxsdp_b200/csrc -o scripts/jacobi_bench.bin scripts/jacobi_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include "jacobi.cuh"
using namespace pb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(512, 1) k_bench(const double* T, int k, int variant, const double* W, double* Uout, double* dout, long long* cyc) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int lda = (k + 2) | 1;
    const int m = (k + 1) & ~1;
    double* A = reinterpret_cast<double*>(raw);
    double* B = A + lda * lda;
    double* U = B + lda * lda;
    JacobiScratch js = jacobi_carve(U + lda * lda, lda);
    for (int idx = threadIdx.x; idx < lda * lda; idx += blockDim.x) { A[idx] = 0.0; B[idx] = 0.0; }
    __syncthreads();
    for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x) { int r = idx % k, c = idx / k; A[r + c * lda] = T[r + c * k]; }
    __syncthreads();
    long long t0 = clock64();
    const double* D;
    if (variant == 0) { jacobi_eigh_smem(k, A, lda, U, lda, js); D = A; }
    else if (variant == 1) D = jacobi_eigh_smem_fast(m, A, B, lda, U, lda, js);
    else D = jacobi_eigh_smem_warm(m, k, A, B, lda, U, lda, W, js);
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = (long long)js.red[38]; }
    for (int idx = threadIdx.x; idx < lda * lda; idx += blockDim.x) Uout[idx] = U[idx];
    for (int i = threadIdx.x; i < k; i += blockDim.x) dout[i] = D[i + i * lda];
}

int main(int argc, char** argv) {
    int k = argc > 1 ? atoi(argv[1]) : 25;
    const int lda = (k + 2) | 1;
    std::vector<double> T(k * k, 0.0), T2(k * k, 0.0);
    srand(1);
    for (int i = 0; i < k; ++i) {
        T[i + i * k] = 10.0 * rand() / RAND_MAX - 5.0;
        if (i + 1 < k) { double e = 3.0 * rand() / RAND_MAX + 0.1; T[i + (i + 1) * k] = e; T[i + 1 + i * k] = e; }
    }
    T2 = T;
    for (int i = 0; i < k; ++i) { T2[i + i * k] += 1e-3 * (rand() / (double)RAND_MAX - 0.5); if (i + 1 < k) { double d = 1e-3 * (rand() / (double)RAND_MAX - 0.5); T2[i + (i + 1) * k] += d; T2[i + 1 + i * k] += d; } }
    double *dT, *dT2, *dU, *dW, *dd; long long* cyc;
    CK(cudaMalloc(&dT, 8 * k * k)); CK(cudaMalloc(&dT2, 8 * k * k)); CK(cudaMalloc(&dU, 8 * lda * lda)); CK(cudaMalloc(&dW, 8 * lda * lda)); CK(cudaMalloc(&dd, 8 * k));
    CK(cudaMallocManaged(&cyc, 16));
    CK(cudaMemcpy(dT, T.data(), 8 * k * k, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dT2, T2.data(), 8 * k * k, cudaMemcpyHostToDevice));
    size_t smem = 8 * (3 * lda * lda) + jacobi_scratch_bytes(lda) + 64;
    CK(cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<double> d0(k), d1(k);
    for (int rep = 0; rep < 2; ++rep) {
        for (int variant = 0; variant < 2; ++variant) {
            k_bench<<<1, 512, smem>>>(dT, k, variant, nullptr, dU, dd, cyc); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy((variant ? d1 : d0).data(), dd, 8 * k, cudaMemcpyDeviceToHost));
            printf("k=%d variant %d (%s): %lld cycles = %.1f us, rounds with rotations %lld\n", k, variant, variant ? "fast cold" : "reference jacobi", cyc[0], cyc[0] / 1965.0, cyc[1]);
        }
        double md = 0; std::vector<double> a = d0, b = d1; std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
        for (int i = 0; i < k; ++i) md = fmax(md, fabs(a[i] - b[i]));
        printf("   max eigenvalue difference fast vs reference: %.3e\n", md);
        // warm: eigenvectors of T (in dU from the fast cold run) as the start for the perturbed T2
        CK(cudaMemcpy(dW, dU, 8 * lda * lda, cudaMemcpyDeviceToDevice));
        k_bench<<<1, 512, smem>>>(dT2, k, 2, dW, dU, dd, cyc); CK(cudaDeviceSynchronize());
        printf("k=%d variant 2 (warm, |dT| ~ 1e-3): %lld cycles = %.1f us, rounds with rotations %lld\n", k, cyc[0], cyc[0] / 1965.0, cyc[1]);
        k_bench<<<1, 512, smem>>>(dT2, k, 1, nullptr, dU, dd, cyc); CK(cudaDeviceSynchronize());
        printf("k=%d variant 1 on the perturbed matrix (cold): %lld cycles = %.1f us, rounds %lld\n", k, cyc[0], cyc[0] / 1965.0, cyc[1]);
    }
    return 0;
}
