// runtime.cuh — device/pinned memory caches and the device-side problem ingest (declarations).
//
// Ingest = reference src/scaling.jl:2-58 (`preprocess!`: variable permutation [PSD blocks | SOC blocks | free],
// `norm_scaling`: sqrt(2)/2 on the off-diagonal svec columns of A, G and c), the construction of M = [A; G] and
// M' (reference src/pdhg.jl:95-142, structs.jl:153-157) and the norms of src/pdhg.jl:14-16,121-133 — all of it on
// the GPU: the caller's CSC arrays are uploaded as they are and nothing on the host is proportional to n.
#pragma once
#include <vector>

#include "../../include/proxsdp_b200_types.h"
#include "common.cuh"

namespace pb {

// host<->device traffic of the current solve (reported as h2d/d2h bytes by the bench)
extern thread_local long long g_h2d_bytes, g_d2h_bytes;

// Device memory comes from a per-process cache of cudaMalloc blocks (a freed block is kept and handed to the next
// request of the same size class: back-to-back solves of one process re-use their buffers instead of paying
// cudaMalloc / cudaFree — ~10 ms per chambolle_pock call at n = 2 001 000).  PROXSDP_B200_MALLOC=plain disables
// the cache, PROXSDP_B200_CACHE_MB bounds it (default 24576).
cudaError_t pb_malloc(void** p, size_t bytes);
void pb_free(void* p);
// pinned host memory, cached the same way (cudaMallocHost costs ~10 ms per call)
void* pb_host_alloc(size_t bytes);
void pb_host_free(void* p);
void pb_cache_trim();      // release everything that is cached and unused

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DBuf& operator=(DBuf&& o) noexcept {
        if (this != &o) { if (p) pb_free(p); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DBuf() { if (p) pb_free(p); }
    void release() { if (p) { pb_free(p); p = nullptr; } n = 0; }
    // uninitialised storage
    void alloc_raw(size_t count) {
        if (p) { pb_free(p); p = nullptr; }
        n = count;
        size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
        cudaError_t e = pb_malloc(reinterpret_cast<void**>(&p), bytes);
        if (e != cudaSuccess) throw CudaError(-4, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    }
    // zero-filled storage (memset on the legacy stream: ordered against every blocking stream)
    void alloc(size_t count) {
        alloc_raw(count);
        PB_CUDA(cudaMemsetAsync(p, 0, std::max<size_t>(count, 1) * sizeof(T), 0));
    }
    void upload(const std::vector<T>& h) { upload(h.data(), h.size()); }
    void upload(const T* h, size_t count) {
        alloc_raw(count);
        if (count) PB_CUDA(cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice));
        g_h2d_bytes += (long long)(count * sizeof(T));
    }
    std::vector<T> download() const {
        std::vector<T> h(n);
        if (n) PB_CUDA(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
        g_d2h_bytes += (long long)(n * sizeof(T));
        return h;
    }
};

// M (CSR, one entry per constraint row) and M' (DCSR: only the non-empty rows of M' = columns of M are stored)
struct CsrDev {
    int nrows = 0, ncols = 0, nnz = 0, group = 1, n_long = 0, long_threshold = 1 << 30;
    DBuf<int> rowptr;         // M: nrows + 1 ; M': n_nz + 1 (compact)
    DBuf<int> colidx;
    DBuf<int> long_rows;      // M: row ids ; M': compact indices q (row nz_rows[q])
    DBuf<int> nz_rows;        // M' only: the variable (position) of compact row q, ascending
    int n_nz = 0;
    DBuf<double> val;         // scaled (working) values
    DBuf<double> val_orig;    // the caller's values (slacks and duals are reported un-scaled, pdhg.jl:701-787)
};

struct IngestOut {
    bool identity = true;             // the permutation of preprocess! is the identity
    DBuf<int> ord;                    // position -> user variable   (empty when identity)
    DBuf<int> var_ordering;           // user variable -> position   (empty when identity)
    DBuf<double> c_orig;              // objective in position order, un-scaled
    CsrDev M, Mt;
    double fro2 = 0.0;                // ||M||_F^2 of the scaled matrix
    double norm_c2 = 0.0;             // ||c||^2 of the caller's objective
    long long launches = 0;
};

// cone table in position order (host, tiny): side and svec offset of every PSD cone, psd_end = first position
// after the PSD blocks, listed_end = first position after the SOC blocks
struct ConeTable {
    int n_sdp = 0;
    const int* side = nullptr;
    const long long* off = nullptr;
    long long psd_end = 0, listed_end = 0;
};

// Uploads the problem and builds everything above on `stream`; synchronises the stream once or twice (the host needs
// the two norms and the row counts).  cone_side_d / cone_off_d: device copies of the cone table (n_sdp entries).
void ingest_problem(const proxsdp_problem_t* prob, const ConeTable& cones, const int* cone_side_d,
                    const long long* cone_off_d, bool want_matrices, cudaStream_t stream, IngestOut& out);

// equilibrate! (reference src/equilibration.jl:1-71) and the scaling M <- E M D of src/pdhg.jl:64-93 on the device.
// Decides like the reference (option, min(M) / max(M) against equilibration_limit, equilibration_force); when the
// preconditioner is applied the working values of M and M' are rebuilt from the caller's values, fro2 = ||M||_F^2 of
// the new working matrix.  D is a multiple of the identity by construction (the reference re-sets the column scaling to
// its mean in every step), so it is kept as one number: d[0] on the device, d_host on the host.
struct EquilibrateOut {
    bool applied = false;
    DBuf<double> E;                   // R row scalings
    DBuf<double> d;                   // d[0] = the column scaling
    double d_host = 1.0;
    double fro2 = 0.0;
    long long launches = 0;
};
bool equilibrate_device(CsrDev& M, CsrDev& Mt, long long n, long long R, const proxsdp_options_t& opt,
                        const long long* cone_off_d, int n_sdp, long long psd_end, cudaStream_t stream, EquilibrateOut& out);
// dst[i] = E[i] * src[i] ; dst[i] = d[0] * src[i]   (src == dst allowed)
void launch_eq_mul(const double* src, const double* E, long long len, double* dst, cudaStream_t stream);
void launch_eq_mul_scalar(const double* src, const double* dptr, long long len, double* dst, cudaStream_t stream);

// out[i] = src[perm ? perm[i] : i]
void launch_gather(const double* src, const int* perm, long long n, double* out, cudaStream_t stream);

}  // namespace pb
