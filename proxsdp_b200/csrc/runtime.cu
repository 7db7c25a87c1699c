// runtime.cu — memory caches and the device-side problem ingest (see runtime.cuh).
//
// The only library calls in here are CUB's device-wide radix sort and prefix sum (setup only, once per solve:
// ordering the non-zeros of M = [A; G] by (column, row) and by (row, column)); every other kernel of the ingest and
// everything on the per-iteration path is hand-written.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "runtime.cuh"

namespace pb {

thread_local long long g_h2d_bytes = 0, g_d2h_bytes = 0;

// ---------------------------------------------------------------------------
// caches
// ---------------------------------------------------------------------------
namespace {

struct BlockCache {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;      // size -> block (per device for device memory)
    std::map<void*, size_t> live;                   // blocks handed out
    size_t cached_bytes = 0;
};

struct DevCaches {
    BlockCache dev[16];
    BlockCache host;
};
DevCaches& caches() { static DevCaches c; return c; }

bool cache_enabled() {
    static int state = -1;
    if (state < 0) {
        const char* e = getenv("PROXSDP_B200_MALLOC");
        state = (e && std::string(e) == "plain") ? 0 : 1;
    }
    return state == 1;
}
size_t cache_limit_bytes() {
    static size_t lim = 0;
    if (!lim) {
        const char* e = getenv("PROXSDP_B200_CACHE_MB");
        lim = (size_t)(e ? std::max(0L, atol(e)) : 24576L) << 20;
    }
    return lim;
}
// size classes: exact up to 1 MiB granularity above 1 MiB, 256 B below — a block is re-used for any request of its class
size_t size_class(size_t bytes) {
    if (bytes <= 256) return 256;
    if (bytes <= (1u << 20)) { size_t s = 256; while (s < bytes) s <<= 1; return s; }
    return (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
}

}  // namespace

cudaError_t pb_malloc(void** p, size_t bytes) {
    if (!cache_enabled()) return cudaMalloc(p, bytes);
    int dev = 0;
    cudaGetDevice(&dev);
    BlockCache& c = caches().dev[dev & 15];
    const size_t cls = size_class(bytes);
    {
        std::lock_guard<std::mutex> g(c.mu);
        auto it = c.free_blocks.find(cls);
        if (it != c.free_blocks.end()) {
            *p = it->second;
            c.free_blocks.erase(it);
            c.cached_bytes -= cls;
            c.live[*p] = cls;
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, cls);
    if (e != cudaSuccess) {
        // out of memory: drop the cache and retry once
        cudaGetLastError();
        pb_cache_trim();
        e = cudaMalloc(p, cls);
        if (e != cudaSuccess) return e;
    }
    std::lock_guard<std::mutex> g(c.mu);
    c.live[*p] = cls;
    return cudaSuccess;
}

void pb_free(void* p) {
    if (!p) return;
    if (!cache_enabled()) { cudaFree(p); return; }
    cudaPointerAttributes at{};
    int dev = 0;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess) dev = at.device; else { cudaGetLastError(); cudaGetDevice(&dev); }
    BlockCache& c = caches().dev[dev & 15];
    size_t cls = 0;
    {
        std::lock_guard<std::mutex> g(c.mu);
        auto it = c.live.find(p);
        if (it != c.live.end()) { cls = it->second; c.live.erase(it); }
        if (cls && c.cached_bytes + cls <= cache_limit_bytes()) {
            c.free_blocks.emplace(cls, p);
            c.cached_bytes += cls;
            return;
        }
    }
    cudaFree(p);
}

void* pb_host_alloc(size_t bytes) {
    BlockCache& c = caches().host;
    const size_t cls = size_class(bytes);
    {
        std::lock_guard<std::mutex> g(c.mu);
        auto it = c.free_blocks.find(cls);
        if (it != c.free_blocks.end()) {
            void* p = it->second;
            c.free_blocks.erase(it);
            c.cached_bytes -= cls;
            c.live[p] = cls;
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, cls);
    if (e != cudaSuccess) { cudaGetLastError(); throw CudaError(-4, std::string("cudaMallocHost failed: ") + cudaGetErrorString(e)); }
    std::lock_guard<std::mutex> g(c.mu);
    c.live[p] = cls;
    return p;
}

void pb_host_free(void* p) {
    if (!p) return;
    BlockCache& c = caches().host;
    size_t cls = 0;
    {
        std::lock_guard<std::mutex> g(c.mu);
        auto it = c.live.find(p);
        if (it != c.live.end()) { cls = it->second; c.live.erase(it); }
        if (cls && cache_enabled() && c.cached_bytes + cls <= ((size_t)1 << 30)) {
            c.free_blocks.emplace(cls, p);
            c.cached_bytes += cls;
            return;
        }
    }
    cudaFreeHost(p);
}

void pb_cache_trim() {
    for (BlockCache& c : caches().dev) {
        std::lock_guard<std::mutex> g(c.mu);
        for (auto& kv : c.free_blocks) cudaFree(kv.second);
        c.free_blocks.clear();
        c.cached_bytes = 0;
    }
    BlockCache& h = caches().host;
    std::lock_guard<std::mutex> g(h.mu);
    for (auto& kv : h.free_blocks) cudaFreeHost(kv.second);
    h.free_blocks.clear();
    h.cached_bytes = 0;
}

// ---------------------------------------------------------------------------
// ingest kernels
// ---------------------------------------------------------------------------
namespace {

enum IngestFlag { F_ERR_RANGE = 0, F_ERR_DUP, F_NONIDENT, F_ERR_ROW, F_COUNT };
enum IngestRec { R_FRO2 = 0, R_NORMC2, R_COUNT };
enum IngestCnt { C_NNZ_ROWS = 0, C_LONG_T, C_LONG_M, C_COUNT };

// preprocess! (scaling.jl:2-26), cone-listed part: position q <- variable idx[q]
__global__ void k_perm_mark(const int64_t* __restrict__ idx, long long listed, long long n, long long base,
                            int* __restrict__ used, int* __restrict__ ord, int* __restrict__ flags) {
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= listed) return;
    const long long v = idx[q] - base;
    if (v < 0 || v >= n) { flags[F_ERR_RANGE] = 1; return; }
    if (used) {
        if (atomicExch(used + v, 1) != 0) flags[F_ERR_DUP] = 1;
        ord[q] = (int)v;
    }
    if (v != q) flags[F_NONIDENT] = 1;
}
// identity check without the `used` table is not enough to detect duplicates, but a list with v == q everywhere has none

__global__ void k_unused_flag(const int* __restrict__ used, long long n, int* __restrict__ flag) {
    long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) flag[v] = used[v] ? 0 : 1;
}
// free variables in ascending order behind the listed ones (scaling.jl:17-23)
__global__ void k_fill_free(const int* __restrict__ used, const int* __restrict__ rank, long long n, long long listed,
                            int* __restrict__ ord) {
    long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n && !used[v]) ord[listed + rank[v]] = (int)v;
}
__global__ void k_invert_perm(const int* __restrict__ ord, long long n, int* __restrict__ inv) {
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) inv[ord[j]] = (int)j;
}
__global__ void k_gather(const double* __restrict__ src, const int* __restrict__ perm, long long n, double* __restrict__ out) {
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = src[perm ? perm[i] : i];
}

// one thread per stored entry of A (e < nnzA) or G: column by binary search in colptr, position through the
// permutation, scaling by sqrt(2)/2 on off-diagonal positions; emits both sort keys
struct EntryArgs {
    const int64_t* A_colptr; const int64_t* A_rowval; const double* A_nzval; long long nnzA;
    const int64_t* G_colptr; const int64_t* G_rowval; const double* G_nzval; long long nnzG;
    long long n, p, m, base, psd_end;
    const int* var_ordering;       // nullptr: identity
    const int* cone_side; const long long* cone_off; int n_sdp;
    unsigned long long* key_t; unsigned long long* key_m; unsigned int* id;
    double* val_s; double* val_o;
    int* flags;
};
__global__ void k_entries(EntryArgs a) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.nnzA + a.nnzG) return;
    const bool isA = e < a.nnzA;
    const long long q = isA ? e : e - a.nnzA;
    const int64_t* cp = isA ? a.A_colptr : a.G_colptr;
    // largest column j with cp[j] - base <= q
    long long lo = 0, hi = a.n - 1;
    while (lo < hi) { long long mid = (lo + hi + 1) >> 1; if (cp[mid] - a.base <= q) lo = mid; else hi = mid - 1; }
    const long long col = lo;
    long long row = (isA ? a.A_rowval[q] : a.G_rowval[q]) - a.base;
    const double v = isA ? a.A_nzval[q] : a.G_nzval[q];
    if (row < 0 || row >= (isA ? a.p : a.m)) { a.flags[F_ERR_ROW] = 1; row = 0; }
    if (!isA) row += a.p;
    const long long pos = a.var_ordering ? a.var_ordering[col] : col;
    const bool od = offdiag_position(pos, a.psd_end, a.cone_off, a.n_sdp);
    a.key_t[e] = ((unsigned long long)pos << 32) | (unsigned long long)(unsigned int)row;
    a.key_m[e] = ((unsigned long long)row << 32) | (unsigned long long)(unsigned int)pos;
    a.id[e] = (unsigned int)e;
    a.val_o[e] = v;
    a.val_s[e] = od ? __dmul_rn(v, 0.70710678118654752440 /* sqrt(2)/2 */) : v;
}

// sorted (key, id) -> colidx / val / val_orig, plus the "first entry of its row" flag for M'
__global__ void k_scatter_sorted(const unsigned long long* __restrict__ key, const unsigned int* __restrict__ id,
                                 const double* __restrict__ val_s, const double* __restrict__ val_o, long long nnz,
                                 int* __restrict__ colidx, double* __restrict__ val, double* __restrict__ val_orig,
                                 int* __restrict__ head) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const unsigned long long kk = key[k];
    colidx[k] = (int)(unsigned int)(kk & 0xffffffffULL);
    const unsigned int e = id[k];
    val[k] = val_s[e];
    val_orig[k] = val_o[e];
    if (head) head[k] = (k == 0 || (key[k - 1] >> 32) != (kk >> 32)) ? 1 : 0;
}

// compact rows of M': nz_rows[q] = position, nz_ptr[q] = first entry; count -> cnt[C_NNZ_ROWS]
__global__ void k_dcsr_rows(const unsigned long long* __restrict__ key, const int* __restrict__ head,
                            const int* __restrict__ scan, long long nnz, int* __restrict__ nz_rows,
                            int* __restrict__ nz_ptr, int* __restrict__ cnt) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    if (head[k]) { const int q = scan[k]; nz_rows[q] = (int)(key[k] >> 32); nz_ptr[q] = (int)k; }
    if (k == nnz - 1) { const int tot = scan[k] + head[k]; nz_ptr[tot] = (int)nnz; cnt[C_NNZ_ROWS] = tot; }
}
// rows of M' longer than the threshold (one CTA per row in the SpMV) — compact indices, ascending
__global__ void k_long_flag(const int* __restrict__ ptr, const int* __restrict__ count_dev, int count_host,
                            int threshold, int* __restrict__ flag) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int cnt = count_dev ? *count_dev : count_host;
    if (q < cnt) flag[q] = (ptr[q + 1] - ptr[q] > threshold) ? 1 : 0;
    else if (q < count_host) flag[q] = 0;
}
__global__ void k_compact(const int* __restrict__ flag, const int* __restrict__ scan, int count, int* __restrict__ out,
                          int* __restrict__ cnt_slot) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    if (flag[q]) out[scan[q]] = q;
    if (q == count - 1) *cnt_slot = scan[q] + flag[q];
}
// CSR row pointers of M from the (row, position)-sorted keys
__global__ void k_rowptr_from_keys(const unsigned long long* __restrict__ key, long long nnz, long long R, int* __restrict__ rowptr) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > R) return;
    long long lo = 0, hi = nnz;               // first k with row(key[k]) >= r
    while (lo < hi) { long long mid = (lo + hi) >> 1; if ((long long)(key[mid] >> 32) < r) lo = mid + 1; else hi = mid; }
    rowptr[r] = (int)lo;
}

// deterministic sum of squares: fixed grid, per-block partials folded by one block in block order
__global__ void __launch_bounds__(256) k_sumsq_partial(const double* __restrict__ v, long long n, double* __restrict__ partials) {
    __shared__ double red[40];
    double s0 = 0.0, s1 = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n; i += 2 * stride) { const double a = v[i], b = v[i + stride]; s0 = fma(a, a, s0); s1 = fma(b, b, s1); }
    if (i < n) { const double a = v[i]; s0 = fma(a, a, s0); }
    const double s = block_sum(s0 + s1, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_sum_fold(const double* __restrict__ partials, int nb, double* __restrict__ out) {
    __shared__ double red[40];
    double s = 0.0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) s += partials[b];
    s = block_sum(s, red);
    if (threadIdx.x == 0) *out = s;
}

struct Scratch {      // CUB temporary storage, grown on demand
    DBuf<unsigned char> buf;
    void* get(size_t bytes) { if (buf.n < bytes) buf.alloc_raw(bytes); return buf.p; }
};

void sort_pairs(Scratch& tmp, const unsigned long long* kin, unsigned long long* kout, const unsigned int* vin,
                unsigned int* vout, int n, int end_bit, cudaStream_t s) {
    size_t bytes = 0;
    PB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, end_bit, s));
    void* t = tmp.get(bytes);
    PB_CUDA(cub::DeviceRadixSort::SortPairs(t, bytes, kin, kout, vin, vout, n, 0, end_bit, s));
}
void exclusive_scan(Scratch& tmp, const int* in, int* out, int n, cudaStream_t s) {
    size_t bytes = 0;
    PB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s));
    void* t = tmp.get(bytes);
    PB_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, in, out, n, s));
}

int bits_for(unsigned long long v) { int b = 1; while (b < 64 && (v >> b)) ++b; return b; }

template <class T>
void upload_async(DBuf<T>& d, const T* h, size_t count, cudaStream_t s) {
    d.alloc_raw(count);
    if (count) PB_CUDA(cudaMemcpyAsync(d.p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    g_h2d_bytes += (long long)(count * sizeof(T));
}

// ---------------------------------------------------------------------------
// equilibrate! (reference src/equilibration.jl:1-71) on the device
// ---------------------------------------------------------------------------
// The reference re-sets the column scaling v to its mean in every step (equilibration.jl:56-58), so v is one number and
// D = exp(v) I.  With that the step only sees M through the row sums r_i = sum_j M_ij^2:
//   row_norms_i = (exp(u_i) exp(v))^2 r_i,      sum_j col_norms_j = sum_i row_norms_i,
// and the whole iteration is R independent scalar recurrences coupled by one sum per step.  One CTA walks all steps
// (rows strided over its 1024 threads, the sum folded in a fixed order: bit-reproducible); u and its running average
// live in global memory.  out[0] = 1 if the preconditioner is applied, out[1] = d = exp(mean-averaged v).
struct EqArgs {
    const int* rowptr; const double* val_o; long long R, n, nnz;
    double lb, ub, limit; long long iters; int enabled, force;
    double* u; double* ubar; double* r; double* E; double* out;
};
__global__ void __launch_bounds__(1024) k_equilibrate(EqArgs a) {
    __shared__ double red[40];
    __shared__ double s_min[32], s_max[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- pdhg.jl:66-73: maximum(M) / minimum(M) run over the structural zeros of the sparse matrix as well
    int on = a.enabled;
    if (on) {
        double mn = INFINITY, mx = -INFINITY;
        for (long long k = tid; k < a.nnz; k += blockDim.x) { const double v = a.val_o[k]; mn = fmin(mn, v); mx = fmax(mx, v); }
        for (int o = 16; o; o >>= 1) { mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
        if (lane == 0) { s_min[warp] = mn; s_max[warp] = mx; }
        __syncthreads();
        mn = s_min[0]; mx = s_max[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = fmin(mn, s_min[w]); mx = fmax(mx, s_max[w]); }
        if ((double)a.nnz < (double)a.R * (double)a.n) { mn = fmin(mn, 0.0); mx = fmax(mx, 0.0); }
        if (mn / mx <= a.limit) on = 0;
    }
    if (a.force) on = 1;
    if (tid == 0) { a.out[0] = on ? 1.0 : 0.0; a.out[1] = 1.0; }
    if (!on) return;
    // ---- r_i, u = ubar = 0
    for (long long i = tid; i < a.R; i += blockDim.x) {
        double s0 = 0.0;
        for (int k = a.rowptr[i]; k < a.rowptr[i + 1]; ++k) { const double v = a.val_o[k]; s0 = fma(v, v, s0); }
        a.r[i] = s0; a.u[i] = 0.0; a.ubar[i] = 0.0;
    }
    const double alpha2 = sqrt((double)a.n / (double)a.R), beta2 = sqrt((double)a.R / (double)a.n), gamma = 0.1;
    double v = 0.0, vbar = 0.0;
    for (long long it = 1; it <= a.iters; ++it) {
        const double step = 2.0 / (gamma * ((double)it + 1.0));
        const double dv = exp(v);
        const double w2 = 2.0 / ((double)it + 2.0), w1 = (double)it / ((double)it + 2.0);
        double S = 0.0;
        for (long long i = tid; i < a.R; i += blockDim.x) {
            double u = a.u[i];
            const double e = exp(u) * dv;
            const double rn = e * e * a.r[i];
            S += rn;
            u -= step * (rn - alpha2 + gamma * u);
            u = fmin(a.ub, fmax(u, a.lb));
            a.u[i] = u;
            a.ubar[i] = w2 * u + w1 * a.ubar[i];
        }
        S = block_sum(S, red);
        // mean over the columns of  v_j - step (col_norms_j - beta2 + gamma v_j)
        v = v - step * (S / (double)a.n - beta2 + gamma * v);
        v = fmin(a.ub, fmax(v, 0.0));
        vbar = w2 * v + w1 * vbar;
    }
    for (long long i = tid; i < a.R; i += blockDim.x) a.E[i] = exp(a.ubar[i]);
    if (tid == 0) a.out[1] = exp(vbar);
}

// working values of M (transposed == 0: row = constraint, colidx = position) or M' (transposed == 1: compact row =
// position nz_rows[q], colidx = constraint) from the caller's values: ((E_i v) d), then norm_scaling's sqrt(2)/2 on
// the off-diagonal positions — the order of pdhg.jl:80 (E * M * D) and scaling.jl:28-58.  One warp per row.
__global__ void k_eq_apply(const int* __restrict__ rowptr, int nrows, const int* __restrict__ nz_rows,
                           const int* __restrict__ colidx, const double* __restrict__ val_o, double* __restrict__ val,
                           const double* __restrict__ E, const double* __restrict__ dptr, int transposed, long long psd_end,
                           const long long* __restrict__ cone_off, int n_sdp) {
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= nrows) return;
    const int rowid = nz_rows ? nz_rows[wid] : (int)wid;
    const double d = *dptr;
    for (int k = rowptr[wid] + lane; k < rowptr[wid + 1]; k += 32) {
        const int i = transposed ? colidx[k] : rowid;
        const long long pos = transposed ? rowid : colidx[k];
        const double w = __dmul_rn(__dmul_rn(E[i], val_o[k]), d);
        val[k] = offdiag_position(pos, psd_end, cone_off, n_sdp) ? __dmul_rn(w, 0.70710678118654752440) : w;
    }
}
__global__ void k_eq_mul(const double* __restrict__ src, const double* __restrict__ E, long long len, double* __restrict__ dst) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) dst[i] = __dmul_rn(E[i], src[i]);
}
__global__ void k_eq_mul_scalar(const double* __restrict__ src, const double* __restrict__ dptr, long long len, double* __restrict__ dst) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double d = *dptr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) dst[i] = __dmul_rn(d, src[i]);
}

}  // namespace

void launch_eq_mul(const double* src, const double* E, long long len, double* dst, cudaStream_t stream) {
    if (len > 0) k_eq_mul<<<(int)std::min<long long>((len + 255) / 256, 148 * 8), 256, 0, stream>>>(src, E, len, dst);
}
void launch_eq_mul_scalar(const double* src, const double* dptr, long long len, double* dst, cudaStream_t stream) {
    if (len > 0) k_eq_mul_scalar<<<(int)std::min<long long>((len + 255) / 256, 148 * 8), 256, 0, stream>>>(src, dptr, len, dst);
}

bool equilibrate_device(CsrDev& M, CsrDev& Mt, long long n, long long R, const proxsdp_options_t& opt,
                        const long long* cone_off_d, int n_sdp, long long psd_end, cudaStream_t stream, EquilibrateOut& out) {
    out.applied = false;
    if (n <= 0 || R <= 0 || M.nnz <= 0) return false;       // nothing to scale
    DBuf<double> u, ubar, r, partials;
    u.alloc_raw((size_t)R); ubar.alloc_raw((size_t)R); r.alloc_raw((size_t)R);
    out.E.alloc_raw((size_t)R); out.d.alloc_raw(4);
    EqArgs a{};
    a.rowptr = M.rowptr.p; a.val_o = M.val_orig.p; a.R = R; a.n = n; a.nnz = M.nnz;
    a.lb = opt.equilibration_lb; a.ub = opt.equilibration_ub; a.limit = opt.equilibration_limit;
    a.iters = opt.equilibration_iters; a.enabled = opt.equilibration ? 1 : 0; a.force = opt.equilibration_force ? 1 : 0;
    a.u = u.p; a.ubar = ubar.p; a.r = r.p; a.E = out.E.p; a.out = out.d.p + 2;
    k_equilibrate<<<1, 1024, 0, stream>>>(a);
    out.launches++;
    double h[2] = {0.0, 1.0};
    PB_CUDA(cudaMemcpyAsync(h, out.d.p + 2, sizeof(double) * 2, cudaMemcpyDeviceToHost, stream));
    PB_CUDA(cudaStreamSynchronize(stream));
    g_d2h_bytes += 16;
    if (h[0] == 0.0) { out.E.release(); out.d.release(); return false; }
    out.applied = true;
    out.d_host = h[1];
    PB_CUDA(cudaMemcpyAsync(out.d.p, out.d.p + 3, sizeof(double), cudaMemcpyDeviceToDevice, stream));    // d at slot 0
    k_eq_apply<<<ceil_div((long long)M.nrows * 32, 256), 256, 0, stream>>>(M.rowptr.p, M.nrows, nullptr, M.colidx.p, M.val_orig.p,
                                                                           M.val.p, out.E.p, out.d.p, 0, psd_end, cone_off_d, n_sdp);
    k_eq_apply<<<ceil_div((long long)Mt.n_nz * 32, 256), 256, 0, stream>>>(Mt.rowptr.p, Mt.n_nz, Mt.nz_rows.p, Mt.colidx.p, Mt.val_orig.p,
                                                                           Mt.val.p, out.E.p, out.d.p, 1, psd_end, cone_off_d, n_sdp);
    const int red_blocks = 148 * 4;
    partials.alloc_raw((size_t)red_blocks + 1);
    k_sumsq_partial<<<red_blocks, 256, 0, stream>>>(Mt.val.p, Mt.nnz, partials.p);
    k_sum_fold<<<1, 256, 0, stream>>>(partials.p, red_blocks, partials.p + red_blocks);
    out.launches += 4;
    PB_CUDA(cudaMemcpyAsync(&out.fro2, partials.p + red_blocks, sizeof(double), cudaMemcpyDeviceToHost, stream));
    PB_CUDA(cudaStreamSynchronize(stream));      // also keeps u / ubar / r alive until the kernel is done
    g_d2h_bytes += 8;
    return true;
}

void launch_gather(const double* src, const int* perm, long long n, double* out, cudaStream_t stream) {
    if (n <= 0) return;
    k_gather<<<(int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, stream>>>(src, perm, n, out);
}

void ingest_problem(const proxsdp_problem_t* prob, const ConeTable& cones, const int* cone_side_d,
                    const long long* cone_off_d, bool want_matrices, cudaStream_t stream, IngestOut& out) {
    const long long n = prob->n, p = prob->p, m = prob->m, R = p + m, base = prob->index_base;
    const long long listed = cones.listed_end;
    Scratch tmp;
    DBuf<int> flags, cnt;
    DBuf<double> rec, partials;
    flags.alloc(F_COUNT); cnt.alloc(C_COUNT); rec.alloc(R_COUNT);
    int* h_small = static_cast<int*>(pb_host_alloc(256));
    double* h_rec = reinterpret_cast<double*>(h_small + 16);
    struct HostFree { void* p; ~HostFree() { pb_host_free(p); } } host_free{h_small};

    // ---- objective: upload as given, ||c||^2 before any permutation / scaling (pdhg.jl:16)
    DBuf<double> c_user;
    upload_async(c_user, prob->c, (size_t)n, stream);
    const int red_blocks = 148 * 4;
    partials.alloc_raw((size_t)red_blocks);
    if (n > 0) {
        k_sumsq_partial<<<red_blocks, 256, 0, stream>>>(c_user.p, n, partials.p);
        k_sum_fold<<<1, 256, 0, stream>>>(partials.p, red_blocks, rec.p + R_NORMC2);
        out.launches += 2;
    }

    // ---- permutation (scaling.jl:2-26).  Pass 1 only checks for the identity (what JuMP/MOI produce for matrix
    // variables created first): no per-variable tables are built in that case.
    DBuf<int64_t> idx_d;
    if (listed > 0) {
        idx_d.alloc_raw((size_t)listed);
        const long long n_sdp_idx = cones.psd_end;
        if (n_sdp_idx > 0) PB_CUDA(cudaMemcpyAsync(idx_d.p, prob->sdp_idx, sizeof(int64_t) * (size_t)n_sdp_idx, cudaMemcpyHostToDevice, stream));
        if (listed > n_sdp_idx) PB_CUDA(cudaMemcpyAsync(idx_d.p + n_sdp_idx, prob->soc_idx, sizeof(int64_t) * (size_t)(listed - n_sdp_idx), cudaMemcpyHostToDevice, stream));
        g_h2d_bytes += (long long)sizeof(int64_t) * listed;
        k_perm_mark<<<ceil_div(listed, 256), 256, 0, stream>>>(idx_d.p, listed, n, base, nullptr, nullptr, flags.p);
        out.launches++;
    }
    PB_CUDA(cudaMemcpyAsync(h_small, flags.p, sizeof(int) * F_COUNT, cudaMemcpyDeviceToHost, stream));
    PB_CUDA(cudaStreamSynchronize(stream));
    if (h_small[F_ERR_RANGE]) throw CudaError(-3, "variable index out of range in a cone");
    out.identity = h_small[F_NONIDENT] == 0;
    if (!out.identity) {
        DBuf<int> used, uflag, urank;
        used.alloc((size_t)n);
        out.ord.alloc_raw((size_t)n); out.var_ordering.alloc_raw((size_t)n);
        uflag.alloc_raw((size_t)n); urank.alloc_raw((size_t)n);
        k_perm_mark<<<ceil_div(listed, 256), 256, 0, stream>>>(idx_d.p, listed, n, base, used.p, out.ord.p, flags.p);
        k_unused_flag<<<ceil_div(n, 256), 256, 0, stream>>>(used.p, n, uflag.p);
        exclusive_scan(tmp, uflag.p, urank.p, (int)n, stream);
        k_fill_free<<<ceil_div(n, 256), 256, 0, stream>>>(used.p, urank.p, n, listed, out.ord.p);
        k_invert_perm<<<ceil_div(n, 256), 256, 0, stream>>>(out.ord.p, n, out.var_ordering.p);
        out.launches += 5;
        PB_CUDA(cudaMemcpyAsync(h_small, flags.p, sizeof(int) * F_COUNT, cudaMemcpyDeviceToHost, stream));
        PB_CUDA(cudaStreamSynchronize(stream));
        if (h_small[F_ERR_DUP]) throw CudaError(-3, "a variable appears in two cones");
        out.c_orig.alloc_raw((size_t)n);
        launch_gather(c_user.p, out.ord.p, n, out.c_orig.p, stream);
        out.launches++;
    } else {
        out.c_orig = std::move(c_user);
    }
    idx_d.release();

    // ---- M = [A; G] and M' (pdhg.jl:95-142)
    const long long nnzA = (p > 0 && prob->A_colptr && n > 0) ? prob->A_colptr[n] - base : 0;
    const long long nnzG = (m > 0 && prob->G_colptr && n > 0) ? prob->G_colptr[n] - base : 0;
    const long long nnz = nnzA + nnzG;
    if (nnz >= (1LL << 31) - 64) throw CudaError(-1, "too many non-zeros for 32-bit indices");
    CsrDev& M = out.M; CsrDev& Mt = out.Mt;
    M.nrows = (int)R; M.ncols = (int)n; M.nnz = (int)nnz; M.long_threshold = 4096;
    Mt.nrows = (int)n; Mt.ncols = (int)R; Mt.nnz = (int)nnz; Mt.long_threshold = 256;
    if (want_matrices && nnz > 0) {
        DBuf<int64_t> Acp, Arv, Gcp, Grv;
        DBuf<double> Anz, Gnz;
        if (nnzA > 0) { upload_async(Acp, prob->A_colptr, (size_t)n + 1, stream); upload_async(Arv, prob->A_rowval, (size_t)nnzA, stream); upload_async(Anz, prob->A_nzval, (size_t)nnzA, stream); }
        if (nnzG > 0) { upload_async(Gcp, prob->G_colptr, (size_t)n + 1, stream); upload_async(Grv, prob->G_rowval, (size_t)nnzG, stream); upload_async(Gnz, prob->G_nzval, (size_t)nnzG, stream); }
        DBuf<unsigned long long> key_t, key_m, key_s;
        DBuf<unsigned int> id, id_s;
        DBuf<double> val_s, val_o;
        DBuf<int> head, scan;
        key_t.alloc_raw((size_t)nnz); key_m.alloc_raw((size_t)nnz); key_s.alloc_raw((size_t)nnz);
        id.alloc_raw((size_t)nnz); id_s.alloc_raw((size_t)nnz);
        val_s.alloc_raw((size_t)nnz); val_o.alloc_raw((size_t)nnz);
        head.alloc_raw((size_t)nnz); scan.alloc_raw((size_t)nnz);
        EntryArgs ea{};
        ea.A_colptr = Acp.p; ea.A_rowval = Arv.p; ea.A_nzval = Anz.p; ea.nnzA = nnzA;
        ea.G_colptr = Gcp.p; ea.G_rowval = Grv.p; ea.G_nzval = Gnz.p; ea.nnzG = nnzG;
        ea.n = n; ea.p = p; ea.m = m; ea.base = base; ea.psd_end = cones.psd_end;
        ea.var_ordering = out.identity ? nullptr : out.var_ordering.p;
        ea.cone_side = cone_side_d; ea.cone_off = cone_off_d; ea.n_sdp = cones.n_sdp;
        ea.key_t = key_t.p; ea.key_m = key_m.p; ea.id = id.p; ea.val_s = val_s.p; ea.val_o = val_o.p; ea.flags = flags.p;
        k_entries<<<ceil_div(nnz, 256), 256, 0, stream>>>(ea);
        out.launches++;
        const int key_bits_t = 32 + bits_for((unsigned long long)std::max<long long>(n, 1));
        const int key_bits_m = 32 + bits_for((unsigned long long)std::max<long long>(R, 1));
        // ---- M' : entries ordered by (position, row)
        sort_pairs(tmp, key_t.p, key_s.p, id.p, id_s.p, (int)nnz, key_bits_t, stream);
        Mt.colidx.alloc_raw((size_t)nnz); Mt.val.alloc_raw((size_t)nnz); Mt.val_orig.alloc_raw((size_t)nnz);
        k_scatter_sorted<<<ceil_div(nnz, 256), 256, 0, stream>>>(key_s.p, id_s.p, val_s.p, val_o.p, nnz, Mt.colidx.p, Mt.val.p,
                                                                Mt.val_orig.p, head.p);
        exclusive_scan(tmp, head.p, scan.p, (int)nnz, stream);
        const long long max_rows = std::min<long long>(nnz, n);
        Mt.nz_rows.alloc_raw((size_t)max_rows); Mt.rowptr.alloc_raw((size_t)max_rows + 1);
        k_dcsr_rows<<<ceil_div(nnz, 256), 256, 0, stream>>>(key_s.p, head.p, scan.p, nnz, Mt.nz_rows.p, Mt.rowptr.p, cnt.p);
        // long rows of M' (head / scan are free again: re-used as flag / rank over the compact rows)
        k_long_flag<<<ceil_div(max_rows, 256), 256, 0, stream>>>(Mt.rowptr.p, cnt.p + C_NNZ_ROWS, (int)max_rows, Mt.long_threshold, head.p);
        exclusive_scan(tmp, head.p, scan.p, (int)max_rows, stream);
        Mt.long_rows.alloc_raw((size_t)max_rows);
        k_compact<<<ceil_div(max_rows, 256), 256, 0, stream>>>(head.p, scan.p, (int)max_rows, Mt.long_rows.p, cnt.p + C_LONG_T);
        // ||M||_F^2 of the scaled matrix, in the (deterministic) sorted order
        k_sumsq_partial<<<red_blocks, 256, 0, stream>>>(Mt.val.p, nnz, partials.p);
        k_sum_fold<<<1, 256, 0, stream>>>(partials.p, red_blocks, rec.p + R_FRO2);
        out.launches += 8;
        // ---- M : entries ordered by (row, position)
        sort_pairs(tmp, key_m.p, key_s.p, id.p, id_s.p, (int)nnz, key_bits_m, stream);
        M.colidx.alloc_raw((size_t)nnz); M.val.alloc_raw((size_t)nnz); M.val_orig.alloc_raw((size_t)nnz);
        k_scatter_sorted<<<ceil_div(nnz, 256), 256, 0, stream>>>(key_s.p, id_s.p, val_s.p, val_o.p, nnz, M.colidx.p, M.val.p,
                                                                M.val_orig.p, nullptr);
        M.rowptr.alloc_raw((size_t)R + 1);
        k_rowptr_from_keys<<<ceil_div(R + 1, 256), 256, 0, stream>>>(key_s.p, nnz, R, M.rowptr.p);
        DBuf<int> lflag, lscan;
        lflag.alloc_raw((size_t)R); lscan.alloc_raw((size_t)R);
        k_long_flag<<<ceil_div(R, 256), 256, 0, stream>>>(M.rowptr.p, nullptr, (int)R, M.long_threshold, lflag.p);
        exclusive_scan(tmp, lflag.p, lscan.p, (int)R, stream);
        M.long_rows.alloc_raw((size_t)R);
        k_compact<<<ceil_div(R, 256), 256, 0, stream>>>(lflag.p, lscan.p, (int)R, M.long_rows.p, cnt.p + C_LONG_M);
        out.launches += 4;
        PB_CUDA(cudaMemcpyAsync(h_small, flags.p, sizeof(int) * F_COUNT, cudaMemcpyDeviceToHost, stream));
        PB_CUDA(cudaMemcpyAsync(h_small + 8, cnt.p, sizeof(int) * C_COUNT, cudaMemcpyDeviceToHost, stream));
        PB_CUDA(cudaMemcpyAsync(h_rec, rec.p, sizeof(double) * R_COUNT, cudaMemcpyDeviceToHost, stream));
        PB_CUDA(cudaStreamSynchronize(stream));      // also keeps the temporaries above alive until their kernels are done
        if (h_small[F_ERR_ROW]) throw CudaError(-3, "row index out of range in A or G");
        Mt.n_nz = h_small[8 + C_NNZ_ROWS]; Mt.n_long = h_small[8 + C_LONG_T]; M.n_long = h_small[8 + C_LONG_M];
        const double avg = R > 0 ? (double)nnz / (double)R : 0.0;
        M.group = avg <= 1.5 ? 1 : avg <= 3 ? 2 : avg <= 6 ? 4 : avg <= 12 ? 8 : avg <= 24 ? 16 : 32;
    } else {
        if (want_matrices) {
            M.rowptr.alloc((size_t)R + 1); M.colidx.alloc(1); M.val.alloc(1); M.val_orig.alloc(1); M.long_rows.alloc(1);
            Mt.rowptr.alloc(1); Mt.colidx.alloc(1); Mt.val.alloc(1); Mt.val_orig.alloc(1); Mt.long_rows.alloc(1); Mt.nz_rows.alloc(1);
        }
        PB_CUDA(cudaMemcpyAsync(h_rec, rec.p, sizeof(double) * R_COUNT, cudaMemcpyDeviceToHost, stream));
        PB_CUDA(cudaStreamSynchronize(stream));
    }
    g_d2h_bytes += 64;
    out.fro2 = h_rec[R_FRO2];
    out.norm_c2 = h_rec[R_NORMC2];
}

}  // namespace pb
