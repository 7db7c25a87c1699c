// solver.cu — host side of the B200 ProxSDP hot path: setup, per-iteration launch
// sequence, the reference's scalar control logic, result assembly, and the C ABI.
//
// Restates reference src/pdhg.jl:1-530 (`chambolle_pock`) with every array operation of
// the loop executed by the kernels in kernels_vec.cuh / lanczos.cuh / fulleig.cuh and
// exactly one device->host read-back (the scalar record) per iteration.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <numeric>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>

#include "../../include/proxsdp_b200.h"
#include "common.cuh"
#include "runtime.cuh"
#include "printing.hpp"
#include "spectral.hpp"
#include "fulleig.cuh"
#include "jacobi.cuh"
#include "kernels_vec.cuh"
#include "kernels_result.cuh"
#include "lanczos.cuh"
#include "lanczos_cl.cuh"
#include "lanczos_cl3.cuh"

namespace pb {

__global__ void k_scale_offdiag_copy(const double* __restrict__ src, long long n, long long psd_end,
                                     const long long* __restrict__ cone_off, int n_sdp, double num, int divide,
                                     double* __restrict__ dst);
__global__ void k_scale_copy(const double* __restrict__ src, double a, long long n, double* __restrict__ dst);

static thread_local std::string g_last_error;

// ---------------------------------------------------------------------------
// NCCL, bound at run time (dlopen) so that the library loads on machines without it; only the
// sharded (multi-GPU) entry points need it.
// ---------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi& nccl_api() {
    static NcclApi api;
    if (api.handle) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
    if (!api.handle) throw CudaError(-6, "NCCL not found (dlopen libnccl.so.2 failed): multi-GPU entry points are unavailable");
    auto sym = [&](const char* n) { void* f = dlsym(api.handle, n); if (!f) throw CudaError(-6, std::string("NCCL symbol missing: ") + n); return f; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    return api;
}
#define PB_NCCL(expr)                                                                              \
    do {                                                                                           \
        ncclResult_t _r = (expr);                                                                  \
        if (_r != ncclSuccess) throw pb::CudaError(-7, std::string(#expr " failed: ") + nccl_api().GetErrorString(_r)); \
    } while (0)


// the strip width (64-double chunks per warp) selects the kernel instance: cpw = ceil(ceil(n / 64) / 16);
// <strip width, streamed rows per load chunk (two chunks in flight), rows per butterfly>: the widest choice that ptxas
// fits into 128 registers without spilling on the per-step path.  Deeper load batches were measured and are slower
// (profiles/r2_rb_sweep.txt): the mat-vec is not limited by the bytes in flight.
static const void* lanczos_cl3_kernel(int cpw) {
    switch (cpw) {
        case 1: return (const void*)k_lanczos_cl3<1, 3, 9>;
        case 2: return (const void*)k_lanczos_cl3<2, 2, 8>;
        // sides > 2048 (matrices of 34 MB and more): the variants with the L2 prefetch of the slab rows compiled in
        // (profiles/r2_l2_prefetch.txt; at 32 MB it gains nothing and its code costs the per-step path a few per cent)
        case 3: return (const void*)k_lanczos_cl3<3, 1, 8, false, true>;
        case 4: return (const void*)k_lanczos_cl3<4, 1, 6, false, true>;
        case 5: return (const void*)k_lanczos_cl3<5, 1, 6, false, true>;
        case 6: return (const void*)k_lanczos_cl3<6, 1, 6, false, true>;
        case 8: return (const void*)k_lanczos_cl3<8, 1, 4, false, true>;
        default: return nullptr;
    }
}

static bool g_timing = getenv("PROXSDP_B200_TIMING") != nullptr;
static double now_s();
struct StageTimer {
    double t; const char* what;
    explicit StageTimer(const char* w);
    void lap(const char* next);
};
static double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

StageTimer::StageTimer(const char* w) : t(now_s()), what(w) {}
void StageTimer::lap(const char* next) {
    if (g_timing) { cudaDeviceSynchronize(); double n = now_s(); fprintf(stderr, "[timing] %-28s %8.3f ms\n", what, 1e3 * (n - t)); t = n; }
    what = next;
}

// Julia's max() propagates NaN
static double jl_max(double a, double b) { return (a != a || b != b) ? NAN : std::max(a, b); }

// CircularVector (reference src/structs.jl:2-30)
struct Circ {
    std::vector<double> v;
    long long l = 0;
    void init(long long len) { l = len; v.assign((size_t)len, 0.0); }
    static long long mod1(long long i, long long l) { return ((i - 1) % l + l) % l; }
    double get(long long i) const { return v[(size_t)mod1(i, l)]; }
    void set(long long i, double x) { v[(size_t)mod1(i, l)] = x; }
    double max_abs_diff() const {   // structs.jl:14-20 (includes the wrap seam)
        double val = 0.0;
        for (long long i = 1; i <= l; ++i) {
            double d = std::fabs(get(i) - get(i - 1));
            if (d > val) val = d;
        }
        return val;
    }
};

// splitmix64 + Box-Muller substitute for Julia's MersenneTwister stream
// (reference src/eigsolver.jl:392-411); identical to the oracle's generator.
static void eig_resid_default(long long n, long long seed, long long init, double* out) {
    uint64_t s = (uint64_t)seed;
    auto next = [&]() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    };
    auto uni = [&]() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); };
    if (init == 3) {
        double nn = 0.0;
        for (long long i = 0; i < n; ++i) {
            double u1 = 1.0 - uni();
            double u2 = uni();
            out[i] = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925286766559 * u2);
        }
        for (long long i = 0; i < n; ++i) nn += out[i] * out[i];
        nn = std::sqrt(nn);
        for (long long i = 0; i < n; ++i) out[i] /= nn;
    } else if (init == 2) {
        for (long long i = 0; i < n; ++i) out[i] = uni();
    } else if (init == 1) {
        for (long long i = 0; i < n; ++i) out[i] = 1.0;
    } else {
        for (long long i = 0; i < n; ++i) out[i] = 0.0;
    }
}

}  // namespace pb
struct proxsdp_b200_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
};
namespace pb {

// ---------------------------------------------------------------------------
// per-cone device state
// ---------------------------------------------------------------------------
struct ConeDev {
    int side = 0, ld = 0;
    long long off = 0;          // svec offset in x
    bool small = false;
    DBuf<double> X;             // ld x ld, large cones only
    DBuf<double> Vfull;         // ld x ld eigenvectors for the block-Jacobi path (allocated lazily)
    DBuf<double> Vtmp;          // ld x ld scratch of the warm start (A W)
    // implicit operator (opt.implicit_psd_operator): CSR by row of the sparsity pattern of mat(M'y + c) inside this cone
    DBuf<int> imp_rowptr, imp_col, imp_pos;
    DBuf<double> imp_coef;
    int imp_cap = 0;            // largest number of pattern entries in one CTA's rows (cluster split)
    bool imp_tried = false, imp_ready = false;
    bool lowrank_valid = false; // x holds svec(Y diag(lam) Y') of the previous Krylov projection (Y, kept_* still on the device)
    bool used_implicit = false; // this iteration's eigsolve ran on the implicit operator (X was not formed)
    bool announced_resident = false;
    bool have_V = false;        // Vfull holds the eigenvectors of this cone's previous full projection
    long long full_calls = 0;
    DBuf<double> Y;             // ld x (Kmax)
    DBuf<double> resid, vals, kept_lam;
    DBuf<double> ritz_ws[2];    // warm start of the K x K Ritz eigenproblem (ping-pong between launches)
    long long ritz_launches = 0;
    DBuf<int> info, kept_idx, nkept;
};

class Solver {
  public:
    Solver(const proxsdp_problem_t* prob, const proxsdp_options_t* opt_in, bool cones_only = false, bool force_large = false,
           const proxsdp_shard_t* shard = nullptr);
    ~Solver();
    void solve(proxsdp_result_t* out);
    // stepwise form of the same loop (bench / step-level seam): begin -> run(max_steps)* -> finish
    void begin(proxsdp_result_t* out);
    bool run(long long max_steps, bool flush_l2);   // true once the loop has terminated
    void finish(proxsdp_result_t* out);
    long long iterations_done() const { return iter_; }
    double time_psd_ms() const { return time_psd_ms_; }
    double time_lanczos_ms_ = 0, time_post_ms_ = 0, time_flush_ms_ = 0;
    long long lanczos_timed_calls = 0;
    // step-level seam: one psd_projection! on x (device x_[cur_] -> x_[1-cur_]) with tau = 0
    void psd_projection_launch(long long iter, double tau, bool force_full);
    void sync_scalars(bool iteration_end = false);
    void reset_scalars();
    void fallback_projection(long long iter);
    void launch_soc_only();
    bool force_large_ = false;

    // exposed for the seams
    proxsdp_options_t opt;
    long long n = 0, p = 0, m = 0, R = 0;
    int n_sdp = 0, n_soc = 0;
    std::vector<ConeDev> cones;
    std::vector<long long> target_rank, current_rank;
    std::vector<double> min_eig;
    DBuf<double> x_[2], Mty_[2], y_[2], Mx_[2], c_, b_, h_;
    int cur_ = 0;
    double* scal_host = nullptr;
    long long launches = 0;
    long long lanczos_matvecs = 0, lanczos_calls = 0, full_eig_calls = 0, linesearch_trials = 0;
    cudaStream_t stream = nullptr;
    int scal_len = 0;

    // device full eigendecomposition of the matrix currently in cone.X (destroys X):
    // eigenvalues -> host vector (unsorted), eigenvectors in cone.Vfull
    // warm: start from the eigenvector basis of the cone's previous decomposition (consecutive PDHG iterates have nearly
    // the same eigenvectors, so W' A W is nearly diagonal and 2-3 sweeps replace ~10); every 32nd call starts cold
    std::vector<double> full_eig_device(ConeDev& cd, bool warm = false);
    int bj_warm_ = 1;          // PROXSDP_B200_BJ_WARM=0 disables the warm start
    void lanczos_launch(ConeDev& cd, int cone_idx, int nev, int K, int maxiter, double tol);
    bool lanczos_launch_cluster(ConeDev& cd, int cone_idx, int nev, int K, int maxiter, double tol);
    bool lanczos_launch_cluster3(ConeDev& cd, int cone_idx, int nev, int K, int maxiter, double tol);
    bool lanczos_launch_implicit(ConeDev& cd, int cone_idx, int nev, int K, int maxiter, double tol, double tau);
    bool lanczos_launch_resident(ConeDev& cd, int cone_idx, int nev, int K, int maxiter, double tol);
    bool implicit_pattern(ConeDev& cd);
    double last_tau_ = 0.0;
    long long implicit_calls = 0;
    double lz_stop_above_ = 1e300;   // cl3 kernel: early exit once the top Ritz value exceeds this (cone_feas only)
    int small_fast_ = 0;       // small cones: double-buffered Jacobi when three matrices fit shared memory
    int small_warm_ = 1;       // small cones: warm-start the Jacobi sweeps from the previous projection's eigenvectors
    long long small_warm_calls_ = 0;
    DBuf<double> small_warm_d_;
    int lz_kernel_ = 3;        // 3 = third-generation cluster kernel (lanczos_cl3.cuh), 2 = second generation (lanczos_cl.cuh)
    int lz_coop_ = 1;          // cluster kernels are launched with cudaLaunchAttributeCooperative (PROXSDP_B200_LZ_COOP=0: plain launch)
    bool lz_demoted_ = false;
    int lz_arrow_ = 0;         // cl3 kernel: 1 = arrowhead thick restart + dense Ritz solves (first-round behaviour)
    int lz_poll_ns_ = 0;       // cl3 kernel: nanoseconds of back-off between polls of the exchange words
    int lz_resident_ = 1;      // mid-size cones: single-cluster kernel with X resident in distributed shared memory (PROXSDP_B200_LZ_RESIDENT=0: off)
    int lz_resident16_ = 1;    // ... also with the non-portable cluster size 16 (PROXSDP_B200_LZ_RESIDENT=8: portable size only, =16: that size only)
    int lz_resident8_ = 1;
    int lz_resident_ok_[2] = {0, 0};
    long long resident_calls = 0;
    long long lz_spin_limit_ = 4000000000LL;   // cycles (~2 s) a spin loop of the cluster kernels waits for a peer
    int lz_bi_memory_ = 0;     // PROXSDP_B200_RITZ_MEM=1: once a bisection Ritz solve has declined, the later analyses of that launch go straight
                               // to the dense solver (experiment; off: on gpp500-1 it changes which copies of multiple eigenvalues are found)
    DBuf<unsigned int> nzmask_;   // structural-nonzero bitmap of (M'y, c): the streaming kernels skip the zeros (kernels_vec.cuh)
    int lz_pf_ = -1;           // cl3 kernel: slab rows prefetched into L2 ahead of the loads (-1: 8 rows when the matrix exceeds L2)
    int lz_strict_ = 0;        // cl3 kernel: FP64 alpha + two Gram-Schmidt passes on every step (PROXSDP_B200_LZ_STRICT=1)
    int lz_xres_ = 0;          // >= 1: cap the resident slab rows of the cl3 symv at lz_xres_ - 1 (experiments / tests)
    size_t lz_cl3_smem_max_ = 0;
    int lz_mode_ = 0;          // 0 = cluster-replicated kernel when it fits, 1 = row-distributed kernel only
    int lz_cluster_ = 8;       // cluster size of the replicated kernel
    int lz_bi_ = 1;            // leading Ritz pairs by bisection + twisted vectors (dense Jacobi as fallback)
    int lz_warm_ = 1;          // warm-start the Ritz eigenproblem from the previous eigsolve
    bool lz_cluster_warned_ = false;
    long long lz_cluster_launches_ = 0;
    size_t lz_cl_smem_max_ = 0;

  private:
    void setup_host(const proxsdp_problem_t* prob);
    double spectral_norm_device();
    bool logging() const { return opt.log_verbose && rank_ == 0; }     // sharded runs: rank 0 writes the log
    void log_progress(double dual_feas_val);
    void launch_spmv(const CsrDev& A, const double* x, double* y);
    void launch_full_projection_large(int k);
    void launch_reconstruct(ConeDev& cd, double* x_out);
    void launch_post_eig(double tau0, bool first_pass);
    void launch_dual_trial(int trial, double tau0);
    void launch_ladder(int trial0, int T, double tau0);
    long long linesearch_continue(double tau0, bool& exhausted, double& last_tau);
    void host_residuals(long long k);
  public:
    // step-level seams (proxsdp_b200_dual_step / proxsdp_b200_residuals): one linesearch! / dual_step!, one
    // compute_residual! + compute_gap!, on state the caller has placed in the ping-pong buffers
    long long seam_dual_step(double primal_step, double primal_step_old, double theta, double beta, double dual_step, double* out4);
    void seam_residuals(double primal_step, double dual_step, double beta, double norm_b, double norm_h, double norm_c, double* out8);
  private:
    void cache_solution(double c_factor, proxsdp_result_t* out);
    // get_duals + dual_feas (pdhg.jl:701-732) on the device: dual_cone = c_factor * c + M'y (un-scaled, off-diagonals / 2)
    // is left in res_dc_d_ (position order); returns the dual infeasibility measure
    double dual_feas_device(const double* y_dev, double c_factor);
    void rank_increment_rule(int idx);
    void record_trace(proxsdp_result_t* out);
    bool krylov_eligible(int k, long long iter) const;

    bool cones_only_ = false;
    int dev_ = 0, num_sms_ = 148;
    int l2_bytes_ = 126 << 20;
    size_t smem_optin_ = 0;
    // host problem data: only the right-hand sides (R doubles) and the cone table stay on the host
    std::vector<double> b_host_, h_host_;
    bool identity_ = false;                        // variable permutation is the identity (no ord / var_ordering)
    std::vector<long long> soc_off_h_;
    std::vector<int> soc_len_h_;
    long long psd_end_ = 0;                        // first index after the PSD blocks
    long long listed_end_ = 0;                     // first index after the SOC blocks (free variables follow)
    // device problem (built by ingest_problem, runtime.cu)
    CsrDev M_, Mt_;
    DBuf<int> ord_d_, var_ordering_d_;             // empty when identity_
    DBuf<double> c_orig_d_, b_orig_d_, h_orig_d_;  // un-scaled objective (position order) and right-hand sides
    DBuf<double> eq_E_d_, eq_d_d_;                 // equilibrate!: row scalings E (R), column scaling d[0] (D = d I)
    DBuf<double> res_x_d_, res_dc_d_, res_user_d_, res_slack_d_, feas_d_, scal_scratch_d_;
    double* scal_target_ = nullptr;                // scalar record the eigen kernels write to (scal_d_ inside the loop)
    // the iteration record reaches the host through mapped page-locked memory (scal_host) + a sequence word
    double* scal_host_dev_ = nullptr;
    unsigned long long* pub_seq_host_ = nullptr;
    unsigned long long* pub_seq_dev_ = nullptr;
    unsigned long long pub_seq_ = 0;
    bool scal_clean_ = false;                      // the header of the device record is already zero (reset by the last publish)
    bool spin_sync_ = true;                        // PROXSDP_B200_SYNC=memcpy: cudaMemcpyAsync + cudaStreamSynchronize instead
    int fused_ladder_ = 1;                         // PROXSDP_B200_LADDER_FUSED=0: one launch pair per trial (first-round path)
    DBuf<double> ls_long_sums_, ls_partial_d_;
    double* feas_host_ = nullptr;
    DBuf<int> cone_side_d_, small_ids_d_;
    DBuf<long long> cone_off_d_, soc_off_d_;
    DBuf<int> soc_len_d_;
    DBuf<double> soc_gap_d_, scal_d_, partials_d_, out_min_d_, offnorm_d_;
    DBuf<unsigned int> counters_d_;
    DBuf<double> bj_Q_;
    DBuf<int> bj_rot_;
    int bj_inner_sweeps_ = 2;      // cyclic Jacobi sweeps on a 64 x 64 pivot block per visit (PROXSDP_B200_BJ_INNER)
    std::vector<int> small_ids_, large_ids_;
    int max_small_side_ = 0, Kmax_ = 25;
    ReduceWs ws_{};
    int reduce_blocks_ = 0;
    // Params (reference src/structs.jl:159-192)
    long long rank_update_ = 0, update_cont_ = 0, iter_ = 0, stop_reason_ = 0;
    std::string stop_reason_string_ = "Not optimized";
    double primal_step_ = 0, primal_step_old_ = 0, dual_step_ = 0, theta_ = 1, beta_ = 1, adapt_level_ = 0.9;
    long long window_ = 200;
    double time0_ = 0, norm_c_ = 0, norm_b_ = 0, norm_h_ = 0;
    double dual_feasibility_ = -1.0;
    bool dual_feasibility_check_ = false, certificate_search_ = false, certificate_found_ = false;
    long long certificate_search_min_iter_ = 0;
    // Residuals
    Circ dual_gap_, prim_obj_, dual_obj_, feasibility_, primal_residual_, dual_residual_, comb_residual_;
    double equa_feasibility_ = 0, ineq_feasibility_ = 0;
    double soc_gap_max_ = -1.0;
    // timing
    cudaEvent_t ev_psd0_ = nullptr, ev_psd1_ = nullptr;
    double time_psd_ms_ = 0;
    long long n_psd_ = 0;
    int ladder_ = 4;
    long long trace_mv0_ = 0, trace_ls0_ = 0;
    // sharded solves: whole-problem scalars are combined across ranks
    proxsdp_b200_comm* comm_ = nullptr;
    int nranks_ = 1, rank_ = 0;
    long long global_n_ = -1, global_R_ = -1, global_p_ = -1, global_m_ = -1;
    bool global_has_soc_ = false;
    DBuf<double> gather_d_, red_d_;
    void host_reduce(double* vals, int count, int op);          // op 0 = sum, 1 = max (rank-ordered, deterministic)
    bool sharded() const { return nranks_ > 1; }
    // loop-carried state of chambolle_pock's main loop
    long long ada_count_ = 0, k_next_ = 1;
    bool have_cached_ = false, loop_done_ = false;
    proxsdp_result_t* out_ = nullptr;
    double t_loop_accum_ = 0;
    cudaEvent_t ev_lz0_ = nullptr, ev_lz1_ = nullptr, ev_post1_ = nullptr, ev_fl0_ = nullptr, ev_fl1_ = nullptr;
    bool lz_timed_ = false;
    DBuf<double> flush_buf_;
    DBuf<uint4> lz_xbuf_, lz_vx_;
    DBuf<double> lz_wg_;
    DBuf<unsigned int> lz_bar_;
    DBuf<uint4> lz3_wg_, lz3_apart_;
    unsigned long long lz3_epoch_ = 0;
    DBuf<long long> lz_prof_;
    unsigned long long lz_epoch_ = 0;
};

// ---------------------------------------------------------------------------
// construction / setup  (pdhg.jl:7-142)
// ---------------------------------------------------------------------------
Solver::Solver(const proxsdp_problem_t* prob, const proxsdp_options_t* opt_in, bool cones_only, bool force_large,
               const proxsdp_shard_t* shard)
    : opt(*opt_in), force_large_(force_large), cones_only_(cones_only) {
    if (shard && shard->nranks > 1) {
        if (!shard->comm) throw CudaError(-1, "sharded solve needs a communicator");
        comm_ = static_cast<proxsdp_b200_comm*>(shard->comm);
        nranks_ = comm_->nranks; rank_ = comm_->rank;
        if (nranks_ != shard->nranks || rank_ != shard->rank) throw CudaError(-1, "shard descriptor does not match the communicator");
        global_n_ = shard->global_n; global_R_ = shard->global_p + shard->global_m;
        global_p_ = shard->global_p; global_m_ = shard->global_m;
        opt.device_id = comm_->device;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw CudaError(-5, "no CUDA device available");
    dev_ = (int)opt.device_id;
    if (dev_ < 0 || dev_ >= ndev) throw CudaError(-1, "device_id out of range");
    StageTimer st0("ctor: device");
    PB_CUDA(cudaSetDevice(dev_));
    {
        // cudaGetDeviceProperties costs milliseconds: two attributes, cached per process and device
        static std::mutex mu;
        static int sms[64], smem[64], l2b[64];
        static bool have[64] = {false};
        std::lock_guard<std::mutex> g(mu);
        const int d = dev_ & 63;
        if (!have[d]) {
            PB_CUDA(cudaDeviceGetAttribute(&sms[d], cudaDevAttrMultiProcessorCount, dev_));
            PB_CUDA(cudaDeviceGetAttribute(&smem[d], cudaDevAttrMaxSharedMemoryPerBlockOptin, dev_));
            PB_CUDA(cudaDeviceGetAttribute(&l2b[d], cudaDevAttrL2CacheSize, dev_));
            size_t lim = 0;
            if (cudaDeviceGetLimit(&lim, cudaLimitStackSize) == cudaSuccess && lim < 4096) cudaDeviceSetLimit(cudaLimitStackSize, 4096);
            have[d] = true;
        }
        num_sms_ = sms[d]; smem_optin_ = (size_t)smem[d]; l2_bytes_ = l2b[d];
    }
    st0.lap("ctor: stream+events");
    PB_CUDA(cudaStreamCreate(&stream));   // blocking stream: ordered against the synchronous setup copies on the legacy stream
    PB_CUDA(cudaEventCreate(&ev_psd0_));
    PB_CUDA(cudaEventCreate(&ev_psd1_));
    for (cudaEvent_t* e : {&ev_lz0_, &ev_lz1_, &ev_post1_, &ev_fl0_, &ev_fl1_}) PB_CUDA(cudaEventCreate(e));
    if (const char* e = getenv("PROXSDP_B200_LADDER")) ladder_ = std::max(1, atoi(e));
    if (const char* e = getenv("PROXSDP_B200_LANCZOS")) lz_mode_ = (std::string(e) == "rows") ? 1 : 0;
    if (const char* e = getenv("PROXSDP_B200_CLUSTER")) lz_cluster_ = std::max(1, std::min(LZC_MAXC, atoi(e)));
    if (const char* e = getenv("PROXSDP_B200_RITZ_WARM")) lz_warm_ = atoi(e);
    if (const char* e = getenv("PROXSDP_B200_LZ_KERNEL")) lz_kernel_ = atoi(e);
    if (const char* e = getenv("PROXSDP_B200_LZ_XRES")) lz_xres_ = atoi(e);
    if (const char* e = getenv("PROXSDP_B200_LZ_STRICT")) lz_strict_ = atoi(e) != 0 ? 1 : 0;
    if (const char* e = getenv("PROXSDP_B200_LZ_POLL_NS")) lz_poll_ns_ = std::max(0, atoi(e));
    if (const char* e = getenv("PROXSDP_B200_LZ_PF")) lz_pf_ = std::max(0, atoi(e));
    if (const char* e = getenv("PROXSDP_B200_RITZ_MEM")) lz_bi_memory_ = atoi(e);
    if (const char* e = getenv("PROXSDP_B200_LZ_SPIN_S")) lz_spin_limit_ = (long long)(std::max(1.0, atof(e)) * 2.0e9);
    if (const char* e = getenv("PROXSDP_B200_LZ_RESIDENT")) { const int v = atoi(e); lz_resident_ = v != 0; lz_resident16_ = (v != 8); lz_resident8_ = (v != 16); }
    if (const char* e = getenv("PROXSDP_B200_LZ_ARROW")) lz_arrow_ = atoi(e) != 0 ? 1 : 0;
    // Nsight Compute cannot replay a launch that carries the cooperative attribute together with a cluster dimension
    // (it reports LaunchFailed and tears the process down): under an injected profiler the launch is a plain cluster
    // launch, co-residency then rests on cudaOccupancyMaxActiveClusters and the in-kernel time-outs
    if (getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("NV_NSIGHT_INJECTION_PORT_BASE")) lz_coop_ = 0;
    if (const char* e = getenv("PROXSDP_B200_LZ_COOP")) lz_coop_ = atoi(e) != 0 ? 1 : 0;
    if (const char* e = getenv("PROXSDP_B200_BJ_INNER")) bj_inner_sweeps_ = std::max(1, atoi(e));
    if (const char* e = getenv("PROXSDP_B200_BJ_WARM")) bj_warm_ = atoi(e) != 0 ? 1 : 0;
    if (const char* e = getenv("PROXSDP_B200_RITZ_BI")) lz_bi_ = atoi(e);
    g_h2d_bytes = 0; g_d2h_bytes = 0;
    st0.lap("ctor: done");
    setup_host(prob);
}

Solver::~Solver() {
    // the device blocks of this solver go back to the per-process cache (runtime.cu): nothing may still be running on them
    if (stream) cudaStreamSynchronize(stream);
    if (scal_host) pb_host_free(scal_host);
    if (feas_host_) pb_host_free(feas_host_);
    if (pub_seq_host_) pb_host_free(pub_seq_host_);
    if (ev_psd0_) cudaEventDestroy(ev_psd0_);
    if (ev_psd1_) cudaEventDestroy(ev_psd1_);
    for (cudaEvent_t e : {ev_lz0_, ev_lz1_, ev_post1_, ev_fl0_, ev_fl1_}) if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
}

// ---------------------------------------------------------------------------
// approx_norm = false: sigma_max(M) (reference src/pdhg.jl:107-118, Arpack.svds(M, nsv = 1))
// ---------------------------------------------------------------------------
// y[row_ids ? row_ids[w] : w] = sum_k val[k] x[colidx[k]] over the entries of (compact) row w; one warp per row
__global__ void k_spmv_warp_rows(const int* __restrict__ rowptr, int nrows, const int* __restrict__ row_ids,
                                 const int* __restrict__ colidx, const double* __restrict__ val,
                                 const double* __restrict__ x, double* __restrict__ y) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= nrows) return;
    const int lane = threadIdx.x & 31;
    double s0 = 0.0;
    for (int k = rowptr[w] + lane; k < rowptr[w + 1]; k += 32) s0 = fma(val[k], x[colidx[k]], s0);
    s0 = warp_sum(s0);
    if (lane == 0) y[row_ids ? row_ids[w] : (int)w] = s0;
}

// Largest singular value of the working M = sqrt(lambda_max(M M')): restarted Lanczos with full re-orthogonalisation on
// vectors of length R.  The two sparse products of every step run on the device (M' as DCSR into a zero-initialised
// n-vector, then M); the R-vectors, the small Rayleigh quotient and the restart live on the host — this is setup code
// (a few dozen products of 20 nnz bytes each), the reference spends it inside ARPACK.  Returns -1 when 60 restarts do
// not bring the residual below 1e-11 lambda (the caller keeps the Frobenius norm, like the reference's `catch`).
double Solver::spectral_norm_device() {
    if (R <= 0 || n <= 0 || M_.nnz <= 0) return 0.0;
    DBuf<double> v_d, t_d, w_d;
    v_d.alloc_raw((size_t)R); w_d.alloc_raw((size_t)R); t_d.alloc((size_t)n);
    std::vector<double> V0((size_t)R);
    eig_resid_default(R, 1234, 3, V0.data());
    auto matvec = [&](const double* vin, double* wout) {
        PB_CUDA(cudaMemcpyAsync(v_d.p, vin, sizeof(double) * (size_t)R, cudaMemcpyHostToDevice, stream));
        if (Mt_.n_nz > 0)
            k_spmv_warp_rows<<<ceil_div((long long)Mt_.n_nz * 32, 256), 256, 0, stream>>>(Mt_.rowptr.p, Mt_.n_nz, Mt_.nz_rows.p, Mt_.colidx.p,
                                                                                     Mt_.val.p, v_d.p, t_d.p);
        k_spmv_warp_rows<<<ceil_div((long long)R * 32, 256), 256, 0, stream>>>(M_.rowptr.p, (int)R, nullptr, M_.colidx.p, M_.val.p, t_d.p, w_d.p);
        launches += 2;
        PB_CUDA(cudaMemcpyAsync(wout, w_d.p, sizeof(double) * (size_t)R, cudaMemcpyDeviceToHost, stream));
        PB_CUDA(cudaStreamSynchronize(stream));
        g_h2d_bytes += 8 * R; g_d2h_bytes += 8 * R;
    };
    return lanczos_sigma_max(R, V0, matvec);
}

void Solver::setup_host(const proxsdp_problem_t* prob) {
    StageTimer st("setup: cone table + ingest");
    n = prob->n; p = prob->p; m = prob->m; R = p + m;
    const long long base = prob->index_base;
    n_sdp = (int)prob->n_sdp; n_soc = (int)prob->n_soc;
    if (n < 0 || p < 0 || m < 0 || (base != 0 && base != 1)) throw CudaError(-1, "invalid problem sizes");
    if ((opt.equilibration || opt.equilibration_force) && sharded())
        throw CudaError(-2, "equilibration needs the row norms of the whole constraint matrix: not available on the sharded path");
    if (n >= (1LL << 31) - 64 || R >= (1LL << 31) - 64) throw CudaError(-1, "problem too large for 32-bit indices");

    // norms of the right-hand sides before any scaling (pdhg.jl:14-15); ||c|| comes back from the device ingest
    auto nrm2 = [](const double* v, long long len) { double s = 0; for (long long i = 0; i < len; ++i) s += v[i] * v[i]; return std::sqrt(s); };
    norm_b_ = nrm2(prob->b, p); norm_h_ = nrm2(prob->h, m);

    // cone table in position order (preprocess!, scaling.jl:2-26: [PSD blocks | SOC blocks | free variables])
    cones.resize((size_t)n_sdp);
    soc_off_h_.resize((size_t)n_soc); soc_len_h_.resize((size_t)n_soc);
    std::vector<int> side_h((size_t)n_sdp);
    std::vector<long long> off_h((size_t)n_sdp);
    long long pos = 0;
    for (int k = 0; k < n_sdp; ++k) {
        const long long side = prob->sdp_side[k], tri = side * (side + 1) / 2;
        if (side < 1 || prob->sdp_ptr[k + 1] - prob->sdp_ptr[k] != tri || prob->sdp_ptr[k] - prob->sdp_ptr[0] != pos)
            throw CudaError(-3, "sdp cone length mismatch");
        cones[k].side = (int)side; cones[k].off = pos;
        side_h[k] = (int)side; off_h[k] = pos;
        pos += tri;
    }
    psd_end_ = pos;
    for (int k = 0; k < n_soc; ++k) {
        soc_off_h_[k] = pos;
        soc_len_h_[k] = (int)(prob->soc_ptr[k + 1] - prob->soc_ptr[k]);
        if (soc_len_h_[k] < 1 || prob->soc_ptr[k] - prob->soc_ptr[0] != pos - psd_end_) throw CudaError(-3, "soc cone length mismatch");
        pos += soc_len_h_[k];
    }
    listed_end_ = pos;
    if (listed_end_ > n) throw CudaError(-3, "the cones list more variables than the problem has");
    cone_side_d_.upload(side_h); cone_off_d_.upload(off_h);

    // ---- device ingest: permutation, sqrt(2)/2 scaling, M and M', ||c||, ||M||_F (runtime.cu)
    ConeTable ct{};
    ct.n_sdp = n_sdp; ct.side = side_h.data(); ct.off = off_h.data(); ct.psd_end = psd_end_; ct.listed_end = listed_end_;
    proxsdp_problem_t pr = *prob;
    if (n_sdp > 0) pr.sdp_idx = prob->sdp_idx + prob->sdp_ptr[0];
    if (n_soc > 0) pr.soc_idx = prob->soc_idx + prob->soc_ptr[0];
    IngestOut ing;
    ingest_problem(&pr, ct, cone_side_d_.p, cone_off_d_.p, !cones_only_, stream, ing);
    launches += ing.launches;
    identity_ = ing.identity;
    ord_d_ = std::move(ing.ord); var_ordering_d_ = std::move(ing.var_ordering);
    c_orig_d_ = std::move(ing.c_orig);
    M_ = std::move(ing.M); Mt_ = std::move(ing.Mt);
    norm_c_ = std::sqrt(ing.norm_c2);
    b_host_.assign(prob->b, prob->b + p); h_host_.assign(prob->h, prob->h + m);
    st.lap("setup: params + vectors");

    // diagonal preconditioning (pdhg.jl:64-93, equilibration.jl); like the reference this mutates opt.equilibration
    if ((opt.equilibration || opt.equilibration_force) && !cones_only_) {
        EquilibrateOut eq;
        opt.equilibration = equilibrate_device(M_, Mt_, n, R, opt, cone_off_d_.p, n_sdp, psd_end_, stream, eq) ? 1 : 0;
        launches += eq.launches;
        if (opt.equilibration) { eq_E_d_ = std::move(eq.E); eq_d_d_ = std::move(eq.d); ing.fro2 = eq.fro2; }
    } else {
        opt.equilibration = 0;
    }
    const bool equilibrated = opt.equilibration != 0;

    // step size: 1 / ||M||_F (pdhg.jl:121-133)
    double fro = ing.fro2;
    double any_cone = (n_soc > 0 || n_sdp > 0) ? 1.0 : 0.0, any_soc = n_soc > 0 ? 1.0 : 0.0;
    if (sharded()) {
        // whole-problem norms: sqrt of the rank-ordered sum of the per-rank sums of squares
        double sums[4] = {norm_b_ * norm_b_, norm_h_ * norm_h_, norm_c_ * norm_c_, fro};
        host_reduce(sums, 4, 0);
        norm_b_ = std::sqrt(sums[0]); norm_h_ = std::sqrt(sums[1]); norm_c_ = std::sqrt(sums[2]); fro = sums[3];
        double flags[2] = {any_cone, any_soc};
        host_reduce(flags, 2, 1);
        any_cone = flags[0]; any_soc = flags[1];
    }
    global_has_soc_ = any_soc != 0.0;
    fro = std::sqrt(fro);
    if (!opt.approx_norm && !cones_only_) {
        double sv = spectral_norm_device();
        if (sharded()) {
            // the shards are independent blocks (no row of one rank touches another rank's variables): M is block diagonal
            // up to a permutation and sigma_max(M) = max over the ranks; a failure on any rank fails all
            double v[2] = {sv, sv < 0.0 ? 1.0 : 0.0};
            host_reduce(v, 2, 1);
            sv = v[1] != 0.0 ? -1.0 : v[0];
        }
        if (sv >= 0.0) fro = sv;
        else fprintf(stderr, "    WARNING: Failed to compute spectral norm of M, shifting to Frobenius norm\n");
    }
    if (fro < 1e-10) fro = 1.0;
    primal_step_ = 1.0 / fro; primal_step_old_ = primal_step_; dual_step_ = primal_step_;
    const double cte = std::sqrt(2.0) / 2.0;

    // ---- Params (pdhg.jl:7-31) ----
    theta_ = opt.initial_theta; adapt_level_ = opt.initial_adapt_level; window_ = opt.convergence_window;
    beta_ = opt.initial_beta;
    long long r0 = opt.initial_target_rank > 0 ? opt.initial_target_rank : 2;
    target_rank.assign((size_t)n_sdp, r0); current_rank.assign((size_t)n_sdp, r0); min_eig.assign((size_t)n_sdp, 0.0);
    if (opt.max_iter <= 0) opt.max_iter_local = (any_cone != 0.0) ? opt.max_iter_conic : opt.max_iter_lp;
    else opt.max_iter_local = opt.max_iter;
    for (Circ* c : {&dual_gap_, &prim_obj_, &dual_obj_, &feasibility_, &primal_residual_, &dual_residual_, &comb_residual_})
        c->init(2 * window_);

    // ---- device vectors ----
    // working objective = c in position order with norm_scaling applied (scaling.jl:28-58: off-diagonal svec entries
    // *= sqrt(2)/2); the un-scaled copy stays for the dual cone of the result (pdhg.jl:701-710)
    c_.alloc_raw((size_t)n);
    if (n > 0) {
        const double* c_src = c_orig_d_.p;
        if (equilibrated) { launch_eq_mul_scalar(c_orig_d_.p, eq_d_d_.p, n, c_.p, stream); launches++; c_src = c_.p; }    // c = D c (pdhg.jl:86)
        k_scale_offdiag_copy<<<std::max(1, std::min(num_sms_ * 8, ceil_div(n, 256))), 256, 0, stream>>>(
            c_src, n, psd_end_, cone_off_d_.p, n_sdp, cte, 0, c_.p);
        launches++;
    }
    b_.upload(b_host_); h_.upload(h_host_);
    b_orig_d_.upload(b_host_); h_orig_d_.upload(h_host_);
    if (equilibrated) {      // rhs = E rhs (pdhg.jl:83-85); the uploads above went through the legacy stream: ordered
        launch_eq_mul(b_.p, eq_E_d_.p, p, b_.p, stream);
        launch_eq_mul(h_.p, eq_E_d_.p + p, m, h_.p, stream);
        launches += (p > 0) + (m > 0);
    }
    for (int q = 0; q < 2; ++q) { x_[q].alloc((size_t)n); Mty_[q].alloc((size_t)n); y_[q].alloc((size_t)R); Mx_[q].alloc((size_t)R); }
    if (n > 0 && !getenv("PROXSDP_B200_NO_NZMASK")) {
        // (Mty is only ever written at the non-empty rows of M', c is fixed: the bitmap is built once)
        nzmask_.alloc((size_t)((n + 31) / 32 + 1));
        k_mask_from_c<<<std::max(1, std::min(num_sms_ * 8, ceil_div(n, 256))), 256, 0, stream>>>(c_.p, n, nzmask_.p);
        launches++;
        if (Mt_.n_nz > 0) { k_mask_from_rows<<<ceil_div(Mt_.n_nz, 256), 256, 0, stream>>>(Mt_.nz_rows.p, Mt_.n_nz, nzmask_.p); launches++; }
    }
    st.lap("setup: cones");

    // cones
    long long kmax = 2 * std::max<long long>(opt.max_target_rank_krylov_eigs, r0) + 1;
    kmax = std::max<long long>(kmax, opt.eigsolver_min_lanczos);
    Kmax_ = (int)kmax;
    long long roff = 0;
    for (int k = 0; k < n_sdp; ++k) {
        ConeDev& cd = cones[k];
        // Cones of side <= SMALL_CONE_MAX take the fused one-CTA full-eigendecomposition kernel — unless the options make
        // them Krylov-capable (min_size_krylov_eigs below the side, prox_operators.jl:46-49): those are laid out as
        // large cones so that the reference's truncated projection, target-rank and min_eig tracking apply to them.
        const bool krylov_capable = !opt.full_eig_decomp && cd.side > opt.min_size_krylov_eigs && cd.side > 1;
        cd.small = cd.side <= SMALL_CONE_MAX && !force_large_ && !krylov_capable;
        cd.ld = (cd.side + 63) & ~63;     // multiple of the block-Jacobi pivot size (no dummy blocks)
        if (cd.small) {
            small_ids_.push_back(k);
            max_small_side_ = std::max(max_small_side_, cd.side);
        } else {
            large_ids_.push_back(k);
            cd.X.alloc((size_t)cd.ld * cd.ld);
            cd.Y.alloc((size_t)cd.ld * (size_t)(Kmax_ + 1));
            cd.vals.alloc((size_t)cd.ld);
            cd.kept_lam.alloc((size_t)cd.ld);
            cd.kept_idx.alloc((size_t)cd.ld);
            cd.info.alloc(8);
            cd.nkept.alloc(1);
            std::vector<double> rs((size_t)cd.side);
            if (prob->eig_resid) std::copy(prob->eig_resid + roff, prob->eig_resid + roff + cd.side, rs.begin());
            else eig_resid_default(cd.side, opt.eigsolver_resid_seed, opt.krylovkit_resid_init, rs.data());
            cd.resid.upload(rs);
        }
        roff += cd.side;
    }
    st.lap("setup: misc buffers+attrs");
    small_ids_d_.upload(small_ids_);
    soc_off_d_.upload(soc_off_h_); soc_len_d_.upload(soc_len_h_); soc_gap_d_.alloc((size_t)std::max(n_soc, 1));
    out_min_d_.alloc((size_t)std::max(n_sdp, 1));
    offnorm_d_.alloc(2);

    scal_len = S_HEADER + 3 * n_sdp;
    scal_d_.alloc((size_t)scal_len);
    scal_host = static_cast<double*>(pb_host_alloc(sizeof(double) * (size_t)scal_len));
    pub_seq_host_ = static_cast<unsigned long long*>(pb_host_alloc(64));
    *pub_seq_host_ = 0;
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&scal_host_dev_), scal_host, 0) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&pub_seq_dev_), pub_seq_host_, 0) != cudaSuccess) {
        cudaGetLastError();
        spin_sync_ = false;
    }
    if (const char* e = getenv("PROXSDP_B200_SYNC")) spin_sync_ = spin_sync_ && std::string(e) != "memcpy";
    if (const char* e = getenv("PROXSDP_B200_LADDER_FUSED")) fused_ladder_ = atoi(e) != 0 ? 1 : 0;
    scal_scratch_d_.alloc((size_t)scal_len);
    scal_target_ = scal_d_.p;
    feas_d_.alloc(8);
    feas_host_ = static_cast<double*>(pb_host_alloc(sizeof(double) * (size_t)(8 + std::max(n_sdp, 1))));
    reduce_blocks_ = num_sms_ * 8;
    partials_d_.alloc((size_t)reduce_blocks_ * 8);
    counters_d_.alloc(16);
    ws_.partials = partials_d_.p; ws_.counters = counters_d_.p; ws_.max_blocks = reduce_blocks_;

    if (max_small_side_ > 1) {
        small_fast_ = small_cone_smem_bytes(max_small_side_, 1) <= smem_optin_ ? 1 : 0;
        if (const char* e = getenv("PROXSDP_B200_SMALL_FAST")) small_fast_ = (atoi(e) != 0 && small_fast_) ? 1 : 0;
        if (const char* e = getenv("PROXSDP_B200_SMALL_WARM")) small_warm_ = atoi(e) != 0 ? 1 : 0;
        size_t sb = small_cone_smem_bytes(max_small_side_, small_fast_);
        if (sb > smem_optin_) throw CudaError(-4, "small-cone kernel needs more shared memory than the device offers");
        PB_CUDA(cudaFuncSetAttribute(k_small_cone_proj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
    }
    PB_CUDA(cudaFuncSetAttribute(k_lanczos, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin_));
    {
        cudaFuncAttributes fa{};
        PB_CUDA(cudaFuncGetAttributes(&fa, k_lanczos_cl));
        lz_cl_smem_max_ = smem_optin_ - fa.sharedSizeBytes;
        PB_CUDA(cudaFuncSetAttribute(k_lanczos_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lz_cl_smem_max_));
    }
    // the seven instances of the cl3 kernel have the same static shared memory; their attributes are set once per
    // process and device (7 x get + set cost ~10 ms per solver construction otherwise)
    {
        static thread_local int attr_dev = -1;
        static thread_local size_t attr_smem_max = 0;
        static thread_local bool attr_nonportable = false;
        if (attr_dev != dev_ || (lz_cluster_ > 8 && !attr_nonportable)) {
            size_t mx = smem_optin_;
            std::vector<const void*> fns;
            for (int cpw = 1; cpw <= 8; ++cpw) if (const void* fn = lanczos_cl3_kernel(cpw)) fns.push_back(fn);
            fns.push_back((const void*)k_lanczos_cl3<1, 1, 1, 1>);
            fns.push_back((const void*)k_lanczos_cl3<1, 1, 1, 2>);
            for (const void* fn : fns) {
                cudaFuncAttributes fa{};
                PB_CUDA(cudaFuncGetAttributes(&fa, fn));
                mx = std::min(mx, smem_optin_ - fa.sharedSizeBytes);
            }
            for (const void* fn : fns) {
                PB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
                if (lz_cluster_ > 8) PB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            }
            // the resident single-cluster variant may use a cluster of 16 CTAs
            if (cudaFuncSetAttribute((const void*)k_lanczos_cl3<1, 1, 1, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) cudaGetLastError();
            attr_dev = dev_; attr_smem_max = mx; attr_nonportable = lz_cluster_ > 8;
        }
        lz_cl3_smem_max_ = attr_smem_max;
    }
    if (lz_cluster_ > 8) PB_CUDA(cudaFuncSetAttribute(k_lanczos_cl, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    st.lap("done");
    PB_CUDA(cudaFuncSetAttribute(k_bj_pair_eig, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bj_pair_smem_bytes()));
    PB_CUDA(cudaFuncSetAttribute(k_bj_apply<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BJ_APPLY_SMEM));
    PB_CUDA(cudaFuncSetAttribute(k_bj_apply<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BJ_APPLY_SMEM));
}

// ---------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------
void Solver::launch_spmv(const CsrDev& A, const double* x, double* y) {
    if (A.nrows == 0) return;
    const double* poison = scal_d_.p + S_POISON;
    long long threads = (long long)A.nrows * A.group;
    int blocks = ceil_div(threads, 256);
    switch (A.group) {
        case 1: k_spmv_csr<1><<<blocks, 256, 0, stream>>>(A.nrows, A.rowptr.p, A.colidx.p, A.val.p, x, y, A.long_threshold, poison); break;
        case 2: k_spmv_csr<2><<<blocks, 256, 0, stream>>>(A.nrows, A.rowptr.p, A.colidx.p, A.val.p, x, y, A.long_threshold, poison); break;
        case 4: k_spmv_csr<4><<<blocks, 256, 0, stream>>>(A.nrows, A.rowptr.p, A.colidx.p, A.val.p, x, y, A.long_threshold, poison); break;
        case 8: k_spmv_csr<8><<<blocks, 256, 0, stream>>>(A.nrows, A.rowptr.p, A.colidx.p, A.val.p, x, y, A.long_threshold, poison); break;
        case 16: k_spmv_csr<16><<<blocks, 256, 0, stream>>>(A.nrows, A.rowptr.p, A.colidx.p, A.val.p, x, y, A.long_threshold, poison); break;
        default: k_spmv_csr<32><<<blocks, 256, 0, stream>>>(A.nrows, A.rowptr.p, A.colidx.p, A.val.p, x, y, A.long_threshold, poison); break;
    }
    launches++;
    if (A.n_long > 0) {
        k_spmv_long<<<A.n_long, 512, 0, stream>>>(A.long_rows.p, A.rowptr.p, A.colidx.p, A.val.p, x, y, poison);
        launches++;
    }
}

bool Solver::krylov_eligible(int k, long long iter) const {
    // prox_operators.jl:46-49
    return !opt.full_eig_decomp && target_rank[(size_t)k] <= opt.max_target_rank_krylov_eigs &&
           cones[(size_t)k].side > opt.min_size_krylov_eigs && (iter % opt.full_eig_freq) > opt.full_eig_len;
}

bool Solver::lanczos_launch_cluster(ConeDev& cd, int cone_idx, int nev, int K, int maxiter, double tol) {
    if (lz_mode_ == 1 || K > LZC_KMAX) return false;
    const int nside = cd.side;
    if (nside < 64) return false;              // tiny cones (min_size_krylov_eigs lowered by the user): one CTA suffices, row kernel
    const int C = lz_cluster_;
    const int vn_max = (nside + C - 1) / C;
    // grid: as many clusters as can be co-resident, but no more CTAs than there are 8-row slabs
    int want = std::max(C, ((std::max(1, (nside + 7) / 8) + C - 1) / C) * C);
    cudaLaunchConfig_t cfg{};
    // The kernel spins on words written by other CTAs, so the whole grid has to be co-resident: the cooperative attribute
    // makes the launch FAIL (instead of spinning into its 2 s time-out) when another tenant of the device holds SMs.
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.blockDim = dim3(LZ_THREADS); cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = lz_coop_ ? 2 : 1;
    // shared memory depends on rows_max = ceil(n / G); G depends on occupancy, which depends on shared memory:
    // size for the smallest plausible grid first (largest rows_max), then shrink
    int G = std::min(want, (num_sms_ / C) * C);
    if (G < C) return false;
    int rows_max = (nside + G - 1) / G;
    size_t smem = lanczos_cl_smem_bytes(K, rows_max, vn_max, nside, C);
    if (smem > lz_cl_smem_max_) return false;
    cfg.gridDim = dim3((unsigned)G); cfg.dynamicSmemBytes = smem;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, k_lanczos_cl, &cfg) != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return false;
    }
    if (max_clusters * C < G) {
        G = max_clusters * C;
        rows_max = (nside + G - 1) / G;
        smem = lanczos_cl_smem_bytes(K, rows_max, vn_max, nside, C);
        if (smem > lz_cl_smem_max_) return false;
        cfg.gridDim = dim3((unsigned)G); cfg.dynamicSmemBytes = smem;
    }
    if ((rows_max + LZ_NW - 1) / LZ_NW + 1 > LZ_TMAX) return false;
    if ((size_t)cd.ld * (size_t)(K + 1) > cd.Y.n) cd.Y.alloc((size_t)cd.ld * (size_t)(K + 1));
    if (lz_wg_.n < (size_t)2 * cd.ld) lz_wg_.alloc((size_t)2 * cd.ld);
    if (lz_bar_.n == 0) lz_bar_.alloc(4);
    const size_t ws_len = 1 + (size_t)lanczos_kp(Kmax_ > K ? Kmax_ : K) * (size_t)lanczos_kp(Kmax_ > K ? Kmax_ : K);
    for (int q = 0; q < 2; ++q) if (cd.ritz_ws[q].n < ws_len) { cd.ritz_ws[q].alloc(ws_len); cd.ritz_launches = 0; }
    PB_CUDA(cudaMemsetAsync(lz_bar_.p, 0, sizeof(unsigned int), stream));
    LanczosClArgs a{};
    a.X = cd.X.p; a.n = nside; a.ld = cd.ld; a.x0 = cd.resid.p; a.Y = cd.Y.p;
    a.wg = lz_wg_.p; a.bar = lz_bar_.p;
    const int flip = (int)(cd.ritz_launches & 1);
    // every 32nd eigsolve restarts the Ritz basis from the identity so rounding in the accumulated
    // rotations cannot build up over a long solve
    a.ritz_rd = (lz_warm_ && cd.ritz_launches > 0 && (cd.ritz_launches % 32) != 0) ? cd.ritz_ws[flip].p : nullptr;
    a.ritz_wr = lz_warm_ ? cd.ritz_ws[1 - flip].p : nullptr;
    cd.ritz_launches++;
    a.nev = nev; a.K = K; a.maxiter = maxiter; a.tol = tol;
    a.rows_max = rows_max; a.vn_max = vn_max; a.use_bi = lz_bi_;
    a.vals = cd.vals.p; a.info = cd.info.p; a.scal = scal_target_; a.cone = cone_idx;
    if (getenv("PROXSDP_B200_LZ_PROF")) { if (lz_prof_.n == 0) lz_prof_.alloc(8 + 8 * 256); a.prof = lz_prof_.p; }
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_lanczos_cl, a);
    if (e != cudaSuccess && lz_coop_ && (e == cudaErrorNotSupported || e == cudaErrorInvalidValue)) {
        cudaGetLastError();
        lz_coop_ = 0; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, k_lanczos_cl, a);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (!lz_cluster_warned_) { fprintf(stderr, "[proxsdp_b200] cluster Lanczos launch failed (%s); using the row-distributed kernel\n", cudaGetErrorString(e)); lz_cluster_warned_ = true; }
        lz_mode_ = 1;
        return false;
    }
    launches += 1;
    lz_cluster_launches_++;
    return true;
}

// third-generation cluster kernel (lanczos_cl3.cuh): fused alpha + one Gram-Schmidt pass, strip symv with X partly
// resident in shared memory, spill-free step.  Returns false when the configuration does not fit; the caller then
// tries the second-generation kernel.
bool Solver::lanczos_launch_cluster3(ConeDev& cd, int cone_idx, int nev, int K, int maxiter, double tol) {
    if (lz_mode_ == 1 || lz_kernel_ != 3 || K > LZC_KMAX) return false;
    const int nside = cd.side;
    if (nside < 64) return false;
    const int C = lz_cluster_;
    const int vn_max = (nside + C - 1) / C;
    const int cpr = lanczos_cpr(nside);
    int cpw = (cpr + LZ_NW - 1) / LZ_NW;
    if (cpw == 7) cpw = 8;
    if (cpw > 8) return false;                 // side > 8192: the strip symv has no instantiation that wide
    if ((vn_max + LZ_NW - 1) / LZ_NW > 64) return false;      // the gather keeps two rows per lane in registers
    int want = std::max(C, ((std::max(1, (nside + 7) / 8) + C - 1) / C) * C);
    cudaLaunchConfig_t cfg{};
    // The kernel spins on words written by other CTAs, so the whole grid has to be co-resident: the cooperative attribute
    // makes the launch FAIL (instead of spinning into its 2 s time-out) when another tenant of the device holds SMs.
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.blockDim = dim3(LZ_THREADS); cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = lz_coop_ ? 2 : 1;
    int G = std::min(std::min(want, (num_sms_ / C) * C), (LZ3_GMAX / C) * C);
    if (G < C) return false;
    // resident slab rows: whatever shared memory the basis replica and the Ritz scratch leave, at most the shortest slab
    const size_t rowbytes = (size_t)cpr * 64 * sizeof(double);
    auto fit_resident = [&](int grid, Lz3Layout& Lout, size_t& bytes) -> bool {
        Lz3Layout L0 = lanczos_cl3_layout(K, 0, vn_max, nside, C);
        if ((size_t)L0.total * sizeof(double) > lz_cl3_smem_max_) return false;
        long long nres = (long long)((lz_cl3_smem_max_ - (size_t)L0.total * sizeof(double)) / rowbytes);
        nres = std::min<long long>(nres, nside / grid);
        if (lz_xres_ >= 1) nres = std::min<long long>(nres, lz_xres_ - 1);      // PROXSDP_B200_LZ_XRES = rows + 1: cap for experiments
        Lout = lanczos_cl3_layout(K, (int)std::max<long long>(nres, 0), vn_max, nside, C);
        bytes = (size_t)Lout.total * sizeof(double);
        return true;
    };
    Lz3Layout L{};
    size_t smem = 0;
    if (!fit_resident(G, L, smem)) return false;
    cfg.gridDim = dim3((unsigned)G); cfg.dynamicSmemBytes = smem;
    const void* kfn = lanczos_cl3_kernel(cpw);
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kfn, &cfg) != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return false;
    }
    if (max_clusters * C < G) {
        G = max_clusters * C;
        if (!fit_resident(G, L, smem)) return false;
    }
    cfg.gridDim = dim3((unsigned)G); cfg.dynamicSmemBytes = smem;
    const int rows_max = (nside + G - 1) / G;
    if ((size_t)cd.ld * (size_t)(K + 1) > cd.Y.n) cd.Y.alloc((size_t)cd.ld * (size_t)(K + 1));
    // flagged-exchange buffers (shared by all cones of this solver; launches are stream ordered); tags are unique over
    // launches, freshly zeroed buffers carry tag 0, which is never used
    const unsigned long long bound = (unsigned long long)K * (unsigned long long)(std::max(maxiter, 1) + 1) + 16ULL;
    if (lz3_wg_.n < (size_t)2 * cd.ld || lz3_apart_.n == 0 || lz3_epoch_ + bound >= 0xFFFFFFF0ULL) {
        PB_CUDA(cudaStreamSynchronize(stream));
        if (lz3_wg_.n < (size_t)2 * cd.ld) lz3_wg_.alloc((size_t)2 * cd.ld); else PB_CUDA(cudaMemset(lz3_wg_.p, 0, lz3_wg_.n * sizeof(uint4)));
        if (lz3_apart_.n == 0) lz3_apart_.alloc(2 * LZ3_GMAX); else PB_CUDA(cudaMemset(lz3_apart_.p, 0, lz3_apart_.n * sizeof(uint4)));
        lz3_epoch_ = 0;
    }
    const size_t ws_len = 1 + (size_t)lanczos_kp(Kmax_ > K ? Kmax_ : K) * (size_t)lanczos_kp(Kmax_ > K ? Kmax_ : K);
    for (int q = 0; q < 2; ++q) if (cd.ritz_ws[q].n < ws_len) { cd.ritz_ws[q].alloc(ws_len); cd.ritz_launches = 0; }
    LanczosCl3Args a{};
    a.X = cd.X.p; a.n = nside; a.ld = cd.ld; a.x0 = cd.resid.p; a.Y = cd.Y.p;
    a.wg = lz3_wg_.p; a.apart = lz3_apart_.p; a.epoch_base = (unsigned int)lz3_epoch_;
    a.res_begin_off = L.nres;
    const int flip = (int)(cd.ritz_launches & 1);
    a.ritz_rd = (lz_warm_ && cd.ritz_launches > 0 && (cd.ritz_launches % 32) != 0) ? cd.ritz_ws[flip].p : nullptr;
    a.ritz_wr = lz_warm_ ? cd.ritz_ws[1 - flip].p : nullptr;
    a.nev = nev; a.K = K; a.maxiter = maxiter; a.tol = tol;
    a.vn_max = vn_max; a.use_bi = lz_bi_; a.stop_above = lz_stop_above_; a.strict = lz_strict_; a.eager = (opt.krylovkit_eager && lz_stop_above_ >= 1e300) ? 1 : 0; a.poll_ns = lz_poll_ns_; a.arrow_restart = lz_arrow_; a.debug = getenv("PROXSDP_B200_LZ_DEBUG") ? 1 : 0;
    a.rbase = nside / G; a.rrem = nside % G; a.vbase = nside / C; a.vrem = nside % C;
    a.vals = cd.vals.p; a.info = cd.info.p; a.scal = scal_target_; a.cone = cone_idx;
    a.L = L; a.spin_limit = lz_spin_limit_; a.bi_memory = lz_bi_memory_;
    // a matrix that does not stay in L2 between two mat-vecs is pulled in by the TMA engine a few rows ahead of the loads
    {
        // whole slab when the matrix fits L2 (the loads then hit the near L2 partition); a window of ~16 MB over the grid otherwise
        const size_t mbytes = (size_t)nside * (size_t)cd.ld * sizeof(double);
        int pf = 0;
        if (cpw >= 3) pf = mbytes <= (size_t)l2_bytes_ * 3 / 4 ? 32 : (int)std::min<size_t>(32, std::max<size_t>(2, ((size_t)16 << 20) / (rowbytes * (size_t)G) + 1));
        a.pf_rows = lz_pf_ >= 0 ? std::min(lz_pf_, 32) : pf;
    }
    if (getenv("PROXSDP_B200_LZ_PROF")) { if (lz_prof_.n == 0) lz_prof_.alloc(8 + 8 * 256); a.prof = lz_prof_.p; }
    void* kargs[] = {&a};
    cudaError_t e = cudaLaunchKernelExC(&cfg, kfn, kargs);
    if (e != cudaSuccess && lz_coop_ && (e == cudaErrorNotSupported || e == cudaErrorInvalidValue)) {
        // this driver does not combine the cooperative attribute with cluster launches: plain launch, co-residency is
        // then only what cudaOccupancyMaxActiveClusters promised (and the in-kernel time-outs remain the guard)
        cudaGetLastError();
        lz_coop_ = 0; cfg.numAttrs = 1;
        e = cudaLaunchKernelExC(&cfg, kfn, kargs);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (!lz_cluster_warned_) { fprintf(stderr, "[proxsdp_b200] cl3 Lanczos launch failed (%s); using the second-generation kernel\n", cudaGetErrorString(e)); lz_cluster_warned_ = true; }
        lz_kernel_ = 2;
        return false;
    }
    if (getenv("PROXSDP_B200_DEBUG") && cd.ritz_launches == 0)
        fprintf(stderr, "[lanczos] cl3 kernel: side %d, K %d, grid %d x %d, strip width %d chunks, %d of %d slab rows resident in shared memory, %zu KB shared memory\n",
                nside, K, G / C, C, cpw, L.nres, rows_max, smem >> 10);
    cd.ritz_launches++;
    lz3_epoch_ += bound;
    launches += 1;
    lz_cluster_launches_++;
    return true;
}

// Sparsity pattern of S = mat(M'y + c) inside one cone, as CSR by matrix row (built once per solve, on first use): entry
// (row, col, svec position, coefficient) with coefficient 1 on the diagonal and 1/sqrt(2) off it (prox_operators.jl:1-16).
// M'y + c vanishes outside supp(c) ∪ {non-empty rows of M'}, whatever y is.
bool Solver::implicit_pattern(ConeDev& cd) {
    if (cd.imp_tried) return cd.imp_ready;
    cd.imp_tried = true;
    const long long tri = (long long)cd.side * (cd.side + 1) / 2;
    std::vector<double> cblk((size_t)tri);
    PB_CUDA(cudaMemcpy(cblk.data(), c_orig_d_.p + cd.off, sizeof(double) * (size_t)tri, cudaMemcpyDeviceToHost));
    g_d2h_bytes += (long long)sizeof(double) * tri;
    std::vector<int> nzr((size_t)Mt_.n_nz);
    if (Mt_.n_nz > 0) PB_CUDA(cudaMemcpy(nzr.data(), Mt_.nz_rows.p, sizeof(int) * (size_t)Mt_.n_nz, cudaMemcpyDeviceToHost));
    std::vector<long long> pos;
    for (long long k = 0; k < tri; ++k) if (cblk[(size_t)k] != 0.0) pos.push_back(k);
    for (int r : nzr) if (r >= cd.off && r < cd.off + tri) pos.push_back((long long)r - cd.off);
    std::sort(pos.begin(), pos.end());
    pos.erase(std::unique(pos.begin(), pos.end()), pos.end());
    if ((long long)pos.size() * 2 > (1LL << 30)) return false;
    const int nside = cd.side;
    std::vector<int> rp((size_t)nside + 1, 0);
    struct E { int row, col, pos; double coef; };
    std::vector<E> ent;
    ent.reserve(pos.size() * 2);
    const double isq2 = 1.0 / std::sqrt(2.0);
    for (long long k : pos) {
        long long j = (long long)((std::sqrt(8.0 * (double)k + 1.0) - 1.0) * 0.5);
        while ((j + 1) * (j + 2) / 2 <= k) ++j;
        while (j * (j + 1) / 2 > k) --j;
        const long long i = k - j * (j + 1) / 2;
        if (i == j) ent.push_back({(int)i, (int)j, (int)k, 1.0});
        else { ent.push_back({(int)i, (int)j, (int)k, isq2}); ent.push_back({(int)j, (int)i, (int)k, isq2}); }
    }
    for (const E& e : ent) rp[(size_t)e.row + 1]++;
    for (int r = 0; r < nside; ++r) rp[(size_t)r + 1] += rp[(size_t)r];
    std::vector<int> next(rp.begin(), rp.end() - 1), col(ent.size()), ps(ent.size());
    std::vector<double> cf(ent.size());
    for (const E& e : ent) { const int d = next[(size_t)e.row]++; col[(size_t)d] = e.col; ps[(size_t)d] = e.pos; cf[(size_t)d] = e.coef; }
    // entries per CTA of the cluster split (rows v0 .. v0 + vn)
    const int C = lz_cluster_;
    int cap = 0;
    for (int c = 0; c < C; ++c) {
        const int vb = nside / C, vr = nside % C;
        const int v0 = c * vb + std::min(c, vr), vn = vb + (c < vr ? 1 : 0);
        cap = std::max(cap, rp[(size_t)(v0 + vn)] - rp[(size_t)v0]);
    }
    cd.imp_cap = std::max(cap, 1);
    cd.imp_rowptr.upload(rp); cd.imp_col.upload(col); cd.imp_pos.upload(ps); cd.imp_coef.upload(cf);
    cd.imp_ready = true;
    return true;
}

// One eigsolve on the implicit operator  Y diag(lam) Y' - tau mat(M'y + c): a single thread-block cluster, no grid
// exchange, no dense matrix (lanczos_cl3.cuh, IMP = true).  Returns false when the configuration does not fit.
bool Solver::lanczos_launch_implicit(ConeDev& cd, int cone_idx, int nev, int K, int maxiter, double tol, double tau) {
    if (lz_mode_ == 1 || lz_kernel_ != 3 || K > LZC_KMAX || cd.side < 64) return false;
    if (!implicit_pattern(cd)) return false;
    const int nside = cd.side;
    const int C = lz_cluster_;
    const int vn_max = (nside + C - 1) / C;
    if ((vn_max + LZ_NW - 1) / LZ_NW > 64) return false;
    // (the previous projection kept at most its own target rank pairs, which never exceeds the current one)
    Lz3Layout L = lanczos_cl3_layout(K, 0, vn_max, nside, C, cd.imp_cap, std::min(nev, LZ3_RMAX));
    const size_t smem = (size_t)L.total * sizeof(double);
    if (smem > lz_cl3_smem_max_) {
        if (getenv("PROXSDP_B200_DEBUG")) fprintf(stderr, "[lanczos] implicit operator: %zu KB of shared memory needed, %zu available: dense path\n", smem >> 10, lz_cl3_smem_max_ >> 10);
        return false;
    }
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.blockDim = dim3(LZ_THREADS); cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = lz_coop_ ? 2 : 1;
    cfg.gridDim = dim3((unsigned)C); cfg.dynamicSmemBytes = smem;
    if ((size_t)cd.ld * (size_t)(K + 1) > cd.Y.n) return false;      // (Y must not be re-allocated: it holds the operator)
    const size_t ws_len = 1 + (size_t)lanczos_kp(Kmax_ > K ? Kmax_ : K) * (size_t)lanczos_kp(Kmax_ > K ? Kmax_ : K);
    for (int q = 0; q < 2; ++q) if (cd.ritz_ws[q].n < ws_len) { cd.ritz_ws[q].alloc(ws_len); cd.ritz_launches = 0; }
    LanczosCl3Args a{};
    a.X = nullptr; a.n = nside; a.ld = cd.ld; a.x0 = cd.resid.p; a.Y = cd.Y.p;
    a.wg = nullptr; a.apart = nullptr; a.epoch_base = 0; a.res_begin_off = 0;
    const int flip = (int)(cd.ritz_launches & 1);
    a.ritz_rd = (lz_warm_ && cd.ritz_launches > 0 && (cd.ritz_launches % 32) != 0) ? cd.ritz_ws[flip].p : nullptr;
    a.ritz_wr = lz_warm_ ? cd.ritz_ws[1 - flip].p : nullptr;
    a.nev = nev; a.K = K; a.maxiter = maxiter; a.tol = tol;
    a.vn_max = vn_max; a.use_bi = lz_bi_; a.stop_above = 1e300; a.strict = lz_strict_; a.eager = (opt.krylovkit_eager && lz_stop_above_ >= 1e300) ? 1 : 0; a.poll_ns = 0; a.arrow_restart = lz_arrow_; a.debug = 0;
    a.rbase = nside / C; a.rrem = nside % C; a.vbase = nside / C; a.vrem = nside % C;
    a.vals = cd.vals.p; a.info = cd.info.p; a.scal = scal_target_; a.cone = cone_idx;
    a.L = L; a.spin_limit = lz_spin_limit_; a.bi_memory = lz_bi_memory_;
    a.imp_Y = cd.Y.p; a.imp_kept_idx = cd.kept_idx.p; a.imp_kept_lam = cd.kept_lam.p; a.imp_nkept = cd.nkept.p;
    a.imp_rowptr = cd.imp_rowptr.p; a.imp_col = cd.imp_col.p; a.imp_pos = cd.imp_pos.p; a.imp_coef = cd.imp_coef.p;
    a.imp_Mty = Mty_[cur_].p + cd.off; a.imp_c = c_.p + cd.off; a.imp_tau = tau;
    const void* kfn = (const void*)k_lanczos_cl3<1, 1, 1, 1>;
    void* kargs[] = {&a};
    cudaError_t e = cudaLaunchKernelExC(&cfg, kfn, kargs);
    if (e != cudaSuccess && lz_coop_ && (e == cudaErrorNotSupported || e == cudaErrorInvalidValue)) {
        cudaGetLastError();
        cfg.numAttrs = 1;
        e = cudaLaunchKernelExC(&cfg, kfn, kargs);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (getenv("PROXSDP_B200_DEBUG")) fprintf(stderr, "[lanczos] implicit operator: launch failed (%s): dense path\n", cudaGetErrorString(e));
        return false;
    }
    cd.ritz_launches++;
    launches += 1;
    lz_cluster_launches_++;
    implicit_calls++;
    return true;
}

// Mid-size cones: the whole matrix fits the distributed shared memory of ONE cluster — CTA rank c keeps
// rows [c n/C, (c+1) n/C) of X next to its slice of the Krylov basis for the whole eigsolve (staged once by the TMA
// engine), so a mat-vec is a shared-memory GEMV and every exchange of the step stays inside the cluster (DSMEM): no
// grid-wide exchange through L2 at all.  Cluster size 8 up to side ~420, 16 (non-portable) up to side ~580 in FP64.
bool Solver::lanczos_launch_resident(ConeDev& cd, int cone_idx, int nev, int K, int maxiter, double tol) {
    if (!lz_resident_ || lz_mode_ == 1 || lz_kernel_ != 3 || K > LZC_KMAX || cd.side < 64) return false;
    const int nside = cd.side;
    const void* kfn = (const void*)k_lanczos_cl3<1, 1, 1, 2>;
    const int cpr = lanczos_cpr(nside);
    for (int C : {8, 16}) {
        if (C == 16 && !lz_resident16_) break;
        if (C == 8 && !lz_resident8_) continue;
        const int vn_max = (nside + C - 1) / C;
        if (vn_max > 4 * LZ_NW) continue;                 // the resident GEMV keeps at most 4 rows per warp
        Lz3Layout L = lanczos_cl3_layout(K, vn_max, vn_max, nside, C, 0, 1);
        const size_t smem = (size_t)L.total * sizeof(double);
        if (smem > lz_cl3_smem_max_) continue;
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeCooperative;
        attr[1].val.cooperative = 1;
        cfg.blockDim = dim3(LZ_THREADS); cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = lz_coop_ ? 2 : 1;
        cfg.gridDim = dim3((unsigned)C); cfg.dynamicSmemBytes = smem;
        if (lz_resident_ok_[C == 16] == 0) {              // first use of this cluster size: can one such cluster be resident at all?
            int max_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, kfn, &cfg) != cudaSuccess || max_clusters < 1) { cudaGetLastError(); lz_resident_ok_[C == 16] = -1; }
            else lz_resident_ok_[C == 16] = 1;
        }
        if (lz_resident_ok_[C == 16] < 0) continue;
        if ((size_t)cd.ld * (size_t)(K + 1) > cd.Y.n) cd.Y.alloc((size_t)cd.ld * (size_t)(K + 1));
        const size_t ws_len = 1 + (size_t)lanczos_kp(Kmax_ > K ? Kmax_ : K) * (size_t)lanczos_kp(Kmax_ > K ? Kmax_ : K);
        for (int q = 0; q < 2; ++q) if (cd.ritz_ws[q].n < ws_len) { cd.ritz_ws[q].alloc(ws_len); cd.ritz_launches = 0; }
        LanczosCl3Args a{};
        a.X = cd.X.p; a.n = nside; a.ld = cd.ld; a.x0 = cd.resid.p; a.Y = cd.Y.p;
        a.wg = nullptr; a.apart = nullptr; a.epoch_base = 0; a.res_begin_off = 0;
        const int flip = (int)(cd.ritz_launches & 1);
        a.ritz_rd = (lz_warm_ && cd.ritz_launches > 0 && (cd.ritz_launches % 32) != 0) ? cd.ritz_ws[flip].p : nullptr;
        a.ritz_wr = lz_warm_ ? cd.ritz_ws[1 - flip].p : nullptr;
        a.nev = nev; a.K = K; a.maxiter = maxiter; a.tol = tol;
        a.vn_max = vn_max; a.use_bi = lz_bi_; a.stop_above = lz_stop_above_; a.strict = lz_strict_; a.eager = (opt.krylovkit_eager && lz_stop_above_ >= 1e300) ? 1 : 0; a.poll_ns = 0; a.arrow_restart = lz_arrow_;
        a.debug = getenv("PROXSDP_B200_LZ_DEBUG") ? 1 : 0;
        a.rbase = nside / C; a.rrem = nside % C; a.vbase = nside / C; a.vrem = nside % C;
        a.vals = cd.vals.p; a.info = cd.info.p; a.scal = scal_target_; a.cone = cone_idx;
        a.L = L; a.spin_limit = lz_spin_limit_; a.bi_memory = lz_bi_memory_;
        if (getenv("PROXSDP_B200_LZ_PROF")) { if (lz_prof_.n == 0) lz_prof_.alloc(8 + 8 * 256); a.prof = lz_prof_.p; }
        void* kargs[] = {&a};
        cudaError_t e = cudaLaunchKernelExC(&cfg, kfn, kargs);
        if (e != cudaSuccess && lz_coop_ && (e == cudaErrorNotSupported || e == cudaErrorInvalidValue)) {
            cudaGetLastError();
            cfg.numAttrs = 1;
            e = cudaLaunchKernelExC(&cfg, kfn, kargs);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            lz_resident_ok_[C == 16] = -1;
            if (getenv("PROXSDP_B200_DEBUG")) fprintf(stderr, "[lanczos] resident single-cluster kernel, cluster size %d: launch failed (%s)\n", C, cudaGetErrorString(e));
            continue;
        }
        if (getenv("PROXSDP_B200_DEBUG") && !cd.announced_resident) {
            cd.announced_resident = true;
            fprintf(stderr, "[lanczos] cl3 resident kernel: side %d, K %d, ONE cluster of %d CTAs, %d rows of X per CTA in shared memory, %zu KB shared memory\n",
                    nside, K, C, vn_max, smem >> 10);
        }
        (void)cpr;
        cd.ritz_launches++;
        launches += 1;
        lz_cluster_launches_++;
        resident_calls++;
        return true;
    }
    return false;
}

void Solver::lanczos_launch(ConeDev& cd, int cone_idx, int nev, int K, int maxiter, double tol) {
    if (lanczos_launch_resident(cd, cone_idx, nev, K, maxiter, tol)) return;
    if (lanczos_launch_cluster3(cd, cone_idx, nev, K, maxiter, tol)) return;
    if (lanczos_launch_cluster(cd, cone_idx, nev, K, maxiter, tol)) return;
    const int nside = cd.side;
    int G = std::min(num_sms_, std::max(1, (nside + 7) / 8));
    int rows_max = (nside + G - 1) / G;
    if ((rows_max + LZ_NW - 1) / LZ_NW + 1 > LZ_TMAX) throw CudaError(-4, "Lanczos kernel: cone side too large for the row-slab layout");
    int jac_inplace = 0;
    size_t smem = lanczos_smem_bytes(K, rows_max, nside, 0);
    if (smem > smem_optin_) { jac_inplace = 1; smem = lanczos_smem_bytes(K, rows_max, nside, 1); }
    if (smem > smem_optin_) throw CudaError(-4, "Lanczos kernel: vector + basis slab do not fit shared memory for this (n, K)");
    if ((size_t)cd.ld * (size_t)(K + 1) > cd.Y.n) cd.Y.alloc((size_t)cd.ld * (size_t)(K + 1));
    // flagged-exchange buffers (shared by all cones of this solver; launches are stream ordered)
    const size_t need_x = (size_t)2 * (size_t)(K + 2) * (size_t)G, need_v = (size_t)2 * (size_t)cd.ld;
    const unsigned long long bound = 4ULL * (unsigned long long)K * (unsigned long long)std::max(maxiter, 1) + 16ULL;
    if (need_x > lz_xbuf_.n || need_v > lz_vx_.n || lz_epoch_ + bound >= 0xFFFFFFF0ULL) {
        PB_CUDA(cudaStreamSynchronize(stream));
        if (need_x > lz_xbuf_.n) lz_xbuf_.alloc(need_x); else PB_CUDA(cudaMemset(lz_xbuf_.p, 0, lz_xbuf_.n * sizeof(uint4)));
        if (need_v > lz_vx_.n) lz_vx_.alloc(need_v); else PB_CUDA(cudaMemset(lz_vx_.p, 0, lz_vx_.n * sizeof(uint4)));
        lz_epoch_ = 0;      // freshly zeroed buffers carry flag 0, which is never used
    }
    LanczosArgs a{};
    a.X = cd.X.p; a.n = nside; a.ld = cd.ld; a.x0 = cd.resid.p; a.Y = cd.Y.p;
    a.xbuf = lz_xbuf_.p; a.vx = lz_vx_.p; a.epoch_base = (unsigned int)lz_epoch_;
    lz_epoch_ += bound;
    a.nev = nev; a.K = K; a.maxiter = maxiter; a.tol = tol;
    a.rows_max = rows_max; a.jac_inplace = jac_inplace;
    a.vals = cd.vals.p; a.info = cd.info.p; a.scal = scal_target_; a.cone = cone_idx; a.n_cones_total = n_sdp;
    if (getenv("PROXSDP_B200_LZ_PROF")) { if (lz_prof_.n == 0) lz_prof_.alloc(8 + 8 * 256); a.prof = lz_prof_.p; }
    void* args[] = {&a};
    PB_CUDA(cudaLaunchCooperativeKernel((void*)k_lanczos, dim3(G), dim3(LZ_THREADS), args, smem, stream));
    launches++;
}

void Solver::launch_reconstruct(ConeDev& cd, double* x_out) {
    int nt = (cd.side + 31) / 32;
    int tiles = nt * (nt + 1) / 2;
    k_reconstruct_svec<<<tiles, dim3(32, 8), 0, stream>>>(cd.Y.p, cd.ld, cd.side, cd.kept_idx.p, cd.kept_lam.p,
                                                          cd.nkept.p, x_out + cd.off, scal_d_.p + S_POISON);
    launches++;
}

// Block-Jacobi eigendecomposition of cd.X (destroyed).  Returns the eigenvalues (unsorted);
// eigenvectors are the columns of cd.Vfull.
std::vector<double> Solver::full_eig_device(ConeDev& cd, bool warm) {
    const int nside = cd.side, ld = cd.ld, NP = ld;
    if (cd.Vfull.n < (size_t)ld * ld) { cd.Vfull.alloc((size_t)ld * ld); cd.have_V = false; }
    warm = warm && bj_warm_ && cd.have_V && (cd.full_calls % 32) != 0;
    cd.full_calls++;
    int nb = NP / BJ_B;
    int mplayers = (nb + 1) & ~1;
    int npairs = mplayers / 2;
    if (bj_Q_.n < (size_t)npairs * BJ_QSTRIDE) bj_Q_.alloc((size_t)npairs * BJ_QSTRIDE);
    if (bj_rot_.n < 1) bj_rot_.alloc(1);
    {
        long long tot = (long long)NP * NP;
        if (warm) {
            // A <- W' (A W) with W = the previous eigenvectors (orthogonal to working precision); V starts as W
            if (cd.Vtmp.n < (size_t)ld * ld) cd.Vtmp.alloc_raw((size_t)ld * ld);
            k_bj_pad<<<ceil_div(tot, 256), 256, 0, stream>>>(cd.X.p, ld, nside, NP);
            k_gemm64<false><<<dim3(NP / 64, NP / 64), 256, 0, stream>>>(cd.X.p, cd.Vfull.p, cd.Vtmp.p, NP, ld);
            k_gemm64<true><<<dim3(NP / 64, NP / 64), 256, 0, stream>>>(cd.Vfull.p, cd.Vtmp.p, cd.X.p, NP, ld);
            launches += 3;
        } else {
            k_bj_init<<<ceil_div(tot, 256), 256, 0, stream>>>(cd.X.p, cd.Vfull.p, ld, nside, NP);
            launches++;
        }
    }
    cd.have_V = true;
    std::vector<double> offn(2);
    const int chunks = (NP + BJ_P - 1) / BJ_P;
    for (int sweep = 0; sweep < 60; ++sweep) {
        PB_CUDA(cudaMemsetAsync(bj_rot_.p, 0, sizeof(int), stream));
        for (int round = 0; round < std::max(1, mplayers - 1); ++round) {
            BjRound pr{round, mplayers, nb};
            k_bj_pair_eig<<<npairs, 512, bj_pair_smem_bytes(), stream>>>(cd.X.p, ld, pr, bj_inner_sweeps_, bj_Q_.p, bj_rot_.p);
            k_bj_apply<0><<<dim3(npairs, chunks, 2), 256, BJ_APPLY_SMEM, stream>>>(cd.X.p, cd.Vfull.p, ld, NP, pr, bj_Q_.p);
            k_bj_apply<1><<<dim3(npairs, chunks, 1), 256, BJ_APPLY_SMEM, stream>>>(cd.X.p, cd.Vfull.p, ld, NP, pr, bj_Q_.p);
            launches += 3;
        }
        k_bj_offnorm<<<std::min(reduce_blocks_, ceil_div((long long)nside * nside, 256)), 256, 0, stream>>>(
            cd.X.p, ld, nside, offnorm_d_.p, ws_);
        launches++;
        PB_CUDA(cudaMemcpyAsync(offn.data(), offnorm_d_.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        PB_CUDA(cudaStreamSynchronize(stream));
        if (getenv("PROXSDP_B200_DEBUG")) fprintf(stderr, "[bj] n=%d sweep %d off2=%.3e tot2=%.3e\n", nside, sweep, offn[0], offn[1]);
        if (!(offn[0] > 1e-30 * offn[1])) break;     // off^2 <= 1e-30 * ||A||_F^2
    }
    DBuf<double> diag;
    diag.alloc((size_t)nside);
    k_bj_diag<<<ceil_div(nside, 256), 256, 0, stream>>>(cd.X.p, ld, nside, diag.p);
    launches++;
    PB_CUDA(cudaStreamSynchronize(stream));
    full_eig_calls++;
    return diag.download();
}

// full_eig! for a large cone (prox_operators.jl:111-126) on the matrix in cones[k].X
void Solver::launch_full_projection_large(int k) {
    ConeDev& cd = cones[(size_t)k];
    std::vector<double> w = full_eig_device(cd, /*warm=*/true);
    std::vector<int> idx;
    std::vector<double> lam;
    long long rk = 0;
    for (int i = 0; i < cd.side; ++i)
        if (w[(size_t)i] > 0.0) { idx.push_back(i); lam.push_back(w[(size_t)i]); if (w[(size_t)i] > opt.tol_psd) rk++; }
    int nk = (int)idx.size();
    if (nk > 0) {
        PB_CUDA(cudaMemcpyAsync(cd.kept_idx.p, idx.data(), sizeof(int) * (size_t)nk, cudaMemcpyHostToDevice, stream));
        PB_CUDA(cudaMemcpyAsync(cd.kept_lam.p, lam.data(), sizeof(double) * (size_t)nk, cudaMemcpyHostToDevice, stream));
    }
    PB_CUDA(cudaMemcpyAsync(cd.nkept.p, &nk, sizeof(int), cudaMemcpyHostToDevice, stream));
    double rec[3] = {(double)rk, 0.0, -1.0};
    PB_CUDA(cudaMemcpyAsync(scal_d_.p + S_HEADER + 3 * k, rec, sizeof(rec), cudaMemcpyHostToDevice, stream));
    int nt = (cd.side + 31) / 32;
    int tiles = nt * (nt + 1) / 2;
    k_reconstruct_svec<<<tiles, dim3(32, 8), 0, stream>>>(cd.Vfull.p, cd.ld, cd.side, cd.kept_idx.p, cd.kept_lam.p,
                                                          cd.nkept.p, x_[1 - cur_].p + cd.off, nullptr);
    launches++;
    PB_CUDA(cudaStreamSynchronize(stream));   // idx/lam/nk/rec are stack/host temporaries
}

// psd_projection! (prox_operators.jl:33-66) fused with the primal gradient step (pdhg.jl:622)
// reads x_[cur_], Mty_[cur_], c_ ; writes the PSD part of x_[1-cur_]
void Solver::psd_projection_launch(long long iter, double tau, bool force_full) {
    const double* x = x_[cur_].p; const double* Mty = Mty_[cur_].p;
    double* xn = x_[1 - cur_].p;
    PB_CUDA(cudaEventRecord(ev_psd0_, stream));
    lz_timed_ = false;
    if (!small_ids_.empty()) {
        SmallConeArgs a{};
        a.cone_ids = small_ids_d_.p; a.cone_side = cone_side_d_.p; a.cone_off = cone_off_d_.p;
        a.x = x; a.Mty = Mty; a.c = c_.p; a.tau = tau; a.tol_psd = opt.tol_psd; a.x_out = xn; a.scal = scal_d_.p;
        a.mode = 0; a.scale = 1.0; a.out_min = out_min_d_.p;
        size_t sb = max_small_side_ > 1 ? small_cone_smem_bytes(max_small_side_, small_fast_) : 0;
        a.fast = small_fast_;
        if (small_fast_ && small_warm_ && max_small_side_ > 1) {
            const int mm = small_cone_m(max_small_side_);
            const size_t stride = (size_t)mm * (size_t)(mm | 1);
            if (small_warm_d_.n < stride * small_ids_.size()) { small_warm_d_.alloc(stride * small_ids_.size()); small_warm_calls_ = 0; }
            a.warm = small_warm_d_.p; a.warm_stride = (long long)stride;
            a.warm_read = (small_warm_calls_ > 0 && (small_warm_calls_ % 32) != 0) ? 1 : 0;
            small_warm_calls_++;
        }
        k_small_cone_proj<<<(int)small_ids_.size(), small_fast_ ? 512 : 256, sb, stream>>>(a);
        launches++;
        full_eig_calls += (long long)small_ids_.size();
    }
    std::vector<int> full_now;
    last_tau_ = tau;
    for (int k : large_ids_) {
        ConeDev& cd = cones[(size_t)k];
        int nt = (cd.side + 31) / 32;
        int tiles = nt * (nt + 1) / 2;
        const bool krylov = !force_full && krylov_eligible(k, iter);
        const int nev = (int)target_rank[(size_t)k];
        const int K = (int)std::max<long long>(2 * nev + 1, opt.eigsolver_min_lanczos);   // eigsolver.jl:794
        // implicit operator: x_k is still the low-rank product of the previous Krylov projection, so the matrix
        // x_k - tau (M'y + c) need not be formed at all (opt.implicit_psd_operator)
        cd.used_implicit = false;
        if (krylov && opt.implicit_psd_operator && cd.lowrank_valid && nev <= LZ3_RMAX && tau != 0.0) {
            const bool timed = !lz_timed_;
            if (timed) PB_CUDA(cudaEventRecord(ev_lz0_, stream));
            cd.used_implicit = lanczos_launch_implicit(cd, k, nev, K, (int)opt.krylovkit_max_iter, opt.krylovkit_tol, tau);
            if (cd.used_implicit && timed) { PB_CUDA(cudaEventRecord(ev_lz1_, stream)); lz_timed_ = true; }
        }
        if (!cd.used_implicit) {
            k_svec_to_mat<true><<<tiles, dim3(32, 8), 0, stream>>>(x + cd.off, Mty + cd.off, c_.p + cd.off, tau, 1.0,
                                                                   cd.side, cd.ld, cd.X.p, nzmask_.p, (long long)cd.off);
            launches++;
        }
        if (krylov) {
            if (!cd.used_implicit) {
                const bool timed = !lz_timed_;     // the per-kernel timer covers the first eigsolve of the iteration
                if (timed) PB_CUDA(cudaEventRecord(ev_lz0_, stream));
                lanczos_launch(cd, k, nev, K, (int)opt.krylovkit_max_iter, opt.krylovkit_tol);
                if (timed) { PB_CUDA(cudaEventRecord(ev_lz1_, stream)); lz_timed_ = true; }
            }
            cd.lowrank_valid = true;           // (cleared again by the fallback when the eigsolve did not converge)
            lanczos_calls++;
            k_lanczos_select<<<1, 32, 0, stream>>>(cd.vals.p, cd.info.p, nev, cd.kept_idx.p, cd.kept_lam.p,
                                                   cd.nkept.p, scal_d_.p, k);
            launches++;
            launch_reconstruct(cd, xn);
        } else {
            cd.lowrank_valid = false;
            full_now.push_back(k);
        }
    }
    for (int k : full_now) launch_full_projection_large(k);
    PB_CUDA(cudaEventRecord(ev_psd1_, stream));
}

void Solver::launch_dual_trial(int trial, double tau0) {
    DualArgs d{};
    d.y = y_[cur_].p; d.Mx = Mx_[1 - cur_].p; d.Mx_old = Mx_[cur_].p; d.b = b_.p; d.h = h_.p; d.y_new = y_[1 - cur_].p;
    d.p = (int)p; d.m = (int)m;
    d.tau0 = tau0; d.decay = opt.linsearch_decay; d.tau_old = primal_step_old_; d.beta = beta_;
    d.sigma_fixed = dual_step_; d.trial = trial; d.use_theta = opt.line_search_flag ? 1 : 0;
    int blocks = std::max(1, std::min(reduce_blocks_, ceil_div(R, 256)));
    k_dual_trial<<<blocks, 256, 0, stream>>>(d, scal_d_.p, ws_);
    MtArgs t{};
    t.N = (int)n; t.rowptr = Mt_.rowptr.p; t.colidx = Mt_.colidx.p; t.val = Mt_.val.p; t.long_threshold = Mt_.long_threshold;
    t.nz_rows = Mt_.nz_rows.p; t.n_nz = Mt_.n_nz;
    if (Mt_.n_long > 0) {
        k_spmv_mt_long<<<Mt_.n_long, 512, 0, stream>>>(Mt_.long_rows.p, Mt_.nz_rows.p, Mt_.rowptr.p, Mt_.colidx.p, Mt_.val.p,
                                                       y_[1 - cur_].p, Mty_[1 - cur_].p, scal_d_.p);
        launches++;
    }
    t.y_new = y_[1 - cur_].p; t.Mty = Mty_[cur_].p; t.Mty_new = Mty_[1 - cur_].p;
    t.beta = beta_; t.delta = opt.delta; t.trial = trial;
    t.do_test = opt.line_search_flag ? (sharded() ? 2 : 1) : 0;
    int blocks2 = std::max(1, std::min(reduce_blocks_, ceil_div(Mt_.n_nz, 256)));
    k_spmv_mt_norm<<<blocks2, 256, 0, stream>>>(t, scal_d_.p, ws_);
    launches += 2;
    if (sharded() && opt.line_search_flag) {
        if (red_d_.n < (size_t)nranks_ * 16) red_d_.alloc((size_t)nranks_ * 16);
        PB_NCCL(nccl_api().AllGather(scal_d_.p + S_YNORM2, red_d_.p, 2, ncclDouble, comm_->comm, stream));
        k_ls_decide<<<1, 32, 0, stream>>>(red_d_.p, nranks_, scal_d_.p, beta_, opt.delta, trial);
        launches++;
    }
}

// trials [trial0, trial0 + T) of the line search in two launches (kernels_vec.cuh: k_ls_ladder / k_ls_apply)
void Solver::launch_ladder(int trial0, int T, double tau0) {
    LadderArgs a{};
    a.y = y_[cur_].p; a.Mx = Mx_[1 - cur_].p; a.Mx_old = Mx_[cur_].p; a.b = b_.p; a.h = h_.p;
    a.p = (int)p; a.m = (int)m;
    a.tau0 = tau0; a.decay = opt.linsearch_decay; a.tau_old = primal_step_old_; a.beta = beta_;
    a.sigma_fixed = dual_step_; a.delta = opt.delta;
    a.use_theta = opt.line_search_flag ? 1 : 0; a.do_test = opt.line_search_flag ? (sharded() ? 2 : 1) : 0;
    if (sharded() && ls_partial_d_.n == 0) ls_partial_d_.alloc(2 * LS_MAXT);
    a.partial_out = ls_partial_d_.p;
    a.ntrials = std::max(1, std::min(T, LS_MAXT));
    a.nz_rows = Mt_.nz_rows.p; a.nz_ptr = Mt_.rowptr.p; a.colidx = Mt_.colidx.p; a.val = Mt_.val.p; a.n_nz = Mt_.n_nz;
    a.long_threshold = Mt_.long_threshold; a.long_rows = Mt_.long_rows.p; a.n_long = Mt_.n_long;
    if (Mt_.n_long > 0 && ls_long_sums_.n < (size_t)Mt_.n_long * LS_MAXT) ls_long_sums_.alloc((size_t)Mt_.n_long * LS_MAXT);
    a.long_sums = ls_long_sums_.p;
    a.Mty = Mty_[cur_].p; a.y_new = y_[1 - cur_].p; a.Mty_new = Mty_[1 - cur_].p;
    if (Mt_.n_long > 0) {
        k_ls_long_rows<<<Mt_.n_long, 512, 0, stream>>>(a, scal_d_.p, trial0);
        launches++;
    }
    const long long work = std::max<long long>(R, Mt_.n_nz);
    const int blocks = std::max(1, std::min(reduce_blocks_ / 2, ceil_div(work, 256)));
    k_ls_ladder<<<blocks, 256, 0, stream>>>(a, scal_d_.p, ws_, trial0);
    if (a.do_test == 2) {
        // one exchange for the whole ladder: 2 * LS_MAXT doubles per rank
        if (red_d_.n < (size_t)nranks_ * 2 * LS_MAXT) red_d_.alloc((size_t)nranks_ * 2 * LS_MAXT);
        PB_NCCL(nccl_api().AllGather(ls_partial_d_.p, red_d_.p, 2 * LS_MAXT, ncclDouble, comm_->comm, stream));
        k_ls_decide_gathered<<<1, 32, 0, stream>>>(a, red_d_.p, nranks_, scal_d_.p, trial0);
        launches++;
    }
    k_ls_apply<<<blocks, 256, 0, stream>>>(a, scal_d_.p, trial0);
    launches += 2;
}

// everything after the eigen-solves: SOC projection, Mx, dual step / linesearch ladder, residuals
void Solver::launch_post_eig(double tau0, bool first_pass) {
    double* xn = x_[1 - cur_].p;
    const double* poison = scal_d_.p + S_POISON;
    if (first_pass && psd_end_ < n) {
        long long cnt = n - psd_end_;
        k_primal_tail<<<std::min(reduce_blocks_, ceil_div(cnt, 256)), 256, 0, stream>>>(
            x_[cur_].p, Mty_[cur_].p, c_.p, primal_step_, psd_end_, n, xn);
        launches++;
    }
    if (first_pass && iter_ == 1 && opt.advanced_initialization) {
        // pdhg.jl:138-142 seeds pair.x = tau*c but leaves pair.x_old = 0, and compute_residual!
        // (residuals.jl:41-48) reads x_old: once the primal step has consumed x, the old buffer becomes 0.
        PB_CUDA(cudaMemsetAsync(x_[cur_].p, 0, sizeof(double) * (size_t)std::max<long long>(n, 1), stream));
    }
    if (n_soc > 0) {
        k_soc_project<<<n_soc, 256, 0, stream>>>(xn, soc_off_d_.p, soc_len_d_.p, soc_gap_d_.p, poison);
        k_soc_gap_max<<<1, 256, 0, stream>>>(soc_gap_d_.p, n_soc, scal_d_.p);
        launches += 2;
    }
    launch_spmv(M_, xn, Mx_[1 - cur_].p);
    int ntr = opt.line_search_flag ? (int)std::min<long long>(ladder_, std::max<long long>(opt.max_linsearch_steps, 1)) : 1;
    if (fused_ladder_) launch_ladder(0, ntr, tau0);
    else for (int t = 0; t < ntr; ++t) launch_dual_trial(t, tau0);
    int blocksN = std::max(1, std::min(std::min(reduce_blocks_, 4 * num_sms_), ceil_div(n, 512)));      // one resident wave
    k_residual_primal<<<blocksN, 256, 0, stream>>>(n, xn, x_[cur_].p, Mty_[1 - cur_].p, Mty_[cur_].p, c_.p, scal_d_.p, ws_, nzmask_.p);
    int blocksR = std::max(1, std::min(reduce_blocks_, ceil_div(R, 256)));
    k_residual_dual<<<blocksR, 256, 0, stream>>>((int)p, (int)m, beta_, opt.line_search_flag ? 1 : 0, dual_step_,
                                                 y_[1 - cur_].p, y_[cur_].p, Mx_[1 - cur_].p, Mx_[cur_].p, b_.p, h_.p,
                                                 scal_d_.p, ws_);
    launches += 2;
}

void Solver::launch_soc_only() {
    if (n_soc == 0) return;
    k_soc_project<<<n_soc, 256, 0, stream>>>(x_[1 - cur_].p, soc_off_d_.p, soc_len_d_.p, soc_gap_d_.p, scal_d_.p + S_POISON);
    k_soc_gap_max<<<1, 256, 0, stream>>>(soc_gap_d_.p, n_soc, scal_d_.p);
    launches += 2;
}

void Solver::reset_scalars() {
    const double soc_init = (sharded() && n_soc == 0) ? -1.0e300 : 0.0;
    k_scal_reset<<<ceil_div(scal_len, 256), 256, 0, stream>>>(scal_d_.p, scal_len, soc_init, time0_ > 0 ? now_s() - time0_ : 0.0);
    launches++;
}

// Krylov fallback (prox_operators.jl:55-57): cones whose eigsolve reported converged == 0 are
// redone with the full eigendecomposition; the others only need their (skipped) reconstruction.
void Solver::fallback_projection(long long iter) {
    PB_CUDA(cudaMemsetAsync(scal_d_.p + S_POISON, 0, sizeof(double), stream));
    if (sharded()) PB_CUDA(cudaMemsetAsync(scal_d_.p + S_LS_ACCEPTED, 0, sizeof(double) * (S_LS_EVALS + 1), stream));   // other ranks ran their trials on partial sums
    for (int kk : large_ids_) {
        if (!krylov_eligible(kk, iter)) continue;
        if (scal_host[S_HEADER + 3 * kk + 2] == 0.0) {
            if (!lz_demoted_ && lz_mode_ == 0) {
                // converged == 0 is either KrylovKit's "nothing converged" or an eigsolve that gave up because a peer CTA
                // never answered (info[0] == 0: part of the grid was not resident).  The latter would repeat its 2 s
                // time-out on every iteration: switch to the cooperative row-distributed kernel for the rest of the solve.
                int info[4] = {1, 0, 0, 0};
                PB_CUDA(cudaMemcpyAsync(info, cones[(size_t)kk].info.p, sizeof(info), cudaMemcpyDeviceToHost, stream));
                PB_CUDA(cudaStreamSynchronize(stream));
                if (info[0] == 0 && info[2] > 0) {
                    fprintf(stderr, "[proxsdp_b200] a cluster eigsolve timed out waiting for a peer CTA (device shared with another "
                                    "tenant?): using the cooperative row-distributed Lanczos kernel from here on\n");
                    lz_mode_ = 1; lz_demoted_ = true;
                }
            }
            {
                ConeDev& cf = cones[(size_t)kk];
                if (cf.used_implicit) {      // the eigsolve ran without the dense matrix: the exact projection needs it
                    int nt = (cf.side + 31) / 32;
                    k_svec_to_mat<true><<<nt * (nt + 1) / 2, dim3(32, 8), 0, stream>>>(x_[cur_].p + cf.off, Mty_[cur_].p + cf.off, c_.p + cf.off,
                                                                                       last_tau_, 1.0, cf.side, cf.ld, cf.X.p);
                    launches++;
                }
                cf.lowrank_valid = false;
            }
            launch_full_projection_large(kk);
        } else {
            ConeDev& cd = cones[(size_t)kk];
            k_lanczos_select<<<1, 32, 0, stream>>>(cd.vals.p, cd.info.p, (int)target_rank[(size_t)kk], cd.kept_idx.p,
                                                   cd.kept_lam.p, cd.nkept.p, scal_d_.p, kk);
            launches++;
            launch_reconstruct(cd, x_[1 - cur_].p);
        }
    }
}

void Solver::sync_scalars(bool iteration_end) {
    if (sharded()) {
        // whole-problem record: one all-gather of the per-rank headers, folded in rank order on every rank
        if (gather_d_.n < (size_t)nranks_ * S_HEADER) gather_d_.alloc((size_t)nranks_ * S_HEADER);
        PB_NCCL(nccl_api().AllGather(scal_d_.p, gather_d_.p, S_HEADER, ncclDouble, comm_->comm, stream));
        k_fold_header<<<1, 32, 0, stream>>>(gather_d_.p, nranks_, scal_d_.p);
        launches++;
    }
    if (!spin_sync_) {
        PB_CUDA(cudaMemcpyAsync(scal_host, scal_d_.p, sizeof(double) * (size_t)scal_len, cudaMemcpyDeviceToHost, stream));
        g_d2h_bytes += (long long)sizeof(double) * scal_len;
        PB_CUDA(cudaStreamSynchronize(stream));
        scal_clean_ = false;
    } else {
        // the last kernel of the sequence writes the record into mapped host memory and bumps a sequence word; the
        // host spins on it (a D2H copy + stream synchronisation costs 25-40 us per iteration, this ~3).  Sharded runs
        // re-initialise the record every iteration (elapsed time, SOC identity), so nothing is reset here for them.
        ++pub_seq_;
        const int mode = (iteration_end && !sharded()) ? (opt.line_search_flag ? 1 : 2) : 0;
        k_publish_record<<<1, 128, 0, stream>>>(scal_d_.p, scal_len, mode, scal_host_dev_, pub_seq_dev_, pub_seq_);
        launches++;
        g_d2h_bytes += (long long)sizeof(double) * scal_len;
        volatile unsigned long long* seq = pub_seq_host_;
        unsigned long long spins = 0;
        while (*seq != pub_seq_) {
            if ((++spins & 0x3fff) == 0) {
                cudaError_t q = cudaStreamQuery(stream);
                if (q == cudaSuccess) { if (*seq == pub_seq_) break; if (spins > (1ULL << 26)) throw CudaError(-100, "iteration record never arrived"); }
                else if (q != cudaErrorNotReady) throw CudaError(-100 - (int)q, std::string("kernel failed: ") + cudaGetErrorString(q));
            }
        }
        scal_clean_ = mode != 0 && scal_host[S_POISON] == 0.0 && (!opt.line_search_flag || scal_host[S_LS_ACCEPTED] != 0.0);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw CudaError(-100 - (int)e, std::string("kernel launch failed: ") + cudaGetErrorString(e));
}

// host-level all-reduce of a few doubles (setup norms, rare convergence flags): all-gather + rank-ordered fold
void Solver::host_reduce(double* vals, int count, int op) {
    if (!sharded()) return;
    if (count > 16) throw CudaError(-1, "host_reduce: too many values");
    if (red_d_.n < (size_t)(nranks_ + 1) * 16) red_d_.alloc((size_t)(nranks_ + 1) * 16);
    double* send = red_d_.p + (size_t)nranks_ * 16;
    PB_CUDA(cudaMemcpyAsync(send, vals, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, stream));
    PB_NCCL(nccl_api().AllGather(send, red_d_.p, (size_t)count, ncclDouble, comm_->comm, stream));
    std::vector<double> all((size_t)nranks_ * (size_t)count);
    PB_CUDA(cudaMemcpyAsync(all.data(), red_d_.p, sizeof(double) * all.size(), cudaMemcpyDeviceToHost, stream));
    PB_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i < count; ++i) {
        double v = all[(size_t)i];
        for (int r = 1; r < nranks_; ++r) {
            double o = all[(size_t)r * count + i];
            v = (op == 0) ? v + o : jl_max(v, o);
        }
        vals[i] = v;
    }
}

// ---------------------------------------------------------------------------
// result assembly (pdhg.jl:678-787), on the device
// ---------------------------------------------------------------------------
double Solver::dual_feas_device(const double* y_dev, double c_factor) {
    // get_duals (pdhg.jl:701-710): dual_cone = c + A'y_eq + G'y_in (un-scaled), off-diagonals / 2
    if (res_dc_d_.n < (size_t)n) res_dc_d_.alloc_raw((size_t)n);
    double* dc = res_dc_d_.p;
    const int blocks_n = std::max(1, std::min(reduce_blocks_, ceil_div(n, 256)));
    if (n > 0) {
        k_scale_copy<<<blocks_n, 256, 0, stream>>>(c_orig_d_.p, c_factor, n, dc);
        launches++;
        if (Mt_.n_nz > 0) {
            k_dual_cone_rows<<<std::max(1, std::min(reduce_blocks_, ceil_div((long long)Mt_.n_nz * 32, 256))), 256, 0, stream>>>(
                Mt_.n_nz, Mt_.nz_rows.p, Mt_.rowptr.p, Mt_.colidx.p, Mt_.val_orig.p, y_dev, dc);
            launches++;
        }
        if (n_sdp > 0) {
            k_scale_offdiag_copy<<<blocks_n, 256, 0, stream>>>(dc, n, psd_end_, cone_off_d_.p, n_sdp, 2.0, 1, dc);
            launches++;
        }
    }
    // dual_feas (pdhg.jl:716-732): inequality multipliers, free-variable part
    PB_CUDA(cudaMemsetAsync(feas_d_.p, 0, sizeof(double) * FS_COUNT, stream));
    {
        const long long work = std::max<long long>(m, n - listed_end_);
        k_feas_ineq_tail<<<std::max(1, std::min(reduce_blocks_, ceil_div(work, 256))), 256, 0, stream>>>(
            y_dev + p, m, dc, listed_end_, n, feas_d_.p, ws_);
        launches++;
    }
    // cone_feas (pdhg.jl:678-699): min eigenvalue of mat(dual_cone) with off-diagonals / sqrt(2); SOC: t - ||v||
    if (n_soc > 0) {
        k_feas_soc<<<n_soc, 256, 0, stream>>>(dc, soc_off_d_.p, soc_len_d_.p, soc_gap_d_.p);
        k_feas_max<<<1, 256, 0, stream>>>(soc_gap_d_.p, nullptr, n_soc, 0, feas_d_.p, FS_SOC);
        launches += 2;
    }
    if (!small_ids_.empty()) {
        SmallConeArgs a{};
        a.cone_ids = small_ids_d_.p; a.cone_side = cone_side_d_.p; a.cone_off = cone_off_d_.p;
        a.x = dc; a.Mty = dc; a.c = dc; a.tau = 0.0; a.tol_psd = opt.tol_psd; a.x_out = nullptr;
        a.scal = scal_scratch_d_.p; a.mode = 1; a.scale = 1.0; a.out_min = out_min_d_.p;
        size_t sb = max_small_side_ > 1 ? small_cone_smem_bytes(max_small_side_, small_fast_) : 0;
        a.fast = small_fast_;
        k_small_cone_proj<<<(int)small_ids_.size(), small_fast_ ? 512 : 256, sb, stream>>>(a);
        k_feas_max<<<1, 256, 0, stream>>>(out_min_d_.p, small_ids_d_.p, (int)small_ids_.size(), 1, feas_d_.p, FS_SMALL_PSD);
        launches += 2;
    }
    double cone_viol = 0.0;
    // large cones: lambda_min(Z) = -lambda_max(-Z) by Lanczos (nev = 1) on the negated matrix.  The eigen kernels write
    // their bookkeeping to a scratch record so that the live per-iteration record is left alone.
    scal_target_ = scal_scratch_d_.p;
    struct Restore { double*& t; double* v; double& stop; ~Restore() { t = v; stop = 1e300; } } restore{scal_target_, scal_d_.p, lz_stop_above_};
    for (int k : large_ids_) {
        ConeDev& cd = cones[(size_t)k];
        int nt = (cd.side + 31) / 32;
        int tiles = nt * (nt + 1) / 2;
        k_svec_to_mat<false><<<tiles, dim3(32, 8), 0, stream>>>(dc + cd.off, nullptr, nullptr, 0.0, -1.0, cd.side, cd.ld, cd.X.p);
        launches++;
        // Only the sign of lambda_min relative to tol_feasibility_dual is consumed (dual_feasible_user_tol,
        // pdhg.jl:712-732), so the extreme eigenvalue is resolved to 1e-3 of that tolerance, not to 1e-10 ...
        int K = (int)std::max<long long>(opt.eigsolver_min_lanczos, 3);
        const double tol_ev = std::max(1e-10, 1e-3 * std::min(opt.tol_feasibility_dual, 1.0));
        // ... and a Ritz value of -Z above the tolerance already proves the violation (Ritz values are lower bounds
        // of lambda_max), so the eigsolve may stop there instead of converging an interior-looking extreme pair
        lz_stop_above_ = opt.tol_feasibility_dual;
        PB_CUDA(cudaMemsetAsync(scal_scratch_d_.p, 0, sizeof(double) * (size_t)scal_len, stream));
        lanczos_launch(cd, k, 1, K, 60, tol_ev);
        lz_stop_above_ = 1e300;
        int info[4] = {0, 0, 0, 0};
        double top = 0.0;
        PB_CUDA(cudaMemcpyAsync(info, cd.info.p, sizeof(info), cudaMemcpyDeviceToHost, stream));
        PB_CUDA(cudaMemcpyAsync(&top, cd.vals.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
        PB_CUDA(cudaStreamSynchronize(stream));
        g_d2h_bytes += 24;
        double lmin;
        if (info[0] >= 1 && info[1] >= 1) {
            lmin = -top;
        } else {
            // not converged (or the launch gave up): the reference takes the exact spectrum here (eigen!, pdhg.jl:685)
            std::vector<double> w = full_eig_device(cd);      // X still holds -Z: Lanczos only reads it
            cd.have_V = false;                                // (the basis of -Z is no warm start for the next projection)
            double mx = w.empty() ? 0.0 : w[0];
            for (double v : w) mx = std::max(mx, v);
            lmin = -mx;
        }
        cone_viol = std::max(cone_viol, -std::min(0.0, lmin));
    }
    PB_CUDA(cudaMemcpyAsync(feas_host_, feas_d_.p, sizeof(double) * FS_COUNT, cudaMemcpyDeviceToHost, stream));
    PB_CUDA(cudaStreamSynchronize(stream));
    g_d2h_bytes += (long long)sizeof(double) * FS_COUNT;
    cone_viol = std::max(cone_viol, std::max(feas_host_[FS_SOC], feas_host_[FS_SMALL_PSD]));
    return std::max(cone_viol, std::max(feas_host_[FS_INEQ], feas_host_[FS_ZERO]));
}

// dst = src with the off-diagonal svec entries of every PSD block divided by num (divide != 0: fix_diag_scaling,
// pdhg.jl:734-743) or multiplied by num (divide == 0: norm_scaling, scaling.jl:28-58); src == dst is allowed
__global__ void k_scale_offdiag_copy(const double* __restrict__ src, long long n, long long psd_end,
                                     const long long* __restrict__ cone_off, int n_sdp, double num, int divide,
                                     double* __restrict__ dst) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        double v = src[k];
        if (offdiag_position(k, psd_end, cone_off, n_sdp)) v = divide ? v / num : __dmul_rn(v, num);
        dst[k] = v;
    }
}

__global__ void k_scale_copy(const double* __restrict__ src, double a, long long n, double* __restrict__ dst) {
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __dmul_rn(a, src[i]);
}

// cache_solution (pdhg.jl:745-787).  Like the reference it rescales pair.x IN PLACE.  Slacks, duals and the dual
// feasibility are computed by the kernels of kernels_result.cuh; every output vector is downloaded once, straight
// into the caller's buffer (pinned or pageable).
void Solver::cache_solution(double c_factor, proxsdp_result_t* out) {
    StageTimer st("finish: scaling + slack");
    for (ConeDev& cd : cones) cd.lowrank_valid = false;      // x is rescaled in place below: no longer svec(Y lam Y')
    const int blocks_n = std::max(1, std::min(reduce_blocks_, ceil_div(n, 256)));
    if (n_sdp > 0 && n > 0) {
        k_scale_offdiag_copy<<<blocks_n, 256, 0, stream>>>(x_[cur_].p, n, psd_end_, cone_off_d_.p, n_sdp, std::sqrt(2.0), 1, x_[cur_].p);
        launches++;
    }
    if (opt.equilibration && eq_E_d_.p) {        // remove equilibrating (pdhg.jl:751-755), in place like the reference
        launch_eq_mul_scalar(x_[cur_].p, eq_d_d_.p, n, x_[cur_].p, stream);
        launch_eq_mul(y_[cur_].p, eq_E_d_.p, R, y_[cur_].p, stream);
        launches += (n > 0) + (R > 0);
    }
    const double* xw = x_[cur_].p;
    const double* yw = y_[cur_].p;
    if (res_slack_d_.n < (size_t)R) res_slack_d_.alloc_raw((size_t)R);
    if (R > 0 && !cones_only_) {
        k_slack<<<std::max(1, std::min(reduce_blocks_, ceil_div(R * 32, 256))), 256, 0, stream>>>(
            (int)R, M_.rowptr.p, M_.colidx.p, M_.val_orig.p, xw, b_orig_d_.p, (int)p, h_orig_d_.p, res_slack_d_.p);
        launches++;
    }
    st.lap("finish: dual_feas");
    double dfeas = dual_feas_device(yw, c_factor);
    host_reduce(&dfeas, 1, 1);
    st.lap("finish: outputs");
    out->status = stop_reason_;
    snprintf(out->status_string, PROXSDP_STATUS_STRING_LEN, "%s", stop_reason_string_.c_str());
    auto d2h = [&](double* dst, const double* src, long long cnt) {
        if (!dst || cnt <= 0) return;
        PB_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost, stream));
        g_d2h_bytes += (long long)sizeof(double) * cnt;
    };
    if (identity_) {
        d2h(out->primal, xw, n);
        d2h(out->dual_cone, res_dc_d_.p, n);
    } else {
        // user order: result[i] = v[var_ordering[i]] (pdhg.jl:768-769)
        if (res_user_d_.n < (size_t)(2 * n)) res_user_d_.alloc_raw((size_t)(2 * n));
        if (out->primal) { launch_gather(xw, var_ordering_d_.p, n, res_user_d_.p, stream); launches++; d2h(out->primal, res_user_d_.p, n); }
        if (out->dual_cone) { launch_gather(res_dc_d_.p, var_ordering_d_.p, n, res_user_d_.p + n, stream); launches++; d2h(out->dual_cone, res_user_d_.p + n, n); }
    }
    d2h(out->dual_eq, yw, p);
    d2h(out->dual_in, yw + p, m);
    d2h(out->slack_eq, res_slack_d_.p, p);
    d2h(out->slack_in, res_slack_d_.p + p, m);
    PB_CUDA(cudaStreamSynchronize(stream));
    out->primal_residual = equa_feasibility_;
    out->dual_residual = ineq_feasibility_;
    out->objval = prim_obj_.get(iter_);
    out->dual_objval = dual_obj_.get(iter_);
    out->gap = dual_gap_.get(iter_);
    out->time = now_s() - time0_;
    out->iter = iter_;
    long long fr = 0;
    for (long long r : current_rank) fr += r;
    if (sharded()) { double f = (double)fr; host_reduce(&f, 1, 0); fr = (long long)f; }
    out->final_rank = fr;
    out->primal_feasible_user_tol = feasibility_.get(iter_) <= opt.tol_feasibility;
    out->dual_feasible_user_tol = dfeas <= opt.tol_feasibility_dual;
    out->certificate_found = certificate_found_;
    out->result_count = 1;
    out->final_primal_res = primal_residual_.get(iter_);
    out->final_dual_res = dual_residual_.get(iter_);
    st.lap("done");
}

void Solver::rank_increment_rule(int idx) {
    // pdhg.jl:271-279 / 294-302
    if (opt.freeze_target_rank) return;
    if (current_rank[(size_t)idx] + opt.rank_slack >= target_rank[(size_t)idx]) {
        if (min_eig[(size_t)idx] > opt.tol_psd) {
            long long t = opt.rank_increment == 0 ? opt.rank_increment_factor * target_rank[(size_t)idx]
                                                  : opt.rank_increment_factor + target_rank[(size_t)idx];
            target_rank[(size_t)idx] = std::min<long long>(t, cones[(size_t)idx].side);
        }
    }
}

void Solver::record_trace(proxsdp_result_t* out) {
    if (!out->trace || out->trace_len >= opt.trace_cap) return;
    double* row = out->trace + out->trace_len * PROXSDP_TRACE_COLS;
    double tr = 0, cr = 0, me = 0;
    for (int i = 0; i < n_sdp; ++i) {
        tr += (double)target_rank[(size_t)i]; cr += (double)current_rank[(size_t)i];
        if (i == 0 || min_eig[(size_t)i] < me) me = min_eig[(size_t)i];
    }
    row[0] = (double)iter_; row[1] = prim_obj_.get(iter_); row[2] = dual_obj_.get(iter_); row[3] = dual_gap_.get(iter_);
    row[4] = feasibility_.get(iter_); row[5] = primal_residual_.get(iter_); row[6] = dual_residual_.get(iter_);
    row[7] = primal_step_; row[8] = beta_; row[9] = tr; row[10] = cr; row[11] = me;
    row[12] = (double)(lanczos_matvecs - trace_mv0_); row[13] = (double)(linesearch_trials - trace_ls0_);
    out->trace_len++;
}

long long Solver::seam_dual_step(double primal_step, double primal_step_old, double theta, double beta, double dual_step, double* out4) {
    primal_step_ = primal_step; primal_step_old_ = primal_step_old; theta_ = theta; beta_ = beta; dual_step_ = dual_step;
    reset_scalars();
    const double tau0 = opt.line_search_flag ? primal_step_ * std::sqrt(1.0 + theta_) : primal_step_;   // pdhg.jl:541
    int ntr = opt.line_search_flag ? (int)std::min<long long>(ladder_, std::max<long long>(opt.max_linsearch_steps, 1)) : 1;
    if (fused_ladder_) launch_ladder(0, ntr, tau0);
    else for (int t = 0; t < ntr; ++t) launch_dual_trial(t, tau0);
    sync_scalars();
    long long evals = (long long)scal_host[S_LS_EVALS];
    bool exhausted = false;
    double last_tau = 0.0;
    if (opt.line_search_flag && scal_host[S_LS_ACCEPTED] == 0.0) evals = linesearch_continue(tau0, exhausted, last_tau);
    if (opt.line_search_flag) {       // linesearch! epilogue (pdhg.jl:577-579)
        primal_step_ = scal_host[S_TAU];
        theta_ = (exhausted ? last_tau : primal_step_) / primal_step_old_;
        primal_step_old_ = primal_step_;
        dual_step_ = beta_ * primal_step_;
    } else {                          // dual_step! (pdhg.jl:606)
        primal_step_old_ = primal_step_;
    }
    out4[0] = primal_step_; out4[1] = theta_; out4[2] = dual_step_; out4[3] = primal_step_old_;
    return std::max<long long>(evals, 1);
}

void Solver::seam_residuals(double primal_step, double dual_step, double beta, double norm_b, double norm_h, double norm_c, double* out8) {
    primal_step_ = primal_step; dual_step_ = dual_step; beta_ = beta; norm_b_ = norm_b; norm_h_ = norm_h; norm_c_ = norm_c;
    reset_scalars();
    // the kernels take the accepted step from the record (what the line search leaves there)
    double rec[3] = {1.0, 0.0, primal_step};
    PB_CUDA(cudaMemcpyAsync(scal_d_.p + S_LS_ACCEPTED, rec, sizeof(rec), cudaMemcpyHostToDevice, stream));
    int blocksN = std::max(1, std::min(std::min(reduce_blocks_, 4 * num_sms_), ceil_div(n, 512)));      // one resident wave
    k_residual_primal<<<blocksN, 256, 0, stream>>>(n, x_[1].p, x_[0].p, Mty_[1].p, Mty_[0].p, c_.p, scal_d_.p, ws_);
    int blocksR = std::max(1, std::min(reduce_blocks_, ceil_div(R, 256)));
    k_residual_dual<<<blocksR, 256, 0, stream>>>((int)p, (int)m, beta_, 0, dual_step_, y_[1].p, y_[0].p, Mx_[1].p, Mx_[0].p,
                                                 b_.p, h_.p, scal_d_.p, ws_);
    launches += 2;
    sync_scalars();
    iter_ = 1;
    host_residuals(1);
    out8[0] = primal_residual_.get(1); out8[1] = dual_residual_.get(1); out8[2] = comb_residual_.get(1);
    out8[3] = equa_feasibility_; out8[4] = ineq_feasibility_; out8[5] = prim_obj_.get(1); out8[6] = dual_obj_.get(1);
    out8[7] = dual_gap_.get(1);
}

// The line search after its first ladder of trials came back without an accepted step (pdhg.jl:543-571): further
// ladders, and — when max_linsearch_steps is exhausted — the reference's exit state (last trial's vectors, the step
// decayed once more).  Returns the number of trials evaluated.
long long Solver::linesearch_continue(double tau0, bool& exhausted, double& last_tau) {
    long long t = std::min<long long>(ladder_, std::max<long long>(opt.max_linsearch_steps, 1));
    while (scal_host[S_LS_ACCEPTED] == 0.0 && t < opt.max_linsearch_steps) {
        if (fused_ladder_) {
            // the next batch of trials, again side by side
            const int T = (int)std::min<long long>(ladder_, opt.max_linsearch_steps - t);
            launch_ladder((int)t, T, tau0);
            sync_scalars();
            t = (long long)scal_host[S_LS_EVALS];
        } else {
            launch_dual_trial((int)t, tau0);
            sync_scalars();
            ++t;
        }
    }
    if (scal_host[S_LS_ACCEPTED] == 0.0) {
        // loop exhausted: the reference keeps the last trial's y/Mty and the once-more decayed step
        double tau = tau0;
        for (long long q = 0; q < opt.max_linsearch_steps; ++q) { if (q == opt.max_linsearch_steps - 1) last_tau = tau; tau *= opt.linsearch_decay; }
        exhausted = true;
        double rec[3] = {1.0, (double)(opt.max_linsearch_steps - 1), tau};
        PB_CUDA(cudaMemcpyAsync(scal_d_.p + S_LS_ACCEPTED, rec, sizeof(rec), cudaMemcpyHostToDevice, stream));
        PB_CUDA(cudaStreamSynchronize(stream));
        scal_host[S_LS_ACCEPTED] = 1.0; scal_host[S_TAU] = tau;
    }
    return t;
}

// compute_residual! (residuals.jl:37-71) and compute_gap! (residuals.jl:2-35): the scalar part, on the record of
// iteration k that the residual kernels filled
void Solver::host_residuals(long long k) {
    {
        double den = jl_max(jl_max(scal_host[S_RES_P_DEN], norm_b_), jl_max(norm_h_, 1.0));
        double pr = std::sqrt((double)(sharded() ? global_n_ : n)) * scal_host[S_RES_P_NUM] / den;
        double den2 = jl_max(jl_max(scal_host[S_RES_D_DEN], norm_c_), 1.0);
        double dr = std::sqrt((double)(sharded() ? global_R_ : R)) * scal_host[S_RES_D_NUM] / den2;
        primal_residual_.set(k, pr); dual_residual_.set(k, dr); comb_residual_.set(k, jl_max(pr, dr));
    }
    {
        const long long gp = sharded() ? global_p_ : p, gm = sharded() ? global_m_ : m;
        if (gp > 0) equa_feasibility_ = scal_host[S_EQ_MAX] / (1.0 + norm_b_);
        if (gm > 0) ineq_feasibility_ = scal_host[S_IN_MAX] / (1.0 + norm_h_);
        feasibility_.set(k, std::max(equa_feasibility_, ineq_feasibility_));
        double po = scal_host[S_PRIM_OBJ], dobj = 0.0;
        if (gp > 0) dobj -= scal_host[S_BY];
        if (gm > 0) dobj -= scal_host[S_HY];
        prim_obj_.set(k, po); dual_obj_.set(k, dobj);
        dual_gap_.set(k, std::fabs(po - dobj) / (1.0 + std::fabs(po) + std::fabs(dobj)));
    }
}

// ---------------------------------------------------------------------------
// chambolle_pock main loop (pdhg.jl:145-530)
// ---------------------------------------------------------------------------
void Solver::log_progress(double dual_feas_val) {      // print_progress (printing.jl:99-151)
    long long tr = 0;
    for (long long r : target_rank) tr += r;
    plog::emit(plog::progress(iter_, prim_obj_.get(iter_), dual_gap_.get(iter_), feasibility_.get(iter_), primal_residual_.get(iter_),
                            dual_residual_.get(iter_), tr, now_s() - time0_, dual_obj_.get(iter_), dual_feas_val,
                            opt.extended_log != 0, opt.extended_log2 != 0, opt.log_repeat_header != 0));
}

void Solver::solve(proxsdp_result_t* out) {
    begin(out);
    run(-1, false);
    finish(out);
}

void Solver::begin(proxsdp_result_t* out) {
    time0_ = now_s();
    out_ = out;
    out->trace_len = 0;
    ada_count_ = 0; have_cached_ = false; loop_done_ = false; k_next_ = 1; t_loop_accum_ = 0;
    if (logging()) {                              // pdhg.jl:43-52
        std::vector<long long> sides, lens;
        for (const ConeDev& cd : cones) sides.push_back(cd.side);
        for (int l : soc_len_h_) lens.push_back(l);
        std::string hdr = plog::header_1();
        hdr += plog::parameters(opt.tol_gap, opt.tol_feasibility, opt.tol_primal, opt.tol_dual, opt.tol_soc, opt.tol_psd,
                               n_soc >= 1, n_sdp >= 1, opt.max_iter_local, opt.time_limit);
        hdr += plog::constraints(p, m);
        if (n_soc + n_sdp > 0) hdr += plog::prob_data(lens, sides);
        hdr += plog::header_2(opt.extended_log != 0, opt.extended_log2 != 0, true);
        plog::emit(hdr);
    }

    // advanced initialisation (pdhg.jl:138-142): x = tau*c ; Mx = M x ; Mx_old = M*0 = 0
    if (opt.advanced_initialization) {
        if (n > 0) {
            k_scale_copy<<<std::max(1, std::min(reduce_blocks_, ceil_div(n, 256))), 256, 0, stream>>>(c_.p, primal_step_, n, x_[cur_].p);
            launches++;
        }
        PB_CUDA(cudaMemsetAsync(scal_d_.p, 0, sizeof(double) * (size_t)scal_len, stream));
        // pdhg.jl:140 writes a.Mx; the first primal_step! overwrites it, and a.Mx_old = M*x_old = 0.
        launch_spmv(M_, x_[cur_].p, Mx_[cur_].p);
        PB_CUDA(cudaMemsetAsync(Mx_[cur_].p, 0, sizeof(double) * (size_t)std::max<long long>(R, 1), stream));
    }
    out->time_setup = 0.0;   // filled by the caller (constructor time)
}

// Runs up to max_steps (< 0: unbounded) further iterations of the CP loop (pdhg.jl:145-484).
bool Solver::run(long long max_steps, bool flush_l2) {
    proxsdp_result_t* out = out_;
    if (loop_done_) return true;
    auto append = [&](const char* s) { stop_reason_string_ += s; };
    long long& ada_count = ada_count_;
    bool& have_cached = have_cached_;
    double t_loop0 = now_s();
    struct LoopTimer { double& acc; double t0; ~LoopTimer() { acc += now_s() - t0; } } loop_timer{t_loop_accum_, t_loop0};
    if (flush_l2 && flush_buf_.n == 0) flush_buf_.alloc((size_t)24 << 20);   // 192 MiB > 126 MB L2

    const long long kmax = 2 * opt.max_iter_local;
    long long steps = 0;
    loop_done_ = true;      // cleared again if we leave because of max_steps
    for (long long k = k_next_; k <= kmax; ++k) {
        if (max_steps >= 0 && steps >= max_steps) { loop_done_ = false; break; }
        ++steps;
        k_next_ = k + 1;
        iter_ = k;
        trace_mv0_ = lanczos_matvecs; trace_ls0_ = linesearch_trials;
        if (flush_l2) {
            PB_CUDA(cudaEventRecord(ev_fl0_, stream));
            PB_CUDA(cudaMemsetAsync(flush_buf_.p, 0, flush_buf_.n * sizeof(double), stream));
            PB_CUDA(cudaEventRecord(ev_fl1_, stream));
        }
        // ------------------------------------------------------------------ device work
        if (!scal_clean_) reset_scalars();
        double tau_primal = primal_step_;                       // pdhg.jl:622 uses the current step
        psd_projection_launch(k, tau_primal, false);
        double tau0 = opt.line_search_flag ? primal_step_ * std::sqrt(1.0 + theta_) : primal_step_;   // pdhg.jl:541
        launch_post_eig(tau0, true);
        PB_CUDA(cudaEventRecord(ev_post1_, stream));
        sync_scalars(true);
        {
            float ms = 0.f;
            if (n_sdp > 0) {
                if (cudaEventElapsedTime(&ms, ev_psd0_, ev_psd1_) == cudaSuccess) time_psd_ms_ += ms;
                n_psd_++;
            }
            if (lz_timed_ && cudaEventElapsedTime(&ms, ev_lz0_, ev_lz1_) == cudaSuccess) { time_lanczos_ms_ += ms; lanczos_timed_calls++; }
            if (cudaEventElapsedTime(&ms, ev_psd1_, ev_post1_) == cudaSuccess) time_post_ms_ += ms;
            if (flush_l2 && cudaEventElapsedTime(&ms, ev_fl0_, ev_fl1_) == cudaSuccess) time_flush_ms_ += ms;
        }
        // ---- Krylov fallback (prox_operators.jl:55-57): redo the failed cones exactly, then the tail
        if (scal_host[S_POISON] != 0.0) {
            fallback_projection(k);
            launch_post_eig(tau0, false);
            double keep_ops = scal_host[S_NUMOPS];
            sync_scalars(true);
            scal_host[S_NUMOPS] = keep_ops;
        }
        // ---- linesearch beyond the speculative ladder (pdhg.jl:543-571)
        long long evals = (long long)scal_host[S_LS_EVALS];
        bool ls_exhausted = false;
        double ls_last_tau = 0.0;
        if (opt.line_search_flag && scal_host[S_LS_ACCEPTED] == 0.0) {
            const double keep_ops = scal_host[S_NUMOPS];
            evals = linesearch_continue(tau0, ls_exhausted, ls_last_tau);
            int blocksN = std::max(1, std::min(std::min(reduce_blocks_, 4 * num_sms_), ceil_div(n, 512)));      // one resident wave
            k_residual_primal<<<blocksN, 256, 0, stream>>>(n, x_[1 - cur_].p, x_[cur_].p, Mty_[1 - cur_].p, Mty_[cur_].p, c_.p, scal_d_.p, ws_, nzmask_.p);
            int blocksR = std::max(1, std::min(reduce_blocks_, ceil_div(R, 256)));
            k_residual_dual<<<blocksR, 256, 0, stream>>>((int)p, (int)m, beta_, 1, dual_step_, y_[1 - cur_].p, y_[cur_].p,
                                                         Mx_[1 - cur_].p, Mx_[cur_].p, b_.p, h_.p, scal_d_.p, ws_);
            launches += 2;
            sync_scalars(true);
            scal_host[S_NUMOPS] = keep_ops;
        }
        linesearch_trials += std::max<long long>(evals, 1);
        lanczos_matvecs += (long long)scal_host[S_NUMOPS];

        // ------------------------------------------------------------------ host scalars
        // linesearch! epilogue (pdhg.jl:577-579) / dual_step! (pdhg.jl:606)
        if (opt.line_search_flag) {
            primal_step_ = scal_host[S_TAU];
            // theta belongs to the last trial that was evaluated (pdhg.jl:544): when the loop runs out of trials the
            // step is decayed once more (pdhg.jl:569) but theta is not recomputed
            theta_ = (ls_exhausted ? ls_last_tau : primal_step_) / primal_step_old_;
            primal_step_old_ = primal_step_;
            dual_step_ = beta_ * primal_step_;
        } else {
            primal_step_old_ = primal_step_;
        }
        for (int q = 0; q < n_sdp; ++q) {
            current_rank[(size_t)q] = (long long)scal_host[S_HEADER + 3 * q + 0];
            min_eig[(size_t)q] = scal_host[S_HEADER + 3 * q + 1];
        }
        soc_gap_max_ = (n_soc > 0 || global_has_soc_) ? scal_host[S_SOC_GAP] : -1.0;
        host_residuals(k);
        cur_ = 1 - cur_;    // keep-old copies (residuals.jl:65-68) are a pointer swap

        if ((opt.check_dual_feas && (k % opt.check_dual_feas_freq) == 0) ||
            (opt.log_verbose && opt.log_freq > 0 && (k % opt.log_freq) == 0 && opt.extended_log2)) {     // pdhg.jl:166-173
            double f = stop_reason_ == 6 ? 0.0 : 1.0;
            dual_feasibility_ = dual_feas_device(y_[cur_].p, f);
            host_reduce(&dual_feasibility_, 1, 1);
            dual_feasibility_check_ = true;
        } else {
            dual_feasibility_check_ = false;
        }
        if (logging() && opt.log_freq > 0 && (k % opt.log_freq) == 0) log_progress(dual_feasibility_);   // pdhg.jl:176-178
        record_trace(out);

        if (iter_ < certificate_search_min_iter_) continue;                   // pdhg.jl:180-182

        if (opt.certificate_search && certificate_search_) {                  // pdhg.jl:184-244
            if (stop_reason_ == 6) {
                if (dual_obj_.get(k) > +opt.certificate_obj_tol) {
                    dual_feasibility_ = dual_feas_device(y_[cur_].p, 0.0);
                    host_reduce(&dual_feasibility_, 1, 1);
                    dual_feasibility_check_ = true;
                    if (dual_feasibility_ < opt.tol_feasibility_dual) {
                        certificate_found_ = true;
                        append(" [Dual ray found]");
                        if (logging()) plog::emit(plog::note("Dual ray found"));
                        break;
                    }
                }
            } else {
                if (prim_obj_.get(k) < -opt.certificate_obj_tol) {
                    if (feasibility_.get(iter_) < opt.tol_feasibility) {
                        certificate_found_ = true;
                        append(" [Primal ray found]");
                        if (logging()) plog::emit(plog::note("Primal ray found"));
                        break;
                    }
                }
            }
            double cr = comb_residual_.get(k);
            if ((prim_obj_.get(k) < -opt.certificate_fail_tol && dual_obj_.get(k) < -opt.certificate_fail_tol &&
                 feasibility_.get(iter_) < -opt.certificate_fail_tol) || cr != cr) {
                append(" [Failed to find certificate]");
                if (logging()) plog::emit(plog::note("Failed to finds certificate"));
                break;
            }
        }

        // convergence check (pdhg.jl:247-332)
        rank_update_ += 1;
        const double gap_k = dual_gap_.get(iter_), feas_k = feasibility_.get(iter_);
        const double pr_k = primal_residual_.get(k), dr_k = dual_residual_.get(k);
        if (gap_k <= opt.tol_gap && feas_k <= opt.tol_feasibility &&
            (!opt.check_dual_feas || dual_feasibility_ < opt.tol_feasibility_dual)) {
            bool conv_rank = true;                                            // residuals.jl:88-101
            for (int q = 0; q < n_sdp; ++q)
                if (!(cones[(size_t)q].side < opt.min_size_krylov_eigs || target_rank[(size_t)q] > opt.max_target_rank_krylov_eigs ||
                      min_eig[(size_t)q] < opt.tol_psd)) { conv_rank = false; break; }
            bool conv_soc = !((n_soc > 0 || global_has_soc_) && soc_gap_max_ >= opt.tol_soc);      // residuals.jl:73-86
            if (sharded()) { double f = conv_rank ? 0.0 : 1.0; host_reduce(&f, 1, 1); conv_rank = (f == 0.0); }
            if (conv_rank && conv_soc && iter_ > opt.min_iter) {
                if (!certificate_search_) {
                    stop_reason_ = 1;
                    stop_reason_string_ = "Optimal solution found";
                } else {
                    append(" [Failed to find certificate - type 2]");
                    if (logging()) plog::emit(plog::note("Failed to find certificate - type 2"));
                    break;
                }
                break;
            } else if (rank_update_ > window_) {
                update_cont_ += 1;
                if (update_cont_ > 0) {
                    for (int q = 0; q < n_sdp; ++q) rank_increment_rule(q);
                    rank_update_ = 0; update_cont_ = 0;
                }
            }
        } else if (k > window_ && comb_residual_.get(k - window_) < comb_residual_.get(k) && rank_update_ > window_) {
            update_cont_ += 1;
            if (update_cont_ > opt.divergence_min_update) {
                bool any_room = false;
                for (int q = 0; q < n_sdp; ++q) {
                    if (target_rank[(size_t)q] < cones[(size_t)q].side) any_room = true;
                    rank_increment_rule(q);
                }
                if (sharded()) { double f = any_room ? 1.0 : 0.0; host_reduce(&f, 1, 1); any_room = (f != 0.0); }
                if (any_room) { rank_update_ = 0; update_cont_ = 0; }     // the counters are whole-problem state
            }
        } else if (pr_k > opt.tol_primal && dr_k < opt.tol_dual && k > window_) {
            ada_count += 1;
            if (ada_count > opt.adapt_window) {
                ada_count = 0;
                if (opt.line_search_flag) { beta_ *= (1.0 - adapt_level_); primal_step_ /= std::sqrt(1.0 - adapt_level_); }
                else { primal_step_ /= (1.0 - adapt_level_); dual_step_ *= (1.0 - adapt_level_); }
                adapt_level_ *= opt.adapt_decay;
            }
        } else if (pr_k < opt.tol_primal && dr_k > opt.tol_dual && k > window_) {
            ada_count += 1;
            if (ada_count > opt.adapt_window) {
                ada_count = 0;
                if (opt.line_search_flag) { beta_ /= (1.0 - adapt_level_); primal_step_ *= std::sqrt(1.0 - adapt_level_); }
                else { primal_step_ *= (1.0 - adapt_level_); dual_step_ /= (1.0 - adapt_level_); }
                adapt_level_ *= opt.adapt_decay;
            }
        }

        auto start_cert_infeas = [&]() {       // certificate_infeasibility (pdhg.jl:655-676)
            if (logging()) plog::emit(plog::note("Begin search for infeasibility certificate"));
            PB_CUDA(cudaMemsetAsync(c_.p, 0, sizeof(double) * (size_t)std::max<long long>(n, 1), stream));
            certificate_search_min_iter_ = iter_ + 2 * opt.convergence_window + iter_ / 5 + 1000;
            certificate_search_ = true;
            opt.time_limit *= 1.1;
            opt.max_iter_local = opt.max_iter_local + opt.max_iter_local / 10;
            cache_solution(1.0, out); have_cached = true;
        };
        auto start_cert_dual_infeas = [&]() {  // certificate_dual_infeasibility (pdhg.jl:639-653)
            if (logging()) plog::emit(plog::note("Begin search for dual infeasibility certificate"));
            std::fill(b_host_.begin(), b_host_.end(), 0.0);
            std::fill(h_host_.begin(), h_host_.end(), 0.0);
            PB_CUDA(cudaMemsetAsync(b_.p, 0, sizeof(double) * (size_t)std::max<long long>(p, 1), stream));
            PB_CUDA(cudaMemsetAsync(h_.p, 0, sizeof(double) * (size_t)std::max<long long>(m, 1), stream));
            certificate_search_min_iter_ = iter_ + 2 * opt.convergence_window + iter_ / 5 + 1000;
            certificate_search_ = true;
            opt.time_limit *= 1.1;
            opt.max_iter_local = opt.max_iter_local + opt.max_iter_local / 10;
            cache_solution(1.0, out); have_cached = true;
        };

        // max_iter or time limit (pdhg.jl:335-382)
        const double elapsed_k = sharded() ? scal_host[S_ELAPSED] : now_s() - time0_;      // sharded: max over ranks, so all ranks stop together
        if (iter_ >= opt.max_iter_local || elapsed_k >= opt.time_limit) {
            if (iter_ > opt.min_iter_time_infeas && dual_gap_.max_abs_diff() < opt.infeas_stable_gap_tol &&
                dual_gap_.get(k) > opt.infeas_limit_gap_tol) {
                if (feasibility_.get(iter_) <= opt.tol_feasibility / 100) {
                    stop_reason_ = 5;
                    stop_reason_string_ = "Problem declared unbounded due to lack of improvement";
                    if (opt.certificate_search && !certificate_search_) start_cert_dual_infeas();
                    else if (opt.certificate_search && certificate_search_) {}
                    else break;
                } else if (feasibility_.get(iter_) > opt.infeas_feasibility_tol) {
                    stop_reason_ = 6;
                    stop_reason_string_ = "Problem declared infeasible due to lack of improvement";
                    if (opt.certificate_search && !certificate_search_) start_cert_infeas();
                    else if (opt.certificate_search && certificate_search_) {}
                    else break;
                }
            } else if (iter_ >= opt.max_iter_local) {
                stop_reason_ = 3;
                stop_reason_string_ = "Iteration limit of " + std::to_string(opt.max_iter_local) + " was hit";
                if (opt.warn_on_limit && rank_ == 0) fprintf(stderr, "Warning:     WARNING: Iteration limit hit.\n");   // @warn, pdhg.jl:369-371
            } else {
                stop_reason_ = 2;
                char buf[160];
                snprintf(buf, sizeof(buf), "Time limit hit, limit: %g time: %g", opt.time_limit, now_s() - time0_);
                stop_reason_string_ = buf;
                if (opt.warn_on_limit && rank_ == 0) plog::emit("    WARNING: Time limit hit.\n");                          // pdhg.jl:375-377
            }
            if (iter_ >= opt.max_iter_local || elapsed_k >= opt.time_limit) break;
        }

        if (opt.certificate_search && certificate_search_) continue;          // pdhg.jl:385-387

        const double dobj_k = dual_obj_.get(k), pobj_k = prim_obj_.get(k);
        char buf[200];
        if ((iter_ > opt.min_iter_max_obj && dobj_k > opt.max_obj) || dobj_k != dobj_k) {        // pdhg.jl:390-405
            stop_reason_ = 6;
            snprintf(buf, sizeof(buf), "Infeasible: |Dual objective| = %g > maximum allowed = %g", dobj_k, opt.max_obj);
            stop_reason_string_ = buf;
            if (opt.certificate_search && !certificate_search_) start_cert_infeas(); else break;
        }
        if ((iter_ > opt.min_iter_max_obj && pobj_k < -opt.max_obj) || pobj_k != pobj_k) {        // pdhg.jl:408-422
            stop_reason_ = 5;
            snprintf(buf, sizeof(buf), "Unbounded: |Primal objective| = %g > maximum allowed = %g", pobj_k, opt.max_obj);
            stop_reason_string_ = buf;
            if (opt.certificate_search && !certificate_search_) start_cert_dual_infeas(); else break;
        }
        if (iter_ > opt.min_iter_max_obj && dual_gap_.get(k) > opt.infeas_limit_gap_tol &&       // pdhg.jl:425-444
            feasibility_.get(iter_) > opt.infeas_feasibility_tol &&
            feasibility_.max_abs_diff() < opt.infeas_stable_feasibility_tol) {
            stop_reason_ = 6;
            snprintf(buf, sizeof(buf), "Infeasible: feasibility stalled at %g", feasibility_.get(iter_));
            stop_reason_string_ = buf;
            if (opt.certificate_search && !certificate_search_) start_cert_infeas(); else break;
        }
        if (iter_ > opt.min_iter_max_obj && dual_gap_.get(k) > 1 - opt.infeas_gap_tol &&        // pdhg.jl:447-483
            dual_gap_.max_abs_diff() < opt.infeas_stable_gap_tol) {
            if (std::fabs(dobj_k) > std::fabs(pobj_k) && feasibility_.get(iter_) > opt.infeas_feasibility_tol) {
                stop_reason_ = 6;
                stop_reason_string_ = "Infeasible: duality gap stalled at 100 % with |Dual objective| >> |Primal objective|";
                if (opt.certificate_search && !certificate_search_) start_cert_infeas(); else break;
            } else if (std::fabs(pobj_k) > std::fabs(dobj_k) && feasibility_.get(iter_) <= opt.tol_feasibility) {
                stop_reason_ = 5;
                stop_reason_string_ = "Unbounded: duality gap stalled at 100 % with |Dual objective| << |Primal objective|";
                if (opt.certificate_search && !certificate_search_) start_cert_dual_infeas(); else break;
            }
        }
    }
    return loop_done_;
}

void Solver::finish(proxsdp_result_t* out) {
    if (lz_prof_.n) {
        PB_CUDA(cudaStreamSynchronize(stream));
        std::vector<long long> pr = lz_prof_.download();
        const char* nm_rows[8] = {"symv", "fold", "pass1+xchg", "upd1+pass2+xchg", "upd2+beta", "publish+vxchg", "ritz", "loop-top"};
        const char* nm_cl[8] = {"symv", "fold+gridxchg+gather(+local)", "gs dots+push", "cluster.sync", "beta+publish v", "ritz", "gs reduce+update", "loop-top"};
        const char** nm = lz_cluster_launches_ > 0 ? nm_cl : nm_rows;
        long long tot = 0; for (int i = 0; i < 8; ++i) tot += pr[i];
        fprintf(stderr, "[lz-prof] ritz: bisection accepted %lld, dense jacobi %lld, cycles up to the decision %lld\n", pr[8], pr[9], pr[10]);
        fprintf(stderr, "[lz-prof] ritz_top_bi cycles: setup %lld, multisection %lld, vectors %lld, checks %lld\n", pr[11], pr[12], pr[13], pr[14]);
        const double mv = (double)std::max<long long>(lanczos_matvecs, 1);
        for (int i = 0; i < 8; ++i) fprintf(stderr, "[lz-prof] %-30s %12lld cyc  %5.1f%%  (%.2f us/matvec)\n", nm[i], pr[i], 100.0 * pr[i] / std::max<long long>(tot, 1), pr[i] / 1965.0 / mv);
        if (lz_kernel_ == 3 && lz_cluster_launches_ > 0) {
            fprintf(stderr, "[lz-prof] us/matvec per profiled CTA (0, C-1, G/2, G-1):\n");
            const char* nm2[10] = {"symv+reduce+publish", "exchange+alpha+local", "gs dots+push", "cluster.sync 1", "reduce+update+publish", "cluster.sync 2", "ritz", " symv: warp 0 loads+FMA+park", " symv: wait for slowest warp", " symv: row reduce+publish"};
            const int sl[10] = {0, 1, 2, 3, 6, 4, 5, 15, 16, 17};
            for (int q = 0; q < 10; ++q) {
                fprintf(stderr, "[lz-prof]   %-32s", nm2[q]);
                for (int r = 0; r < 4; ++r) fprintf(stderr, " %6.2f", pr[(size_t)32 * r + sl[q]] / 1965.0 / mv);
                fprintf(stderr, "\n");
            }
        }
    }
    out->time_loop = t_loop_accum_;
    if (opt.log_verbose) {                        // pdhg.jl:486-505
        const double time_ = now_s() - time0_;
        double val = -1.0;
        if (opt.extended_log2) {
            val = dual_feas_device(y_[cur_].p, stop_reason_ == 6 ? 0.0 : 1.0);
            host_reduce(&val, 1, 1);
        }
        if (logging()) {
            log_progress(val);
            long long max_rank = 0;
            for (long long r : current_rank) max_rank = std::max(max_rank, r);
            plog::emit(plog::result(stop_reason_string_, time_, prim_obj_.get(iter_), dual_obj_.get(iter_), dual_gap_.get(iter_),
                                  equa_feasibility_, ineq_feasibility_, max_rank));
        }
    }
    // results (pdhg.jl:486-529)
    if (opt.certificate_search && certificate_search_) {
        if (certificate_found_) {
            cache_solution(stop_reason_ == 6 ? 0.0 : 1.0, out);     // pdhg.jl:515-517: c_orig .*= 0 for a dual ray
        } else if (!have_cached_) {
            cache_solution(1.0, out);
        }
    } else {
        cache_solution(1.0, out);
    }
    out->time_psd_proj = time_psd_ms_ * 1e-3;
    out->n_psd_proj = n_psd_;
    out->lanczos_matvecs = lanczos_matvecs;
    out->lanczos_calls = lanczos_calls;
    out->full_eig_calls = full_eig_calls;
    out->linesearch_trials = linesearch_trials;
    out->gpu_launches = launches;
    out->time_lanczos = time_lanczos_ms_ * 1e-3;
    out->time_rest = time_post_ms_ * 1e-3;
    out->time_l2_flush = time_flush_ms_ * 1e-3;
    out->lanczos_timed_calls = lanczos_timed_calls;
    out->h2d_bytes = g_h2d_bytes;
    out->d2h_bytes = g_d2h_bytes;
    out->implicit_calls = implicit_calls;
    if (out->target_rank) for (int q = 0; q < n_sdp; ++q) out->target_rank[q] = target_rank[(size_t)q];
}

}  // namespace pb

// ===========================================================================
// C ABI
// ===========================================================================
using namespace pb;

template <class F>
static int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const CudaError& e) {
        g_last_error = e.what();
        cudaGetLastError();
        return e.code;
    } catch (const std::bad_alloc&) {
        g_last_error = "host allocation failed";
        return -4;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return -1;
    }
}

// a problem that consists of variable cones only (no rows, zero objective)
struct ConesOnlyProblem {
    std::vector<int64_t> sdp_ptr, sdp_idx, soc_ptr, soc_idx;
    std::vector<double> zeros;
    proxsdp_problem_t pr{};
    ConesOnlyProblem(int64_t n_sdp, const int64_t* sides, int64_t n_soc, const int64_t* lens, const double* resid) {
        int64_t N = 0;
        sdp_ptr.assign((size_t)n_sdp + 1, 0);
        for (int64_t k = 0; k < n_sdp; ++k) { N += sides[k] * (sides[k] + 1) / 2; sdp_ptr[(size_t)k + 1] = N; }
        int64_t Npsd = N;
        soc_ptr.assign((size_t)n_soc + 1, 0);
        for (int64_t k = 0; k < n_soc; ++k) { N += lens[k]; soc_ptr[(size_t)k + 1] = N - Npsd; }
        sdp_idx.resize((size_t)Npsd);
        std::iota(sdp_idx.begin(), sdp_idx.end(), (int64_t)0);
        soc_idx.resize((size_t)(N - Npsd));
        std::iota(soc_idx.begin(), soc_idx.end(), Npsd);
        zeros.assign((size_t)std::max<int64_t>(N, 1), 0.0);
        pr.n = N; pr.p = 0; pr.m = 0; pr.index_base = 0;
        pr.b = zeros.data(); pr.h = zeros.data(); pr.c = zeros.data();
        pr.n_sdp = n_sdp; pr.sdp_side = sides; pr.sdp_ptr = sdp_ptr.data(); pr.sdp_idx = sdp_idx.data();
        pr.n_soc = n_soc; pr.soc_ptr = soc_ptr.data(); pr.soc_idx = soc_idx.data();
        pr.eig_resid = resid;
    }
};

extern "C" {

int proxsdp_b200_solve(const proxsdp_problem_t* problem, const proxsdp_options_t* options, proxsdp_result_t* result) {
    if (!problem || !options || !result) { g_last_error = "null argument"; return -1; }
    return guarded([&]() {
        double t0 = now_s();
        Solver s(problem, options);
        double t_setup = now_s() - t0;
        s.solve(result);
        result->time_setup = t_setup;
        result->time += t_setup;    // the reference's clock (p.time0, pdhg.jl:13) starts before the Init block
    });
}

int proxsdp_b200_comm_unique_id(char id[128]) {
    if (!id) { g_last_error = "null argument"; return -1; }
    return guarded([&]() {
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
        ncclUniqueId uid;
        PB_NCCL(nccl_api().GetUniqueId(&uid));
        std::memcpy(id, &uid, 128);
    });
}

int proxsdp_b200_comm_create(const char id[128], int64_t rank, int64_t nranks, int64_t device_id,
                             proxsdp_b200_comm_t** comm) {
    if (!id || !comm || nranks < 1 || rank < 0 || rank >= nranks) { g_last_error = "invalid argument"; return -1; }
    *comm = nullptr;
    return guarded([&]() {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw CudaError(-5, "no CUDA device available");
        if (device_id < 0 || device_id >= ndev) throw CudaError(-1, "device_id out of range");
        PB_CUDA(cudaSetDevice((int)device_id));
        std::unique_ptr<proxsdp_b200_comm> c(new proxsdp_b200_comm());
        c->rank = (int)rank; c->nranks = (int)nranks; c->device = (int)device_id;
        ncclUniqueId uid;
        std::memcpy(&uid, id, 128);
        PB_NCCL(nccl_api().CommInitRank(&c->comm, (int)nranks, uid, (int)rank));
        *comm = c.release();
    });
}

int proxsdp_b200_comm_destroy(proxsdp_b200_comm_t* comm) {
    if (!comm) return 0;
    return guarded([&]() {
        if (comm->comm) nccl_api().CommDestroy(comm->comm);
        delete comm;
    });
}

int proxsdp_b200_solve_sharded(const proxsdp_problem_t* problem, const proxsdp_options_t* options,
                               const proxsdp_shard_t* shard, proxsdp_result_t* result) {
    if (!problem || !options || !shard || !result) { g_last_error = "null argument"; return -1; }
    return guarded([&]() {
        double t0 = now_s();
        Solver s(problem, options, false, false, shard);
        double t_setup = now_s() - t0;
        s.solve(result);
        result->time_setup = t_setup;
        result->time += t_setup;
    });
}

struct proxsdp_b200_handle {
    std::unique_ptr<Solver> s;
    // run() records the trace and (certificate search, pdhg.jl:639-676) an early cached solution through
    // this library-owned result; finish() copies it into the caller's buffers.
    proxsdp_result_t scratch{};
    std::vector<double> primal, dual_cone, dual_eq, dual_in, slack_eq, slack_in, trace;
    std::vector<int64_t> target_rank;
    double t_setup = 0.0;
};

int proxsdp_b200_create(const proxsdp_problem_t* problem, const proxsdp_options_t* options,
                        proxsdp_b200_handle_t** handle) {
    if (!problem || !options || !handle) { g_last_error = "null argument"; return -1; }
    *handle = nullptr;
    return guarded([&]() {
        std::unique_ptr<proxsdp_b200_handle> h(new proxsdp_b200_handle());
        double t0 = now_s();
        h->s.reset(new Solver(problem, options));
        Solver& s = *h->s;
        h->primal.assign((size_t)s.n, 0.0); h->dual_cone.assign((size_t)s.n, 0.0);
        h->dual_eq.assign((size_t)s.p, 0.0); h->slack_eq.assign((size_t)s.p, 0.0);
        h->dual_in.assign((size_t)s.m, 0.0); h->slack_in.assign((size_t)s.m, 0.0);
        h->target_rank.assign((size_t)std::max(s.n_sdp, 1), 0);
        h->trace.assign((size_t)std::max<int64_t>(options->trace_cap, 0) * PROXSDP_TRACE_COLS, 0.0);
        proxsdp_result_t& r = h->scratch;
        r.primal = h->primal.data(); r.dual_cone = h->dual_cone.data(); r.dual_eq = h->dual_eq.data();
        r.dual_in = h->dual_in.data(); r.slack_eq = h->slack_eq.data(); r.slack_in = h->slack_in.data();
        r.target_rank = h->target_rank.data();
        r.trace = h->trace.empty() ? nullptr : h->trace.data();
        h->s->begin(&h->scratch);
        PB_CUDA(cudaStreamSynchronize(h->s->stream));
        h->t_setup = now_s() - t0;
        *handle = h.release();
    });
}

int proxsdp_b200_iterate(proxsdp_b200_handle_t* handle, int64_t max_steps, int64_t flush_l2,
                         int64_t* steps_done, int64_t* finished, double* device_ms) {
    if (!handle || !handle->s) { g_last_error = "null handle"; return -1; }
    return guarded([&]() {
        Solver& s = *handle->s;
        cudaEvent_t e0, e1;
        PB_CUDA(cudaEventCreate(&e0)); PB_CUDA(cudaEventCreate(&e1));
        long long it0 = s.iterations_done();
        PB_CUDA(cudaEventRecord(e0, s.stream));
        bool done = s.run(max_steps, flush_l2 != 0);
        PB_CUDA(cudaEventRecord(e1, s.stream));
        PB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        PB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (steps_done) *steps_done = s.iterations_done() - it0;
        if (finished) *finished = done ? 1 : 0;
        if (device_ms) *device_ms = (double)ms;
    });
}

int proxsdp_b200_counters(proxsdp_b200_handle_t* handle, int64_t* counts, double* times_ms) {
    if (!handle || !handle->s) { g_last_error = "null handle"; return -1; }
    Solver& s = *handle->s;
    if (counts) {
        counts[0] = s.iterations_done(); counts[1] = s.launches; counts[2] = s.lanczos_matvecs; counts[3] = s.lanczos_calls;
        counts[4] = s.lanczos_timed_calls; counts[5] = s.full_eig_calls; counts[6] = s.linesearch_trials; counts[7] = 0;
        for (long long r : s.target_rank) counts[7] += r;
    }
    if (times_ms) {
        times_ms[0] = s.time_psd_ms(); times_ms[1] = s.time_lanczos_ms_; times_ms[2] = s.time_post_ms_; times_ms[3] = s.time_flush_ms_;
    }
    return 0;
}

int proxsdp_b200_finish(proxsdp_b200_handle_t* handle, proxsdp_result_t* result) {
    if (!handle || !handle->s || !result) { g_last_error = "null argument"; return -1; }
    return guarded([&]() {
        proxsdp_result_t& r = handle->scratch;
        Solver& s = *handle->s;
        s.finish(&r);
        double* trace_dst = result->trace; int64_t* tr_dst = result->target_rank;
        double *d0 = result->primal, *d1 = result->dual_cone, *d2 = result->dual_eq, *d3 = result->dual_in,
               *d4 = result->slack_eq, *d5 = result->slack_in;
        *result = r;
        result->primal = d0; result->dual_cone = d1; result->dual_eq = d2; result->dual_in = d3;
        result->slack_eq = d4; result->slack_in = d5; result->trace = trace_dst; result->target_rank = tr_dst;
        auto cp = [](double* dst, const std::vector<double>& src) { if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(double)); };
        cp(d0, handle->primal); cp(d1, handle->dual_cone); cp(d2, handle->dual_eq); cp(d3, handle->dual_in);
        cp(d4, handle->slack_eq); cp(d5, handle->slack_in);
        if (trace_dst && r.trace_len > 0) std::memcpy(trace_dst, handle->trace.data(), sizeof(double) * (size_t)r.trace_len * PROXSDP_TRACE_COLS);
        if (tr_dst) for (int q = 0; q < s.n_sdp; ++q) tr_dst[q] = handle->target_rank[(size_t)q];
        result->time_setup = handle->t_setup;
    });
}

int proxsdp_b200_destroy(proxsdp_b200_handle_t* handle) {
    if (!handle) return 0;
    return guarded([&]() { delete handle; });
}

int proxsdp_b200_psd_project(int64_t n_sdp, const int64_t* sides, double* x, const int64_t* target_rank,
                             const proxsdp_options_t* options, int64_t iter, int64_t mode, const double* resid,
                             int64_t* current_rank, double* min_eig, int64_t* converged, int64_t* numops,
                             int64_t repeat, double* ms_per_call) {
    if (n_sdp < 0 || !sides || !x || !target_rank || !options) { g_last_error = "null argument"; return -1; }
    return guarded([&]() {
        ConesOnlyProblem cp(n_sdp, sides, 0, nullptr, resid);
        proxsdp_options_t o = *options;
        int64_t max_tr = 2;
        for (int64_t k = 0; k < n_sdp; ++k) max_tr = std::max(max_tr, target_rank[k]);
        o.initial_target_rank = max_tr;      // sizes the Lanczos workspaces
        Solver s(&cp.pr, &o, /*cones_only=*/true);
        for (int64_t k = 0; k < n_sdp; ++k) s.target_rank[(size_t)k] = target_rank[k];
        PB_CUDA(cudaMemcpy(s.x_[0].p, x, sizeof(double) * (size_t)cp.pr.n, cudaMemcpyHostToDevice));
        cudaEvent_t e0, e1;
        PB_CUDA(cudaEventCreate(&e0)); PB_CUDA(cudaEventCreate(&e1));
        int64_t reps = std::max<int64_t>(repeat, 1);
        double total_ms = 0.0;
        for (int64_t r = 0; r < reps; ++r) {
            s.cur_ = 0;
            s.reset_scalars();
            PB_CUDA(cudaEventRecord(e0, s.stream));
            s.psd_projection_launch(iter, 0.0, mode == 1);
            PB_CUDA(cudaEventRecord(e1, s.stream));
            s.sync_scalars();
            float ms = 0.f;
            PB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            total_ms += ms;
            if (s.scal_host[S_POISON] != 0.0) {
                double ops = s.scal_host[S_NUMOPS];
                s.fallback_projection(iter);
                s.sync_scalars();
                s.scal_host[S_NUMOPS] = ops;
            }
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        PB_CUDA(cudaMemcpy(x, s.x_[1].p, sizeof(double) * (size_t)cp.pr.n, cudaMemcpyDeviceToHost));
        for (int64_t k = 0; k < n_sdp; ++k) {
            if (current_rank) current_rank[k] = (int64_t)s.scal_host[S_HEADER + 3 * k + 0];
            if (min_eig) min_eig[k] = s.scal_host[S_HEADER + 3 * k + 1];
            if (converged) converged[k] = (int64_t)s.scal_host[S_HEADER + 3 * k + 2];
        }
        if (numops) *numops = (int64_t)s.scal_host[S_NUMOPS];
        if (ms_per_call) *ms_per_call = total_ms / (double)reps;
    });
}

int proxsdp_b200_soc_project(int64_t n_soc, const int64_t* lens, double* x) {
    if (n_soc < 0 || (n_soc > 0 && (!lens || !x))) { g_last_error = "null argument"; return -1; }
    return guarded([&]() {
        ConesOnlyProblem cp(0, nullptr, n_soc, lens, nullptr);
        proxsdp_options_t o{};
        o.eigsolver_min_lanczos = 25; o.max_target_rank_krylov_eigs = 16; o.convergence_window = 200;
        o.approx_norm = 1;
        Solver s(&cp.pr, &o, /*cones_only=*/true);
        if (n_soc == 0) return;
        PB_CUDA(cudaMemcpy(s.x_[1].p, x, sizeof(double) * (size_t)cp.pr.n, cudaMemcpyHostToDevice));
        s.cur_ = 0;
        s.reset_scalars();
        s.launch_soc_only();
        s.sync_scalars();
        PB_CUDA(cudaMemcpy(x, s.x_[1].p, sizeof(double) * (size_t)cp.pr.n, cudaMemcpyDeviceToHost));
    });
}

// a problem without cones and without objective: only the rows M = [A; G] (taken as they are), b, h
static proxsdp_problem_t rows_only_problem(const proxsdp_problem_t* rows, const proxsdp_step_state_t* st, std::vector<double>& zeros,
                                           std::vector<int64_t>& colptr0) {
    proxsdp_problem_t pr{};
    pr.n = st->n; pr.p = st->p; pr.m = st->m;
    zeros.assign((size_t)std::max<int64_t>(st->n, 1), 0.0);
    if (rows) {
        pr.index_base = rows->index_base;
        pr.A_colptr = rows->A_colptr; pr.A_rowval = rows->A_rowval; pr.A_nzval = rows->A_nzval;
        pr.G_colptr = rows->G_colptr; pr.G_rowval = rows->G_rowval; pr.G_nzval = rows->G_nzval;
    } else {
        colptr0.assign((size_t)st->n + 1, 0);
        pr.index_base = 0;
        pr.A_colptr = colptr0.data(); pr.G_colptr = colptr0.data();
    }
    pr.b = st->b; pr.h = st->h; pr.c = st->c ? st->c : zeros.data();
    return pr;
}

int proxsdp_b200_dual_step(const proxsdp_problem_t* rows, const proxsdp_options_t* options, const proxsdp_step_state_t* st,
                           double* y_new, double* Mty_new, double* scalars_out, int64_t* trials) {
    if (!rows || !options || !st || !y_new || !Mty_new) { g_last_error = "null argument"; return -1; }
    return guarded([&]() {
        std::vector<double> zeros; std::vector<int64_t> cp0;
        proxsdp_problem_t pr = rows_only_problem(rows, st, zeros, cp0);
        pr.c = zeros.data();
        Solver s(&pr, options);
        const size_t Rb = sizeof(double) * (size_t)(st->p + st->m), Nb = sizeof(double) * (size_t)st->n;
        // linesearch! reads pair.y, a.Mx, a.Mx_old and a.Mty_old (the current M'y)
        s.cur_ = 0;
        if (Rb) {
            PB_CUDA(cudaMemcpy(s.y_[0].p, st->y, Rb, cudaMemcpyHostToDevice));
            PB_CUDA(cudaMemcpy(s.Mx_[1].p, st->Mx, Rb, cudaMemcpyHostToDevice));
            PB_CUDA(cudaMemcpy(s.Mx_[0].p, st->Mx_old, Rb, cudaMemcpyHostToDevice));
        }
        if (Nb) PB_CUDA(cudaMemcpy(s.Mty_[0].p, st->Mty, Nb, cudaMemcpyHostToDevice));
        double out4[4];
        long long ev = s.seam_dual_step(st->primal_step, st->primal_step_old, st->theta, st->beta, st->dual_step, out4);
        if (Rb) PB_CUDA(cudaMemcpy(y_new, s.y_[1].p, Rb, cudaMemcpyDeviceToHost));
        if (Nb) PB_CUDA(cudaMemcpy(Mty_new, s.Mty_[1].p, Nb, cudaMemcpyDeviceToHost));
        if (scalars_out) for (int i = 0; i < 4; ++i) scalars_out[i] = out4[i];
        if (trials) *trials = ev;
    });
}

int proxsdp_b200_residuals(const proxsdp_options_t* options, const proxsdp_step_state_t* st, double* out) {
    if (!options || !st || !out) { g_last_error = "null argument"; return -1; }
    return guarded([&]() {
        std::vector<double> zeros; std::vector<int64_t> cp0;
        proxsdp_problem_t pr = rows_only_problem(nullptr, st, zeros, cp0);
        pr.c = zeros.data();                 // the working objective is uploaded below exactly as given
        Solver s(&pr, options);
        const size_t Rb = sizeof(double) * (size_t)(st->p + st->m), Nb = sizeof(double) * (size_t)st->n;
        s.cur_ = 0;     // new iterates in slot 1, old ones in slot 0 (the layout at the end of an iteration)
        if (Nb) {
            PB_CUDA(cudaMemcpy(s.x_[1].p, st->x, Nb, cudaMemcpyHostToDevice));
            PB_CUDA(cudaMemcpy(s.x_[0].p, st->x_old, Nb, cudaMemcpyHostToDevice));
            PB_CUDA(cudaMemcpy(s.Mty_[1].p, st->Mty, Nb, cudaMemcpyHostToDevice));
            PB_CUDA(cudaMemcpy(s.Mty_[0].p, st->Mty_old, Nb, cudaMemcpyHostToDevice));
            PB_CUDA(cudaMemcpy(s.c_.p, st->c, Nb, cudaMemcpyHostToDevice));
        }
        if (Rb) {
            PB_CUDA(cudaMemcpy(s.y_[1].p, st->y, Rb, cudaMemcpyHostToDevice));
            PB_CUDA(cudaMemcpy(s.y_[0].p, st->y_old, Rb, cudaMemcpyHostToDevice));
            PB_CUDA(cudaMemcpy(s.Mx_[1].p, st->Mx, Rb, cudaMemcpyHostToDevice));
            PB_CUDA(cudaMemcpy(s.Mx_[0].p, st->Mx_old, Rb, cudaMemcpyHostToDevice));
        }
        s.seam_residuals(st->primal_step, st->dual_step, st->beta, st->norm_b, st->norm_h, st->norm_c, out);
    });
}

int proxsdp_b200_lanczos(int64_t n, const double* A, const double* x0, int64_t howmany, int64_t krylovdim,
                         int64_t maxiter, double tol, double* vals, double* vecs, int64_t* nvals,
                         int64_t* converged, int64_t* numops, int64_t* numiter, int64_t repeat, double* ms_per_call) {
    if (n < 1 || !A || !x0 || !vals || howmany < 1 || krylovdim < howmany) { g_last_error = "invalid argument"; return -1; }
    return guarded([&]() {
        int64_t side = n;
        ConesOnlyProblem cp(1, &side, 0, nullptr, x0);
        proxsdp_options_t o{};
        o.eigsolver_min_lanczos = krylovdim; o.max_target_rank_krylov_eigs = howmany; o.convergence_window = 200;
        o.approx_norm = 1; o.krylovkit_resid_init = 3;
        Solver s(&cp.pr, &o, /*cones_only=*/true, /*force_large=*/true);
        ConeDev& cd = s.cones[0];
        PB_CUDA(cudaMemcpy2D(cd.X.p, sizeof(double) * (size_t)cd.ld, A, sizeof(double) * (size_t)n,
                             sizeof(double) * (size_t)n, (size_t)n, cudaMemcpyHostToDevice));
        cudaEvent_t e0, e1;
        PB_CUDA(cudaEventCreate(&e0)); PB_CUDA(cudaEventCreate(&e1));
        int64_t reps = std::max<int64_t>(repeat, 1);
        double total_ms = 0.0;
        for (int64_t r = 0; r < reps; ++r) {
            // An untimed launch runs first when a time is wanted: the events and the timed launch are enqueued while it
            // executes, so the elapsed time is the kernel's, not the host's launch latency on an idle stream.
            if (ms_per_call && repeat > 1) s.lanczos_launch(cd, 0, (int)howmany, (int)krylovdim, (int)maxiter, tol);
            s.reset_scalars();
            PB_CUDA(cudaEventRecord(e0, s.stream));
            s.lanczos_launch(cd, 0, (int)howmany, (int)krylovdim, (int)maxiter, tol);
            PB_CUDA(cudaEventRecord(e1, s.stream));
            s.sync_scalars();
            float ms = 0.f;
            PB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            total_ms += ms;
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        std::vector<int> info = cd.info.download();
        std::vector<double> v = cd.vals.download();
        int nv = info[0];
        for (int i = 0; i < nv; ++i) vals[i] = v[(size_t)i];
        if (vecs)
            PB_CUDA(cudaMemcpy2D(vecs, sizeof(double) * (size_t)n, cd.Y.p, sizeof(double) * (size_t)cd.ld,
                                 sizeof(double) * (size_t)n, (size_t)nv, cudaMemcpyDeviceToHost));
        if (nvals) *nvals = nv;
        if (converged) *converged = info[1];
        if (numops) *numops = info[2];
        if (numiter) *numiter = info[3];
        if (ms_per_call) *ms_per_call = total_ms / (double)reps;
    });
}

int proxsdp_b200_eigh(int64_t n, const double* A, double* w, double* Z) {
    if (n < 1 || !A || !w) { g_last_error = "invalid argument"; return -1; }
    return guarded([&]() {
        int64_t side = n;
        ConesOnlyProblem cp(1, &side, 0, nullptr, nullptr);
        proxsdp_options_t o{};
        o.eigsolver_min_lanczos = 25; o.max_target_rank_krylov_eigs = 16; o.convergence_window = 200;
        o.approx_norm = 1; o.krylovkit_resid_init = 1;
        Solver s(&cp.pr, &o, /*cones_only=*/true, /*force_large=*/true);
        ConeDev& cd = s.cones[0];
        PB_CUDA(cudaMemcpy2D(cd.X.p, sizeof(double) * (size_t)cd.ld, A, sizeof(double) * (size_t)n,
                             sizeof(double) * (size_t)n, (size_t)n, cudaMemcpyHostToDevice));
        std::vector<double> ev = s.full_eig_device(cd);
        std::vector<int> order((size_t)n);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ev[(size_t)a] < ev[(size_t)b]; });
        for (int64_t i = 0; i < n; ++i) w[i] = ev[(size_t)order[(size_t)i]];
        if (Z) {
            std::vector<double> V((size_t)n * (size_t)n);
            PB_CUDA(cudaMemcpy2D(V.data(), sizeof(double) * (size_t)n, cd.Vfull.p, sizeof(double) * (size_t)cd.ld,
                                 sizeof(double) * (size_t)n, (size_t)n, cudaMemcpyDeviceToHost));
            for (int64_t j = 0; j < n; ++j)
                std::memcpy(Z + j * n, V.data() + (size_t)order[(size_t)j] * (size_t)n, sizeof(double) * (size_t)n);
        }
    });
}

// page-locked host memory from the library's cache: result / input vectors that live in it are copied at full PCIe
// rate without staging (and without first-touch page faults on a fresh 16 MB result vector)
void* proxsdp_b200_host_alloc(int64_t bytes) {
    if (bytes < 0) { g_last_error = "invalid argument"; return nullptr; }
    void* p = nullptr;
    int rc = guarded([&]() {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw CudaError(-5, "no CUDA device available");
        p = pb_host_alloc((size_t)std::max<int64_t>(bytes, 1));
    });
    return rc == 0 ? p : nullptr;
}
int proxsdp_b200_host_free(void* ptr) {
    return guarded([&]() { pb_host_free(ptr); });
}
int proxsdp_b200_trim_caches(void) {
    return guarded([&]() { cudaDeviceSynchronize(); pb_cache_trim(); });
}

int proxsdp_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
const char* proxsdp_b200_last_error(void) { return g_last_error.c_str(); }
const char* proxsdp_b200_version(void) { return "proxsdp_b200 0.1.0 (sm_100a)"; }
int64_t proxsdp_b200_sizeof_problem(void) { return (int64_t)sizeof(proxsdp_problem_t); }
int64_t proxsdp_b200_sizeof_options(void) { return (int64_t)sizeof(proxsdp_options_t); }
int64_t proxsdp_b200_sizeof_result(void) { return (int64_t)sizeof(proxsdp_result_t); }
int64_t proxsdp_b200_sizeof_shard(void) { return (int64_t)sizeof(proxsdp_shard_t); }

}  // extern "C"
