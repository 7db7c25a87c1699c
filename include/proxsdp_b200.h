/*
 * proxsdp_b200.h — C ABI of the B200-native ProxSDP hot path (libproxsdp_b200.so).
 *
 * The reference (mariohsouto/ProxSDP.jl) exposes no FFI; the drop-in boundary is the
 * plain Julia call `chambolle_pock(aff, con, opt)::Result` (reference src/pdhg.jl:1-5,
 * called only from src/MOI_wrapper.jl:310).  A Julia maintainer binds these entry
 * points with `ccall`; the Python host mirror binds them with
 * ctypes (proxsdp_b200/solver.py).  Plain pointers and sizes only.
 *
 * Return value of every function: 0 on success, negative on failure
 *   -1  invalid argument            -2  unsupported on this path (equilibration when sharded)
 *   -3  malformed cones             -4  out of device/host memory
 *   -5  no CUDA device / extension  <= -100  CUDA runtime error (-100 - cudaError_t)
 * The solver's own outcome (optimal, limits, infeasible, ...) is NOT an error: it is
 * reported in proxsdp_result_t.status like the reference does (src/structs.jl:60-81,
 * src/MOI_wrapper.jl:381-399).  proxsdp_b200_last_error() gives the message of the last
 * failure on the calling thread.
 *
 * Threading: like the reference (global TimerOutputs / logger, src/MOI_wrapper.jl:304-315)
 * a solve is synchronous and one solve runs per process at a time.  Host arrays are
 * borrowed for the duration of the call only; the library never retains them and never
 * calls back into the host language.
 */
#ifndef PROXSDP_B200_H
#define PROXSDP_B200_H

#include "proxsdp_b200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Whole `chambolle_pock` (reference src/pdhg.jl:1-530): setup (preprocess!, norm_scaling,
 * step size), the PDHG loop with every per-iteration array operation on the GPU, the
 * reference's control logic on the host, and result assembly (cache_solution). */
int proxsdp_b200_solve(const proxsdp_problem_t* problem, const proxsdp_options_t* options,
                       proxsdp_result_t* result);

/* The same solve in three calls, for callers that want to own the loop (the reference's
 * `for k in 1:2*opt.max_iter_local` body, src/pdhg.jl:145-484, is one `iterate` step):
 *   create   — everything before the loop (src/pdhg.jl:7-142): copies the problem to the device;
 *   iterate  — up to max_steps (< 0: until termination) further PDHG iterations; *finished = 1 once the
 *              reference's loop would have exited; *device_ms = CUDA-event time of this call on the
 *              solver's stream; flush_l2 != 0 writes a buffer larger than L2 before every iteration
 *              (measurement hygiene only);
 *   finish   — result assembly (`cache_solution`, src/pdhg.jl:745-787) into caller buffers;
 *   destroy  — releases all device memory.
 * create + iterate(-1) + finish + destroy is exactly proxsdp_b200_solve. */
typedef struct proxsdp_b200_handle proxsdp_b200_handle_t;
int proxsdp_b200_create(const proxsdp_problem_t* problem, const proxsdp_options_t* options,
                        proxsdp_b200_handle_t** handle);
int proxsdp_b200_iterate(proxsdp_b200_handle_t* handle, int64_t max_steps, int64_t flush_l2,
                         int64_t* steps_done, int64_t* finished, double* device_ms);
/* Running totals since create.  counts[8]: iterations, kernel launches, Lanczos mat-vecs, Lanczos calls,
 * Lanczos calls covered by the kernel timer, full eigendecompositions, linesearch trials, sum(target_rank).
 * times_ms[4] (CUDA events on the solver's stream): psd_projection!, Lanczos kernel, rest of the
 * iteration, L2 flush.  Either pointer may be NULL. */
int proxsdp_b200_counters(proxsdp_b200_handle_t* handle, int64_t* counts, double* times_ms);
int proxsdp_b200_finish(proxsdp_b200_handle_t* handle, proxsdp_result_t* result);
int proxsdp_b200_destroy(proxsdp_b200_handle_t* handle);

/* One `psd_projection!` (reference src/prox_operators.jl:33-66) over concatenated svec
 * blocks.  x: in/out (sum of tri_len doubles, working-space scaling: off-diagonals carry
 * sqrt(2)).  mode 0 = the reference's dispatch (Krylov when eligible), 1 = force the full
 * eigendecomposition.  resid: optional concatenated Lanczos start vectors (NULL = library
 * default).  Outputs per cone: current_rank, min_eig, converged (KrylovKit info.converged;
 * -1 when the full path ran); numops = total Lanczos mat-vecs.  The projection is executed
 * `repeat` (>= 1) times on device-resident data and *ms_per_call (may be NULL) receives the
 * mean CUDA-event time of one call — the "ms per eig-projection" metric. */
int proxsdp_b200_psd_project(int64_t n_sdp, const int64_t* sides, double* x,
                             const int64_t* target_rank, const proxsdp_options_t* options,
                             int64_t iter, int64_t mode, const double* resid,
                             int64_t* current_rank, double* min_eig, int64_t* converged,
                             int64_t* numops, int64_t repeat, double* ms_per_call);

/* `soc_projection!` (reference src/prox_operators.jl:138-158) over concatenated SOC blocks
 * (lens[k] entries each, the first being t).  x: in/out. */
int proxsdp_b200_soc_project(int64_t n_soc, const int64_t* lens, double* x);

/* `linesearch!` (options->line_search_flag != 0) or `dual_step!` (reference src/pdhg.jl:532-609) on explicit state.
 * rows: only n, p, m, index_base, A, G of the problem are read — M = [A; G] is taken AS IS (working representation, no
 * cone scaling), b and h come from `st`.  In: st->y, Mx, Mx_old, Mty (the current M'y, called Mty_old inside
 * linesearch!), primal_step, primal_step_old, theta, beta, dual_step.  Out: y_new (p + m), Mty_new (n) and
 * scalars_out[4] = {primal_step, theta, dual_step, primal_step_old} as the function leaves them; *trials = line-search
 * trials evaluated. */
int proxsdp_b200_dual_step(const proxsdp_problem_t* rows, const proxsdp_options_t* options,
                           const proxsdp_step_state_t* st, double* y_new, double* Mty_new,
                           double* scalars_out, int64_t* trials);

/* `compute_residual!` + `compute_gap!` (reference src/residuals.jl:2-71) on explicit state.  out[8] =
 * {primal_residual, dual_residual, comb_residual, equa_feasibility, ineq_feasibility, prim_obj, dual_obj, dual_gap}. */
int proxsdp_b200_residuals(const proxsdp_options_t* options, const proxsdp_step_state_t* st, double* out);

/* Kernel-level seam of the eigen back-end: KrylovKit.eigsolve(A, x0, howmany, :LR,
 * Lanczos(orth, krylovdim, maxiter, tol)) as used at reference src/eigsolver.jl:802-812.
 * A: n x n column-major, both triangles valid.  vals: krylovdim doubles; vecs: n x krylovdim
 * column-major.  `repeat` as above; *ms_per_call = mean device time of one eigsolve. */
int proxsdp_b200_lanczos(int64_t n, const double* A, const double* x0, int64_t howmany,
                         int64_t krylovdim, int64_t maxiter, double tol,
                         double* vals, double* vecs, int64_t* nvals, int64_t* converged,
                         int64_t* numops, int64_t* numiter, int64_t repeat, double* ms_per_call);

/* Full symmetric eigendecomposition on the device (replaces LinearAlgebra.eigen!, reference
 * src/prox_operators.jl:113, src/pdhg.jl:685).  A: n x n column-major (both triangles).
 * w: n eigenvalues ascending; Z: n x n eigenvectors (may be NULL). */
int proxsdp_b200_eigh(int64_t n, const double* A, double* w, double* Z);

/* Multi-GPU: one process per GPU.  The communicator wraps NCCL: rank 0 obtains a
 * 128-byte id, the host language broadcasts it (torch.distributed / MPI.jl), every rank calls comm_create.
 * proxsdp_b200_solve_sharded is proxsdp_b200_solve on this rank's blocks with all whole-problem scalars
 * combined across ranks: one all-gather of the 19-double iteration record per iteration plus one of the two
 * line-search norms per trial; no vector ever crosses NVLink.  Results are this rank's pieces. */
typedef struct proxsdp_b200_comm proxsdp_b200_comm_t;
int proxsdp_b200_comm_unique_id(char id[128]);
int proxsdp_b200_comm_create(const char id[128], int64_t rank, int64_t nranks, int64_t device_id,
                             proxsdp_b200_comm_t** comm);
int proxsdp_b200_comm_destroy(proxsdp_b200_comm_t* comm);
int proxsdp_b200_solve_sharded(const proxsdp_problem_t* local_problem, const proxsdp_options_t* options,
                               const proxsdp_shard_t* shard, proxsdp_result_t* local_result);

/* Page-locked host memory from the library's per-process cache (cudaMallocHost blocks that are kept and re-used).
 * Optional: any host pointer is accepted by every entry point; vectors that live in this memory (a shim would put
 * `Result.primal` / `Result.dual_cone` and `AffineSets.c` there) cross PCIe without a staging copy.  The library
 * itself also draws its device buffers from a per-process cache so that back-to-back solves do not pay
 * cudaMalloc / cudaFree; proxsdp_b200_trim_caches() returns everything that is cached and unused to the driver. */
void* proxsdp_b200_host_alloc(int64_t bytes);
int proxsdp_b200_host_free(void* ptr);
int proxsdp_b200_trim_caches(void);

/* Library / device information. */
int proxsdp_b200_device_count(void);
const char* proxsdp_b200_last_error(void);
const char* proxsdp_b200_version(void);
/* sizeof() of the three PODs, so a binding can verify its struct layout at load time. */
int64_t proxsdp_b200_sizeof_problem(void);
int64_t proxsdp_b200_sizeof_options(void);
int64_t proxsdp_b200_sizeof_result(void);
int64_t proxsdp_b200_sizeof_shard(void);

#ifdef __cplusplus
}
#endif
#endif /* PROXSDP_B200_H */
