/*
 * proxsdp_b200_types.h — plain-old-data layouts of the drop-in boundary.
 *
 * The reference has no FFI: the seam is the Julia call
 *     chambolle_pock(affine_sets::AffineSets, conic_sets::ConicSets, opt::Options)::Result
 * (reference src/pdhg.jl:1-5, only caller src/MOI_wrapper.jl:310).  These structs
 * are the flat C images of the Julia structs that cross that seam:
 *     AffineSets  src/structs.jl:32-42     -> proxsdp_problem_t (n,p,m,A,G,b,h,c)
 *     ConicSets   src/structs.jl:44-58     -> proxsdp_problem_t (sdp_ and soc_ tables)
 *     Options     src/options.jl:1-132     -> proxsdp_options_t (same names, same order)
 *     Result      src/structs.jl:60-81     -> proxsdp_result_t
 * Everything is 8-byte wide (int64_t / double / pointer) so that a Julia
 * `struct` with Int64/Float64/Ptr fields, a ctypes.Structure and this header
 * agree without padding rules.  Julia Bool fields travel as int64_t 0/1.
 *
 * No torch / CUDA types appear here.
 */
#ifndef PROXSDP_B200_TYPES_H
#define PROXSDP_B200_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- problem: AffineSets + ConicSets (structs.jl:32-58) ------------------
 * min c'x  s.t.  A x = b (p rows), G x <= h (m rows),
 *                x[sdp_idx[sdp_ptr[k]..sdp_ptr[k+1])] in PSD-triangle(sdp_side[k]),
 *                x[soc_idx[soc_ptr[k]..soc_ptr[k+1])] in SOC, the rest free.
 * A, G are SparseMatrixCSC{Float64,Int64}: colptr (n+1), rowval (nnz), nzval (nnz);
 * index_base = 1 when the arrays come straight from Julia (1-based), 0 otherwise.
 * Variable index lists (sdp_idx, soc_idx) use the same base.  All arrays are
 * borrowed for the duration of the call and never written (the reference mutates
 * its AffineSets in place, pdhg.jl:58,95,647-663; the library works on copies).
 */
typedef struct proxsdp_problem {
    int64_t n;            /* number of primal variables            */
    int64_t p;            /* equality rows                         */
    int64_t m;            /* inequality rows                       */
    int64_t index_base;   /* 0 or 1                                */
    const int64_t* A_colptr; const int64_t* A_rowval; const double* A_nzval;
    const int64_t* G_colptr; const int64_t* G_rowval; const double* G_nzval;
    const double* b;      /* (p) */
    const double* h;      /* (m) */
    const double* c;      /* (n) */
    int64_t n_sdp;
    const int64_t* sdp_side;  /* (n_sdp)   SDPSet.sq_side                     */
    const int64_t* sdp_ptr;   /* (n_sdp+1) offsets into sdp_idx (0-based)     */
    const int64_t* sdp_idx;   /* SDPSet.vec_i concatenated, tri_len each      */
    int64_t n_soc;
    const int64_t* soc_ptr;   /* (n_soc+1) offsets into soc_idx (0-based)     */
    const int64_t* soc_idx;   /* SOCSet.idx concatenated                      */
    /* Optional Lanczos start vectors (EigSolverAlloc.resid, eigsolver.jl:392-411),
     * concatenated per PSD cone (sum of sdp_side doubles).  NULL => the library
     * draws the documented splitmix64/Box-Muller substitute for Julia's
     * MersenneTwister(eigsolver_resid_seed) stream (not reproducible outside Julia). */
    const double* eig_resid;
} proxsdp_problem_t;

/* ---- Options (options.jl:1-132), same field names and order ------------- */
typedef struct proxsdp_options {
    /* printing */
    int64_t log_verbose;
    int64_t log_freq;
    int64_t timer_verbose;
    int64_t timer_file;
    int64_t disable_julia_logger;
    /* time */
    double  time_limit;
    int64_t warn_on_limit;
    int64_t extended_log;
    int64_t extended_log2;
    int64_t log_repeat_header;
    /* tolerances */
    double  tol_gap;
    double  tol_feasibility;
    double  tol_feasibility_dual;
    double  tol_primal;
    double  tol_dual;
    double  tol_psd;
    double  tol_soc;
    int64_t check_dual_feas;
    int64_t check_dual_feas_freq;
    double  max_obj;
    int64_t min_iter_max_obj;
    /* infeasibility check */
    int64_t min_iter_time_infeas;
    double  infeas_gap_tol;
    double  infeas_limit_gap_tol;
    double  infeas_stable_gap_tol;
    double  infeas_feasibility_tol;
    double  infeas_stable_feasibility_tol;
    int64_t certificate_search;
    double  certificate_obj_tol;
    double  certificate_fail_tol;
    /* beta bounds (dead in the reference) */
    double  min_beta;
    double  max_beta;
    double  initial_beta;
    /* adaptive steps */
    double  initial_adapt_level;
    double  adapt_decay;
    int64_t adapt_window;
    /* PDHG */
    int64_t convergence_window;
    int64_t convergence_check;
    int64_t max_iter;
    int64_t min_iter;
    int64_t divergence_min_update;
    int64_t max_iter_lp;
    int64_t max_iter_conic;
    int64_t max_iter_local;
    int64_t advanced_initialization;
    /* linesearch */
    int64_t line_search_flag;
    int64_t max_linsearch_steps;
    double  delta;
    double  initial_theta;
    double  linsearch_decay;
    /* spectral decomposition */
    int64_t full_eig_decomp;
    int64_t max_target_rank_krylov_eigs;
    int64_t min_size_krylov_eigs;
    int64_t warm_start_eig;
    int64_t rank_increment;
    int64_t rank_increment_factor;
    /* eigsolver selection */
    int64_t eigsolver;
    int64_t eigsolver_min_lanczos;
    int64_t eigsolver_resid_seed;
    /* Arpack */
    double  arpack_tol;
    int64_t arpack_resid_init;
    int64_t arpack_reset_resid;
    int64_t arpack_max_iter;
    /* KrylovKit */
    int64_t krylovkit_reset_resid;
    int64_t krylovkit_resid_init;
    double  krylovkit_tol;
    int64_t krylovkit_max_iter;
    int64_t krylovkit_eager;
    int64_t krylovkit_verbose;
    /* rank heuristics */
    int64_t reduce_rank;
    int64_t rank_slack;
    int64_t full_eig_freq;
    int64_t full_eig_len;
    /* equilibration */
    int64_t equilibration;
    int64_t equilibration_iters;
    double  equilibration_lb;
    double  equilibration_ub;
    double  equilibration_limit;
    int64_t equilibration_force;
    /* norm */
    int64_t approx_norm;
    /* ---- extensions (not in options.jl; zero = reference behaviour) ---- */
    int64_t initial_target_rank;   /* 0 => 2 (pdhg.jl:19); used by the rank-sweep config */
    int64_t freeze_target_rank;    /* 1 => never bump target_rank (rank-sweep measurement) */
    int64_t device_id;             /* CUDA device ordinal (product only)                   */
    int64_t trace_cap;             /* record up to this many iterations into result trace  */
    int64_t implicit_psd_operator; /* 1 => Krylov projections of large cones apply the matrix IMPLICITLY as
                                      Y diag(lam) Y' - tau mat(M'y + c) (low rank + sparse) whenever the
                                      previous projection left a low-rank iterate: no dense n x n matrix is formed or
                                      read, the eigsolve runs inside one thread-block cluster (product only)            */
} proxsdp_options_t;

/* ---- Result (structs.jl:60-81) ------------------------------------------
 * Vector outputs are caller-allocated with the sizes shown; any may be NULL. */
#define PROXSDP_STATUS_STRING_LEN 256
#define PROXSDP_TRACE_COLS 14
typedef struct proxsdp_result {
    int64_t status;       /* 0 not called, 1 optimal, 2 time limit, 3 iteration limit,
                             5 dual infeasible/unbounded, 6 infeasible (MOI_wrapper.jl:381-399) */
    char    status_string[PROXSDP_STATUS_STRING_LEN];
    double* primal;       /* (n) user order, unscaled */
    double* dual_cone;    /* (n) */
    double* dual_eq;      /* (p) */
    double* dual_in;      /* (m) */
    double* slack_eq;     /* (p) */
    double* slack_in;     /* (m) */
    double  primal_residual;   /* = equality feasibility  (pdhg.jl:774) */
    double  dual_residual;     /* = inequality feasibility (pdhg.jl:775) */
    double  objval;
    double  dual_objval;
    double  gap;
    double  time;
    int64_t iter;
    int64_t final_rank;
    int64_t primal_feasible_user_tol;
    int64_t dual_feasible_user_tol;
    int64_t certificate_found;
    int64_t result_count;
    /* ---- measurement extras (not part of the reference Result) ---- */
    double  final_primal_res;  /* residuals.primal_residual[iter] (fixed-point residual) */
    double  final_dual_res;    /* residuals.dual_residual[iter]                           */
    double  time_setup;        /* seconds before the first iteration                      */
    double  time_loop;         /* seconds inside the CP loop                              */
    double  time_psd_proj;     /* seconds inside psd_projection! ("sdp proj" section)    */
    int64_t n_psd_proj;        /* number of psd_projection! calls                         */
    int64_t lanczos_matvecs;   /* sum of KrylovKit numops over all calls                  */
    int64_t lanczos_calls;
    int64_t full_eig_calls;
    int64_t linesearch_trials;
    int64_t gpu_launches;      /* kernels launched by the product path (0 for the oracle) */
    int64_t* target_rank;      /* (n_sdp) final target rank per cone, may be NULL         */
    /* optional per-iteration trace, row-major (trace_cap x PROXSDP_TRACE_COLS):
       iter, prim_obj, dual_obj, gap, feasibility, primal_res, dual_res,
       primal_step, beta, sum(target_rank), sum(current_rank), min(min_eig),
       lanczos mat-vecs this iteration, linesearch trials this iteration */
    double* trace;
    int64_t trace_len;
    /* per-section device timers (CUDA events on the solver's stream; the analogue of the
       reference's TimerOutputs sections, pdhg.jl:626 "sdp proj" > "eigs" etc.) */
    double  time_lanczos;        /* seconds inside the Lanczos kernel (first eigsolve of each iteration) */
    double  time_rest;           /* seconds from the end of psd_projection! to the end of the iteration's kernels */
    double  time_l2_flush;       /* seconds spent in the optional between-iteration L2 flush (bench only) */
    int64_t lanczos_timed_calls; /* launches covered by time_lanczos                                      */
    int64_t h2d_bytes;           /* host->device bytes copied during this solve                           */
    int64_t d2h_bytes;           /* device->host bytes copied during this solve                           */
    int64_t implicit_calls;      /* eigsolves that ran on the implicit low-rank + sparse operator         */
} proxsdp_result_t;

/* ---- step-level seams: the slice of PrimalDual / AuxiliaryData / Params (structs.jl:83-192) that `linesearch!` /
 * `dual_step!` (pdhg.jl:532-609) and `compute_residual!` / `compute_gap!` (residuals.jl:2-71) read.  All vectors are in
 * the solver's WORKING representation (permuted, off-diagonal svec entries scaled): these seams exist for kernel-level
 * parity tests, the same way the reference's functions operate on its internal state. */
typedef struct proxsdp_step_state {
    int64_t n, p, m;                       /* variables, equality rows, inequality rows */
    const double* b; const double* h; const double* c;
    const double* x; const double* x_old;              /* (n)      */
    const double* y; const double* y_old;              /* (p + m)  */
    const double* Mx; const double* Mx_old;            /* (p + m)  */
    const double* Mty; const double* Mty_old;          /* (n)      */
    double primal_step, primal_step_old, dual_step, theta, beta;
    double norm_b, norm_h, norm_c;
} proxsdp_step_state_t;

/* ---- sharded solves (independent blocks of a stacked problem, one block set per GPU) --------------
 * The reference is single-process; this is the B200 extension of SURVEY.md section 8(e).  Each rank passes
 * the sub-problem made of its blocks (variables, cones and rows that are coupled only among themselves:
 * proxsdp_b200/sharding.py builds them) plus this descriptor.  Everything that the reference computes over
 * the WHOLE problem — ||b||, ||c||, ||h||, ||M||_F, sqrt(N), sqrt(R), the line-search norms, the residual /
 * feasibility maxima, the objective dot products, the convergence flags — is combined across ranks, so all
 * ranks walk through exactly the iterations of the un-sharded solve and stop together. */
typedef void (*proxsdp_reduce_fn)(double* vals, int64_t count, int64_t op /* 0 = sum, 1 = max */, void* ctx);
typedef struct proxsdp_shard {
    int64_t rank, nranks;
    int64_t global_n, global_p, global_m;   /* sizes of the whole problem */
    void*   comm;                           /* proxsdp_b200_comm_t* (product); NULL for the oracle       */
    proxsdp_reduce_fn reduce;               /* oracle only (test infrastructure): host all-reduce callback */
    void*   reduce_ctx;
} proxsdp_shard_t;

#ifdef __cplusplus
}
#endif
#endif /* PROXSDP_B200_TYPES_H */
