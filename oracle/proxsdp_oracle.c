/*
 * proxsdp_oracle.c — CPU restatement of ProxSDP's chambolle_pock.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path;
 * it is the checker the CUDA path is compared against (tests/, smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py).  The product library
 * (proxsdp_b200/csrc) never links or calls it.
 *
 * The real reference (Julia) cannot run in the build container (no julia binary,
 * no depot, no network) and its Krylov arithmetic lives in KrylovKit.jl, which is
 * not under /root/reference.  PARITY STATUS:
 *   - exact-projection path (full eigendecomposition): pinned against the
 *     reference's own known-answer tests (test/moi_proxsdp_unit.jl:43-47,89-93,
 *     132-136,173-177,209-217,255-265,329,336; test/moi_mimo.jl:71-75;
 *     test/moi_sdplib.jl:53-56) and a numpy/LAPACK mirror (oracle/oracle_np.py).
 *   - Krylov path: PARITY UNPINNED (the reference asserts nothing about it beyond
 *     "X is PSD to 1e-4" on two n=124 problems); checked against the exact path
 *     and SDPLIB optimal values instead.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/proxsdp_b200_types.h"

/* from oracle_eig.c */
extern int oracle_lanczos_eager;
int oracle_eigh(int64_t n, double* A, double* w, double* Z);
int oracle_lanczos(int64_t n, const double* A, const double* x0, int64_t howmany,
                   int64_t krylovdim, int64_t maxiter, double tol,
                   double* vals, double* vecs, int64_t* nvals_out,
                   int64_t* converged_out, int64_t* numops_out, int64_t* numiter_out);

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------- */
/* sparse CSC                                                                */
/* ------------------------------------------------------------------------- */
typedef struct {
    int64_t nr, nc, nnz;
    int64_t* colptr;   /* nc+1, 0-based */
    int64_t* rowidx;   /* nnz,  0-based */
    double* val;
} csc_t;

static csc_t csc_alloc(int64_t nr, int64_t nc, int64_t nnz) {
    csc_t a;
    a.nr = nr; a.nc = nc; a.nnz = nnz;
    a.colptr = (int64_t*)calloc((size_t)nc + 1, sizeof(int64_t));
    a.rowidx = (int64_t*)calloc((size_t)(nnz > 0 ? nnz : 1), sizeof(int64_t));
    a.val = (double*)calloc((size_t)(nnz > 0 ? nnz : 1), sizeof(double));
    return a;
}
static void csc_free(csc_t* a) { free(a->colptr); free(a->rowidx); free(a->val); }

static csc_t csc_from_input(int64_t nr, int64_t nc, const int64_t* colptr, const int64_t* rowval,
                            const double* nzval, int64_t base) {
    int64_t nnz = (nr > 0 && colptr) ? (colptr[nc] - base) : 0;
    csc_t a = csc_alloc(nr, nc, nnz);
    if (nnz == 0) return a;
    for (int64_t j = 0; j <= nc; ++j) a.colptr[j] = colptr[j] - base;
    for (int64_t k = 0; k < nnz; ++k) { a.rowidx[k] = rowval[k] - base; a.val[k] = nzval[k]; }
    return a;
}
static csc_t csc_copy(const csc_t* s) {
    csc_t a = csc_alloc(s->nr, s->nc, s->nnz);
    memcpy(a.colptr, s->colptr, sizeof(int64_t) * ((size_t)s->nc + 1));
    if (s->nnz) {
        memcpy(a.rowidx, s->rowidx, sizeof(int64_t) * (size_t)s->nnz);
        memcpy(a.val, s->val, sizeof(double) * (size_t)s->nnz);
    }
    return a;
}
/* A[:, ord]  (scaling.jl:24) */
static csc_t csc_permute_cols(const csc_t* s, const int64_t* ord) {
    csc_t a = csc_alloc(s->nr, s->nc, s->nnz);
    int64_t pos = 0;
    for (int64_t j = 0; j < s->nc; ++j) {
        int64_t src = ord[j];
        a.colptr[j] = pos;
        for (int64_t k = s->colptr[src]; k < s->colptr[src + 1]; ++k) {
            a.rowidx[pos] = s->rowidx[k];
            a.val[pos] = s->val[k];
            pos++;
        }
    }
    a.colptr[s->nc] = pos;
    return a;
}
/* vcat(A, G)  (pdhg.jl:104) */
static csc_t csc_vstack(const csc_t* A, const csc_t* G) {
    csc_t m = csc_alloc(A->nr + G->nr, A->nc, A->nnz + G->nnz);
    int64_t pos = 0;
    for (int64_t j = 0; j < A->nc; ++j) {
        m.colptr[j] = pos;
        for (int64_t k = A->colptr[j]; k < A->colptr[j + 1]; ++k) {
            m.rowidx[pos] = A->rowidx[k]; m.val[pos] = A->val[k]; pos++;
        }
        for (int64_t k = G->colptr[j]; k < G->colptr[j + 1]; ++k) {
            m.rowidx[pos] = G->rowidx[k] + A->nr; m.val[pos] = G->val[k]; pos++;
        }
    }
    m.colptr[A->nc] = pos;
    return m;
}
/* materialised transpose (pdhg.jl:105,128: Matrices.Mt is a SparseMatrixCSC) */
static csc_t csc_transpose(const csc_t* s) {
    csc_t t = csc_alloc(s->nc, s->nr, s->nnz);
    for (int64_t k = 0; k < s->nnz; ++k) t.colptr[s->rowidx[k] + 1]++;
    for (int64_t j = 0; j < s->nr; ++j) t.colptr[j + 1] += t.colptr[j];
    int64_t* next = (int64_t*)malloc(sizeof(int64_t) * ((size_t)s->nr + 1));
    memcpy(next, t.colptr, sizeof(int64_t) * ((size_t)s->nr + 1));
    for (int64_t j = 0; j < s->nc; ++j)
        for (int64_t k = s->colptr[j]; k < s->colptr[j + 1]; ++k) {
            int64_t q = next[s->rowidx[k]]++;
            t.rowidx[q] = j;
            t.val[q] = s->val[k];
        }
    free(next);
    return t;
}
/* y = A x  — SparseArrays mul!(y, A::SparseMatrixCSC, x): column scatter
 * (pdhg.jl:140-141,556,603,634) */
static void csc_mul(const csc_t* a, const double* x, double* y) {
    #pragma omp parallel for schedule(static) if (a->nr > 16384)
    for (int64_t i = 0; i < a->nr; ++i) y[i] = 0.0;
    for (int64_t j = 0; j < a->nc; ++j) {
        double xj = x[j];
        for (int64_t k = a->colptr[j]; k < a->colptr[j + 1]; ++k) y[a->rowidx[k]] += a->val[k] * xj;
    }
}
/* y += alpha * A' x (gather), used only for result assembly (pdhg.jl:706) */
static void csc_mul_t_add(const csc_t* a, const double* x, double* y) {
    for (int64_t j = 0; j < a->nc; ++j) {
        double s = 0.0;
        for (int64_t k = a->colptr[j]; k < a->colptr[j + 1]; ++k) s += a->val[k] * x[a->rowidx[k]];
        y[j] += s;
    }
}

/* ------------------------------------------------------------------------- */
/* small helpers                                                             */
/* ------------------------------------------------------------------------- */
/* The N-length sweeps are OpenMP-parallel so that the CPU baseline uses every host
 * core (the Julia reference runs these broadcasts single-threaded; only its BLAS
 * calls are threaded — the oracle is therefore a generous stand-in). */
#define PAR_MIN 16384
static double vnorm2(int64_t n, const double* a) {
    double s = 0.0;
    #pragma omp parallel for reduction(+:s) schedule(static) if (n > PAR_MIN)
    for (int64_t i = 0; i < n; ++i) s += a[i] * a[i];
    return sqrt(s);
}
static double vnorminf(int64_t n, const double* a) {
    double s = 0.0;
    int has_nan = 0;
    #pragma omp parallel for reduction(max:s) reduction(|:has_nan) schedule(static) if (n > PAR_MIN)
    for (int64_t i = 0; i < n; ++i) { double t = fabs(a[i]); if (t != t) has_nan = 1; if (t > s) s = t; }
    return has_nan ? NAN : s;
}
static double vdot(int64_t n, const double* a, const double* b) {
    double s = 0.0;
    #pragma omp parallel for reduction(+:s) schedule(static) if (n > PAR_MIN)
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* CircularVector (structs.jl:2-30): index i (1-based, any integer) -> mod1(i, l) */
typedef struct { double* v; int64_t l; } circ_t;
static int64_t mod1_(int64_t i, int64_t l) { int64_t r = ((i - 1) % l + l) % l; return r; /* 0-based slot */ }
static double circ_get(const circ_t* c, int64_t i) { return c->v[mod1_(i, c->l)]; }
static void circ_set(circ_t* c, int64_t i, double x) { c->v[mod1_(i, c->l)] = x; }
/* max_abs_diff (structs.jl:14-20): all l slots including the wrap seam v[1]-v[0] == v[1]-v[l] */
static double circ_max_abs_diff(const circ_t* c) {
    double val = 0.0;
    for (int64_t i = 1; i <= c->l; ++i) {
        double d = fabs(circ_get(c, i) - circ_get(c, i - 1));
        if (d > val) val = d;     /* Julia max(0.0, NaN) = NaN; NaN never occurs before the isnan exits */
    }
    return val;
}

/* Substitute for Julia's MersenneTwister stream (eigsolver.jl:392-411): splitmix64
 * + Box-Muller.  The product's host code uses the identical generator. */
static uint64_t splitmix64_next(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static double sm_uniform(uint64_t* s) { return (double)(splitmix64_next(s) >> 11) * (1.0 / 9007199254740992.0); }
void oracle_eig_resid(int64_t n, int64_t seed, int64_t init, double* out) {
    uint64_t s = (uint64_t)seed;
    if (init == 3) {
        for (int64_t i = 0; i < n; ++i) {
            double u1 = 1.0 - sm_uniform(&s);
            double u2 = sm_uniform(&s);
            out[i] = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
        }
        double nn = vnorm2(n, out);
        for (int64_t i = 0; i < n; ++i) out[i] /= nn;
    } else if (init == 2) {
        for (int64_t i = 0; i < n; ++i) out[i] = sm_uniform(&s);
    } else if (init == 1) {
        for (int64_t i = 0; i < n; ++i) out[i] = 1.0;
    } else {
        for (int64_t i = 0; i < n; ++i) out[i] = 0.0;
    }
}

/* ------------------------------------------------------------------------- */
/* solver state                                                              */
/* ------------------------------------------------------------------------- */
typedef struct {
    /* problem (working copies, permuted + scaled) */
    int64_t n, p, m, R;
    csc_t A, G, A_orig, G_orig, M, Mt;
    double *b, *h, *c, *b_orig, *h_orig, *c_orig;
    double *eq_E, *eq_D;    /* diagonal preconditioners of equilibrate! (NULL when equilibration is off) */
    int64_t n_sdp, n_soc;
    int64_t* sdp_side;
    int64_t* sdp_off;       /* offset of cone k's svec block in x */
    int64_t* soc_off;       /* offset of SOC k in x (t first) */
    int64_t* soc_len;
    int64_t* var_ordering;  /* inverse permutation */
    /* iterates (structs.jl:83-151) */
    double *x, *x_old, *y, *y_old, *Mty, *Mty_old, *Mx, *Mx_old, *y_half, *y_temp;
    double** mat;           /* dense n_k x n_k per cone, column-major */
    double** resid;         /* Lanczos start vectors */
    /* eig workspaces */
    double* eig_w; double* eig_Z; double* lan_vals; double* lan_vecs;
    /* Params (structs.jl:159-192) */
    int64_t *current_rank, *target_rank;
    double* min_eig;
    int64_t rank_update, update_cont, iter, stop_reason;
    char stop_reason_string[PROXSDP_STATUS_STRING_LEN];
    double primal_step, primal_step_old, dual_step, theta, beta, adapt_level;
    int64_t window;
    double time0, norm_c, norm_b, norm_h;
    double dual_feasibility;
    int64_t dual_feasibility_check, certificate_search, certificate_search_min_iter, certificate_found;
    /* Residuals (structs.jl:100-122) */
    circ_t dual_gap, prim_obj, dual_obj, feasibility, primal_residual, dual_residual, comb_residual;
    double equa_feasibility, ineq_feasibility;
    /* eig solver state (EigSolverAlloc) */
    int* eig_converged; int64_t* eig_converged_eigs;
    /* stats */
    double time_psd; int64_t n_psd, lanczos_matvecs, lanczos_calls, full_eig_calls, linesearch_trials;
    int64_t trace_mv0, trace_ls0;
    /* sharded runs (test infrastructure for the multi-GPU host logic): whole-problem scalars are combined
       across ranks through the host callback of proxsdp_shard_t */
    const proxsdp_shard_t* shard;
    int64_t global_n, global_R;
    int global_has_soc;
} state_t;

static void red(const state_t* s, double* vals, int64_t count, int64_t op) {
    if (s->shard && s->shard->nranks > 1 && s->shard->reduce) s->shard->reduce(vals, count, op, s->shard->reduce_ctx);
}

typedef proxsdp_options_t opts_t;

/* psd_vec_to_square (prox_operators.jl:1-16): svec -> upper triangle, off-diag / sqrt_2 */
static void psd_vec_to_square(const state_t* s, const double* v, double sqrt_2) {
    int64_t cont = 0;
    for (int64_t k = 0; k < s->n_sdp; ++k) {
        int64_t n = s->sdp_side[k];
        double* X = s->mat[k];
        #pragma omp parallel for schedule(static, 16) if (n > 256)
        for (int64_t j = 0; j < n; ++j) {
            const double* vj = v + cont + j * (j + 1) / 2;
            for (int64_t i = 0; i <= j; ++i) X[i + j * n] = (i != j) ? vj[i] / sqrt_2 : vj[i];
        }
        cont += n * (n + 1) / 2;
    }
}
/* psd_square_to_vec (prox_operators.jl:17-31) */
static void psd_square_to_vec(const state_t* s, double* v, double sqrt_2) {
    int64_t cont = 0;
    for (int64_t k = 0; k < s->n_sdp; ++k) {
        int64_t n = s->sdp_side[k];
        const double* X = s->mat[k];
        #pragma omp parallel for schedule(static, 16) if (n > 256)
        for (int64_t j = 0; j < n; ++j) {
            double* vj = v + cont + j * (j + 1) / 2;
            for (int64_t i = 0; i <= j; ++i) vj[i] = (i != j) ? X[i + j * n] * sqrt_2 : X[i + j * n];
        }
        cont += n * (n + 1) / 2;
    }
}

/* BLAS.gemm!('N','T', val, v, v, 1, X): full n x n rank-1 update (prox_operators.jl:82,104,119) */
static void rank1_update(int64_t n, double val, const double* v, double* X) {
    #pragma omp parallel for schedule(static) if (n > 256)
    for (int64_t j = 0; j < n; ++j) {
        double t = val * v[j];
        double* col = X + j * n;
        for (int64_t i = 0; i < n; ++i) col[i] += v[i] * t;
    }
}

/* full_eig! (prox_operators.jl:111-126) */
static void full_eig(state_t* s, const opts_t* opt, int64_t idx) {
    int64_t n = s->sdp_side[idx];
    double* X = s->mat[idx];
    s->current_rank[idx] = 0;
    oracle_eigh(n, X, s->eig_w, s->eig_Z);
    s->full_eig_calls++;
    s->min_eig[idx] = 0.0;
    memset(X, 0, sizeof(double) * (size_t)n * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        if (s->eig_w[i] > 0.0) {
            rank1_update(n, s->eig_w[i], s->eig_Z + (size_t)i * n, X);
            if (s->eig_w[i] > opt->tol_psd) s->current_rank[idx]++;
        }
    }
}

/* krylovkit_eig! wrapper + solver (prox_operators.jl:89-109, eigsolver.jl:798-823) */
static void krylov_eig(state_t* s, const opts_t* opt, int64_t idx) {
    int64_t n = s->sdp_side[idx];
    double* X = s->mat[idx];
    int64_t nev = s->target_rank[idx];
    int64_t ncv = 2 * nev + 1;
    if (ncv < opt->eigsolver_min_lanczos) ncv = opt->eigsolver_min_lanczos;   /* eigsolver.jl:794 */
    int64_t nvals = 0, conv = 0, numops = 0, numiter = 0;
    s->eig_converged[idx] = 1;
    oracle_lanczos_eager = opt->krylovkit_eager != 0;                            /* eigsolver.jl:809 */
    oracle_lanczos(n, X, s->resid[idx], nev, ncv, opt->krylovkit_max_iter, opt->krylovkit_tol,
                   s->lan_vals, s->lan_vecs, &nvals, &conv, &numops, &numiter);
    oracle_lanczos_eager = 0;
    s->lanczos_calls++;
    s->lanczos_matvecs += numops;
    s->eig_converged_eigs[idx] = conv;
    if (conv == 0) s->eig_converged[idx] = 0;                                   /* eigsolver.jl:816-818 */
    if (s->eig_converged[idx]) {
        memset(X, 0, sizeof(double) * (size_t)n * (size_t)n);
        double mn = s->lan_vals[0];
        for (int64_t i = 1; i < nvals; ++i) if (s->lan_vals[i] < mn) mn = s->lan_vals[i];
        s->min_eig[idx] = mn;
        int64_t lim = nev < conv ? nev : conv;
        for (int64_t i = 0; i < lim; ++i) {
            double val = s->lan_vals[i];
            if (val > 0.0) {
                s->current_rank[idx]++;
                rank1_update(n, val, s->lan_vecs + (size_t)i * n, X);
            }
        }
    }
}

/* psd_projection! (prox_operators.jl:33-66) */
static void psd_projection(state_t* s, const opts_t* opt, double* v) {
    double t0 = now_s();
    for (int64_t k = 0; k < s->n_sdp; ++k) s->min_eig[k] = 0.0;
    psd_vec_to_square(s, v, sqrt(2.0));
    for (int64_t idx = 0; idx < s->n_sdp; ++idx) {
        int64_t n = s->sdp_side[idx];
        s->current_rank[idx] = 0;
        if (n == 1) {
            double* X = s->mat[idx];
            X[0] = X[0] > 0.0 ? X[0] : 0.0;
            s->min_eig[idx] = X[0];
        } else if (!opt->full_eig_decomp &&
                   s->target_rank[idx] <= opt->max_target_rank_krylov_eigs &&
                   n > opt->min_size_krylov_eigs &&
                   (s->iter % opt->full_eig_freq) > opt->full_eig_len) {
            /* eigsolver == 1 (ARPACK) is served by the same Lanczos restatement */
            krylov_eig(s, opt, idx);
            if (!s->eig_converged[idx]) full_eig(s, opt, idx);
        } else {
            full_eig(s, opt, idx);
        }
    }
    psd_square_to_vec(s, v, sqrt(2.0));
    s->time_psd += now_s() - t0;
    s->n_psd++;
}

/* soc_projection! (prox_operators.jl:138-158) */
static void soc_projection(state_t* s, double* v) {
    for (int64_t k = 0; k < s->n_soc; ++k) {
        double* t = v + s->soc_off[k];
        double* w = t + 1;
        int64_t len = s->soc_len[k] - 1;
        double nv = vnorm2(len, w);
        if (nv <= -t[0]) {
            t[0] = 0.0;
            for (int64_t i = 0; i < len; ++i) w[i] = 0.0;
        } else if (nv <= t[0]) {
        } else {
            double val = 0.5 * (1.0 + t[0] / nv);
            for (int64_t i = 0; i < len; ++i) w[i] *= val;
            t[0] = val * nv;
        }
    }
}

/* box_projection! (prox_operators.jl:160-170) */
static void box_projection(const state_t* s, double* v, double step) {
    for (int64_t i = 0; i < s->p; ++i) v[i] = s->b[i];
    for (int64_t i = 0; i < s->m; ++i) {
        double t = v[s->p + i] / step;
        v[s->p + i] = t < s->h[i] ? t : s->h[i];
    }
}

/* primal_step! (pdhg.jl:611-637) */
static void primal_step(state_t* s, const opts_t* opt) {
    #pragma omp parallel for schedule(static) if (s->n > PAR_MIN)
    for (int64_t i = 0; i < s->n; ++i) s->x[i] -= s->primal_step * (s->Mty[i] + s->c[i]);
    if (s->n_sdp >= 1) psd_projection(s, opt, s->x);
    if (s->n_soc >= 1) soc_projection(s, s->x);
    csc_mul(&s->M, s->x, s->Mx);
}

/* linesearch! (pdhg.jl:532-582) */
static void linesearch(state_t* s, const opts_t* opt) {
    int64_t R = s->R;
    s->primal_step = s->primal_step * sqrt(1.0 + s->theta);
    for (int64_t it = 0; it < opt->max_linsearch_steps; ++it) {
        s->linesearch_trials++;
        s->theta = s->primal_step / s->primal_step_old;
        double bt = s->beta * s->primal_step;
        for (int64_t i = 0; i < R; ++i)
            s->y_half[i] = s->y[i] + bt * ((1.0 + s->theta) * s->Mx[i] - s->theta * s->Mx_old[i]);
        memcpy(s->y_temp, s->y_half, sizeof(double) * (size_t)R);
        box_projection(s, s->y_half, bt);
        for (int64_t i = 0; i < R; ++i) s->y_temp[i] -= bt * s->y_half[i];
        csc_mul(&s->Mt, s->y_temp, s->Mty);
        /* in-place norms (pdhg.jl:559-564) */
        #pragma omp parallel for schedule(static) if (s->n > PAR_MIN)
    for (int64_t i = 0; i < s->n; ++i) s->Mty[i] -= s->Mty_old[i];
        for (int64_t i = 0; i < R; ++i) s->y_temp[i] -= s->y_old[i];
        double y_norm = vnorm2(R, s->y_temp);
        double Mty_norm = vnorm2(s->n, s->Mty);
        if (s->shard) {
            double v[2] = {y_norm * y_norm, Mty_norm * Mty_norm};
            red(s, v, 2, 0);
            y_norm = sqrt(v[0]); Mty_norm = sqrt(v[1]);
        }
        if (sqrt(s->beta) * s->primal_step * Mty_norm <= opt->delta * y_norm) {
            break;
        } else {
            s->primal_step *= opt->linsearch_decay;
        }
    }
    #pragma omp parallel for schedule(static) if (s->n > PAR_MIN)
    for (int64_t i = 0; i < s->n; ++i) s->Mty[i] += s->Mty_old[i];
    for (int64_t i = 0; i < R; ++i) s->y_temp[i] += s->y_old[i];
    memcpy(s->y, s->y_temp, sizeof(double) * (size_t)R);
    s->primal_step_old = s->primal_step;
    s->dual_step = s->beta * s->primal_step;
}

/* dual_step! (pdhg.jl:584-609) */
static void dual_step(state_t* s) {
    int64_t R = s->R;
    for (int64_t i = 0; i < R; ++i) s->y_half[i] = s->y[i] + s->dual_step * (2.0 * s->Mx[i] - s->Mx_old[i]);
    memcpy(s->y_temp, s->y_half, sizeof(double) * (size_t)R);
    box_projection(s, s->y_half, s->dual_step);
    for (int64_t i = 0; i < R; ++i) s->y_temp[i] -= s->dual_step * s->y_half[i];
    csc_mul(&s->Mt, s->y_temp, s->Mty);
    memcpy(s->y, s->y_temp, sizeof(double) * (size_t)R);
    s->primal_step_old = s->primal_step;
}

/* compute_residual! (residuals.jl:37-71) */
static void compute_residual(state_t* s) {
    int64_t n = s->n, R = s->R;
    double mx1;
    #pragma omp parallel for schedule(static) if (n > PAR_MIN)
    for (int64_t i = 0; i < n; ++i) s->Mty_old[i] = s->x_old[i] - s->primal_step * s->Mty_old[i];
    #pragma omp parallel for schedule(static) if (n > PAR_MIN)
    for (int64_t i = 0; i < n; ++i) s->x_old[i] = s->x[i] - s->primal_step * s->Mty[i];
    #pragma omp parallel for schedule(static) if (n > PAR_MIN)
    for (int64_t i = 0; i < n; ++i) s->x_old[i] -= s->Mty_old[i];
    mx1 = vnorminf(n, s->Mty_old);
    double nx = vnorminf(n, s->x_old);
    if (s->shard) { double v[2] = {mx1, nx}; red(s, v, 2, 1); mx1 = v[0]; nx = v[1]; }
    if (s->norm_b > mx1) mx1 = s->norm_b;
    if (s->norm_h > mx1) mx1 = s->norm_h;
    if (1.0 > mx1) mx1 = 1.0;
    double pr = sqrt((double)(s->shard ? s->global_n : n)) * nx / mx1;
    circ_set(&s->primal_residual, s->iter, pr);

    for (int64_t i = 0; i < R; ++i) s->Mx_old[i] = s->y_old[i] - s->dual_step * s->Mx_old[i];
    for (int64_t i = 0; i < R; ++i) s->y_old[i] = s->y[i] - s->dual_step * s->Mx[i];
    for (int64_t i = 0; i < R; ++i) s->y_old[i] -= s->Mx_old[i];
    double mx2 = vnorminf(R, s->Mx_old);
    double ny = vnorminf(R, s->y_old);
    if (s->shard) { double v[2] = {mx2, ny}; red(s, v, 2, 1); mx2 = v[0]; ny = v[1]; }
    if (s->norm_c > mx2) mx2 = s->norm_c;
    if (1.0 > mx2) mx2 = 1.0;
    double dr = sqrt((double)(s->shard ? s->global_R : R)) * ny / mx2;
    circ_set(&s->dual_residual, s->iter, dr);
    circ_set(&s->comb_residual, s->iter, pr > dr ? pr : dr);

    memcpy(s->x_old, s->x, sizeof(double) * (size_t)n);
    memcpy(s->y_old, s->y, sizeof(double) * (size_t)R);
    memcpy(s->Mty_old, s->Mty, sizeof(double) * (size_t)n);
    memcpy(s->Mx_old, s->Mx, sizeof(double) * (size_t)R);
}

/* compute_gap! (residuals.jl:2-35) */
static void compute_gap(state_t* s) {
    /* sharded: a rank may hold no equality (inequality) rows while the whole problem does */
    double fe = 0.0, fi = 0.0;
    for (int64_t i = 0; i < s->p; ++i) { double t = fabs(s->Mx[i] - s->b[i]); if (t > fe) fe = t; }
    for (int64_t i = 0; i < s->m; ++i) { double t = s->Mx[s->p + i] - s->h[i]; if (t > fi) fi = t; }
    int64_t gp = s->p, gm = s->m;
    if (s->shard) { double v[2] = {fe, fi}; red(s, v, 2, 1); fe = v[0]; fi = v[1]; gp = s->shard->global_p; gm = s->shard->global_m; }
    if (gp > 0) s->equa_feasibility = fe / (1.0 + s->norm_b);
    if (gm > 0) s->ineq_feasibility = fi / (1.0 + s->norm_h);
    double feas = s->equa_feasibility > s->ineq_feasibility ? s->equa_feasibility : s->ineq_feasibility;
    circ_set(&s->feasibility, s->iter, feas);
    double po = vdot(s->n, s->c, s->x);
    double by = 0.0, hy = 0.0;
    if (s->p > 0) by = vdot(s->p, s->b, s->y);
    if (s->m > 0) hy = vdot(s->m, s->h, s->y + s->p);
    if (s->shard) { double v[3] = {po, by, hy}; red(s, v, 3, 0); po = v[0]; by = v[1]; hy = v[2]; }
    double dobj = 0.0;
    if (gp > 0) dobj -= by;
    if (gm > 0) dobj -= hy;
    circ_set(&s->prim_obj, s->iter, po);
    circ_set(&s->dual_obj, s->iter, dobj);
    circ_set(&s->dual_gap, s->iter, fabs(po - dobj) / (1.0 + fabs(po) + fabs(dobj)));
}

/* soc_convergence (residuals.jl:73-86) */
static int soc_convergence(const state_t* s, const opts_t* opt) {
    int ok = 1;
    for (int64_t k = 0; k < s->n_soc; ++k) {
        const double* t = s->x + s->soc_off[k];
        if (vnorm2(s->soc_len[k] - 1, t + 1) - t[0] >= opt->tol_soc) { ok = 0; break; }
    }
    if (s->shard) { double f = ok ? 0.0 : 1.0; red(s, &f, 1, 1); ok = (f == 0.0); }
    return ok;
}
/* convergedrank (residuals.jl:88-101) */
static int convergedrank(const state_t* s, const opts_t* opt) {
    int ok = 1;
    for (int64_t k = 0; k < s->n_sdp; ++k) {
        if (!(s->sdp_side[k] < opt->min_size_krylov_eigs ||
              s->target_rank[k] > opt->max_target_rank_krylov_eigs ||
              s->min_eig[k] < opt->tol_psd)) { ok = 0; break; }
    }
    if (s->shard) { double f = ok ? 0.0 : 1.0; red(s, &f, 1, 1); ok = (f == 0.0); }
    return ok;
}

/* fix_diag_scaling (pdhg.jl:734-743) */
static void fix_diag_scaling(const state_t* s, double* v, double num) {
    int64_t cont = 0;
    for (int64_t k = 0; k < s->n_sdp; ++k) {
        int64_t n = s->sdp_side[k];
        for (int64_t j = 0; j < n; ++j)
            for (int64_t i = 0; i <= j; ++i) {
                if (i != j) v[cont] /= num;
                cont++;
            }
    }
}

/* cone_feas (pdhg.jl:678-699) — NB the reference folds the SOC violation into
 * sdp_viol (latent quirk, pdhg.jl:695); the returned max is the same. */
static double cone_feas(state_t* s, const double* v, int64_t* cont_out) {
    double viol = 0.0;
    psd_vec_to_square(s, v, sqrt(2.0));
    int64_t cont = 0;
    for (int64_t k = 0; k < s->n_sdp; ++k) {
        int64_t n = s->sdp_side[k];
        cont += n * (n + 1) / 2;
        if (n == 1) {
            double t = -fmin(0.0, s->mat[k][0]);
            if (t > viol) viol = t;
        } else {
            oracle_eigh(n, s->mat[k], s->eig_w, s->eig_Z);
            double t = -fmin(0.0, s->eig_w[0]);
            if (t > viol) viol = t;
        }
    }
    for (int64_t k = 0; k < s->n_soc; ++k) {
        int64_t len = s->soc_len[k];
        double sv = v[cont];
        double t = -fmin(0.0, sv - vnorm2(len - 1, v + cont + 1));
        if (t > viol) viol = t;
        cont += len;
    }
    *cont_out = cont;
    return viol;
}

/* get_duals (pdhg.jl:701-710): dual_cone = c + A' y_eq + G' y_in, then /2 on off-diagonals */
static void get_duals(const state_t* s, const double* y, const double* c, double* dual_cone) {
    for (int64_t i = 0; i < s->n; ++i) dual_cone[i] = c[i];
    csc_mul_t_add(&s->A_orig, y, dual_cone);
    csc_mul_t_add(&s->G_orig, y + s->p, dual_cone);
    fix_diag_scaling(s, dual_cone, 2.0);
}

/* dual_feas (pdhg.jl:712-732) */
static double dual_feas_from(state_t* s, const double* dual_in, const double* dual_cone) {
    double ineq_viol = 0.0;
    if (s->m > 0) {
        double mn = dual_in[0];
        for (int64_t i = 1; i < s->m; ++i) if (dual_in[i] < mn) mn = dual_in[i];
        ineq_viol = -fmin(0.0, mn);
    }
    int64_t cont = 0;
    double cone_viol = cone_feas(s, dual_cone, &cont);
    double zero_viol = 0.0;
    for (int64_t i = cont; i < s->n; ++i) { double t = fabs(dual_cone[i]); if (t > zero_viol) zero_viol = t; }
    double r = cone_viol;
    if (ineq_viol > r) r = ineq_viol;
    if (zero_viol > r) r = zero_viol;
    return r;
}
static double dual_feas_y(state_t* s, const double* y, const double* c) {
    double* dc = (double*)malloc(sizeof(double) * (size_t)(s->n > 0 ? s->n : 1));
    get_duals(s, y, c, dc);
    double r = dual_feas_from(s, y + s->p, dc);
    free(dc);
    red(s, &r, 1, 1);
    return r;
}

/* cache_solution (pdhg.jl:745-787).  NB it rescales pair.x IN PLACE (pdhg.jl:749),
 * which matters when it is called mid-loop by the certificate branches. */
static void cache_solution(state_t* s, const opts_t* opt, const double* c, proxsdp_result_t* out) {
    fix_diag_scaling(s, s->x, sqrt(2.0));
    int64_t n = s->n, p = s->p, m = s->m;
    if (opt->equilibration && s->eq_E) {          /* remove equilibrating (pdhg.jl:751-755) */
        for (int64_t j = 0; j < n; ++j) s->x[j] = s->eq_D[j] * s->x[j];
        for (int64_t i = 0; i < p + m; ++i) s->y[i] = s->eq_E[i] * s->y[i];
    }
    double* slack_eq = (double*)calloc((size_t)(p > 0 ? p : 1), sizeof(double));
    double* slack_in = (double*)calloc((size_t)(m > 0 ? m : 1), sizeof(double));
    double* dual_cone = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    if (p > 0) { csc_mul(&s->A_orig, s->x, slack_eq); for (int64_t i = 0; i < p; ++i) slack_eq[i] -= s->b_orig[i]; }
    if (m > 0) { csc_mul(&s->G_orig, s->x, slack_in); for (int64_t i = 0; i < m; ++i) slack_in[i] -= s->h_orig[i]; }
    get_duals(s, s->y, c, dual_cone);
    double dfeas = dual_feas_from(s, s->y + p, dual_cone);
    red(s, &dfeas, 1, 1);

    out->status = s->stop_reason;
    snprintf(out->status_string, PROXSDP_STATUS_STRING_LEN, "%s", s->stop_reason_string);
    if (out->primal) for (int64_t i = 0; i < n; ++i) out->primal[i] = s->x[s->var_ordering[i]];
    if (out->dual_cone) for (int64_t i = 0; i < n; ++i) out->dual_cone[i] = dual_cone[s->var_ordering[i]];
    if (out->dual_eq) for (int64_t i = 0; i < p; ++i) out->dual_eq[i] = s->y[i];
    if (out->dual_in) for (int64_t i = 0; i < m; ++i) out->dual_in[i] = s->y[p + i];
    if (out->slack_eq) for (int64_t i = 0; i < p; ++i) out->slack_eq[i] = slack_eq[i];
    if (out->slack_in) for (int64_t i = 0; i < m; ++i) out->slack_in[i] = slack_in[i];
    out->primal_residual = s->equa_feasibility;
    out->dual_residual = s->ineq_feasibility;
    out->objval = circ_get(&s->prim_obj, s->iter);
    out->dual_objval = circ_get(&s->dual_obj, s->iter);
    out->gap = circ_get(&s->dual_gap, s->iter);
    out->time = now_s() - s->time0;
    out->iter = s->iter;
    int64_t fr = 0;
    for (int64_t k = 0; k < s->n_sdp; ++k) fr += s->current_rank[k];
    { double f = (double)fr; red(s, &f, 1, 0); fr = (int64_t)f; }
    out->final_rank = fr;
    out->primal_feasible_user_tol = circ_get(&s->feasibility, s->iter) <= opt->tol_feasibility;
    out->dual_feasible_user_tol = dfeas <= opt->tol_feasibility_dual;
    out->certificate_found = s->certificate_found;
    out->result_count = 1;
    out->final_primal_res = circ_get(&s->primal_residual, s->iter);
    out->final_dual_res = circ_get(&s->dual_residual, s->iter);
    free(slack_eq); free(slack_in); free(dual_cone);
}

/* equilibrate! (equilibration.jl:1-71): Ruiz-like diagonal scaling by projected gradient steps on
   sum_ij (E_i M_ij D_j)^2 with averaged iterates; the column scaling is re-set to its mean in every step */
static void eq_box_project(double* y, int64_t len, double lb, double ub) {
    for (int64_t i = 0; i < len; ++i) y[i] = fmin(ub, fmax(y[i], lb));
}
static void equilibrate(const csc_t* M, int64_t n, int64_t R, const opts_t* opt, double* E, double* D) {
    int64_t max_iters = opt->equilibration_iters;
    double lb = opt->equilibration_lb, ub = opt->equilibration_ub;
    double alpha = pow((double)n / (double)R, 0.25), beta = pow((double)R / (double)n, 0.25);
    double alpha2 = alpha * alpha, beta2 = beta * beta, gamma = 0.1;
    double* u = (double*)calloc((size_t)R + 1, sizeof(double));
    double* v = (double*)calloc((size_t)n + 1, sizeof(double));
    double* u_ = (double*)calloc((size_t)R + 1, sizeof(double));
    double* v_ = (double*)calloc((size_t)n + 1, sizeof(double));
    double* row_norms = (double*)calloc((size_t)R + 1, sizeof(double));
    double* col_norms = (double*)calloc((size_t)n + 1, sizeof(double));
    for (int64_t iter = 1; iter <= max_iters; ++iter) {
        for (int64_t i = 0; i < R; ++i) E[i] = exp(u[i]);
        for (int64_t j = 0; j < n; ++j) D[j] = exp(v[j]);
        double step_size = 2.0 / (gamma * ((double)iter + 1.0));
        for (int64_t i = 0; i < R; ++i) row_norms[i] = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            double cn = 0.0;
            for (int64_t q = M->colptr[j]; q < M->colptr[j + 1]; ++q) {
                double w = E[M->rowidx[q]] * (M->val[q] * D[j]);      /* mul!(M_, M, D); mul!(M_, E, M_) */
                row_norms[M->rowidx[q]] += w * w;
                cn += w * w;
            }
            col_norms[j] = cn;
        }
        for (int64_t i = 0; i < R; ++i) u[i] -= step_size * (row_norms[i] - alpha2 + gamma * u[i]);
        eq_box_project(u, R, lb, ub);
        double sum_v = 0.0;
        for (int64_t j = 0; j < n; ++j) { v[j] -= step_size * (col_norms[j] - beta2 + gamma * v[j]); sum_v += v[j]; }
        for (int64_t j = 0; j < n; ++j) v[j] = sum_v / (double)n;
        eq_box_project(v, n, 0.0, ub);
        for (int64_t i = 0; i < R; ++i) u_[i] = 2.0 * u[i] / ((double)iter + 2.0) + (double)iter * u_[i] / ((double)iter + 2.0);
        for (int64_t j = 0; j < n; ++j) v_[j] = 2.0 * v[j] / ((double)iter + 2.0) + (double)iter * v_[j] / ((double)iter + 2.0);
    }
    for (int64_t i = 0; i < R; ++i) E[i] = exp(u_[i]);
    for (int64_t j = 0; j < n; ++j) D[j] = exp(v_[j]);
    free(u); free(v); free(u_); free(v_); free(row_norms); free(col_norms);
}

/* sigma_max(M) = sqrt(lambda_max of the smaller Gram matrix), by the dense symmetric eigensolver below a side of 4000,
   by power iteration on M M' beyond that (the reference: ARPACK svds; any converged method returns the same number) */
static double exact_spectral_norm(const csc_t* M, const csc_t* Mt) {
    int64_t R = M->nr, n = M->nc;
    if (R == 0 || n == 0 || M->nnz == 0) return 0.0;
    int64_t d = (R <= n) ? R : n;                /* side of the smaller Gram matrix: M M' or M' M */
    if (d <= 4000) {
        double* Gm = (double*)calloc((size_t)d * (size_t)d, sizeof(double));
        double* w = (double*)calloc((size_t)d, sizeof(double));
        double* Z = (double*)calloc((size_t)d * (size_t)d, sizeof(double));
        /* Gram[i][j] = sum_c T[i][c] T[j][c]: T = M (columns = variables) for M M', T = M' (columns = constraints) for M' M */
        const csc_t* T = (R <= n) ? M : Mt;
        for (int64_t c = 0; c < T->nc; ++c)
            for (int64_t q1 = T->colptr[c]; q1 < T->colptr[c + 1]; ++q1)
                for (int64_t q2 = T->colptr[c]; q2 < T->colptr[c + 1]; ++q2) {
                    int64_t i = T->rowidx[q1], j = T->rowidx[q2];
                    if (i <= j) Gm[i + j * d] += T->val[q1] * T->val[q2];
                }
        int rc = oracle_eigh(d, Gm, w, Z);
        double lam = rc == 0 ? w[d - 1] : -1.0;
        free(Gm); free(w); free(Z);
        return lam >= 0.0 ? sqrt(lam) : (rc == 0 ? 0.0 : -1.0);
    }
    double* v = (double*)calloc((size_t)R, sizeof(double));
    double* t = (double*)calloc((size_t)n, sizeof(double));
    double* w = (double*)calloc((size_t)R, sizeof(double));
    uint64_t seed = 1234;
    for (int64_t i = 0; i < R; ++i) v[i] = 0.5 + sm_uniform(&seed);
    double nv = vnorm2(R, v), lam = 0.0, lam_old = -1.0;
    for (int64_t i = 0; i < R; ++i) v[i] /= nv;
    int ok = 0;
    for (int it = 0; it < 200000; ++it) {
        csc_mul(Mt, v, t);
        csc_mul(M, t, w);
        lam = vdot(R, v, w);
        nv = vnorm2(R, w);
        if (nv == 0.0) { lam = 0.0; ok = 1; break; }
        for (int64_t i = 0; i < R; ++i) v[i] = w[i] / nv;
        if (fabs(lam - lam_old) <= 1e-15 * fabs(lam)) { ok = 1; break; }
        lam_old = lam;
    }
    free(v); free(t); free(w);
    return ok ? sqrt(fmax(lam, 0.0)) : -1.0;
}

/* certificate_parameters (pdhg.jl:670-676) */
static void certificate_parameters(state_t* s, opts_t* opt) {
    s->certificate_search_min_iter = s->iter + 2 * opt->convergence_window + s->iter / 5 + 1000;
    s->certificate_search = 1;
    opt->time_limit *= 1.1;
    opt->max_iter_local = opt->max_iter_local + opt->max_iter_local / 10;
}
/* certificate_dual_infeasibility (pdhg.jl:639-653) */
static void certificate_dual_infeasibility(state_t* s, opts_t* opt) {
    for (int64_t i = 0; i < s->p; ++i) s->b[i] = 0.0;
    for (int64_t i = 0; i < s->m; ++i) s->h[i] = 0.0;
    certificate_parameters(s, opt);
}
/* certificate_infeasibility (pdhg.jl:655-668) */
static void certificate_infeasibility(state_t* s, opts_t* opt) {
    for (int64_t i = 0; i < s->n; ++i) s->c[i] = 0.0;
    certificate_parameters(s, opt);
}

static void rank_increment_rule(state_t* s, const opts_t* opt, int64_t idx) {
    /* pdhg.jl:271-279 / 294-302 */
    if (opt->freeze_target_rank) return;
    if (s->current_rank[idx] + opt->rank_slack >= s->target_rank[idx]) {
        if (s->min_eig[idx] > opt->tol_psd) {
            int64_t t;
            if (opt->rank_increment == 0) t = opt->rank_increment_factor * s->target_rank[idx];
            else t = opt->rank_increment_factor + s->target_rank[idx];
            if (t > s->sdp_side[idx]) t = s->sdp_side[idx];
            s->target_rank[idx] = t;
        }
    }
}

static int cmp_i64(const void* a, const void* b) {
    int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}

static void record_trace(state_t* s, proxsdp_result_t* out, int64_t cap) {
    if (!out->trace || out->trace_len >= cap) return;
    double* row = out->trace + out->trace_len * PROXSDP_TRACE_COLS;
    int64_t k = s->iter;
    double tr = 0, cr = 0, me = 0;
    for (int64_t i = 0; i < s->n_sdp; ++i) {
        tr += (double)s->target_rank[i]; cr += (double)s->current_rank[i];
        if (i == 0 || s->min_eig[i] < me) me = s->min_eig[i];
    }
    row[0] = (double)k;
    row[1] = circ_get(&s->prim_obj, k);
    row[2] = circ_get(&s->dual_obj, k);
    row[3] = circ_get(&s->dual_gap, k);
    row[4] = circ_get(&s->feasibility, k);
    row[5] = circ_get(&s->primal_residual, k);
    row[6] = circ_get(&s->dual_residual, k);
    row[7] = s->primal_step;
    row[8] = s->beta;
    row[9] = tr; row[10] = cr; row[11] = me;
    row[12] = (double)(s->lanczos_matvecs - s->trace_mv0); row[13] = (double)(s->linesearch_trials - s->trace_ls0);
    out->trace_len++;
}

/* ------------------------------------------------------------------------- */
/* chambolle_pock (pdhg.jl:1-530)                                            */
/* ------------------------------------------------------------------------- */
static int oracle_solve_impl(const proxsdp_problem_t* prob, const proxsdp_options_t* opt_in, const proxsdp_shard_t* shard,
                             proxsdp_result_t* out);
int proxsdp_oracle_solve(const proxsdp_problem_t* prob, const proxsdp_options_t* opt_in, proxsdp_result_t* out) {
    return oracle_solve_impl(prob, opt_in, NULL, out);
}
/* the same solve on this rank's blocks of a stacked problem; whole-problem scalars go through shard->reduce */
int proxsdp_oracle_solve_sharded(const proxsdp_problem_t* prob, const proxsdp_options_t* opt_in, const proxsdp_shard_t* shard,
                                 proxsdp_result_t* out) {
    return oracle_solve_impl(prob, opt_in, (shard && shard->nranks > 1) ? shard : NULL, out);
}
static int oracle_solve_impl(const proxsdp_problem_t* prob, const proxsdp_options_t* opt_in, const proxsdp_shard_t* shard,
                             proxsdp_result_t* out) {
    opts_t optv = *opt_in;          /* the reference mutates opt (pdhg.jl:33-41,673-674) */
    opts_t* opt = &optv;
    state_t S;
    memset(&S, 0, sizeof(S));
    state_t* s = &S;
    int64_t n = prob->n, p = prob->p, m = prob->m, base = prob->index_base;
    s->n = n; s->p = p; s->m = m; s->R = p + m;
    int64_t R = s->R;
    s->shard = shard;
    if (shard) { s->global_n = shard->global_n; s->global_R = shard->global_p + shard->global_m; }
    if ((opt->equilibration || opt->equilibration_force) && shard) return -2;   /* row / column norms of the whole M needed */

    /* Params (pdhg.jl:7-31) */
    s->theta = opt->initial_theta;
    s->adapt_level = opt->initial_adapt_level;
    s->window = opt->convergence_window;
    s->beta = opt->initial_beta;
    s->time0 = now_s();
    s->norm_b = vnorm2(p, prob->b);
    s->norm_h = vnorm2(m, prob->h);
    s->norm_c = vnorm2(n, prob->c);
    if (shard) {
        double v[3] = {s->norm_b * s->norm_b, s->norm_h * s->norm_h, s->norm_c * s->norm_c};
        red(s, v, 3, 0);
        s->norm_b = sqrt(v[0]); s->norm_h = sqrt(v[1]); s->norm_c = sqrt(v[2]);
    }
    s->rank_update = 0; s->stop_reason = 0; s->update_cont = 0;
    snprintf(s->stop_reason_string, PROXSDP_STATUS_STRING_LEN, "Not optimized");
    s->n_sdp = prob->n_sdp; s->n_soc = prob->n_soc;
    int64_t nsd = s->n_sdp, nso = s->n_soc;
    s->target_rank = (int64_t*)calloc((size_t)nsd + 1, sizeof(int64_t));
    s->current_rank = (int64_t*)calloc((size_t)nsd + 1, sizeof(int64_t));
    s->min_eig = (double*)calloc((size_t)nsd + 1, sizeof(double));
    int64_t r0 = opt->initial_target_rank > 0 ? opt->initial_target_rank : 2;
    for (int64_t k = 0; k < nsd; ++k) { s->target_rank[k] = r0; s->current_rank[k] = r0; }
    s->dual_feasibility = -1.0;
    s->dual_feasibility_check = 0;
    s->certificate_search = 0;
    s->certificate_search_min_iter = 0;
    s->certificate_found = 0;
    int64_t ada_count = 0;
    int have_cached = 0;               /* sol = Array{Result}(undef, 0) */

    double any_cone = (nso > 0 || nsd > 0) ? 1.0 : 0.0;
    if (shard) red(s, &any_cone, 1, 1);
    if (opt->max_iter <= 0) {
        if (any_cone != 0.0) opt->max_iter_local = opt->max_iter_conic;
        else opt->max_iter_local = opt->max_iter_lp;
    } else {
        opt->max_iter_local = opt->max_iter;
    }

    /* preprocess! (scaling.jl:2-26) */
    s->sdp_side = (int64_t*)calloc((size_t)nsd + 1, sizeof(int64_t));
    s->sdp_off = (int64_t*)calloc((size_t)nsd + 1, sizeof(int64_t));
    s->soc_off = (int64_t*)calloc((size_t)nso + 1, sizeof(int64_t));
    s->soc_len = (int64_t*)calloc((size_t)nso + 1, sizeof(int64_t));
    int64_t* ord = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    char* used = (char*)calloc((size_t)n + 1, 1);
    int64_t pos = 0, max_side = 1;
    for (int64_t k = 0; k < nsd; ++k) {
        s->sdp_side[k] = prob->sdp_side[k];
        if (s->sdp_side[k] > max_side) max_side = s->sdp_side[k];
        s->sdp_off[k] = pos;
        for (int64_t q = prob->sdp_ptr[k]; q < prob->sdp_ptr[k + 1]; ++q) {
            int64_t v = prob->sdp_idx[q] - base;
            ord[pos++] = v; used[v] = 1;
        }
    }
    for (int64_t k = 0; k < nso; ++k) {
        s->soc_off[k] = pos;
        s->soc_len[k] = prob->soc_ptr[k + 1] - prob->soc_ptr[k];
        for (int64_t q = prob->soc_ptr[k]; q < prob->soc_ptr[k + 1]; ++q) {
            int64_t v = prob->soc_idx[q] - base;
            ord[pos++] = v; used[v] = 1;
        }
    }
    {
        int64_t start = pos;
        for (int64_t v = 0; v < n; ++v) if (!used[v]) ord[pos++] = v;
        qsort(ord + start, (size_t)(pos - start), sizeof(int64_t), cmp_i64);
    }
    if (pos != n) { free(ord); free(used); return -3; }   /* a variable in two cones */
    s->var_ordering = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    for (int64_t j = 0; j < n; ++j) s->var_ordering[ord[j]] = j;     /* sortperm(ord) */

    csc_t A0 = csc_from_input(p, n, prob->A_colptr, prob->A_rowval, prob->A_nzval, base);
    csc_t G0 = csc_from_input(m, n, prob->G_colptr, prob->G_rowval, prob->G_nzval, base);
    s->A = csc_permute_cols(&A0, ord);
    s->G = csc_permute_cols(&G0, ord);
    csc_free(&A0); csc_free(&G0);
    s->c = (double*)calloc((size_t)n + 1, sizeof(double));
    s->c_orig = (double*)calloc((size_t)n + 1, sizeof(double));
    for (int64_t j = 0; j < n; ++j) { s->c[j] = prob->c[ord[j]]; s->c_orig[j] = s->c[j]; }
    s->b = (double*)calloc((size_t)p + 1, sizeof(double));
    s->h = (double*)calloc((size_t)m + 1, sizeof(double));
    s->b_orig = (double*)calloc((size_t)p + 1, sizeof(double));
    s->h_orig = (double*)calloc((size_t)m + 1, sizeof(double));
    for (int64_t i = 0; i < p; ++i) { s->b[i] = prob->b[i]; s->b_orig[i] = prob->b[i]; }
    for (int64_t i = 0; i < m; ++i) { s->h[i] = prob->h[i]; s->h_orig[i] = prob->h[i]; }
    /* copies before scaling (pdhg.jl:59-61) */
    s->A_orig = csc_copy(&s->A);
    s->G_orig = csc_copy(&s->G);

    /* diagonal preconditioning (pdhg.jl:64-93) */
    if (opt->equilibration) {
        csc_t M0 = csc_vstack(&s->A, &s->G);
        /* maximum(M) / minimum(M) of a SparseMatrixCSC run over the structural zeros as well */
        double UB = 0.0, LB = 0.0;
        int have = 0;
        for (int64_t q = 0; q < M0.nnz; ++q) {
            if (!have || M0.val[q] > UB) UB = M0.val[q];
            if (!have || M0.val[q] < LB) LB = M0.val[q];
            have = 1;
        }
        if ((double)M0.nnz < (double)R * (double)n) { if (!have || 0.0 > UB) UB = 0.0; if (!have || 0.0 < LB) LB = 0.0; }
        if (LB / UB <= opt->equilibration_limit) opt->equilibration = 0;
        csc_free(&M0);
    }
    /* (with equilibration_force alone the reference reaches equilibrate! with M undefined, pdhg.jl:74-78: an
       UndefVarError.  Here the forced case runs on M = [A; G] like the unforced one — what the author meant.) */
    if (opt->equilibration_force) opt->equilibration = 1;
    if (opt->equilibration) {
        csc_t M0 = csc_vstack(&s->A, &s->G);
        s->eq_E = (double*)calloc((size_t)R + 1, sizeof(double));
        s->eq_D = (double*)calloc((size_t)n + 1, sizeof(double));
        equilibrate(&M0, n, R, opt, s->eq_E, s->eq_D);
        /* M = E * M * D (left to right), A / G = its row blocks, rhs = E * rhs, c = D * c */
        for (int64_t j = 0; j < n; ++j) {
            for (int64_t q = s->A.colptr[j]; q < s->A.colptr[j + 1]; ++q) s->A.val[q] = (s->eq_E[s->A.rowidx[q]] * s->A.val[q]) * s->eq_D[j];
            for (int64_t q = s->G.colptr[j]; q < s->G.colptr[j + 1]; ++q) s->G.val[q] = (s->eq_E[p + s->G.rowidx[q]] * s->G.val[q]) * s->eq_D[j];
            s->c[j] = s->eq_D[j] * s->c[j];
        }
        for (int64_t i = 0; i < p; ++i) s->b[i] = s->eq_E[i] * s->b_orig[i];
        for (int64_t i = 0; i < m; ++i) s->h[i] = s->eq_E[p + i] * s->h_orig[i];
        csc_free(&M0);
    }

    /* norm_scaling (scaling.jl:28-58) */
    {
        double cte = sqrt(2.0) / 2.0;
        int64_t cont = 0;
        for (int64_t k = 0; k < nsd; ++k) {
            int64_t side = s->sdp_side[k];
            for (int64_t j = 0; j < side; ++j)
                for (int64_t i = 0; i <= j; ++i) {
                    if (i != j) {
                        for (int64_t q = s->A.colptr[cont]; q < s->A.colptr[cont + 1]; ++q) s->A.val[q] *= cte;
                        for (int64_t q = s->G.colptr[cont]; q < s->G.colptr[cont + 1]; ++q) s->G.val[q] *= cte;
                        s->c[cont] *= cte;
                    }
                    cont++;
                }
        }
    }

    /* PrimalDual, AuxiliaryData (structs.jl:83-151) */
    s->x = (double*)calloc((size_t)n + 1, sizeof(double));
    s->x_old = (double*)calloc((size_t)n + 1, sizeof(double));
    s->Mty = (double*)calloc((size_t)n + 1, sizeof(double));
    s->Mty_old = (double*)calloc((size_t)n + 1, sizeof(double));
    s->y = (double*)calloc((size_t)R + 1, sizeof(double));
    s->y_old = (double*)calloc((size_t)R + 1, sizeof(double));
    s->Mx = (double*)calloc((size_t)R + 1, sizeof(double));
    s->Mx_old = (double*)calloc((size_t)R + 1, sizeof(double));
    s->y_half = (double*)calloc((size_t)R + 1, sizeof(double));
    s->y_temp = (double*)calloc((size_t)R + 1, sizeof(double));
    s->mat = (double**)calloc((size_t)nsd + 1, sizeof(double*));
    s->resid = (double**)calloc((size_t)nsd + 1, sizeof(double*));
    s->eig_converged = (int*)calloc((size_t)nsd + 1, sizeof(int));
    s->eig_converged_eigs = (int64_t*)calloc((size_t)nsd + 1, sizeof(int64_t));
    {
        int64_t roff = 0;
        for (int64_t k = 0; k < nsd; ++k) {
            int64_t side = s->sdp_side[k];
            s->mat[k] = (double*)calloc((size_t)side * (size_t)side, sizeof(double));
            s->resid[k] = (double*)calloc((size_t)side, sizeof(double));
            if (prob->eig_resid) memcpy(s->resid[k], prob->eig_resid + roff, sizeof(double) * (size_t)side);
            else oracle_eig_resid(side, opt->eigsolver_resid_seed, opt->krylovkit_resid_init, s->resid[k]);
            roff += side;
        }
    }
    int64_t max_ncv = 2 * opt->max_target_rank_krylov_eigs + 1;
    if (max_ncv < opt->eigsolver_min_lanczos) max_ncv = opt->eigsolver_min_lanczos;
    if (max_ncv < 2 * r0 + 1) max_ncv = 2 * r0 + 1;
    s->eig_w = (double*)calloc((size_t)max_side, sizeof(double));
    s->eig_Z = (double*)calloc((size_t)max_side * (size_t)max_side, sizeof(double));
    s->lan_vals = (double*)calloc((size_t)max_ncv + 1, sizeof(double));
    s->lan_vecs = (double*)calloc((size_t)max_side * (size_t)(max_ncv + 1), sizeof(double));

    int64_t wl = 2 * s->window;
    circ_t* cs[7] = {&s->dual_gap, &s->prim_obj, &s->dual_obj, &s->feasibility,
                     &s->primal_residual, &s->dual_residual, &s->comb_residual};
    for (int q = 0; q < 7; ++q) { cs[q]->l = wl; cs[q]->v = (double*)calloc((size_t)wl, sizeof(double)); }

    /* M, Mt, step sizes (pdhg.jl:104-133) */
    s->M = csc_vstack(&s->A, &s->G);
    s->Mt = csc_transpose(&s->M);
    double spectral_norm = vnorm2(s->M.nnz, s->M.val);       /* approx_norm = true: Frobenius */
    if (shard) { double v = spectral_norm * spectral_norm; red(s, &v, 1, 0); spectral_norm = sqrt(v); }
    if (!opt->approx_norm) {
        /* pdhg.jl:107-118: the largest singular value of M (Arpack.svds, or a dense svd for fewer than two rows /
           columns; the Frobenius norm stays when that fails) */
        double sv = exact_spectral_norm(&s->M, &s->Mt);
        if (shard) {
            /* the shards are independent blocks (no row of one rank touches another rank's variables), so M is block
               diagonal up to a permutation and sigma_max(M) = max over the ranks; a failure on any rank fails all */
            double v[2] = {sv, sv < 0.0 ? 1.0 : 0.0};
            red(s, v, 2, 1);
            sv = v[1] != 0.0 ? -1.0 : v[0];
        }
        if (sv >= 0.0) spectral_norm = sv;
        else fprintf(stderr, "    WARNING: Failed to compute spectral norm of M, shifting to Frobenius norm\n");
    }
    if (spectral_norm < 1e-10) spectral_norm = 1.0;
    s->primal_step = 1.0 / spectral_norm;
    s->primal_step_old = s->primal_step;
    s->dual_step = s->primal_step;

    /* advanced initialisation (pdhg.jl:138-142) */
    if (opt->advanced_initialization) {
        for (int64_t i = 0; i < n; ++i) s->x[i] = s->primal_step * s->c[i];
        csc_mul(&s->M, s->x, s->Mx);
        csc_mul(&s->M, s->x_old, s->Mx_old);
    }
    out->time_setup = now_s() - s->time0;
    out->trace_len = 0;
    double t_loop0 = now_s();

    /* fixed-point loop (pdhg.jl:145-484) */
    int64_t kmax = 2 * opt->max_iter_local;
    for (int64_t k = 1; k <= kmax; ++k) {
        s->iter = k;
        s->trace_mv0 = s->lanczos_matvecs; s->trace_ls0 = s->linesearch_trials;
        primal_step(s, opt);
        if (opt->line_search_flag) linesearch(s, opt); else dual_step(s);
        compute_residual(s);
        compute_gap(s);

        /* (the log itself, printing.jl, is not restated here: only its effect on dual_feasibility_check) */
        if ((opt->check_dual_feas && (k % opt->check_dual_feas_freq) == 0) ||
            (opt->log_verbose && opt->log_freq > 0 && (k % opt->log_freq) == 0 && opt->extended_log2)) {       /* pdhg.jl:166-173 */
            double f = s->stop_reason == 6 ? 0.0 : 1.0;
            double* cc = (double*)malloc(sizeof(double) * ((size_t)n + 1));
            for (int64_t i = 0; i < n; ++i) cc[i] = f * s->c_orig[i];
            s->dual_feasibility = dual_feas_y(s, s->y, cc);
            free(cc);
            s->dual_feasibility_check = 1;
        } else {
            s->dual_feasibility_check = 0;
        }
        record_trace(s, out, opt->trace_cap);

        if (s->iter < s->certificate_search_min_iter) continue;                    /* pdhg.jl:180-182 */

        if (opt->certificate_search && s->certificate_search) {                    /* pdhg.jl:184-244 */
            if (s->stop_reason == 6) {
                if (circ_get(&s->dual_obj, k) > +opt->certificate_obj_tol) {
                    double* cc = (double*)calloc((size_t)n + 1, sizeof(double));
                    s->dual_feasibility = dual_feas_y(s, s->y, cc);
                    free(cc);
                    s->dual_feasibility_check = 1;
                    if (s->dual_feasibility < opt->tol_feasibility_dual) {
                        s->certificate_found = 1;
                        strncat(s->stop_reason_string, " [Dual ray found]",
                                PROXSDP_STATUS_STRING_LEN - strlen(s->stop_reason_string) - 1);
                        break;
                    }
                }
            } else {
                if (circ_get(&s->prim_obj, k) < -opt->certificate_obj_tol) {
                    if (circ_get(&s->feasibility, s->iter) < opt->tol_feasibility) {
                        s->certificate_found = 1;
                        strncat(s->stop_reason_string, " [Primal ray found]",
                                PROXSDP_STATUS_STRING_LEN - strlen(s->stop_reason_string) - 1);
                        break;
                    }
                }
            }
            double cr = circ_get(&s->comb_residual, k);
            if ((circ_get(&s->prim_obj, k) < -opt->certificate_fail_tol &&
                 circ_get(&s->dual_obj, k) < -opt->certificate_fail_tol &&
                 circ_get(&s->feasibility, s->iter) < -opt->certificate_fail_tol) || cr != cr) {
                strncat(s->stop_reason_string, " [Failed to find certificate]",
                        PROXSDP_STATUS_STRING_LEN - strlen(s->stop_reason_string) - 1);
                break;
            }
        }

        /* convergence check (pdhg.jl:247-332) */
        s->rank_update += 1;
        double gap_k = circ_get(&s->dual_gap, s->iter), feas_k = circ_get(&s->feasibility, s->iter);
        double pr_k = circ_get(&s->primal_residual, k), dr_k = circ_get(&s->dual_residual, k);
        if (gap_k <= opt->tol_gap && feas_k <= opt->tol_feasibility &&
            (!opt->check_dual_feas || s->dual_feasibility < opt->tol_feasibility_dual)) {
            if (convergedrank(s, opt) && soc_convergence(s, opt) && s->iter > opt->min_iter) {
                if (!s->certificate_search) {
                    s->stop_reason = 1;
                    snprintf(s->stop_reason_string, PROXSDP_STATUS_STRING_LEN, "Optimal solution found");
                } else {
                    strncat(s->stop_reason_string, " [Failed to find certificate - type 2]",
                            PROXSDP_STATUS_STRING_LEN - strlen(s->stop_reason_string) - 1);
                    break;
                }
                break;
            } else if (s->rank_update > s->window) {
                s->update_cont += 1;
                if (s->update_cont > 0) {
                    for (int64_t idx = 0; idx < nsd; ++idx) rank_increment_rule(s, opt, idx);
                    s->rank_update = 0; s->update_cont = 0;
                }
            }
        } else if (k > s->window && circ_get(&s->comb_residual, k - s->window) < circ_get(&s->comb_residual, k) &&
                   s->rank_update > s->window) {
            s->update_cont += 1;
            if (s->update_cont > opt->divergence_min_update) {
                double any_room = 0.0;
                for (int64_t idx = 0; idx < nsd; ++idx) {
                    if (s->target_rank[idx] < s->sdp_side[idx]) any_room = 1.0;
                    rank_increment_rule(s, opt, idx);
                }
                if (shard) red(s, &any_room, 1, 1);
                if (any_room != 0.0) { s->rank_update = 0; s->update_cont = 0; }
            }
        } else if (pr_k > opt->tol_primal && dr_k < opt->tol_dual && k > s->window) {
            ada_count += 1;
            if (ada_count > opt->adapt_window) {
                ada_count = 0;
                if (opt->line_search_flag) {
                    s->beta *= (1.0 - s->adapt_level);
                    s->primal_step /= sqrt(1.0 - s->adapt_level);
                } else {
                    s->primal_step /= (1.0 - s->adapt_level);
                    s->dual_step *= (1.0 - s->adapt_level);
                }
                s->adapt_level *= opt->adapt_decay;
            }
        } else if (pr_k < opt->tol_primal && dr_k > opt->tol_dual && k > s->window) {
            ada_count += 1;
            if (ada_count > opt->adapt_window) {
                ada_count = 0;
                if (opt->line_search_flag) {
                    s->beta /= (1.0 - s->adapt_level);
                    s->primal_step *= sqrt(1.0 - s->adapt_level);
                } else {
                    s->primal_step *= (1.0 - s->adapt_level);
                    s->dual_step /= (1.0 - s->adapt_level);
                }
                s->adapt_level *= opt->adapt_decay;
            }
        }

        /* max_iter or time limit (pdhg.jl:335-382) */
        double elapsed_k = now_s() - s->time0;
        if (shard) red(s, &elapsed_k, 1, 1);      /* all ranks must leave the loop together */
        if (s->iter >= opt->max_iter_local || elapsed_k >= opt->time_limit) {
            if (s->iter > opt->min_iter_time_infeas &&
                circ_max_abs_diff(&s->dual_gap) < opt->infeas_stable_gap_tol &&
                circ_get(&s->dual_gap, k) > opt->infeas_limit_gap_tol) {
                if (circ_get(&s->feasibility, s->iter) <= opt->tol_feasibility / 100) {
                    s->stop_reason = 5;
                    snprintf(s->stop_reason_string, PROXSDP_STATUS_STRING_LEN,
                             "Problem declared unbounded due to lack of improvement");
                    if (opt->certificate_search && !s->certificate_search) {
                        certificate_dual_infeasibility(s, opt);
                        cache_solution(s, opt, s->c_orig, out); have_cached = 1;
                    } else if (opt->certificate_search && s->certificate_search) {
                    } else {
                        break;
                    }
                } else if (circ_get(&s->feasibility, s->iter) > opt->infeas_feasibility_tol) {
                    s->stop_reason = 6;
                    snprintf(s->stop_reason_string, PROXSDP_STATUS_STRING_LEN,
                             "Problem declared infeasible due to lack of improvement");
                    if (opt->certificate_search && !s->certificate_search) {
                        certificate_infeasibility(s, opt);
                        cache_solution(s, opt, s->c_orig, out); have_cached = 1;
                    } else if (opt->certificate_search && s->certificate_search) {
                    } else {
                        break;
                    }
                }
            } else if (s->iter >= opt->max_iter_local) {
                s->stop_reason = 3;
                snprintf(s->stop_reason_string, PROXSDP_STATUS_STRING_LEN,
                         "Iteration limit of %lld was hit", (long long)opt->max_iter_local);
            } else {
                s->stop_reason = 2;
                snprintf(s->stop_reason_string, PROXSDP_STATUS_STRING_LEN,
                         "Time limit hit, limit: %g time: %g", opt->time_limit, now_s() - s->time0);
            }
            if (s->iter >= opt->max_iter_local || elapsed_k >= opt->time_limit) break;
        }

        if (opt->certificate_search && s->certificate_search) continue;            /* pdhg.jl:385-387 */

        double dobj_k = circ_get(&s->dual_obj, k), pobj_k = circ_get(&s->prim_obj, k);
        /* dual objective growing too much (pdhg.jl:390-405) */
        if ((s->iter > opt->min_iter_max_obj && dobj_k > opt->max_obj) || dobj_k != dobj_k) {
            s->stop_reason = 6;
            snprintf(s->stop_reason_string, PROXSDP_STATUS_STRING_LEN,
                     "Infeasible: |Dual objective| = %g > maximum allowed = %g", dobj_k, opt->max_obj);
            if (opt->certificate_search && !s->certificate_search) {
                certificate_infeasibility(s, opt);
                cache_solution(s, opt, s->c_orig, out); have_cached = 1;
            } else {
                break;
            }
        }
        /* primal objective growing too much (pdhg.jl:408-422) */
        if ((s->iter > opt->min_iter_max_obj && pobj_k < -opt->max_obj) || pobj_k != pobj_k) {
            s->stop_reason = 5;
            snprintf(s->stop_reason_string, PROXSDP_STATUS_STRING_LEN,
                     "Unbounded: |Primal objective| = %g > maximum allowed = %g", pobj_k, opt->max_obj);
            if (opt->certificate_search && !s->certificate_search) {
                certificate_dual_infeasibility(s, opt);
                cache_solution(s, opt, s->c_orig, out); have_cached = 1;
            } else {
                break;
            }
        }
        /* stalled feasibility with meaningful gap (pdhg.jl:425-444) */
        if (s->iter > opt->min_iter_max_obj &&
            circ_get(&s->dual_gap, k) > opt->infeas_limit_gap_tol &&
            circ_get(&s->feasibility, s->iter) > opt->infeas_feasibility_tol &&
            circ_max_abs_diff(&s->feasibility) < opt->infeas_stable_feasibility_tol) {
            s->stop_reason = 6;
            snprintf(s->stop_reason_string, PROXSDP_STATUS_STRING_LEN,
                     "Infeasible: feasibility stalled at %g", circ_get(&s->feasibility, s->iter));
            if (opt->certificate_search && !s->certificate_search) {
                certificate_infeasibility(s, opt);
                cache_solution(s, opt, s->c_orig, out); have_cached = 1;
            } else {
                break;
            }
        }
        /* stalled gap at 100 % (pdhg.jl:447-483) */
        if (s->iter > opt->min_iter_max_obj &&
            circ_get(&s->dual_gap, k) > 1 - opt->infeas_gap_tol &&
            circ_max_abs_diff(&s->dual_gap) < opt->infeas_stable_gap_tol) {
            if (fabs(dobj_k) > fabs(pobj_k) && circ_get(&s->feasibility, s->iter) > opt->infeas_feasibility_tol) {
                s->stop_reason = 6;
                snprintf(s->stop_reason_string, PROXSDP_STATUS_STRING_LEN,
                         "Infeasible: duality gap stalled at 100 %% with |Dual objective| >> |Primal objective|");
                if (opt->certificate_search && !s->certificate_search) {
                    certificate_infeasibility(s, opt);
                    cache_solution(s, opt, s->c_orig, out); have_cached = 1;
                } else {
                    break;
                }
            } else if (fabs(pobj_k) > fabs(dobj_k) && circ_get(&s->feasibility, s->iter) <= opt->tol_feasibility) {
                s->stop_reason = 5;
                snprintf(s->stop_reason_string, PROXSDP_STATUS_STRING_LEN,
                         "Unbounded: duality gap stalled at 100 %% with |Dual objective| << |Primal objective|");
                if (opt->certificate_search && !s->certificate_search) {
                    certificate_dual_infeasibility(s, opt);
                    cache_solution(s, opt, s->c_orig, out); have_cached = 1;
                } else {
                    break;
                }
            }
        }
    }
    out->time_loop = now_s() - t_loop0;

    /* results (pdhg.jl:486-529) */
    if (opt->certificate_search && s->certificate_search) {
        if (s->certificate_found) {
            if (s->stop_reason == 6) for (int64_t i = 0; i < n; ++i) s->c_orig[i] *= 0.0;
            cache_solution(s, opt, s->c_orig, out);
        } else if (!have_cached) {
            cache_solution(s, opt, s->c_orig, out);
        }
    } else {
        cache_solution(s, opt, s->c_orig, out);
    }
    out->time_psd_proj = s->time_psd;
    out->n_psd_proj = s->n_psd;
    out->lanczos_matvecs = s->lanczos_matvecs;
    out->lanczos_calls = s->lanczos_calls;
    out->full_eig_calls = s->full_eig_calls;
    out->linesearch_trials = s->linesearch_trials;
    out->gpu_launches = 0;
    out->time_lanczos = 0.0; out->time_rest = 0.0; out->time_l2_flush = 0.0;
    out->lanczos_timed_calls = 0; out->h2d_bytes = 0; out->d2h_bytes = 0;
    if (out->target_rank) for (int64_t k = 0; k < nsd; ++k) out->target_rank[k] = s->target_rank[k];

    /* free */
    csc_free(&s->A); csc_free(&s->G); csc_free(&s->A_orig); csc_free(&s->G_orig); csc_free(&s->M); csc_free(&s->Mt);
    free(s->b); free(s->h); free(s->c); free(s->b_orig); free(s->h_orig); free(s->c_orig); free(s->eq_E); free(s->eq_D);
    free(s->sdp_side); free(s->sdp_off); free(s->soc_off); free(s->soc_len); free(s->var_ordering);
    free(s->x); free(s->x_old); free(s->y); free(s->y_old); free(s->Mty); free(s->Mty_old);
    free(s->Mx); free(s->Mx_old); free(s->y_half); free(s->y_temp);
    for (int64_t k = 0; k < nsd; ++k) { free(s->mat[k]); free(s->resid[k]); }
    free(s->mat); free(s->resid); free(s->eig_converged); free(s->eig_converged_eigs);
    free(s->eig_w); free(s->eig_Z); free(s->lan_vals); free(s->lan_vecs);
    free(s->target_rank); free(s->current_rank); free(s->min_eig);
    for (int q = 0; q < 7; ++q) free(cs[q]->v);
    free(ord); free(used);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* step-level seams for kernel-parity tests                                  */
/* ------------------------------------------------------------------------- */

/* One psd_projection! (prox_operators.jl:33-66) over concatenated svec blocks.
 * mode: 0 = reference dispatch (Krylov if eligible), 1 = force full eig.
 * In/out x (sum tri_len); in target_rank; out current_rank, min_eig, converged
 * (KrylovKit info.converged, -1 when the full path ran), numops. */
int proxsdp_oracle_psd_project(int64_t n_sdp, const int64_t* sides, double* x,
                               const int64_t* target_rank, const proxsdp_options_t* opt,
                               int64_t iter, int64_t mode, const double* resid,
                               int64_t* current_rank, double* min_eig,
                               int64_t* converged, int64_t* numops) {
    state_t S; memset(&S, 0, sizeof(S));
    state_t* s = &S;
    opts_t o = *opt;
    if (mode == 1) o.full_eig_decomp = 1;
    s->n_sdp = n_sdp; s->iter = iter;
    s->sdp_side = (int64_t*)sides;
    s->target_rank = (int64_t*)target_rank;
    s->current_rank = current_rank;
    s->min_eig = min_eig;
    s->mat = (double**)calloc((size_t)n_sdp + 1, sizeof(double*));
    s->resid = (double**)calloc((size_t)n_sdp + 1, sizeof(double*));
    s->eig_converged = (int*)calloc((size_t)n_sdp + 1, sizeof(int));
    s->eig_converged_eigs = (int64_t*)calloc((size_t)n_sdp + 1, sizeof(int64_t));
    int64_t max_side = 1, roff = 0, max_tr = 2;
    for (int64_t k = 0; k < n_sdp; ++k) {
        int64_t side = sides[k];
        if (side > max_side) max_side = side;
        if (target_rank[k] > max_tr) max_tr = target_rank[k];
        s->mat[k] = (double*)calloc((size_t)side * (size_t)side, sizeof(double));
        s->resid[k] = (double*)calloc((size_t)side, sizeof(double));
        if (resid) memcpy(s->resid[k], resid + roff, sizeof(double) * (size_t)side);
        else oracle_eig_resid(side, o.eigsolver_resid_seed, o.krylovkit_resid_init, s->resid[k]);
        roff += side;
        s->eig_converged_eigs[k] = -1;
    }
    int64_t max_ncv = 2 * max_tr + 1;
    if (max_ncv < o.eigsolver_min_lanczos) max_ncv = o.eigsolver_min_lanczos;
    s->eig_w = (double*)calloc((size_t)max_side, sizeof(double));
    s->eig_Z = (double*)calloc((size_t)max_side * (size_t)max_side, sizeof(double));
    s->lan_vals = (double*)calloc((size_t)max_ncv + 1, sizeof(double));
    s->lan_vecs = (double*)calloc((size_t)max_side * (size_t)(max_ncv + 1), sizeof(double));
    psd_projection(s, &o, x);
    for (int64_t k = 0; k < n_sdp; ++k) if (converged) converged[k] = s->eig_converged_eigs[k];
    if (numops) *numops = s->lanczos_matvecs;
    for (int64_t k = 0; k < n_sdp; ++k) { free(s->mat[k]); free(s->resid[k]); }
    free(s->mat); free(s->resid); free(s->eig_converged); free(s->eig_converged_eigs);
    free(s->eig_w); free(s->eig_Z); free(s->lan_vals); free(s->lan_vecs);
    return 0;
}

/* helpers of the two step seams */
static double* dup_vec(const double* v, int64_t n) {
    double* d = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    if (v && n > 0) memcpy(d, v, sizeof(double) * (size_t)n);
    return d;
}
static void circ_init1(circ_t* c) { c->l = 2; c->v = (double*)calloc(2, sizeof(double)); }

/* linesearch! / dual_step! (pdhg.jl:532-609) on explicit state: the counterpart of proxsdp_b200_dual_step.
 * rows: A (p x n) and G (m x n) of the working problem, M = [A; G] taken as is. */
int proxsdp_oracle_dual_step(const proxsdp_problem_t* rows, const proxsdp_options_t* opt, const proxsdp_step_state_t* st,
                             double* y_new, double* Mty_new, double* scalars_out, int64_t* trials) {
    state_t S; memset(&S, 0, sizeof(S));
    state_t* s = &S;
    s->n = st->n; s->p = st->p; s->m = st->m; s->R = st->p + st->m;
    int64_t n = s->n, R = s->R;
    s->A = csc_from_input(s->p, n, rows->A_colptr, rows->A_rowval, rows->A_nzval, rows->index_base);
    s->G = csc_from_input(s->m, n, rows->G_colptr, rows->G_rowval, rows->G_nzval, rows->index_base);
    s->M = csc_vstack(&s->A, &s->G);
    s->Mt = csc_transpose(&s->M);
    s->b = dup_vec(st->b, s->p); s->h = dup_vec(st->h, s->m);
    s->y = dup_vec(st->y, R); s->y_old = dup_vec(st->y, R);          /* pair.y_old == pair.y when linesearch! runs */
    s->Mx = dup_vec(st->Mx, R); s->Mx_old = dup_vec(st->Mx_old, R);
    s->Mty = dup_vec(NULL, n); s->Mty_old = dup_vec(st->Mty, n);
    s->y_half = dup_vec(NULL, R); s->y_temp = dup_vec(NULL, R);
    s->primal_step = st->primal_step; s->primal_step_old = st->primal_step_old; s->theta = st->theta;
    s->beta = st->beta; s->dual_step = st->dual_step;
    if (opt->line_search_flag) linesearch(s, opt); else { s->linesearch_trials = 1; dual_step(s); }
    memcpy(y_new, s->y, sizeof(double) * (size_t)R);
    memcpy(Mty_new, s->Mty, sizeof(double) * (size_t)n);
    if (scalars_out) { scalars_out[0] = s->primal_step; scalars_out[1] = s->theta; scalars_out[2] = s->dual_step; scalars_out[3] = s->primal_step_old; }
    if (trials) *trials = s->linesearch_trials;
    csc_free(&s->A); csc_free(&s->G); csc_free(&s->M); csc_free(&s->Mt);
    free(s->b); free(s->h); free(s->y); free(s->y_old); free(s->Mx); free(s->Mx_old); free(s->Mty); free(s->Mty_old);
    free(s->y_half); free(s->y_temp);
    return 0;
}

/* compute_residual! + compute_gap! (residuals.jl:2-71) on explicit state: the counterpart of proxsdp_b200_residuals.
 * out[8] = primal_residual, dual_residual, comb_residual, equa_feasibility, ineq_feasibility, prim_obj, dual_obj, gap */
int proxsdp_oracle_residuals(const proxsdp_options_t* opt, const proxsdp_step_state_t* st, double* out) {
    (void)opt;
    state_t S; memset(&S, 0, sizeof(S));
    state_t* s = &S;
    s->n = st->n; s->p = st->p; s->m = st->m; s->R = st->p + st->m;
    int64_t n = s->n, R = s->R;
    s->b = dup_vec(st->b, s->p); s->h = dup_vec(st->h, s->m); s->c = dup_vec(st->c, n);
    s->x = dup_vec(st->x, n); s->x_old = dup_vec(st->x_old, n);
    s->y = dup_vec(st->y, R); s->y_old = dup_vec(st->y_old, R);
    s->Mx = dup_vec(st->Mx, R); s->Mx_old = dup_vec(st->Mx_old, R);
    s->Mty = dup_vec(st->Mty, n); s->Mty_old = dup_vec(st->Mty_old, n);
    s->primal_step = st->primal_step; s->dual_step = st->dual_step; s->beta = st->beta;
    s->norm_b = st->norm_b; s->norm_h = st->norm_h; s->norm_c = st->norm_c;
    s->iter = 1;
    circ_init1(&s->dual_gap); circ_init1(&s->prim_obj); circ_init1(&s->dual_obj); circ_init1(&s->feasibility);
    circ_init1(&s->primal_residual); circ_init1(&s->dual_residual); circ_init1(&s->comb_residual);
    compute_residual(s);
    /* compute_residual! has copied the new iterates over the old ones; compute_gap! reads only the new ones */
    compute_gap(s);
    out[0] = circ_get(&s->primal_residual, 1); out[1] = circ_get(&s->dual_residual, 1); out[2] = circ_get(&s->comb_residual, 1);
    out[3] = s->equa_feasibility; out[4] = s->ineq_feasibility;
    out[5] = circ_get(&s->prim_obj, 1); out[6] = circ_get(&s->dual_obj, 1); out[7] = circ_get(&s->dual_gap, 1);
    free(s->b); free(s->h); free(s->c); free(s->x); free(s->x_old); free(s->y); free(s->y_old); free(s->Mx); free(s->Mx_old);
    free(s->Mty); free(s->Mty_old);
    free(s->dual_gap.v); free(s->prim_obj.v); free(s->dual_obj.v); free(s->feasibility.v);
    free(s->primal_residual.v); free(s->dual_residual.v); free(s->comb_residual.v);
    return 0;
}

/* soc_projection! over concatenated SOC blocks (prox_operators.jl:138-158) */
int proxsdp_oracle_soc_project(int64_t n_soc, const int64_t* lens, double* x) {
    state_t S; memset(&S, 0, sizeof(S));
    S.n_soc = n_soc;
    S.soc_len = (int64_t*)lens;
    S.soc_off = (int64_t*)calloc((size_t)n_soc + 1, sizeof(int64_t));
    int64_t off = 0;
    for (int64_t k = 0; k < n_soc; ++k) { S.soc_off[k] = off; off += lens[k]; }
    soc_projection(&S, x);
    free(S.soc_off);
    return 0;
}

int proxsdp_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void proxsdp_oracle_set_num_threads(int nt) {
#ifdef _OPENMP
    omp_set_num_threads(nt);
#else
    (void)nt;
#endif
}
