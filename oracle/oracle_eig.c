/*
 * oracle_eig.c — CPU eigen-solvers used by the ProxSDP oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path;
 * it is the checker the CUDA path is compared against (tests/, smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py).
 *
 * Two routines, each a restatement of an algorithm the reference reaches through
 * a dependency that is NOT under /root/reference:
 *
 *  oracle_eigh     <-> LinearAlgebra.eigen!(Symmetric)  (LAPACK dsyevr; reference
 *                      src/prox_operators.jl:113, src/pdhg.jl:685).  Restated as
 *                      Householder tridiagonalisation + implicit-shift QL
 *                      (EISPACK tred2/tql2 family).  Eigenvalues ascending.
 *
 *  oracle_lanczos  <-> KrylovKit.eigsolve(A, x0, howmany, :LR,
 *                      Lanczos(KrylovDefaults.orth, krylovdim, maxiter, tol, eager=false))
 *                      (reference src/eigsolver.jl:802-812; KrylovKit.jl compat
 *                      0.5.2-0.9 per reference Project.toml:21, exact version unpinned —
 *                      no Manifest.toml).  Restated from the published algorithm:
 *                      Lanczos with full (two-pass Gram-Schmidt) re-orthogonalisation,
 *                      Krylov-Schur thick restart with keep = (3K + 2*converged)/5,
 *                      convergence |beta * U[K,i]| <= tol on the leading Ritz pairs
 *                      sorted by largest real part.  PARITY UNPINNED at the trajectory
 *                      level: only converged quantities are compared.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* small helpers                                                             */
/* ------------------------------------------------------------------------- */
static double dot_(int64_t n, const double* a, const double* b) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}
static double nrm2_(int64_t n, const double* a) { return sqrt(dot_(n, a, a)); }
static void axpy_(int64_t n, double alpha, const double* x, double* y) {
    for (int64_t i = 0; i < n; ++i) y[i] += alpha * x[i];
}

/* y = A x, A symmetric n x n column-major, only the UPPER triangle is read
 * (BLAS dsymv 'U' semantics; reference eigsolver.jl:802 via Symmetric*Vector,
 * multi-threaded like OpenBLAS would be).  scratch: nthreads*n doubles or NULL. */
void oracle_symv_upper(int64_t n, const double* A, const double* x, double* y) {
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
    if (n < 256) nt = 1;
#endif
    if (nt == 1) {
        for (int64_t i = 0; i < n; ++i) y[i] = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            const double* col = A + j * n;
            double xj = x[j], t = 0.0;
            for (int64_t i = 0; i < j; ++i) {
                y[i] += col[i] * xj;
                t += col[i] * x[i];
            }
            y[j] += col[j] * xj + t;
        }
        return;
    }
#ifdef _OPENMP
    double* priv = (double*)calloc((size_t)nt * (size_t)n, sizeof(double));
    #pragma omp parallel num_threads(nt)
    {
        int t_id = omp_get_thread_num();
        double* yp = priv + (size_t)t_id * n;
        /* cyclic column distribution balances the triangle */
        #pragma omp for schedule(static, 8)
        for (int64_t j = 0; j < n; ++j) {
            const double* col = A + j * n;
            double xj = x[j], t = 0.0;
            for (int64_t i = 0; i < j; ++i) {
                yp[i] += col[i] * xj;
                t += col[i] * x[i];
            }
            yp[j] += col[j] * xj + t;
        }
        #pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            double s = 0.0;
            for (int t = 0; t < nt; ++t) s += priv[(size_t)t * n + i];
            y[i] = s;
        }
    }
    free(priv);
#endif
}

/* ------------------------------------------------------------------------- */
/* full symmetric eigendecomposition                                         */
/* ------------------------------------------------------------------------- */

/* implicit-shift QL on a symmetric tridiagonal (d: diagonal n, e: sub-diagonal
 * in e[0..n-2], e[n-1] scratch).  Z (n x n column-major, ldz = n) is multiplied
 * from the right by the accumulated rotations; pass Z = Q from the reduction, or
 * identity for a bare tridiagonal.  Returns 0, or k>0 if no convergence at row k. */
static int tridiag_ql(int64_t n, double* d, double* e, double* Z, int64_t nrowz) {
    if (n <= 1) return 0;
    e[n - 1] = 0.0;
    for (int64_t l = 0; l < n; ++l) {
        int iter = 0;
        int64_t m;
        do {
            for (m = l; m < n - 1; ++m) {
                double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 120) return (int)(l + 1);
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int64_t i;
                int underflow = 0;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i];
                    double b = c * e[i];
                    r = hypot(f, g);
                    e[i + 1] = r;
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        underflow = 1;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    if (Z) {
                        double* zi = Z + i * nrowz;
                        double* zi1 = Z + (i + 1) * nrowz;
                        for (int64_t k = 0; k < nrowz; ++k) {
                            double fz = zi1[k];
                            zi1[k] = s * zi[k] + c * fz;
                            zi[k] = c * zi[k] - s * fz;
                        }
                    }
                }
                if (underflow) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    return 0;
}

/* sort eigenpairs ascending (selection sort on columns; n is small or the
 * O(n^2) swaps are negligible next to the O(n^3) reduction) */
static void sort_eig_ascending(int64_t n, double* d, double* Z, int64_t nrowz) {
    for (int64_t i = 0; i < n - 1; ++i) {
        int64_t k = i;
        double p = d[i];
        for (int64_t j = i + 1; j < n; ++j)
            if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i];
            d[i] = p;
            if (Z) {
                double* a = Z + i * nrowz;
                double* b = Z + k * nrowz;
                for (int64_t r = 0; r < nrowz; ++r) { double t = a[r]; a[r] = b[r]; b[r] = t; }
            }
        }
    }
}

/* A: n x n column-major symmetric; only the UPPER triangle is read on input
 * (like eigen!(Symmetric(.,:U))); A is destroyed.  w: eigenvalues ascending,
 * Z: n x n column-major eigenvectors.  Returns 0 on success. */
int oracle_eigh(int64_t n, double* A, double* w, double* Z) {
    if (n <= 0) return 0;
    if (n == 1) { w[0] = A[0]; Z[0] = 1.0; return 0; }
    /* mirror the upper triangle so the reduction can use whole columns */
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i < j; ++i) A[j + i * n] = A[i + j * n];

    double* e = (double*)calloc((size_t)n, sizeof(double));
    double* vstore = (double*)calloc((size_t)n * (size_t)n, sizeof(double)); /* reflectors */
    double* tau = (double*)calloc((size_t)n, sizeof(double));
    double* pvec = (double*)calloc((size_t)n, sizeof(double));

    /* Householder reduction to tridiagonal: for k = 0..n-3 annihilate A[k+2:n, k] */
    for (int64_t k = 0; k < n - 2; ++k) {
        int64_t len = n - k - 1;            /* length of x = A[k+1:n, k] */
        double* x = A + (k + 1) + k * n;
        double* v = vstore + k * n;         /* v has len entries */
        double xnorm = nrm2_(len, x);
        if (xnorm == 0.0) { e[k] = 0.0; tau[k] = 0.0; continue; }
        double alpha = (x[0] >= 0.0) ? -xnorm : xnorm;
        for (int64_t i = 0; i < len; ++i) v[i] = x[i];
        v[0] -= alpha;
        double vtv = dot_(len, v, v);
        if (vtv == 0.0) { e[k] = alpha; tau[k] = 0.0; continue; }
        double beta = 2.0 / vtv;
        tau[k] = beta;
        e[k] = alpha;
        /* p = beta * A22 v ; A22 = A[k+1:n, k+1:n] */
        #pragma omp parallel for schedule(static) if (len > 128)
        for (int64_t j = 0; j < len; ++j) {
            const double* col = A + (k + 1) + (k + 1 + j) * n;
            pvec[j] = beta * dot_(len, col, v);
        }
        double K = 0.5 * beta * dot_(len, v, pvec);
        for (int64_t i = 0; i < len; ++i) pvec[i] -= K * v[i];   /* q */
        /* A22 -= v q' + q v' */
        #pragma omp parallel for schedule(static) if (len > 128)
        for (int64_t j = 0; j < len; ++j) {
            double* col = A + (k + 1) + (k + 1 + j) * n;
            double qj = pvec[j], vj = v[j];
            for (int64_t i = 0; i < len; ++i) col[i] -= v[i] * qj + pvec[i] * vj;
        }
    }
    for (int64_t i = 0; i < n; ++i) w[i] = A[i + i * n];
    e[n - 2] = A[(n - 1) + (n - 2) * n];

    /* Q = H_0 H_1 ... H_{n-3}; accumulate backwards into Z */
    memset(Z, 0, sizeof(double) * (size_t)n * (size_t)n);
    for (int64_t i = 0; i < n; ++i) Z[i + i * n] = 1.0;
    for (int64_t k = n - 3; k >= 0; --k) {
        if (tau[k] == 0.0) continue;
        int64_t len = n - k - 1;
        const double* v = vstore + k * n;
        double beta = tau[k];
        /* Z[k+1:n, k+1:n] -= beta v (v' Z[k+1:n, k+1:n]) */
        #pragma omp parallel for schedule(static) if (len > 128)
        for (int64_t j = 0; j < len; ++j) {
            double* col = Z + (k + 1) + (k + 1 + j) * n;
            double s = beta * dot_(len, v, col);
            for (int64_t i = 0; i < len; ++i) col[i] -= s * v[i];
        }
    }
    int rc = tridiag_ql(n, w, e, Z, n);
    sort_eig_ascending(n, w, Z, n);
    free(e); free(vstore); free(tau); free(pvec);
    return rc;
}

/* eigen-decomposition of a small dense symmetric matrix (K x K, column-major,
 * full storage), same algorithm; used for the Rayleigh quotient. */
static int small_eigh(int64_t K, double* T, double* D, double* U) {
    return oracle_eigh(K, T, D, U);
}

/* ------------------------------------------------------------------------- */
/* KrylovKit-style Lanczos eigsolve (:LR)                                    */
/* ------------------------------------------------------------------------- */
typedef void (*oracle_matvec_fn)(int64_t n, const double* A, const double* x, double* y);

/* Outputs: vals (capacity krylovdim), vecs (n x krylovdim column-major capacity),
 * *nvals_out = max(howmany', converged) as KrylovKit returns,
 * *converged_out = info.converged, *numops_out = info.numops, *numiter_out.
 * A: n x n column-major, upper triangle meaningful.  Returns 0. */
/* Lanczos(...; eager = true): the Ritz analysis also runs after every expansion step once K >= howmany, and the
 * eigsolve stops as soon as `howmany` pairs have converged (KrylovKit eigsolve, lanczos.jl).  Set by the caller
 * (reference src/eigsolver.jl:809: opt.krylovkit_eager). */
int oracle_lanczos_eager = 0;

int oracle_lanczos(int64_t n, const double* A, const double* x0, int64_t howmany,
                   int64_t krylovdim, int64_t maxiter, double tol,
                   double* vals, double* vecs, int64_t* nvals_out,
                   int64_t* converged_out, int64_t* numops_out, int64_t* numiter_out) {
    if (krylovdim > n) {
        /* KrylovKit itself would just hit an invariant subspace; capacity-wise the
         * basis can never exceed n vectors, keep the nominal krylovdim for the
         * restart arithmetic but allocate for it. */
    }
    int64_t K = krylovdim;
    double* V = (double*)calloc((size_t)n * (size_t)(K + 1), sizeof(double));
    double* r = (double*)calloc((size_t)n, sizeof(double));
    double* alphas = (double*)calloc((size_t)K + 1, sizeof(double));
    double* betas = (double*)calloc((size_t)K + 1, sizeof(double));
    double* T = (double*)calloc((size_t)K * (size_t)K, sizeof(double));
    double* U = (double*)calloc((size_t)K * (size_t)K, sizeof(double));
    double* D = (double*)calloc((size_t)K, sizeof(double));
    double* f = (double*)calloc((size_t)K, sizeof(double));
    double* S = (double*)calloc((size_t)K * (size_t)K, sizeof(double));
    double* hv = (double*)calloc((size_t)K, sizeof(double));
    double* tmpK = (double*)calloc((size_t)K, sizeof(double));
    double* Vnew = (double*)calloc((size_t)n * (size_t)(K + 1), sizeof(double));

    /* initialize (KrylovKit LanczosIterator initialize) */
    double beta0 = nrm2_(n, x0);
    double* v0 = V;
    for (int64_t i = 0; i < n; ++i) v0[i] = x0[i] / beta0;
    oracle_symv_upper(n, A, v0, r);
    double alpha = dot_(n, v0, r);
    axpy_(n, -alpha, v0, r);
    double beta = nrm2_(n, r);
    {   /* second Gram-Schmidt pass (…GramSchmidt2) */
        double da = dot_(n, v0, r);
        alpha += da;
        axpy_(n, -da, v0, r);
        beta = nrm2_(n, r);
    }
    alphas[0] = alpha;
    betas[0] = beta;
    int64_t k = 1;
    int64_t numops = 1, numiter = 1, converged = 0;
    int64_t Kcur = 1;

    for (;;) {
        beta = betas[k - 1];
        Kcur = k;
        if (beta <= tol && Kcur < howmany) howmany = Kcur;
        if (Kcur == K || beta <= tol || (oracle_lanczos_eager && Kcur >= howmany)) {
            if (Kcur == 1) {
                D[0] = alphas[0];
                U[0] = 1.0;
                f[0] = beta;
                converged = (beta <= tol) ? 1 : 0;
            } else {
                memset(T, 0, sizeof(double) * (size_t)Kcur * (size_t)Kcur);
                for (int64_t i = 0; i < Kcur; ++i) T[i + i * Kcur] = alphas[i];
                for (int64_t i = 0; i + 1 < Kcur; ++i) {
                    T[i + (i + 1) * Kcur] = betas[i];
                    T[(i + 1) + i * Kcur] = betas[i];
                }
                small_eigh(Kcur, T, D, U);     /* ascending, U is Kcur x Kcur */
                /* :LR => sort descending: reverse */
                for (int64_t i = 0; i < Kcur / 2; ++i) {
                    int64_t j = Kcur - 1 - i;
                    double t = D[i]; D[i] = D[j]; D[j] = t;
                    double* a = U + i * Kcur;
                    double* b = U + j * Kcur;
                    for (int64_t q = 0; q < Kcur; ++q) { double tt = a[q]; a[q] = b[q]; b[q] = tt; }
                }
                for (int64_t i = 0; i < Kcur; ++i) f[i] = beta * U[(Kcur - 1) + i * Kcur];
                converged = 0;
                while (converged < Kcur && fabs(f[converged]) <= tol) converged++;
            }
            if (converged >= howmany) break;
        }
        if (Kcur < K) {
            /* expand!: V[k] = r/beta; lanczosrecurrence */
            double* vk = V + (size_t)k * n;
            for (int64_t i = 0; i < n; ++i) vk[i] = r[i] / beta;
            oracle_symv_upper(n, A, vk, r);
            axpy_(n, -beta, V + (size_t)(k - 1) * n, r);
            double a = dot_(n, vk, r);
            axpy_(n, -a, vk, r);
            /* full re-orthogonalisation, modified Gram-Schmidt against all basis vectors */
            for (int64_t q = 0; q <= k; ++q) {
                const double* vq = V + (size_t)q * n;
                double s = dot_(n, vq, r);
                axpy_(n, -s, vq, r);
                if (q == k) a += s;
            }
            alphas[k] = a;
            betas[k] = nrm2_(n, r);
            k += 1;
            numops += 1;
        } else {
            if (numiter == maxiter) break;
            int64_t keep = (3 * K + 2 * converged) / 5;
            /* restore Lanczos (tridiagonal) form in the first keep columns:
             * S = diag(D[0..keep)), arrow row a = f[0..keep) */
            memset(S, 0, sizeof(double) * (size_t)K * (size_t)K);
            for (int64_t j = 0; j < keep; ++j) S[j + j * K] = D[j];
            double* arrow = tmpK;
            for (int64_t j = 0; j < keep; ++j) arrow[j] = f[j];
            for (int64_t j = keep - 1; j >= 0; --j) {
                /* reflector P = I - 2uu' on coords 0..j mapping arrow[0..j] -> nu e_j */
                int64_t len = j + 1;
                /* nu = +||x|| (KrylovKit's householder returns a non-negative nu) */
                double sigma = dot_(j, arrow, arrow);
                double xj = arrow[j];
                double nu = sqrt(xj * xj + sigma);
                int have = 0;
                if (!(sigma == 0.0 && xj == nu)) {
                    for (int64_t i = 0; i < j; ++i) hv[i] = arrow[i];
                    hv[j] = (xj < 0.0) ? (xj - nu) : (-sigma / (xj + nu));
                    double hn = nrm2_(len, hv);
                    if (hn > 0.0) {
                        for (int64_t i = 0; i < len; ++i) hv[i] /= hn;
                        have = 1;
                    }
                }
                betas[j] = nu;             /* H[j+1, j] */
                if (have) {
                    /* S[0..j,0..j] <- P S P */
                    for (int64_t c = 0; c < len; ++c) {           /* left: columns */
                        double* col = S + c * K;
                        double s = 2.0 * dot_(len, hv, col);
                        for (int64_t i = 0; i < len; ++i) col[i] -= s * hv[i];
                    }
                    for (int64_t rr = 0; rr < len; ++rr) {          /* right: rows */
                        double s = 0.0;
                        for (int64_t c = 0; c < len; ++c) s += S[rr + c * K] * hv[c];
                        s *= 2.0;
                        for (int64_t c = 0; c < len; ++c) S[rr + c * K] -= s * hv[c];
                    }
                    /* U[:, 0..j] <- U[:, 0..j] P */
                    for (int64_t rr = 0; rr < K; ++rr) {
                        double s = 0.0;
                        for (int64_t c = 0; c < len; ++c) s += U[rr + c * K] * hv[c];
                        s *= 2.0;
                        for (int64_t c = 0; c < len; ++c) U[rr + c * K] -= s * hv[c];
                    }
                }
                alphas[j] = S[j + j * K];
                /* next arrow row = S[j, 0..j-1] */
                for (int64_t c = 0; c < j; ++c) arrow[c] = S[j + c * K];
            }
            /* basistransform!: B <- B * U[:, 0..keep) ; B[keep] = r / beta */
            #pragma omp parallel for schedule(static) if (n > 512)
            for (int64_t i = 0; i < n; ++i) {
                for (int64_t j = 0; j < keep; ++j) {
                    double s = 0.0;
                    for (int64_t q = 0; q < K; ++q) s += V[i + (size_t)q * n] * U[q + j * K];
                    Vnew[i + (size_t)j * n] = s;
                }
            }
            memcpy(V, Vnew, sizeof(double) * (size_t)n * (size_t)keep);
            double* vk = V + (size_t)keep * n;
            for (int64_t i = 0; i < n; ++i) vk[i] = r[i] / beta;
            /* shrink!: factorization of length keep, residual = beta_keep * V[keep] …
             * then immediately expand (next loop turn) with k = keep and the
             * recurrence coefficient betas[keep-1] = nu.  KrylovKit's shrink! sets
             * r = V[keep+1]*βs[keep]; expand! then re-normalises it. */
            for (int64_t i = 0; i < n; ++i) r[i] = vk[i] * betas[keep - 1];
            k = keep;
            numiter += 1;
        }
    }

    int64_t nv = howmany;
    if (converged > nv) nv = converged;
    if (nv > Kcur) nv = Kcur;
    for (int64_t j = 0; j < nv; ++j) vals[j] = D[j];
    /* vectors = B * U[:, j] */
    for (int64_t j = 0; j < nv; ++j) {
        double* out = vecs + (size_t)j * n;
        for (int64_t i = 0; i < n; ++i) out[i] = 0.0;
        for (int64_t q = 0; q < Kcur; ++q) axpy_(n, U[q + j * Kcur], V + (size_t)q * n, out);
    }
    *nvals_out = nv;
    *converged_out = converged;
    *numops_out = numops;
    *numiter_out = numiter;
    free(V); free(r); free(alphas); free(betas); free(T); free(U); free(D); free(f);
    free(S); free(hv); free(tmpK); free(Vnew);
    return 0;
}
