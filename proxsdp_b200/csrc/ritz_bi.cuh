// ritz_bi.cuh — the top few eigenpairs of the Lanczos tridiagonal by bisection + twisted vectors.
//
// KrylovKit diagonalises the whole K x K Rayleigh quotient (`tridiageigh!`) to decide convergence; on a GPU
// a dense K x K Jacobi solve costs ~130 us of pure latency (136 dependent rounds), more than the 25 mat-vecs
// of the eigsolve themselves.  In the common case (first Krylov cycle, the `howmany` wanted Ritz pairs have
// converged) only the leading few eigenpairs of a plain symmetric TRIDIAGONAL are needed:
//     values   : 32-way multisection on the Sturm count (division-free three-term recurrence), one warp per value;
//     vectors  : forward and backward solutions of (T - lambda I) z = 0 glued at the twist index that minimises
//                the single remaining residual gamma_r (Fernando's twisted factorisation), one warp per vector.
// Every pair is verified (residual, mutual orthogonality, eigenvalue gaps); anything doubtful makes the routine
// decline, and the caller falls back to the dense Jacobi solver, which also serves thick restarts (arrowhead
// Rayleigh quotient) and non-converged cycles.  So the accelerated path can only ever return pairs that pass
// the same accuracy bar as the dense solver.
#pragma once
#include "common.cuh"

namespace pb {

constexpr int RITZ_BI_MAXM = 15;      // pairs computed at most (one warp each; warp m computes one extra value)

struct RitzBiScratch {
    double* ie;      // K        1 / e[j]
    double* e2s;     // K        (e[j] / ||T||)^2
    double* zf;      // 16 * K   forward solutions, one row per warp
    double* zb;      // 16 * K   backward solutions
    double* lam;     // 16       eigenvalues (descending)
    int* fail;       // 1
};

__host__ __device__ inline size_t ritz_bi_scratch_doubles(int K) { return 2 * (size_t)K + 32 * (size_t)K + 16 + 2; }

__device__ inline RitzBiScratch ritz_bi_carve(double* base, int K) {
    RitzBiScratch s;
    s.ie = base; base += K;
    s.e2s = base; base += K;
    s.zf = base; base += 16 * (size_t)K;
    s.zb = base; base += 16 * (size_t)K;
    s.lam = base; base += 16;
    s.fail = reinterpret_cast<int*>(base);
    return s;
}

// number of eigenvalues of tridiag(d, e) that are < x  (sign changes of the leading principal minors).
// The minors are those of (T - x I) / ||T||: entries of size <= 2, so the products cannot overflow between two
// rescaling checks (every 8 steps), and the recurrence p_j = dx_j p_{j-1} - e2_{j-1} p_{j-2} has ONE fused
// multiply-add on its dependent chain: e2_{j-1} p_{j-2} is formed a step early.  FP64 FMA latency is 23 cycles on
// B200, a count is ~30 cycles per row; scripts/sturm_check.py checks this arithmetic against LAPACK.
__device__ __forceinline__ int sturm_count(int k, const double* __restrict__ d, const double* __restrict__ e2s, double x, double inv_t) {
    double p0 = 1.0, p1 = (d[0] - x) * inv_t;
    bool s1 = (p1 <= 0.0);                 // sign of the minor, an exact zero counting as a change of sign
    int cnt = s1 ? 1 : 0;
    double t0 = (k > 1) ? e2s[0] * p0 : 0.0;
    // blocks of 8 rows: the 16 shared-memory loads and the 8 shifted diagonals of a block are formed before its
    // chain starts; the chain itself is ONE FMA per row — the sign bookkeeping (a zero minor takes the sign opposite
    // to its predecessor, and the next minor -e^2 p_{j-1} then agrees with it) runs beside it on predicates
    for (int jb = 1; jb < k; jb += 8) {
        double dx[8], ee[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = jb + u;
            dx[u] = (j < k) ? (d[j] - x) * inv_t : 1.0;          // rows past the end: p stays put (dx = 1, e2 = 0)
            ee[u] = (j < k - 1) ? e2s[j] : 0.0;
        }
        const int nb = min(8, k - jb);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double p2 = fma(dx[u], p1, -t0);
            const bool s2 = (p2 == 0.0) ? !s1 : (p2 < 0.0);
            cnt += (u < nb && s2 != s1) ? 1 : 0;
            if (u < nb) s1 = s2;
            t0 = ee[u] * p1;
            p0 = p1; p1 = p2;
        }
        const double a = fabs(p1);
        if (a > 1e100) { p0 *= 1e-100; p1 *= 1e-100; t0 *= 1e-100; }
        else if (a < 1e-100 && a > 0.0) { p0 *= 1e100; p1 *= 1e100; t0 *= 1e100; }
    }
    return cnt;
}

// Computes the m = min(m_want, RITZ_BI_MAXM) largest eigenpairs of the k x k tridiagonal (d: diagonal, e: sub-diagonal,
// both in shared memory).  On success returns m and fills lam_out[0..m) (descending) and the columns 0..m-1 of U
// (leading dimension ldu, rows 0..k-1).  Returns 0 when it declines.  All threads of the block must call it.
__device__ inline int ritz_top_bi(int k, const double* d, const double* e, int m_want, double* lam_out, double* U, int ldu,
                                  RitzBiScratch sc, long long* prof = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long tp = prof ? clock64() : 0;
#define RITZ_TICK(slot) do { if (prof && tid == 0) { long long tn = clock64(); prof[slot] += tn - tp; tp = tn; } } while (0)
    const int m = min(min(m_want, RITZ_BI_MAXM), k - 1);
    if (k < 8 || m < 1) return 0;
    // ---- bounds, scale, smallest coupling (every warp redundantly: no block barrier) ----
    double gl = 1e300, gu = -1e300, emin = 1e300;
    for (int j = lane; j < k; j += 32) {
        const double el = (j > 0) ? fabs(e[j - 1]) : 0.0, er = (j < k - 1) ? fabs(e[j]) : 0.0;
        gl = fmin(gl, d[j] - el - er);
        gu = fmax(gu, d[j] + el + er);
        if (j < k - 1) emin = fmin(emin, er);
    }
    for (int o = 16; o > 0; o >>= 1) {
        gl = fmin(gl, __shfl_xor_sync(0xffffffffu, gl, o));
        gu = fmax(gu, __shfl_xor_sync(0xffffffffu, gu, o));
        emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, o));
    }
    const double tnorm = fmax(fabs(gl), fabs(gu));
    if (!(tnorm > 0.0) || !(tnorm < 1e140)) return 0;     // degenerate: decline
    const double inv_t = 1.0 / tnorm;
    for (int j = tid; j < k - 1; j += blockDim.x) { const double es = e[j] * inv_t; sc.e2s[j] = es * es; }
    if (tid == 0) *sc.fail = 0;
    __syncthreads();
    RITZ_TICK(11);

    // ---- eigenvalue `warp` (0-based from the top) by 32-way multisection ----
    double lam = 0.0;
    if (warp <= m) {
        const int idx = k - 1 - warp;              // want the smallest x with count(x) > idx
        const double wdt = gu - gl;
        double lo = gl - 1e-3 * wdt, hi = gu + 1e-3 * wdt;
        for (int it = 0; it < 16; ++it) {
            const double x = lo + (hi - lo) * ((double)(lane + 1) * (1.0 / 33.0));
            const int c = sturm_count(k, d, sc.e2s, x, inv_t);
            const unsigned int okmask = __ballot_sync(0xffffffffu, c > idx);
            double nlo, nhi;
            if (okmask) {
                const int L = __ffs(okmask) - 1;
                nhi = __shfl_sync(0xffffffffu, x, L);
                const double xm1 = __shfl_sync(0xffffffffu, x, L > 0 ? L - 1 : 0);
                nlo = (L > 0) ? xm1 : lo;
            } else {
                nlo = __shfl_sync(0xffffffffu, x, 31);
                nhi = hi;
            }
            lo = nlo; hi = nhi;
            // a tridiagonal defines its eigenvalues to eps ||T|| only: stop there instead of chasing relative accuracy
            // on eigenvalues that are small next to ||T|| (5 more rounds for nothing; every pair is verified below)
            if (hi - lo <= fmax(2.3e-16 * fmax(fabs(lo), fabs(hi)), 2.3e-16 * tnorm)) break;
        }
        lam = 0.5 * (lo + hi);
        if (lane == 0) sc.lam[warp] = lam;
    }
    RITZ_TICK(12);

    // ---- eigenvector `warp` by the twisted factorisation (Parlett & Dhillon), written on the leading / trailing
    // principal minors of (T - lam I) / ||T|| so that its dependent chains are division-free: lane 0 runs
    // p_j = dx_j p_{j-1} - e_{j-1}^2 p_{j-2} from the top (p_j / p_{j-1} is the pivot D+_j of L D+ L'), lane 1 the same
    // from the bottom (q_j / q_{j+1} = D-_j of U D- U').  gamma_r = D+_r + D-_r - dx_r is the one residual left when both
    // factorisations are glued at row r; the twist goes to the smallest |gamma_r| and
    //     z_j = (-1)^{r-j} e_j ... e_{r-1} p_{j-1} q_{r+1}  (j <= r),   z_j = (-1)^{j-r} e_r ... e_{j-1} q_{j+1} p_{r-1}  (j >= r).
    // Nothing is ever divided by an off-diagonal entry: a (nearly) decoupled tridiagonal — the Rayleigh quotient right
    // after a thick restart with converged Ritz pairs has couplings of 1e-13 — is handled like any other (entries of
    // the scaled matrix are <= 2 in modulus and k <= 101, so the minors stay far from the overflow threshold).
    if (warp < m) {
        double* pf = sc.zf + (size_t)warp * k;      // pf[j] = p_j   (p_{-1} = 1)
        double* pb = sc.zb + (size_t)warp * k;      // pb[j] = q_j   (q_k = 1)
        if (lane == 0) {
            double p0 = 1.0, p1 = (d[0] - lam) * inv_t;
            pf[0] = p1;
            for (int j = 1; j < k; ++j) {
                const double p2 = fma((d[j] - lam) * inv_t, p1, -sc.e2s[j - 1] * p0);
                pf[j] = p2;
                p0 = p1; p1 = p2;
            }
        } else if (lane == 1) {
            double q0 = 1.0, q1 = (d[k - 1] - lam) * inv_t;
            pb[k - 1] = q1;
            for (int j = k - 2; j >= 0; --j) {
                const double q2 = fma((d[j] - lam) * inv_t, q1, -sc.e2s[j] * q0);
                pb[j] = q2;
                q0 = q1; q1 = q2;
            }
        }
        __syncwarp();
        double best = 1e300; int bestr = 0;
        for (int r = lane; r < k; r += 32) {
            const double pm = (r > 0) ? pf[r - 1] : 1.0, qp = (r < k - 1) ? pb[r + 1] : 1.0;
            double ag = fabs(pf[r] / pm + pb[r] / qp - (d[r] - lam) * inv_t);
            if (!(ag < 1e300)) ag = 1e300;                    // NaN / inf (a vanishing minor): never chosen
            if (ag < best) { best = ag; bestr = r; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int orr = __shfl_xor_sync(0xffffffffu, bestr, o);
            if (ob < best || (ob == best && orr < bestr)) { best = ob; bestr = orr; }
        }
        double* u = U + (size_t)warp * ldu;
        {
            const double pm = (bestr > 0) ? pf[bestr - 1] : 1.0, qp = (bestr < k - 1) ? pb[bestr + 1] : 1.0;
            if (lane == 0) {
                double E = qp;
                u[bestr] = pm * qp;
                for (int j = bestr - 1; j >= 0; --j) { E = -(e[j] * inv_t) * E; u[j] = E * ((j > 0) ? pf[j - 1] : 1.0); }
            } else if (lane == 1) {
                double E = pm;
                for (int j = bestr + 1; j < k; ++j) { E = -(e[j - 1] * inv_t) * E; u[j] = E * ((j < k - 1) ? pb[j + 1] : 1.0); }
            }
        }
        __syncwarp();
        double nrm2 = 0.0;
        for (int j = lane; j < k; j += 32) nrm2 = fma(u[j], u[j], nrm2);
        nrm2 = warp_sum(nrm2);
        const double inrm = rsqrt(nrm2);
        for (int j = lane; j < k; j += 32) u[j] *= inrm;
        __syncwarp();
        // residual check || T u - lam u ||_inf <= 1e-13 ||T||
        double res = 0.0;
        for (int j = lane; j < k; j += 32) {
            double t = (d[j] - lam) * u[j];
            if (j > 0) t = fma(e[j - 1], u[j - 1], t);
            if (j < k - 1) t = fma(e[j], u[j + 1], t);
            res = fmax(res, fabs(t));
        }
        for (int o = 16; o > 0; o >>= 1) res = fmax(res, __shfl_xor_sync(0xffffffffu, res, o));
        if (lane == 0 && !(res <= 1e-13 * tnorm && nrm2 < 1e300)) *sc.fail = 1;
    }
    __syncthreads();
    RITZ_TICK(13);
    // ---- gaps and mutual orthogonality ----
    if (tid < m) {
        if (!(sc.lam[tid] - sc.lam[tid + 1] >= 1e-7 * tnorm)) *sc.fail = 1;
    }
    const int npairs = m * (m - 1) / 2;
    for (int pidx = warp; pidx < npairs; pidx += (int)(blockDim.x >> 5)) {
        int b = (int)((1.0 + sqrt(1.0 + 8.0 * (double)pidx)) * 0.5);
        while (b * (b - 1) / 2 > pidx) --b;
        while ((b + 1) * b / 2 <= pidx) ++b;
        const int a_ = pidx - b * (b - 1) / 2;             // a_ < b
        const double* ua = U + (size_t)a_ * ldu;
        const double* ub = U + (size_t)b * ldu;
        double s = 0.0;
        for (int j = lane; j < k; j += 32) s = fma(ua[j], ub[j], s);
        s = warp_sum(s);
        if (lane == 0 && !(fabs(s) <= 1e-12)) *sc.fail = 1;
    }
    __syncthreads();
    RITZ_TICK(14);
    if (*sc.fail) return 0;
    for (int i = tid; i < m; i += blockDim.x) lam_out[i] = sc.lam[i];
    __syncthreads();
    return m;
}

}  // namespace pb
