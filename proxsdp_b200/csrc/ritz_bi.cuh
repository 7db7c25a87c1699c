// ritz_bi.cuh — the top few eigenpairs of the Lanczos tridiagonal by bisection + twisted vectors.
//
// KrylovKit diagonalises the whole K x K Rayleigh quotient (`tridiageigh!`) to decide convergence; on a GPU
// a dense K x K Jacobi solve costs ~130 us of pure latency (136 dependent rounds), more than the 25 mat-vecs
// of the eigsolve themselves.  In the common case (first Krylov cycle, the `howmany` wanted Ritz pairs have
// converged) only the leading few eigenpairs of a plain symmetric TRIDIAGONAL are needed:
//     values   : 32 concurrent Sturm counts per round (division-free three-term recurrence), one warp per value:
//                33-way multisection until the value is isolated, then points clustered around the secant estimate
//                of the characteristic polynomial's root (quadratic shrinkage of the bracket);
//     vectors  : forward and backward solutions of (T - lambda I) z = 0 glued at the twist index that minimises
//                the single remaining residual gamma_r (Fernando's twisted factorisation), one warp per vector.
// Every pair is verified (residual, mutual orthogonality, eigenvalue gaps); anything doubtful makes the routine
// decline, and the caller falls back to the dense Jacobi solver, which also serves thick restarts (arrowhead
// Rayleigh quotient) and non-converged cycles.  So the accelerated path can only ever return pairs that pass
// the same accuracy bar as the dense solver.
#pragma once
#include "common.cuh"

namespace pb {

constexpr int RITZ_BI_MAXM = 47;      // pairs computed at most: warp w takes pairs w, w + 16, w + 32 (one extra VALUE is computed after the last pair)
constexpr int RITZ_BI_PAD = 16;       // padding rows behind the k real ones (the Sturm chain runs in blocks of 8 rows)

struct RitzBiScratch {
    double* es;      // K        e[j] 2^-s                      (s: the power-of-two scale, entries of T 2^-s are <= 1)
    double* e2s;     // K + PAD  (e[j] 2^-s)^2, zero from row k - 1 on
    double* dxs;     // K + PAD  d[j] 2^-s
    double* zf;      // 16 * K   leading principal minors, one row per warp (re-used when a warp takes a second pair)
    double* zb;      // 16 * K   trailing principal minors
    double* lam;     // 48       eigenvalues (descending)
    double* rn;      // 48       residual || T u - lam u ||_inf of every vector
    int* fail;       // 1
};

__host__ __device__ inline size_t ritz_bi_scratch_doubles(int K) { return 3 * (size_t)K + 2 * RITZ_BI_PAD + 32 * (size_t)K + 48 + 48 + 2; }

__device__ inline RitzBiScratch ritz_bi_carve(double* base, int K) {
    RitzBiScratch s;
    s.es = base; base += K;
    s.e2s = base; base += K + RITZ_BI_PAD;
    s.dxs = base; base += K + RITZ_BI_PAD;
    s.zf = base; base += 16 * (size_t)K;
    s.zb = base; base += 16 * (size_t)K;
    s.lam = base; base += 48;
    s.rn = base; base += 48;
    s.fail = reinterpret_cast<int*>(base);
    return s;
}

__device__ __forceinline__ double pow2_double(int e) { return __hiloint2double((1023 + e) << 20, 0); }

// Number of eigenvalues of the scaled tridiagonal (diagonal dxs, squared couplings e2s) that are < xs: the sign
// changes of its leading principal minors p_j = (dxs_j - xs) p_{j-1} - e2s_{j-1} p_{j-2}.  Also returns the last
// minor p_k(xs) — the characteristic polynomial, which the caller interpolates — as pout 2^exout.
// The FP64 pipe is the bottleneck when 16 warps run this at once (two warp instructions per cycle and SM), so the
// chain is pared down to three FP64 instructions per row: the shift of the diagonal, ONE fused multiply-add on the
// dependent chain (e2s_{j-1} p_{j-2} is formed a step early), and that product.  Everything else runs on the integer
// pipe: the sign and zero tests read the bit pattern of the minor (a zero minor takes the sign opposite to its
// predecessor, and the next minor -e^2 p_{j-1} then agrees with it), rows past the end are disabled by predicates on
// uniform values (dx = 1, e2s = 0: p stays put), and the range check every 8 rows looks at the exponent field.
// Entries of the scaled matrix are <= 1 + |xs| <= 2.01, so a block of 8 rows moves the exponent by < 2^13.
__device__ __forceinline__ int sturm_count(int k, const double* __restrict__ dxs, const double* __restrict__ e2s, double xs, double& pout, int& exout) {
    double p0 = 1.0, p1 = dxs[0] - xs;
    int ex = 0;
    bool s1;
    {
        const int hi = __double2hiint(p1), lo = __double2loint(p1);
        s1 = (((hi & 0x7fffffff) | lo) == 0) || (hi < 0);      // p_0 <= 0  (p_{-1} = 1 > 0)
    }
    int cnt = s1 ? 1 : 0;
    double t0 = e2s[0] * p0;
    for (int jb = 1; jb < k; jb += 8) {
        double dx[8], ee[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            dx[u] = (jb + u < k) ? dxs[jb + u] - xs : 1.0;
            ee[u] = e2s[jb + u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double p2 = fma(dx[u], p1, -t0);
            const int hi = __double2hiint(p2), lo = __double2loint(p2);
            const bool zero = ((hi & 0x7fffffff) | lo) == 0;
            const bool s2 = zero ? !s1 : (hi < 0);
            const bool valid = jb + u < k;
            cnt += (valid && s2 != s1) ? 1 : 0;
            s1 = valid ? s2 : s1;
            t0 = ee[u] * p1;
            p0 = p1; p1 = p2;
        }
        const int eb = (__double2hiint(p1) >> 20) & 0x7ff;
        if (eb > 1023 + 400) { const double f = pow2_double(-400); p0 *= f; p1 *= f; t0 *= f; ex += 400; }
        else if (eb < 1023 - 400 && p1 != 0.0) { const double f = pow2_double(400); p0 *= f; p1 *= f; t0 *= f; ex -= 400; }
    }
    pout = p1; exout = ex;
    return cnt;
}

// Eigenvalue number idx (ascending, 0-based) of the scaled tridiagonal, by one warp: the smallest x with count(x) > idx.
// Rounds of 32 concurrent Sturm counts shrink a bracket [lo, hi] with count(lo) <= idx < count(hi):
//   * uniform rounds (33-way multisection) until the bracket isolates the eigenvalue (count(hi) - count(lo) == 1);
//   * then the characteristic polynomial changes sign exactly once inside the bracket, the secant through its two end
//     values lands within O(w^2 / gap) of the root, and the 32 points of the next round are placed around that
//     estimate at distances w 8^-i (i = 1..16 on either side): the bracket shrinks to a few times the secant error,
//     i.e. quadratically — 5 to 7 rounds in all instead of the 11 that plain multisection needs for 53 bits;
//   * a clustered round that shrank the bracket by less than 8x is followed by a uniform one, so the worst case is
//     still geometric.
// Only Sturm counts move the ends of the bracket: the polynomial values steer where the points go, never what is
// accepted.  The caller verifies every pair (residual, orthogonality, gaps) in any case.
__device__ __forceinline__ double ritz_value_warp(int k, const double* __restrict__ dxs, const double* __restrict__ e2s, int idx,
                                                  double lo, double hi, double tnorm_s, long long* rounds = nullptr) {
    const int lane = threadIdx.x & 31;
    int clo = 0, chi = k, elo = 0, ehi = 0;
    double plo = 0.0, phi = 0.0;
    bool vlo = false, vhi = false, cluster = false;
    double xc = 0.0;
    int it = 0;
    for (; it < 48; ++it) {
        const double w = hi - lo;
        double x;
        if (cluster) {
            const int i = (lane < 16) ? lane + 1 : 32 - lane;             // 1 .. 16, ascending abscissae across the lanes
            const double off = w * pow2_double(-3 * i);
            x = fmin(fmax((lane < 16) ? xc - off : xc + off, lo), hi);
        } else {
            x = lo + w * ((double)(lane + 1) * (1.0 / 33.0));
        }
        double p; int ex;
        const int c = sturm_count(k, dxs, e2s, x, p, ex);
        const unsigned int okmask = __ballot_sync(0xffffffffu, c > idx);
        const int L = okmask ? __ffs(okmask) - 1 : 32;                    // first point whose count exceeds idx
        const int srcl = max(L - 1, 0), srch = min(L, 31);
        const double xl = __shfl_sync(0xffffffffu, x, srcl), xh = __shfl_sync(0xffffffffu, x, srch);
        const double pl = __shfl_sync(0xffffffffu, p, srcl), ph = __shfl_sync(0xffffffffu, p, srch);
        const int cl = __shfl_sync(0xffffffffu, c, srcl), ch = __shfl_sync(0xffffffffu, c, srch);
        const int el = __shfl_sync(0xffffffffu, ex, srcl), eh = __shfl_sync(0xffffffffu, ex, srch);
        if (L > 0) { lo = xl; clo = cl; plo = pl; elo = el; vlo = true; }
        if (L < 32) { hi = xh; chi = ch; phi = ph; ehi = eh; vhi = true; }
        const double wn = hi - lo;
        // a tridiagonal defines its eigenvalues to eps ||T|| only: stop there instead of chasing relative accuracy
        // on eigenvalues that are small next to ||T|| (every pair is verified by the caller)
        if (wn <= fmax(2.3e-16 * fmax(fabs(lo), fabs(hi)), 2.3e-16 * tnorm_s)) { ++it; break; }
        bool iso = (chi - clo == 1) && vlo && vhi && elo == ehi && ((plo < 0.0) != (phi < 0.0));
        double fr = 0.5;
        if (iso) {
            const double a = fabs(plo), b = fabs(phi);
            fr = a / (a + b);
            iso = (fr >= 0.0 && fr <= 1.0);                               // false for NaN
        }
        cluster = iso && (!cluster || wn <= 0.125 * w);
        xc = lo + wn * fr;
    }
    if (rounds && lane == 0) *rounds += it;
    return 0.5 * (lo + hi);
}

// One minor chain of the twisted factorisation: from the top (dir == 0: out[j] = p_j, leading minors of T - lam I) or
// from the bottom (dir == 1: out[j] = q_j, trailing minors), the same three-term recurrence as the Sturm count.  Rows
// are taken in blocks of 8 whose operands are loaded before the block's chain starts (out never aliases the inputs).
__device__ __forceinline__ void minor_chain(int k, int dir, const double* __restrict__ dxs, const double* __restrict__ e2s, double xs,
                                            double* __restrict__ out) {
    const int first = dir ? k - 1 : 0;
    double p0 = 1.0, p1 = dxs[first] - xs;
    out[first] = p1;
    for (int sb = 1; sb < k; sb += 8) {
        double dx[8], ee[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int s = min(sb + u, k - 1);
            const int row = dir ? k - 1 - s : s, cpl = dir ? k - 1 - s : s - 1;      // coupling between this row and the previous one
            dx[u] = dxs[row] - xs;
            ee[u] = e2s[cpl];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (sb + u < k) {
                const double p2 = fma(dx[u], p1, -ee[u] * p0);
                const int s = sb + u;
                out[dir ? k - 1 - s : s] = p2;
                p0 = p1; p1 = p2;
            }
        }
    }
}

// Computes the m = min(m_want, RITZ_BI_MAXM) largest eigenpairs of the k x k tridiagonal (d: diagonal, e: sub-diagonal,
// both in shared memory).  On success returns m and fills lam_out[0..m) (descending) and the columns 0..m-1 of U
// (leading dimension ldu, rows 0..k-1).  Returns 0 when it declines.  All threads of the block must call it.
// have > 0: the first `have` pairs (and value number `have`) are still in place from a previous successful call on the
// same tridiagonal (lam_out, U and the scratch untouched since): only the pairs have..m-1 are computed.
__device__ inline int ritz_top_bi(int k, const double* d, const double* e, int m_want, double* lam_out, double* U, int ldu,
                                  RitzBiScratch sc, long long* prof = nullptr, int have = 0) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long tp = prof ? clock64() : 0;
#define RITZ_TICK(slot) do { if (prof && tid == 0) { long long tn = clock64(); prof[slot] += tn - tp; tp = tn; } } while (0)
    const int m = min(min(m_want, RITZ_BI_MAXM), k - 1);
    if (k < 8 || m < 1) return 0;
    if (have >= m) return m;
    // ---- bounds and scale (every warp redundantly: no block barrier) ----
    double gl = 1e300, gu = -1e300;
    for (int j = lane; j < k; j += 32) {
        const double el = (j > 0) ? fabs(e[j - 1]) : 0.0, er = (j < k - 1) ? fabs(e[j]) : 0.0;
        gl = fmin(gl, d[j] - el - er);
        gu = fmax(gu, d[j] + el + er);
    }
    for (int o = 16; o > 0; o >>= 1) {
        gl = fmin(gl, __shfl_xor_sync(0xffffffffu, gl, o));
        gu = fmax(gu, __shfl_xor_sync(0xffffffffu, gu, o));
    }
    const double tnorm = fmax(fabs(gl), fabs(gu));
    if (!(tnorm > 0.0) || !(tnorm < 1e140) || !(tnorm > 1e-140)) return 0;     // degenerate: decline
    // power-of-two scale: T 2^-s has entries <= 1 and the scaling is exact, (d_j - x) 2^-s == d_j 2^-s - x 2^-s
    const int sexp = ((__double2hiint(tnorm) >> 20) & 0x7ff) - 1023 + 1;
    const double inv_t = pow2_double(-sexp), tscale = pow2_double(sexp);
    if (have == 0) {
        for (int j = tid; j < k + RITZ_BI_PAD; j += blockDim.x) {
            const double es = (j < k - 1) ? e[j] * inv_t : 0.0;
            if (j < k) sc.es[j] = es;
            sc.e2s[j] = es * es;
            sc.dxs[j] = (j < k) ? d[j] * inv_t : 0.0;
        }
        if (tid == 0) *sc.fail = 0;
        __syncthreads();
    }
    RITZ_TICK(11);

    // ---- eigenvalues 0 .. m (0-based from the top; one more than the pairs: the gap below the last pair is checked on it),
    // warp w takes the values w, w + 16 ----
    const int nwarps = (int)(blockDim.x >> 5);
    const int nval = m + 1;
    for (int idx = warp; idx < nval; idx += nwarps) {
        if (idx <= have && have > 0) continue;          // values 0 .. have are still there from the previous call
        const double wdt = (gu - gl) * inv_t;
        const double lam_s = ritz_value_warp(k, sc.dxs, sc.e2s, k - 1 - idx, gl * inv_t - 1e-3 * wdt, gu * inv_t + 1e-3 * wdt, tnorm * inv_t,
                                             (prof && idx == 0) ? prof + 18 : nullptr);
        if (lane == 0) sc.lam[idx] = lam_s * tscale;
    }
    // a (nearly) multiple eigenvalue makes the routine decline: find that out now, before the vectors are computed
    __syncthreads();
    {
        const bool close = (tid < m) && !(sc.lam[tid] - sc.lam[tid + 1] >= 1e-7 * tnorm);
        if (__syncthreads_or(close ? 1 : 0)) return 0;
    }
    RITZ_TICK(12);

    // ---- eigenvector `warp` by the twisted factorisation (Parlett & Dhillon), written on the leading / trailing
    // principal minors of (T - lam I) 2^-s so that its dependent chains are division-free: lane 0 runs
    // p_j = dx_j p_{j-1} - e_{j-1}^2 p_{j-2} from the top (p_j / p_{j-1} is the pivot D+_j of L D+ L'), lane 1 the same
    // from the bottom (q_j / q_{j+1} = D-_j of U D- U').  gamma_r = D+_r + D-_r - dx_r is the one residual left when both
    // factorisations are glued at row r; the twist goes to the smallest |gamma_r| and
    //     z_j = (-1)^{r-j} e_j ... e_{r-1} p_{j-1} q_{r+1}  (j <= r),   z_j = (-1)^{j-r} e_r ... e_{j-1} q_{j+1} p_{r-1}  (j >= r).
    // Nothing is ever divided by an off-diagonal entry: a (nearly) decoupled tridiagonal — the Rayleigh quotient right
    // after a thick restart with converged Ritz pairs has couplings of 1e-13 — is handled like any other (entries of
    // the scaled matrix are <= 2 in modulus and k <= 101, so the minors stay far from the overflow threshold).
    for (int idx = warp; idx < m; idx += nwarps) {
        if (idx < have) continue;
        const double lam_s = sc.lam[idx] * inv_t;   // (written by this very warp, or by the previous call)
        double* pf = sc.zf + (size_t)warp * k;      // pf[j] = p_j   (p_{-1} = 1)
        double* pb = sc.zb + (size_t)warp * k;      // pb[j] = q_j   (q_k = 1)
        __syncwarp();                               // (second pair of this warp: the minors of the first one are dead)
        if (lane < 2) minor_chain(k, lane, sc.dxs, sc.e2s, lam_s, lane ? pb : pf);
        __syncwarp();
        double best = 1e300; int bestr = 0;
        for (int r = lane; r < k; r += 32) {
            const double pm = (r > 0) ? pf[r - 1] : 1.0, qp = (r < k - 1) ? pb[r + 1] : 1.0;
            // gamma_r = p_r / p_{r-1} + q_r / q_{r+1} - dx_r over the common denominator: one division
            double ag = fabs((pf[r] * qp + pb[r] * pm - (sc.dxs[r] - lam_s) * (pm * qp)) / (pm * qp));
            if (!(ag < 1e300)) ag = 1e300;                    // NaN / inf (a vanishing minor): never chosen
            if (ag < best) { best = ag; bestr = r; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int orr = __shfl_xor_sync(0xffffffffu, bestr, o);
            if (ob < best || (ob == best && orr < bestr)) { best = ob; bestr = orr; }
        }
        double* u = U + (size_t)idx * ldu;
        {
            const double pm = (bestr > 0) ? pf[bestr - 1] : 1.0, qp = (bestr < k - 1) ? pb[bestr + 1] : 1.0;
            const double* __restrict__ es = sc.es;
            if (lane == 0) {
                // downwards from the twist, 4 rows at a time (operands first, then the chain of products)
                double E = qp;
                u[bestr] = pm * qp;
                int j = bestr - 1;
                for (; j >= 3; j -= 4) {
                    const double e0 = es[j], e1 = es[j - 1], e2 = es[j - 2], e3 = es[j - 3];
                    const double f0 = pf[j - 1], f1 = pf[j - 2], f2 = pf[j - 3], f3 = (j > 3) ? pf[j - 4] : 1.0;
                    E = -e0 * E; const double u0 = E * f0;
                    E = -e1 * E; const double u1 = E * f1;
                    E = -e2 * E; const double u2 = E * f2;
                    E = -e3 * E; const double u3 = E * f3;
                    u[j] = u0; u[j - 1] = u1; u[j - 2] = u2; u[j - 3] = u3;
                }
                for (; j >= 0; --j) { E = -es[j] * E; u[j] = E * ((j > 0) ? pf[j - 1] : 1.0); }
            } else if (lane == 1) {
                double E = pm;
                int j = bestr + 1;
                for (; j + 3 < k; j += 4) {
                    const double e0 = es[j - 1], e1 = es[j], e2 = es[j + 1], e3 = es[j + 2];
                    const double f0 = pb[j + 1], f1 = pb[j + 2], f2 = pb[j + 3], f3 = (j + 4 < k) ? pb[j + 4] : 1.0;
                    E = -e0 * E; const double u0 = E * f0;
                    E = -e1 * E; const double u1 = E * f1;
                    E = -e2 * E; const double u2 = E * f2;
                    E = -e3 * E; const double u3 = E * f3;
                    u[j] = u0; u[j + 1] = u1; u[j + 2] = u2; u[j + 3] = u3;
                }
                for (; j < k; ++j) { E = -es[j - 1] * E; u[j] = E * ((j < k - 1) ? pb[j + 1] : 1.0); }
            }
        }
        __syncwarp();
        double nrm2 = 0.0;
        for (int j = lane; j < k; j += 32) nrm2 = fma(u[j], u[j], nrm2);
        nrm2 = warp_sum(nrm2);
        const double inrm = rsqrt(nrm2);
        for (int j = lane; j < k; j += 32) u[j] *= inrm;
        __syncwarp();
        // residual || T u - lam u ||_inf: must be <= 1e-13 ||T||; its value also decides below which pairs of vectors need
        // their inner product checked at all
        const double lam = lam_s * tscale;
        double res = (nrm2 < 1e300 && nrm2 > 0.0) ? 0.0 : 1e300;
        for (int j = lane; j < k; j += 32) {
            double t = (d[j] - lam) * u[j];
            if (j > 0) t = fma(e[j - 1], u[j - 1], t);
            if (j < k - 1) t = fma(e[j], u[j + 1], t);
            res = fmax(res, fabs(t));
            if (!(fabs(t) <= 1e300)) res = 1e300;           // NaN
        }
        for (int o = 16; o > 0; o >>= 1) res = fmax(res, __shfl_xor_sync(0xffffffffu, res, o));
        if (lane == 0) { sc.rn[idx] = res; if (!(res <= 1e-13 * tnorm)) *sc.fail = 1; }
    }
    __syncthreads();
    RITZ_TICK(13);
    // ---- gaps and mutual orthogonality (one warp per pair of vectors: measured inside the eigsolve kernel, where this
    // code runs once per launch from a cold instruction cache, the compact warp-cooperative loop beats one thread per pair) ----
    // (the gaps were checked right after the values)
    // Two unit vectors with residuals r_a, r_b for eigenvalues lam_a != lam_b satisfy
    //     |u_a . u_b| <= (||r_a||_2 + ||r_b||_2) / |lam_a - lam_b|
    // ((lam_a - lam_b) u_a.u_b = u_a.r_b - u_b.r_a for symmetric T): pairs whose bound is already below the 1e-12 bar need no
    // inner product — with residuals of a few eps ||T|| that is every pair whose eigenvalues are not nearly degenerate.
    const double sqk = sqrt((double)k);
    const int npairs = m * (m - 1) / 2;
#pragma unroll 1
    for (int pidx = warp; pidx < npairs; pidx += (int)(blockDim.x >> 5)) {
        int b = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)pidx)) * 0.5f);
        while (b * (b - 1) / 2 > pidx) --b;
        while ((b + 1) * b / 2 <= pidx) ++b;
        const int a_ = pidx - b * (b - 1) / 2;             // a_ < b
        if (b < have) continue;                            // both vectors were checked by the previous call
        if ((sc.rn[a_] + sc.rn[b]) * sqk <= 1e-12 * fabs(sc.lam[a_] - sc.lam[b])) continue;
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
