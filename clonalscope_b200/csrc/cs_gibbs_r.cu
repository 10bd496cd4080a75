// cs_gibbs_r.cu — nested-CRP collapsed Gibbs sampler, R-compatible RNG mode.
//
// Consumes base R's random stream exactly as R/BayesNonparCluster.R does
// (set.seed :156, per cell R x (runif :167, sample.int :170), sample.int :207, per sweep
// rnorm :226), so that with the reference's seed the device reproduces the reference's
// own Zall.  Base R's generators are restated from their published algorithms
// (Mersenne-Twister MT19937 with R's seeding, sample.int's ProbSampleReplace with its
// descending heap sort `revsort`, inversion rnorm = AS 241 qnorm); see SURVEY.md App. B.
//
// The random stream is inherently serial, so this mode runs the whole chain in ONE
// persistent CTA: thread 0 owns the stream and the sorts, all threads score the
// occupied clusters and do the end-of-sweep reductions.  Two shortcuts keep the serial
// part short without changing a single draw:
//   * the sorted proposal vector c(npoints, alpha) is cached across cells until a cell
//     actually changes cluster (it is the same vector for all R segments of a cell);
//   * the final categorical draw walks the candidates in descending order directly;
//     only if the selected value is tied (where revsort's unstable order decides) is
//     the full heap sort replayed.
#include "cs_gibbs.cuh"
#include "cs_rstream.cuh"

#include <math_constants.h>

namespace {

constexpr int RT = 256; // threads of the persistent CTA

struct RShared {
    double *cumc;   // lcap+1 : sorted, cumulated proposal probabilities
    double *prob;   // lcap+1 : scores -> probabilities of the final draw
    int *permc;     // lcap+1
    int *perm;      // lcap+1
    int *np;        // lcap
    int *occ;       // lcap : occupied labels (0-based), ascending
    double *xrow, *mu, *sig; // R each
    uint32_t *mt;   // 624
};

__device__ __forceinline__ double dnorm_log(double x, double mu, double sigma)
{
    const double z = fabs((x - mu) / sigma);
    return -(CS_LN_SQRT_2PI + 0.5 * z * z + log(sigma));
}

// sum(dnorm(x, u, sigma, log=T), na.rm=T) in segment order
__device__ double cell_loglik(int R, const double *x, const double *u, const double *sig)
{
    double s = 0.0;
    for (int r = 0; r < R; r++) {
        const double xv = x[r];
        if (cs_isna(xv)) continue;
        const double t = dnorm_log(xv, u[r], sig[r]);
        if (!cs_isna(t)) s += t;
    }
    return s;
}

// ---- the pooled stream (POOL): the uniforms of the current and of the next Mersenne-Twister block
// (624 each) lie tempered in shared memory, so that any thread can read the k-th upcoming draw
// (k < 624).  The stream is still consumed in base R's order: only WHO reads a draw changes.
struct RPool {
    uint32_t *raw[2]; // raw[cur]: the state R would hold now; raw[cur ^ 1]: after its next refill
    double *uu[2];    // their tempered outputs as unif_rand() returns them
};

__device__ __forceinline__ double mt_temper(uint32_t y)
{
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    const double v = (double)y * 2.3283064365386963e-10;
    if (v <= 0.0) return 0.5 * 2.328306437080797e-10;
    if (1.0 - v <= 0.0) return 1.0 - 0.5 * 2.328306437080797e-10;
    return v;
}
__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b, uint32_t m)
{
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return m ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
// B = refill(A) by all threads (three dependent phases of <= 227 words), then both blocks tempered
__device__ void pool_build(const RPool &P, int cur, bool temper_cur)
{
    const uint32_t *A = P.raw[cur];
    uint32_t *B = P.raw[cur ^ 1];
    const int tid = threadIdx.x;
    if (temper_cur)
        for (int k = tid; k < 624; k += RT) P.uu[cur][k] = mt_temper(A[k]);
    for (int k = tid; k < 227; k += RT) B[k] = mt_twist(A[k], A[k + 1], A[k + 397]);
    __syncthreads();
    for (int k = 227 + tid; k < 454; k += RT) B[k] = mt_twist(A[k], A[k + 1], B[k - 227]);
    __syncthreads();
    for (int k = 454 + tid; k < 623; k += RT) B[k] = mt_twist(A[k], A[k + 1], B[k - 227]);
    if (tid == 0) B[623] = mt_twist(A[623], B[0], B[396]);
    __syncthreads();
    for (int k = tid; k < 624; k += RT) P.uu[cur ^ 1][k] = mt_temper(B[k]);
    __syncthreads();
}

template <bool POOL>
__global__ void __launch_bounds__(RT)
gibbs_r_kernel(ChainDev D, int lcap, int t0, int t1, int last_sweep, double *__restrict__ scratch)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    RShared S;
    RPool P;
    double *s_logsig, *s_term; // POOL: log(sigma) per segment; one row of terms per warp
    int *s_ind;
    {
        double *d = reinterpret_cast<double *>(s_raw);
        S.cumc = d; d += lcap + 1;
        S.prob = d; d += lcap + 1;
        S.xrow = d; d += D.R;
        S.mu = d; d += D.R;
        S.sig = d; d += D.R;
        s_logsig = d; d += POOL ? D.R : 0;
        s_term = d; d += POOL ? (RT / 32) * D.R : 0;
        P.uu[0] = d; d += POOL ? 624 : 0;
        P.uu[1] = d; d += POOL ? 624 : 0;
        int *p = reinterpret_cast<int *>(d);
        S.permc = p; p += lcap + 1;
        S.perm = p; p += lcap + 1;
        S.np = p; p += lcap;
        S.occ = p; p += lcap;
        s_ind = p; p += POOL ? D.R : 0;
        S.mt = reinterpret_cast<uint32_t *>(p);
        P.raw[0] = S.mt;
        P.raw[1] = S.mt + 624;
    }
    __shared__ int s_cur, s_flip;
    __shared__ int s_kt, s_nocc, s_pos, s_prop_valid, s_err, s_z, s_new;
    __shared__ long long s_rej, s_uused;
    __shared__ double s_red[RT];

    const int N = D.N, R = D.R, tid = threadIdx.x;
    for (int k = tid; k < lcap; k += RT) S.np[k] = D.np[k];
    for (int r = tid; r < R; r += RT) S.sig[r] = D.sig[r];
    for (int k = tid; k < 624; k += RT) S.mt[k] = D.mt[k];
    if (tid == 0) {
        s_kt = D.scal->kt;
        s_pos = D.scal->mt_pos;
        s_prop_valid = 0;
        s_err = 0;
        s_rej = 0;
        s_uused = D.scal->u_used;
        s_cur = 0;
        s_flip = 0;
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int k = 0; k < s_kt; k++)
            if (S.np[k] > 0) S.occ[n++] = k;
        s_nocc = n;
    }
    if (POOL) {
        for (int r = tid; r < R; r += RT) s_logsig[r] = log(S.sig[r]);
        pool_build(P, 0, true);
    }
    __syncthreads();
    // the k-th upcoming uniform (k < 624) and the uniform advance of the stream by n draws
    auto upcoming = [&](int k) {
        const int idx = s_pos + k;
        return idx < 624 ? P.uu[s_cur][idx] : P.uu[s_cur ^ 1][idx - 624];
    };
    auto advance = [&](int n) { // (called by all threads, between barriers)
        __syncthreads();
        if (tid == 0) {
            s_pos += n;
            s_flip = 0;
            if (s_pos > 624) { // the current block is used up: the next one becomes current
                s_pos -= 624;
                s_cur ^= 1;
                s_flip = 1;
            }
        }
        __syncthreads();
        if (s_flip) pool_build(P, s_cur, false);
    };

    for (int tt = t0; tt < t1 && !s_err; tt++) {
        const bool reject_live = D.single && tt == 0;
        for (int ii = 0; ii < N; ii++) {
            const int kt = s_kt;
            const int zold = D.Zs[ii] - 1;
            for (int r = tid; r < R; r += RT) S.xrow[r] = D.X[(size_t)ii * R + r];
            if (!s_prop_valid) { // c(npoints, alpha) / sum   (:170)
                const double sum = (double)N + (D.alpha > 0 ? D.alpha : 0.0);
                for (int k = tid; k <= kt; k += RT) {
                    S.cumc[k] = ((k < kt) ? (double)S.np[k] : D.alpha) / sum;
                    S.permc[k] = k + 1;
                }
            }
            __syncthreads();
            // ---- :165-175 proposal of the new cluster's per-segment states
            for (;;) {
                if (tid == 0) {
                    if (!s_prop_valid) {
                        revsort(S.cumc, S.permc, kt + 1);
                        for (int k = 1; k <= kt; k++) S.cumc[k] += S.cumc[k - 1];
                        s_prop_valid = 1;
                    }
                }
                if (POOL) {
                    // draws 2r and 2r + 1 of this attempt belong to segment r: every segment at once
                    __syncthreads();
                    for (int r = tid; r < R; r += RT) {
                        const double tmpr = 0.3 + (2.5 - 0.3) * upcoming(2 * r);
                        const double u = upcoming(2 * r + 1);
                        int lo = 0, hi = kt; // first j with u <= cumc[j] (cumc does not decrease), else kt
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (u > S.cumc[mid]) lo = mid + 1;
                            else hi = mid;
                        }
                        const int indr = S.permc[lo];
                        S.mu[r] = (indr <= kt) ? D.U[(size_t)(indr - 1) * R + r] : tmpr;
                    }
                    advance(2 * R);
                } else if (tid == 0) {
                    int pos = s_pos;
                    for (int r = 0; r < R; r++) {
                        const double tmpr = 0.3 + (2.5 - 0.3) * unif_rand(S.mt, pos);
                        const double u = unif_rand(S.mt, pos);
                        int j = 0;
                        while (j < kt && u > S.cumc[j]) j++;
                        const int indr = S.permc[j];
                        S.mu[r] = (indr <= kt) ? D.U[(size_t)(indr - 1) * R + r] : tmpr;
                    }
                    s_pos = pos;
                }
                __syncthreads();
                if (!reject_live) break;
                int dup = 0; // :174 any(apply(Ut, 1, identical, mu_new))
                for (int k = tid; k < kt; k += RT) {
                    const double *u = D.U + (size_t)k * R;
                    bool same = true;
                    for (int r = 0; r < R && same; r++) {
                        const double uv = u[r], mv = S.mu[r];
                        if (!(uv == mv) && !(cs_isna(uv) && cs_isna(mv))) same = false;
                    }
                    if (same) dup = 1;
                }
                dup = __syncthreads_or(dup);
                if (!dup) break;
                if (tid == 0) s_rej++;
            }
            // ---- :188 remove the cell, :193-198 score occupied clusters and the new table
            if (tid == 0) S.np[zold] -= 1;
            __syncthreads();
            const int nocc = s_nocc;
            if (POOL) {
                // a warp per candidate: the terms of all segments at once, then their sum in segment
                // order by one lane (the order the serial code adds them in)
                const int lane = tid & 31, wid = tid >> 5;
                double *term = s_term + (size_t)wid * R;
                for (int e = wid; e <= nocc; e += RT / 32) {
                    const int k = e < nocc ? S.occ[e] : -1;
                    const int n = e < nocc ? S.np[k] : 1;
                    if (n != 0) {
                        const double *u = e < nocc ? D.U + (size_t)k * R : S.mu;
                        for (int r = lane; r < R; r += 32) {
                            const double xv = S.xrow[r];
                            double t = CUDART_NAN;
                            if (!cs_isna(xv)) {
                                const double z = fabs((xv - u[r]) / S.sig[r]);
                                t = -(CS_LN_SQRT_2PI + 0.5 * z * z + s_logsig[r]);
                            }
                            term[r] = t;
                        }
                        __syncwarp();
                        if (lane == 0) {
                            double acc = 0.0;
                            for (int r = 0; r < R; r++) {
                                const double t = term[r];
                                if (!cs_isna(t)) acc += t;
                            }
                            S.prob[e] = acc + log(e < nocc ? (double)n : D.beta);
                        }
                        __syncwarp();
                    } else if (lane == 0) {
                        S.prob[e] = -INFINITY;
                    }
                }
            } else {
                for (int e = tid; e <= nocc; e += RT) {
                    if (e < nocc) {
                        const int k = S.occ[e];
                        const int n = S.np[k];
                        S.prob[e] = (n != 0) ? cell_loglik(R, S.xrow, D.U + (size_t)k * R, S.sig) + log((double)n)
                                             : -INFINITY;
                    } else {
                        S.prob[e] = cell_loglik(R, S.xrow, S.mu, S.sig) + log(D.beta);
                    }
                }
            }
            __syncthreads();
            // ---- :200-207 categorical draw
            if (POOL) { // the maximum and the exponentials by all threads; the sum in order by thread 0 below
                double mx = -INFINITY;
                for (int e = tid; e <= nocc; e += RT)
                    if (e == nocc || S.np[S.occ[e]] != 0) mx = fmax(mx, S.prob[e]);
                s_red[tid] = mx;
                __syncthreads();
                for (int off = RT / 2; off >= 1; off >>= 1) {
                    if (tid < off) s_red[tid] = fmax(s_red[tid], s_red[tid + off]);
                    __syncthreads();
                }
                mx = s_red[0];
                __syncthreads();
                for (int e = tid; e <= nocc; e += RT)
                    S.prob[e] = (e == nocc || S.np[S.occ[e]] != 0) ? exp(S.prob[e] - mx) : 0.0;
                __syncthreads();
                if (tid == 0) {
                    double sum = 0.0;
                    for (int e = 0; e <= nocc; e++)
                        if (S.prob[e] > 0) sum += S.prob[e];
                    s_red[0] = sum;
                }
                __syncthreads();
                const double sum = s_red[0];
                for (int e = tid; e <= nocc; e += RT) S.prob[e] /= sum;
                __syncthreads();
            }
            if (tid == 0) {
                int pos = s_pos;
                double u;
                if (POOL) {
                    u = upcoming(0);
                } else {
                    double mx = -INFINITY;
                    for (int e = 0; e <= nocc; e++)
                        if (e == nocc || S.np[S.occ[e]] != 0) mx = fmax(mx, S.prob[e]);
                    double sum = 0.0;
                    for (int e = 0; e <= nocc; e++) {
                        const double p = (e == nocc || S.np[S.occ[e]] != 0) ? exp(S.prob[e] - mx) : 0.0;
                        S.prob[e] = p;
                        if (p > 0) sum += p;
                    }
                    for (int e = 0; e <= nocc; e++) S.prob[e] /= sum;
                    u = unif_rand(S.mt, pos);
                    s_pos = pos;
                }
                // walk candidates in descending order; perm[] marks the ones already taken
                int z = -1;
                {
                    double cum = 0.0, prev = -1.0;
                    int taken = 0;
                    for (int e = 0; e <= nocc; e++) S.perm[e] = 0;
                    bool first = true;
                    for (;;) {
                        int best = -1, mult = 0;
                        double bv = 0.0;
                        for (int e = 0; e <= nocc; e++) {
                            if (S.perm[e]) continue;
                            const double p = S.prob[e];
                            if (p > bv) { bv = p; best = e; mult = 1; }
                            else if (p == bv && p > 0.0) mult++;
                        }
                        if (best < 0) break; // ran out of positive candidates (rounding): replay
                        cum = first ? bv : cum + bv;
                        first = false;
                        taken++;
                        S.perm[best] = 1;
                        const bool last_pos = (taken == kt + 1); // j capped at n-1
                        if (u <= cum || last_pos) {
                            if (mult == 1 && bv != prev) z = (best < nocc) ? S.occ[best] + 1 : kt + 1;
                            break;
                        }
                        prev = bv;
                    }
                }
                if (z < 0) { // tie at the selected position: replay sample.int's full heap sort
                    // expand to the full kt+1 vector (zeros for empty clusters) in cumc/permc,
                    // which invalidates the proposal cache
                    for (int k = 0; k <= kt; k++) { S.cumc[k] = 0.0; S.permc[k] = k + 1; }
                    for (int e = 0; e < nocc; e++) S.cumc[S.occ[e]] = S.prob[e];
                    S.cumc[kt] = S.prob[nocc];
                    revsort(S.cumc, S.permc, kt + 1);
                    for (int k = 1; k <= kt; k++) S.cumc[k] += S.cumc[k - 1];
                    int j = 0;
                    while (j < kt && u > S.cumc[j]) j++;
                    z = S.permc[j];
                    s_prop_valid = 0;
                }
                // ---- :209-218 bookkeeping
                s_new = 0;
                if (z > kt) {
                    if (kt + 1 > lcap) {
                        s_err = CS_ERR_KCAP;
                        z = zold + 1;
                        S.np[zold] += 1;
                    } else {
                        S.np[kt] = 1;
                        S.occ[s_nocc++] = kt;
                        s_kt = kt + 1;
                        s_new = 1;
                        s_prop_valid = 0;
                    }
                } else {
                    S.np[z - 1] += 1;
                    if (z - 1 != zold) s_prop_valid = 0;
                }
                if (z - 1 != zold && S.np[zold] == 0) { // cluster emptied: leave the occupied list
                    int n = s_nocc, w = 0;
                    for (int e = 0; e < n; e++)
                        if (S.occ[e] != zold) S.occ[w++] = S.occ[e];
                    s_nocc = w;
                }
                D.Zs[ii] = z;
                s_z = z;
            }
            if (POOL) advance(1);
            __syncthreads();
            if (s_new) {
                double *urow = D.U + (size_t)(s_z - 1) * R;
                for (int r = tid; r < R; r += RT) urow[r] = S.mu[r];
            }
            __syncthreads();
            if (s_err) break;
        }
        if (s_err) break;

        // ---- :221-234 end of sweep
        const int kt = s_kt, nocc = s_nocc;
        const int last = (tt == last_sweep);
        double *mean = scratch; // nocc x R
        // cluster means: thread per (occupied cluster, segment), cells in order
        for (int e = tid; e < nocc * R; e += RT) {
            const int k = S.occ[e / R], r = e % R;
            double s = 0.0;
            int c = 0;
            for (int i = 0; i < N; i++) {
                if (D.Zs[i] - 1 != k) continue;
                const double v = D.X[(size_t)i * R + r];
                if (!cs_isna(v)) { s += v; c++; }
            }
            mean[e] = c ? s / (double)c : 0.0;
        }
        __syncthreads();
        for (int pass = 0; pass < (last ? 2 : 1); pass++) {
            if (tid == 0) {
                int pos = s_pos;
                for (int e = 0; e < nocc; e++) {
                    const int k = S.occ[e];
                    const double *m = mean + (size_t)e * R;
                    double *u = D.U + (size_t)k * R;
                    if (pass == 1) {
                        for (int r = 0; r < R; r++) u[r] = m[r];
                        continue;
                    }
                    bool allzero = true;
                    for (int r = 0; r < R; r++)
                        if (!(m[r] == 0.0)) { allzero = false; break; }
                    if (allzero) {
                        for (int r = 0; r < R; r++) u[r] = 0.0;
                        continue;
                    }
                    const double sq = sqrt((double)S.np[k]);
                    for (int r = 0; r < R; r++) {
                        const double v = rnorm_r(POOL ? P.raw[s_cur] : S.mt, pos, m[r], S.sig[r] / sq);
                        u[r] = (v < 0.0) ? 0.0 : v;
                    }
                }
                s_pos = pos;
            }
            __syncthreads();
            if (POOL && pass == 0) pool_build(P, s_cur, true); // (the serial draws above moved the raw state)
            if (pass == 0 && last) continue;
            // sigmas (:234 / :269)
            for (int r = tid; r < R; r += RT) {
                double s = 0.0;
                int c = 0;
                for (int i = 0; i < N; i++) {
                    const double x = D.X[(size_t)i * R + r];
                    if (cs_isna(x)) continue;
                    c++;
                    const double d = x - D.U[(size_t)(D.Zs[i] - 1) * R + r];
                    if (!cs_isna(d)) s += d * d;
                }
                S.sig[r] = sqrt((1.0 / (double)c) * s);
                if (POOL) s_logsig[r] = log(S.sig[r]);
            }
            __syncthreads();
        }
        // ---- :238-245 record
        for (int i = tid; i < N; i += RT) D.Zall[(size_t)tt * N + i] = D.Zs[i];
        for (int r = tid; r < R; r += RT) D.sigma_all[(size_t)tt * R + r] = S.sig[r];
        {
            double acc = 0.0;
            for (int i = tid; i < N; i += RT)
                acc += cell_loglik(R, D.X + (size_t)i * R, D.U + (size_t)(D.Zs[i] - 1) * R, S.sig);
            s_red[tid] = acc;
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int i = 0; i < RT; i++) s += s_red[i];
                D.LL[tt] = s;
                D.kt_all[tt] = kt;
                D.u_off[tt] = s_uused;
                D.u_off[tt + 1] = s_uused + nocc;
                if (s_uused + nocc > D.urec_cap) s_err = CS_ERR_UCAP;
            }
            __syncthreads();
        }
        if (!s_err) {
            for (int e = tid; e < nocc * R; e += RT)
                D.u_rows[(size_t)(s_uused + e / R) * R + e % R] = D.U[(size_t)S.occ[e / R] * R + e % R];
            for (int e = tid; e < nocc; e += RT) D.u_labels[s_uused + e] = S.occ[e] + 1;
        }
        __syncthreads();
        if (tid == 0) s_uused += nocc;
        __syncthreads();
    }

    for (int k = tid; k < lcap; k += RT) D.np[k] = S.np[k];
    for (int r = tid; r < R; r += RT) {
        D.sig[r] = S.sig[r];
        D.isig[r] = 1.0 / S.sig[r];
    }
    for (int k = tid; k < 624; k += RT) D.mt[k] = POOL ? P.raw[s_cur][k] : S.mt[k];
    if (tid == 0) {
        D.scal->kt = s_kt;
        D.scal->mt_pos = s_pos;
        D.scal->n_reject += s_rej;
        D.scal->u_used = s_uused;
        if (s_err) D.scal->err = s_err;
    }
}

} // namespace

static size_t rmode_smem(int lcap, int R, bool pool)
{
    size_t b = sizeof(double) * (2 * ((size_t)lcap + 1) + 3 * (size_t)R) + sizeof(int) * (2 * ((size_t)lcap + 1) + 2 * (size_t)lcap) +
               sizeof(uint32_t) * 624 + 64;
    if (pool) b += sizeof(double) * ((size_t)R + (size_t)(RT / 32) * R + 2 * 624) + sizeof(int) * (size_t)R + sizeof(uint32_t) * 624;
    return b;
}
size_t cs_rmode_smem_bytes(int lcap, int R) { return rmode_smem(lcap, R, false); }

// sweeps [t0, t1) of the R-compatible chain; last_sweep = index of the run's final sweep.  The pooled
// kernel needs the 2R + 1 draws of a cell inside two Mersenne-Twister blocks (R <= 311) and room in
// shared memory; otherwise (or with CS_RMODE_POOL=0) thread 0 walks the stream draw by draw.
int cs_rmode_run(const ChainDev &D, int lcap, int t0, int t1, int last_sweep, double *scratch, cudaStream_t st)
{
    const bool want_pool = !(getenv("CS_RMODE_POOL") && atoi(getenv("CS_RMODE_POOL")) == 0);
    const bool pool = want_pool && 2 * D.R + 1 < 624 && rmode_smem(lcap, D.R, true) <= 227 * 1024;
    const size_t smem = rmode_smem(lcap, D.R, pool);
    if (smem > 227 * 1024) {
        cs_set_error("R-compatible sampler: label capacity %d with R=%d needs %zu B of shared memory", lcap, D.R, smem);
        return CS_ERR_UNSUPPORTED;
    }
    auto kern = pool ? gibbs_r_kernel<true> : gibbs_r_kernel<false>;
    CS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<1, RT, smem, st>>>(D, lcap, t0, t1, last_sweep, scratch);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// base R's set.seed() scrambling for the Mersenne-Twister (host side)
void cs_r_set_seed(uint32_t seed, uint32_t mt[624])
{
    uint32_t s = seed;
    for (int j = 0; j < 50; j++) s = 69069u * s + 1u;
    s = 69069u * s + 1u; // the slot holding the position
    for (int j = 0; j < 624; j++) {
        s = 69069u * s + 1u;
        mt[j] = s;
    }
}
