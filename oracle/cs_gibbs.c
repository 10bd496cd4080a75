/*
 * cs_gibbs.c — CPU restatement of the nested-CRP collapsed Gibbs sampler.
 *
 * TEST INFRASTRUCTURE ONLY (see cs_oracle.h).
 *
 * Follows R/BayesNonparCluster.R:21-289 of the reference statement by
 * statement (line numbers cited inline).  Two RNG modes:
 *
 *   CSO_RNG_R      base-R Mersenne Twister consumed in exactly the reference's
 *                  order; reproduces the shipped nonpara_clustering*.rds
 *                  fixtures bit-for-bit (Zall) — this is what pins the oracle.
 *   CSO_RNG_PHILOX the product's counter-based stream.  Same model, same
 *                  sequential within-sweep order, same quirks; only the map
 *                  from random words to draws differs (specified in DESIGN.md
 *                  "Philox mode").  The CUDA path is checked label-for-label
 *                  against this mode.
 *
 * Sums that base R accumulates in `long double` (sum, colMeans, colSums) are
 * accumulated in `long double` here as well.
 */
#include "cs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LN_SQRT_2PI 0.918938533204672741780329736406

static inline int is_na(double x) { return x != x; }

/* dnorm(x, mu, sigma, log = TRUE) as base R computes it */
static inline double dnorm_log(double x, double mu, double sigma)
{
    if (is_na(x) || is_na(mu) || is_na(sigma)) return x + mu + sigma;
    if (sigma < 0) return NAN;
    if (!isfinite(sigma)) return -INFINITY;
    if (!isfinite(x) && mu == x) return NAN;
    if (sigma == 0) return (x == mu) ? INFINITY : -INFINITY;
    double z = (x - mu) / sigma;
    if (!isfinite(z)) return -INFINITY;
    z = fabs(z);
    return -(LN_SQRT_2PI + 0.5 * z * z + log(sigma));
}

/* sum(dnorm(x_i, u, sigma, log=T), na.rm=T) over one cell's R segments */
static double cell_loglik(int R, const double *x, const double *u, const double *sigma)
{
    long double s = 0.0L;
    for (int r = 0; r < R; r++) {
        double t = dnorm_log(x[r], u[r], sigma[r]);
        if (!is_na(t)) s += t;
    }
    return (double)s;
}

void cso_loglik_matrix(int N, int R, int K, const double *X, const double *U,
                       const double *sigma, double *LL)
{
    for (int i = 0; i < N; i++)
        for (int k = 0; k < K; k++)
            LL[(size_t)i * K + k] = cell_loglik(R, X + (size_t)i * R, U + (size_t)k * R, sigma);
}

double cso_total_loglik(int N, int R, const double *X, const double *U, int K,
                        const int *Z, const double *sigma)
{
    (void)K;
    long double s = 0.0L;
    for (int i = 0; i < N; i++)
        s += cell_loglik(R, X + (size_t)i * R, U + (size_t)(Z[i] - 1) * R, sigma);
    return (double)s;
}

/* ---- Philox helpers (DESIGN.md "Philox mode") ---- */
enum { PH_PROPOSAL = 1, PH_ASSIGN = 2, PH_RNORM = 3 };

static inline double u53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) / 9007199254740992.0;
}

static void philox_draw(uint32_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t tag,
                        uint32_t attempt, double *ua, double *ub)
{
    uint32_t ctr[4] = {c0, c1, c2, (tag << 24) | (attempt & 0xFFFFFFu)};
    uint32_t key[2] = {seed, 0x434C4F4Eu};
    uint32_t o[4];
    cso_philox4x32(ctr, key, o);
    *ua = u53(o[0], o[1]);
    *ub = u53(o[2], o[3]);
}

/* Philox mode (DESIGN.md "Philox mode"): the squared standardised distance of a cell to a
 * row is accumulated in FIXED POINT — every segment contributes
 *     llrint( ((x - u) * (65536 / sigma))^2 )          (2^-32 resolution)
 * to a 64-bit integer sum, and q = sum * 2^-32.  Integer addition is associative, so the
 * value does not depend on the order in which a parallel machine adds the segments and can
 * be updated exactly when a single term changes; the CUDA path relies on both. */
static double q_fixed(int R, const double *x, const double *u, const double *isig16)
{
    long long s = 0;
    for (int r = 0; r < R; r++) {
        if (is_na(x[r])) continue;
        double d = (x[r] - u[r]) * isig16[r];
        double t = d * d;
        if (!(t < 4611686018427387904.0)) t = 4611686018427387904.0; /* 2^62 clamp (also NaN) */
        s += llrint(t);
    }
    return (double)s * (1.0 / 4294967296.0);
}

typedef struct {
    int cap;      /* allocated rows */
    int kt;       /* rows in use == number of labels ever created */
    double *U;    /* cap * R */
    double *np;   /* counts */
} state_t;

static int grow(state_t *s, int R)
{
    int ncap = s->cap * 2;
    double *U = (double *)realloc(s->U, sizeof(double) * (size_t)ncap * R);
    if (!U) return -1;
    s->U = U;
    double *np = (double *)realloc(s->np, sizeof(double) * (size_t)ncap);
    if (!np) return -1;
    s->np = np;
    s->cap = ncap;
    return 0;
}

/* end-of-sweep statistics shared by the noisy update and the final override:
 * meanX[k][r] = colMeans(X[Z==k,], na.rm=T) with NaN -> 0   (:223-224) */
static void cluster_means(int N, int R, const double *X, const int *Z, int kt,
                          long double *acc, int *cnt, double *mean)
{
    memset(acc, 0, sizeof(long double) * (size_t)kt * R);
    memset(cnt, 0, sizeof(int) * (size_t)kt * R);
    for (int i = 0; i < N; i++) {
        int k = Z[i] - 1;
        const double *x = X + (size_t)i * R;
        for (int r = 0; r < R; r++)
            if (!is_na(x[r])) { acc[(size_t)k * R + r] += x[r]; cnt[(size_t)k * R + r]++; }
    }
    for (size_t e = 0; e < (size_t)kt * R; e++)
        mean[e] = cnt[e] ? (double)(acc[e] / (long double)cnt[e]) : 0.0;
}

/* sigmas = sqrt((1/colSums(!is.na(X))) * colSums((X - U[Z,])^2, na.rm=T))   (:234) */
static void update_sigmas(int N, int R, const double *X, const int *Z, const double *U,
                          double *sig)
{
    for (int r = 0; r < R; r++) {
        long double s = 0.0L;
        int c = 0;
        for (int i = 0; i < N; i++) {
            double x = X[(size_t)i * R + r];
            if (is_na(x)) continue;
            c++;
            double d = x - U[(size_t)(Z[i] - 1) * R + r];
            if (!is_na(d)) s += d * d;
        }
        sig[r] = sqrt((1.0 / (double)c) * (double)s);
    }
}

int cso_gibbs(int N, int R, const double *X, int K0, const double *U0in, const int *Z0,
              const double *sigmas0, double alpha, double beta, int niter, uint32_t seed,
              int rng_mode, int flags, int *Zall, double *sigma_all, double *LL, double *L0,
              int *kt_all, int64_t ucap_rows, int64_t *u_off, int *u_labels, double *u_rows,
              int64_t *n_reject)
{
    if (N <= 0 || R <= 0 || K0 <= 0 || niter <= 0) return 1;
    for (int i = 0; i < N; i++)
        if (Z0[i] < 1 || Z0[i] > K0) return 2; /* the reference errors on such input */

    int rc = 0;
    state_t st;
    st.cap = (K0 > 64 ? K0 : 64);
    st.U = (double *)malloc(sizeof(double) * (size_t)st.cap * R);
    st.np = (double *)malloc(sizeof(double) * (size_t)st.cap);
    int *Zt = (int *)malloc(sizeof(int) * (size_t)N);
    double *sig = (double *)malloc(sizeof(double) * (size_t)R);
    double *mu_new = (double *)malloc(sizeof(double) * (size_t)R);
    double *isig = (double *)malloc(sizeof(double) * (size_t)R);
    double *prob = NULL, *pk = NULL, *cumc = NULL;
    int *perm = NULL, *permc = NULL;
    int pcap = 0;
    long double *acc = NULL;
    int *cnt = NULL;
    double *mean = NULL;
    cso_mt g;
    int64_t rejects = 0, uused = 0;

    /* L0 (:133) */
    {
        long double s = 0.0L;
        for (int i = 0; i < N; i++)
            s += cell_loglik(R, X + (size_t)i * R, U0in + (size_t)(Z0[i] - 1) * R, sigmas0);
        *L0 = (double)s;
    }
    memcpy(Zt, Z0, sizeof(int) * (size_t)N);
    memcpy(sig, sigmas0, sizeof(double) * (size_t)R);

    /* :141-147  one distinct label -> single unnamed row of ones; the
     * `identical()` rejection test at :174 is live only in that case and only
     * until the end of sweep 1 (Ut gains column names at :238). */
    int single = 1;
    for (int i = 1; i < N; i++)
        if (Z0[i] != Z0[0]) { single = 0; break; }
    if (single) {
        if (Z0[0] != 1) { rc = 3; goto done; } /* npoints would be NA in the reference */
        st.kt = 1;
        for (int r = 0; r < R; r++) st.U[r] = 1.0;
    } else {
        st.kt = K0;
        memcpy(st.U, U0in, sizeof(double) * (size_t)K0 * R);
    }
    /* :150-151 */
    for (int k = 0; k < st.kt; k++) st.np[k] = 0;
    for (int i = 0; i < N; i++) st.np[Zt[i] - 1] += 1;

    cso_set_seed(&g, seed); /* :156 */

    for (int tt = 0; tt < niter; tt++) {
        int reject_live = single && tt == 0;
        for (int r = 0; r < R; r++) isig[r] = (1.0 / sig[r]) * 65536.0;
        for (int ii = 0; ii < N; ii++) {
            const double *x = X + (size_t)ii * R;
            int kt = st.kt;
            if (kt + 1 > pcap) {
                pcap = 2 * (kt + 1) + 64;
                prob = (double *)realloc(prob, sizeof(double) * pcap);
                pk = (double *)realloc(pk, sizeof(double) * pcap);
                cumc = (double *)realloc(cumc, sizeof(double) * pcap);
                perm = (int *)realloc(perm, sizeof(int) * pcap);
                permc = (int *)realloc(permc, sizeof(int) * pcap);
            }
            /* ---- :165-175 proposal of a new cluster's per-segment states ---- */
            uint32_t attempt = 0;
            int have_sorted = 0;
            for (;;) {
                for (int r = 0; r < R; r++) {
                    double tmpr;
                    int indr; /* 1-based index into c(Ut[,r], tmpr) */
                    if (rng_mode == CSO_RNG_R) {
                        tmpr = 0.3 + (2.5 - 0.3) * cso_unif_rand(&g); /* :167 */
                        if ((flags & CSO_FLAG_SORT_ONCE)) {
                            if (!have_sorted) {
                                double sum = 0.0;
                                for (int k = 0; k < kt; k++) { cumc[k] = st.np[k]; if (cumc[k] > 0) sum += cumc[k]; }
                                cumc[kt] = alpha; if (alpha > 0) sum += alpha;
                                for (int k = 0; k <= kt; k++) { cumc[k] /= sum; permc[k] = k + 1; }
                                cso_revsort(cumc, permc, kt + 1);
                                for (int k = 1; k <= kt; k++) cumc[k] += cumc[k - 1];
                                have_sorted = 1;
                            }
                            double u = cso_unif_rand(&g);
                            int j = 0;
                            while (j < kt && u > cumc[j]) j++;
                            indr = permc[j];
                        } else {
                            for (int k = 0; k < kt; k++) prob[k] = st.np[k];
                            prob[kt] = alpha;
                            indr = cso_sample_int_prob(&g, kt + 1, prob, perm); /* :170 */
                        }
                        mu_new[r] = (indr <= kt) ? st.U[(size_t)(indr - 1) * R + r] : tmpr;
                    } else {
                        double ua, ub;
                        philox_draw(seed, (uint32_t)ii, (uint32_t)r, (uint32_t)tt, PH_PROPOSAL,
                                    attempt, &ua, &ub);
                        tmpr = fma(2.5 - 0.3, ua, 0.3);
                        double v = ub * ((double)N + alpha);
                        if (v < (double)N) {
                            int j = (int)v;
                            if (j > N - 1) j = N - 1;
                            mu_new[r] = st.U[(size_t)(Zt[j] - 1) * R + r];
                        } else {
                            mu_new[r] = tmpr;
                        }
                    }
                }
                if (!reject_live) break;
                int dup = 0; /* :174 any(apply(Ut,1,identical,mu_new)) */
                for (int k = 0; k < kt && !dup; k++) {
                    const double *u = st.U + (size_t)k * R;
                    int same = 1;
                    for (int r = 0; r < R; r++) {
                        if (is_na(u[r]) && is_na(mu_new[r])) continue;
                        if (!(u[r] == mu_new[r])) { same = 0; break; }
                    }
                    dup = same;
                }
                if (!dup) break;
                rejects++;
                attempt++;
            }

            /* ---- :188 remove the cell from its cluster ---- */
            st.np[Zt[ii] - 1] -= 1;

            int z;
            if (rng_mode == CSO_RNG_R) {
                /* :193-207 */
                double mx = -INFINITY;
                for (int k = 0; k < kt; k++) {
                    if (st.np[k] != 0) {
                        pk[k] = cell_loglik(R, x, st.U + (size_t)k * R, sig) + log(st.np[k]);
                        if (pk[k] > mx) mx = pk[k];
                    }
                }
                double pnew0 = cell_loglik(R, x, mu_new, sig) + log(beta);
                if (pnew0 > mx) mx = pnew0;
                for (int k = 0; k < kt; k++) prob[k] = (st.np[k] != 0) ? exp(pk[k] - mx) : 0.0;
                prob[kt] = exp(pnew0 - mx);
                z = cso_sample_int_prob(&g, kt + 1, prob, perm);
            } else {
                /* Philox mode: p_k ∝ n_k exp(-Q_k/2), p_new ∝ beta exp(-Q_new/2) with
                 * Q = Σ_r ((x-u)·(1/σ_r))² over non-NA segments; the Σ(log σ + ½log2π)
                 * term is common to every candidate and cancels. */
                double qmin = INFINITY;
                for (int k = 0; k <= kt; k++) {
                    if (k < kt && st.np[k] == 0) continue;
                    const double *u = (k < kt) ? st.U + (size_t)k * R : mu_new;
                    double q = q_fixed(R, x, u, isig);
                    pk[k] = q;
                    if (q < qmin) qmin = q;
                }
                double tot = 0.0;
                for (int k = 0; k <= kt; k++) {
                    if (k < kt && st.np[k] == 0) { prob[k] = 0.0; continue; }
                    double w = exp(-0.5 * (pk[k] - qmin));
                    prob[k] = (k < kt ? st.np[k] : beta) * w;
                    tot += prob[k];
                }
                double ua, ub;
                philox_draw(seed, (uint32_t)ii, 0u, (uint32_t)tt, PH_ASSIGN, 0u, &ua, &ub);
                double t = ua * tot, c = 0.0;
                z = 0;
                int lastpos = 0;
                for (int k = 0; k <= kt; k++) {
                    if (prob[k] <= 0.0) continue;
                    c += prob[k];
                    lastpos = k + 1;
                    if (c > t) { z = k + 1; break; }
                }
                if (z == 0) z = lastpos;
            }
            Zt[ii] = z;
            if (z > kt) { /* :209-212 open a new cluster */
                if (st.kt + 1 > st.cap && grow(&st, R)) { rc = 4; goto done; }
                memcpy(st.U + (size_t)st.kt * R, mu_new, sizeof(double) * (size_t)R);
                st.np[st.kt] = 1;
                st.kt++;
            } else {
                st.np[z - 1] += 1; /* :214 */
            }
        }

        /* ---- :221-232 end of sweep: redraw cluster means ---- */
        int kt = st.kt;
        acc = (long double *)realloc(acc, sizeof(long double) * (size_t)kt * R);
        cnt = (int *)realloc(cnt, sizeof(int) * (size_t)kt * R);
        mean = (double *)realloc(mean, sizeof(double) * (size_t)kt * R);
        cluster_means(N, R, X, Zt, kt, acc, cnt, mean);
        int last = (tt == niter - 1);
        /* run the noisy update first (it consumes random numbers even on the
         * last sweep), then :260-277 overwrites the last sweep noise-free */
        for (int pass = 0; pass < (last ? 2 : 1); pass++) {
            for (int k = 0; k < kt; k++) {
                double *u = st.U + (size_t)k * R;
                if (st.np[k] == 0) {
                    for (int r = 0; r < R; r++) u[r] = NAN;
                    continue;
                }
                const double *m = mean + (size_t)k * R;
                if (pass == 1) { memcpy(u, m, sizeof(double) * (size_t)R); continue; }
                int allzero = 1;
                for (int r = 0; r < R; r++) if (!(m[r] == 0)) { allzero = 0; break; }
                if (allzero) { for (int r = 0; r < R; r++) u[r] = 0.0; continue; } /* :229 */
                double sq = sqrt(st.np[k]);
                for (int r = 0; r < R; r++) {
                    double sd = sig[r] / sq, v;
                    if (rng_mode == CSO_RNG_R) {
                        v = cso_rnorm(&g, m[r], sd); /* :226 */
                    } else if (sd == 0.0) {
                        v = m[r];
                    } else {
                        double ua, ub;
                        philox_draw(seed, (uint32_t)(k + 1), (uint32_t)r, (uint32_t)tt, PH_RNORM, 0u, &ua, &ub);
                        v = m[r] + sd * cso_qnorm(ua);
                    }
                    u[r] = (v < 0) ? 0.0 : v; /* :227 */
                }
            }
            if (pass == 0 && last) continue; /* sigmas of the noisy pass are overwritten too */
            update_sigmas(N, R, X, Zt, st.U, sig); /* :234 / :269 */
        }

        /* ---- :238-245 record ---- */
        memcpy(Zall + (size_t)tt * N, Zt, sizeof(int) * (size_t)N);
        memcpy(sigma_all + (size_t)tt * R, sig, sizeof(double) * (size_t)R);
        kt_all[tt] = kt;
        u_off[tt] = uused;
        for (int k = 0; k < kt; k++) {
            if (st.np[k] == 0) continue;
            if (uused < ucap_rows) {
                u_labels[uused] = k + 1;
                memcpy(u_rows + (size_t)uused * R, st.U + (size_t)k * R, sizeof(double) * (size_t)R);
            }
            uused++;
        }
        LL[tt] = cso_total_loglik(N, R, X, st.U, kt, Zt, sig);
    }
    u_off[niter] = uused;
    if (uused > ucap_rows) rc = 5;

done:
    if (n_reject) *n_reject = rejects;
    free(st.U); free(st.np); free(Zt); free(sig); free(mu_new); free(isig);
    free(prob); free(pk); free(cumc); free(perm); free(permc);
    free(acc); free(cnt); free(mean);
    return rc;
}

/* ------------------------------------------------------------------ */
/* posterior tally: R/AssignCluster.R:37                                */
/*   apply(Zall, 2, function(x) which.max(table(x)[as.character(1:max(Zall))])) */
/* most frequent label per cell over the kept sweeps, ties -> smallest  */
/* ------------------------------------------------------------------ */
void cso_majority_vote(int T, int N, const int *Zall, int *zmaj)
{
    int mx = 0;
    for (size_t e = 0; e < (size_t)T * N; e++)
        if (Zall[e] > mx) mx = Zall[e];
    int *cnt = (int *)calloc((size_t)mx + 1, sizeof(int));
    for (int i = 0; i < N; i++) {
        for (int t = 0; t < T; t++) cnt[Zall[(size_t)t * N + i]]++;
        int best = 0, bc = 0;
        for (int l = 1; l <= mx; l++)
            if (cnt[l] > bc) { bc = cnt[l]; best = l; }
        zmaj[i] = best;
        for (int t = 0; t < T; t++) cnt[Zall[(size_t)t * N + i]] = 0;
    }
    free(cnt);
}
