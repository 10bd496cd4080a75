/*
 * cs_allele.c — CPU restatement of the allele (coverage + major-allele-fraction) sampler.
 *
 * TEST INFRASTRUCTURE ONLY (see cs_oracle.h).
 *
 * Follows R/BayesNonparAlleleCluster.R:20-307 and R/genotype_neighbor.R:7-40 statement
 * by statement.  PARITY UNPINNED: nothing in the reference calls this function and no
 * fixture ships for it, so this restatement is checked only for self-consistency and
 * against the CUDA path.  Quirks kept on purpose:
 *   - npoints = as.numeric(table(Zt)) is not aligned to 1:nrow(U0) (:148): every label
 *     1..K0 must occur in Z0, otherwise the reference fails in sample.int;
 *   - sample(x, 1) with length(x) == 1 samples from 1:x (:161);
 *   - identical(Ut row, mu_new) is never TRUE (Ut carries names), so no rejection (:174);
 *   - genotype_neighbor is always called with maxcp = 6 (:223,:266);
 *   - emptied clusters keep all-zero rows (not NA) in Uall / Ucov / Uallele (:211-213).
 *
 * rng_mode CSO_RNG_R consumes base R's Mersenne-Twister stream in the reference's order;
 * CSO_RNG_PHILOX replaces that stream, draw for draw, by Philox4x32-10 words with a
 * running counter (ctr = (n, 0, 0, 0x05000000), key = (seed, 'CLON'), uniform = u53 of
 * words 0,1) — the algorithm is otherwise identical.
 */
#include "cs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LN_SQRT_2PI 0.918938533204672741780329736406

static inline int is_na(double x) { return x != x; }

typedef struct {
    int mode;
    cso_mt mt;
    uint32_t seed;
    uint64_t n;
} arng;

static double a_unif(arng *g)
{
    if (g->mode == CSO_RNG_R) return cso_unif_rand(&g->mt);
    uint32_t ctr[4] = {(uint32_t)g->n, (uint32_t)(g->n >> 32), 0u, 0x05000000u};
    uint32_t key[2] = {g->seed, 0x434C4F4Eu}, o[4];
    g->n++;
    cso_philox4x32(ctr, key, o);
    return ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6) + 0.5) / 9007199254740992.0;
}

/* R_unif_index(n), sample.kind = "Rejection" */
static int a_unif_index(arng *g, int n)
{
    if (n <= 0) return 0;
    int bits = (int)ceil(log2((double)n));
    for (;;) {
        int64_t v = 0;
        for (int k = 0; k <= bits; k += 16) v = 65536 * v + (int)floor(a_unif(g) * 65536.0);
        if (bits < 64) v &= (((int64_t)1 << bits) - 1);
        if ((double)v < (double)n) return (int)v;
    }
}

static double a_rnorm(arng *g, double mean, double sd)
{
    if (sd == 0.0 || !isfinite(mean)) return mean;
    double u = a_unif(g);
    u = (double)(int)(134217728.0 * u) + a_unif(g);
    return mean + sd * cso_qnorm(u / 134217728.0);
}

/* sample.int(n, 1, prob = p) on a pre-sorted, cumulated vector */
static int pick_sorted(arng *g, int n, const double *cum, const int *perm)
{
    double u = a_unif(g);
    int j = 0;
    while (j < n - 1 && u > cum[j]) j++;
    return perm[j];
}

static void sort_probs(int n, double *p, int *perm)
{
    double sum = 0.0;
    for (int i = 0; i < n; i++)
        if (p[i] > 0) sum += p[i];
    for (int i = 0; i < n; i++) { p[i] /= sum; perm[i] = i + 1; }
    cso_revsort(p, perm, n);
    for (int i = 1; i < n; i++) p[i] += p[i - 1];
}

int cso_genotype_grid(int maxcp, double *mu)
{
    int g = 0;
    for (int ii = 1; ii <= maxcp; ii++)
        for (int j = 0; j <= ii; j++) {
            if (mu) { mu[2 * g] = 0.5 * ii; mu[2 * g + 1] = (j == 0) ? 0.0 : (double)j / (double)ii; }
            g++;
        }
    return g;
}

/* nearest grid state (1-based) of one (coverage, allele) point: R/genotype_neighbor.R:18-24 */
int cso_genotype_neighbor(double cov, double allele, int maxcp)
{
    double mu[2 * 64];
    int G = cso_genotype_grid(maxcp, mu), best = 0;
    double bd = INFINITY;
    for (int i = 0; i < G; i++) {
        double a = cov - mu[2 * i], b = allele - mu[2 * i + 1];
        double d = sqrt(a * a + b * b);
        if (d < bd) { bd = d; best = i; }
    }
    return best + 1;
}

static inline double dnorm_log(double x, double m, double s)
{
    double z = fabs((x - m) / s);
    return -(LN_SQRT_2PI + 0.5 * z * z + log(s));
}

/* sum over non-NA segments of dnorm(x, mean, sd, log=T) */
static double modality_ll(int R, const double *x, const double *mean, const double *sd)
{
    long double s = 0.0L;
    for (int r = 0; r < R; r++)
        if (!is_na(x[r])) s += dnorm_log(x[r], mean[r], sd[r]);
    return (double)s;
}

int cso_gibbs_allele(int N, int R, const double *Xcov_in, const double *Xall, int K0, const int *U0,
                     const int *Z0, const double *sig0_cov, const double *sig0_all, double alpha,
                     double beta, int niter, uint32_t seed, int maxcp, int rng_mode, int *Zall,
                     double *sig_cov_all, double *sig_all_all, double *LL, double *L0, int *kt_all,
                     int64_t ucap_rows, int64_t *u_off, int *u_labels, int *u_state, double *u_cov,
                     double *u_allele)
{
    if (N <= 0 || R <= 0 || K0 <= 0 || niter <= 0 || maxcp < 1 || maxcp > 9) return 1;
    double grid[2 * 64];
    const int G = cso_genotype_grid(maxcp, grid);
    for (int e = 0; e < K0 * R; e++)
        if (U0[e] < 1 || U0[e] > G) return 2;
    int rc = 0, cap = K0 > 64 ? K0 : 64, kt = K0, pcap = 0;
    double *Xcov = (double *)malloc(sizeof(double) * (size_t)N * R);
    for (size_t e = 0; e < (size_t)N * R; e++) Xcov[e] = Xcov_in[e] > 3.0 ? 3.0 : Xcov_in[e]; /* :37 pmin */
    int *Ut = (int *)malloc(sizeof(int) * (size_t)cap * R);
    double *np = (double *)calloc((size_t)cap, sizeof(double));
    int *Zt = (int *)malloc(sizeof(int) * (size_t)N);
    double *sc = (double *)malloc(sizeof(double) * R), *sa = (double *)malloc(sizeof(double) * R);
    double *mc = (double *)malloc(sizeof(double) * R), *ma = (double *)malloc(sizeof(double) * R);
    int *mu_new = (int *)malloc(sizeof(int) * R);
    double *prob = NULL, *cumc = NULL, *pk = NULL;
    int *perm = NULL, *permc = NULL;
    double *Utc = NULL, *Uta = NULL;
    int64_t uused = 0;
    arng g;
    g.mode = rng_mode; g.seed = seed; g.n = 0;
    memcpy(Ut, U0, sizeof(int) * (size_t)K0 * R);
    memcpy(Zt, Z0, sizeof(int) * (size_t)N);
    memcpy(sc, sig0_cov, sizeof(double) * R);
    memcpy(sa, sig0_all, sizeof(double) * R);
    for (int i = 0; i < N; i++) {
        if (Z0[i] < 1 || Z0[i] > K0) { rc = 2; goto done; }
        np[Z0[i] - 1] += 1;
    }
    for (int k = 0; k < K0; k++)
        if (np[k] == 0) { rc = 3; goto done; } /* table(Zt) shorter than nrow(U0): sample.int fails */
    { /* L0 (:131-132): U0a has 0 -> 0.05 and 1 -> 0.95 */
        long double s = 0.0L;
        for (int i = 0; i < N; i++) {
            const int *u = U0 + (size_t)(Z0[i] - 1) * R;
            for (int r = 0; r < R; r++) {
                mc[r] = grid[2 * (u[r] - 1)];
                double a = grid[2 * (u[r] - 1) + 1];
                ma[r] = a == 0.0 ? 0.05 : (a == 1.0 ? 0.95 : a);
            }
            s += modality_ll(R, Xcov + (size_t)i * R, mc, sig0_cov);
        }
        long double s2 = 0.0L;
        for (int i = 0; i < N; i++) {
            const int *u = U0 + (size_t)(Z0[i] - 1) * R;
            for (int r = 0; r < R; r++) {
                double a = grid[2 * (u[r] - 1) + 1];
                ma[r] = a == 0.0 ? 0.05 : (a == 1.0 ? 0.95 : a);
            }
            s2 += modality_ll(R, Xall + (size_t)i * R, ma, sig0_all);
        }
        *L0 = (double)s + (double)s2;
    }
    if (rng_mode == CSO_RNG_R) cso_set_seed(&g.mt, seed); /* :150 */

    for (int tt = 0; tt < niter; tt++) {
        int sorted_valid = 0;
        for (int ii = 0; ii < N; ii++) {
            if (kt + 1 > pcap) {
                pcap = 2 * (kt + 1) + 64;
                prob = (double *)realloc(prob, sizeof(double) * pcap);
                cumc = (double *)realloc(cumc, sizeof(double) * pcap);
                pk = (double *)realloc(pk, sizeof(double) * pcap);
                perm = (int *)realloc(perm, sizeof(int) * pcap);
                permc = (int *)realloc(permc, sizeof(int) * pcap);
            }
            /* :158-166 proposal (the sorted prob vector is the same for all R segments) */
            if (!sorted_valid) {
                for (int k = 0; k < kt; k++) cumc[k] = np[k];
                cumc[kt] = alpha;
                sort_probs(kt + 1, cumc, permc);
                sorted_valid = 1;
            }
            for (int r = 0; r < R; r++) {
                uint64_t used = 0;
                for (int k = 0; k < kt; k++) {
                    int s = Ut[(size_t)k * R + r];
                    if (s >= 1 && s <= G) used |= (uint64_t)1 << (s - 1);
                }
                int cand[64], nc = 0;
                for (int s = 1; s <= G; s++)
                    if (!((used >> (s - 1)) & 1)) cand[nc++] = s;
                int tmpr;
                if (nc == 0) { rc = 6; goto done; }              /* sample(integer(0), 1) errors */
                if (nc == 1) tmpr = 1 + a_unif_index(&g, cand[0]); /* sample(x, 1) == sample(1:x, 1) */
                else tmpr = cand[a_unif_index(&g, nc)];
                int indr = pick_sorted(&g, kt + 1, cumc, permc);
                mu_new[r] = (indr <= kt) ? Ut[(size_t)(indr - 1) * R + r] : tmpr;
            }
            np[Zt[ii] - 1] -= 1; /* :168 */
            const double *xc = Xcov + (size_t)ii * R, *xa = Xall + (size_t)ii * R;
            double mx = -INFINITY;
            for (int k = 0; k <= kt; k++) {
                if (k < kt && np[k] == 0) continue;
                const int *u = (k < kt) ? Ut + (size_t)k * R : mu_new;
                for (int r = 0; r < R; r++) {
                    int s = u[r];
                    mc[r] = (s >= 1 && s <= G) ? grid[2 * (s - 1)] : NAN;
                    ma[r] = (s >= 1 && s <= G) ? grid[2 * (s - 1) + 1] : NAN;
                }
                pk[k] = modality_ll(R, xc, mc, sc) + modality_ll(R, xa, ma, sa) + log(k < kt ? np[k] : beta);
                if (pk[k] > mx) mx = pk[k];
            }
            for (int k = 0; k <= kt; k++) prob[k] = (k < kt && np[k] == 0) ? 0.0 : exp(pk[k] - mx);
            sort_probs(kt + 1, prob, perm);
            int z = pick_sorted(&g, kt + 1, prob, perm); /* :190 */
            int zold = Zt[ii];
            Zt[ii] = z;
            if (z > kt) {
                if (kt + 1 > cap) {
                    cap *= 2;
                    Ut = (int *)realloc(Ut, sizeof(int) * (size_t)cap * R);
                    np = (double *)realloc(np, sizeof(double) * (size_t)cap);
                }
                memcpy(Ut + (size_t)kt * R, mu_new, sizeof(int) * (size_t)R);
                np[kt] = 1;
                kt++;
                sorted_valid = 0;
            } else {
                np[z - 1] += 1;
                if (z != zold) sorted_valid = 0;
            }
        }
        /* :203-231 end of sweep */
        Utc = (double *)realloc(Utc, sizeof(double) * (size_t)kt * R);
        Uta = (double *)realloc(Uta, sizeof(double) * (size_t)kt * R);
        const int last = (tt == niter - 1);
        for (int pass = 0; pass < (last ? 2 : 1); pass++) {
            for (int k = 0; k < kt; k++) {
                int *us = Ut + (size_t)k * R;
                double *uc = Utc + (size_t)k * R, *ua = Uta + (size_t)k * R;
                for (int r = 0; r < R; r++) { us[r] = 0; uc[r] = 0.0; ua[r] = 0.0; }
                if (np[k] == 0) continue;
                for (int r = 0; r < R; r++) {
                    long double s1 = 0.0L, s2 = 0.0L;
                    int c1 = 0, c2 = 0;
                    for (int i = 0; i < N; i++) {
                        if (Zt[i] - 1 != k) continue;
                        double a = Xcov[(size_t)i * R + r], b = Xall[(size_t)i * R + r];
                        if (!is_na(a)) { s1 += a; c1++; }
                        if (!is_na(b)) { s2 += b; c2++; }
                    }
                    mc[r] = c1 ? (double)(s1 / c1) : 0.0;
                    ma[r] = c2 ? (double)(s2 / c2) : 0.0;
                }
                if (pass == 0) {
                    int allzero = 1;
                    for (int r = 0; r < R; r++) if (!(mc[r] == 0)) { allzero = 0; break; }
                    if (allzero) {
                        for (int r = 0; r < R; r++) { mc[r] = 0.0; ma[r] = 0.0; }
                    } else {
                        double sq = sqrt(np[k]);
                        for (int r = 0; r < R; r++) { double v = a_rnorm(&g, mc[r], sc[r] / sq); mc[r] = v < 0 ? 0.0 : v; }
                        for (int r = 0; r < R; r++) { double v = a_rnorm(&g, ma[r], sa[r] / sq); ma[r] = v < 0 ? 0.0 : v; }
                    }
                }
                for (int r = 0; r < R; r++) {
                    us[r] = cso_genotype_neighbor(mc[r], ma[r], 6);
                    uc[r] = mc[r];
                    ua[r] = ma[r];
                }
            }
            if (pass == 0 && last) continue;
            for (int r = 0; r < R; r++) { /* :226-231 */
                long double s1 = 0.0L, s2 = 0.0L;
                int c1 = 0, c2 = 0;
                for (int i = 0; i < N; i++) {
                    double a = Xcov[(size_t)i * R + r], b = Xall[(size_t)i * R + r];
                    size_t o = (size_t)(Zt[i] - 1) * R + r;
                    if (!is_na(a)) { c1++; double d = a - Utc[o]; s1 += d * d; }
                    if (!is_na(b)) { c2++; double d = b - Uta[o]; s2 += d * d; }
                }
                sc[r] = sqrt(1.0 / (double)c1 * (double)s1);
                sa[r] = sqrt(1.0 / (double)c2 * (double)s2);
                if (sc[r] == 0) sc[r] = 0.01;
                if (sa[r] == 0) sa[r] = 0.01;
            }
        }
        memcpy(Zall + (size_t)tt * N, Zt, sizeof(int) * (size_t)N);
        memcpy(sig_cov_all + (size_t)tt * R, sc, sizeof(double) * R);
        memcpy(sig_all_all + (size_t)tt * R, sa, sizeof(double) * R);
        kt_all[tt] = kt;
        u_off[tt] = uused;
        for (int k = 0; k < kt; k++) {
            if (np[k] == 0) continue;
            if (uused < ucap_rows) {
                u_labels[uused] = k + 1;
                memcpy(u_state + (size_t)uused * R, Ut + (size_t)k * R, sizeof(int) * (size_t)R);
                memcpy(u_cov + (size_t)uused * R, Utc + (size_t)k * R, sizeof(double) * (size_t)R);
                memcpy(u_allele + (size_t)uused * R, Uta + (size_t)k * R, sizeof(double) * (size_t)R);
            }
            uused++;
        }
        {
            long double s = 0.0L, s2 = 0.0L;
            for (int i = 0; i < N; i++) s += modality_ll(R, Xcov + (size_t)i * R, Utc + (size_t)(Zt[i] - 1) * R, sc);
            for (int i = 0; i < N; i++) s2 += modality_ll(R, Xall + (size_t)i * R, Uta + (size_t)(Zt[i] - 1) * R, sa);
            LL[tt] = (double)s + (double)s2;
        }
    }
    u_off[niter] = uused;
    if (uused > ucap_rows) rc = 5;
done:
    free(Xcov); free(Ut); free(np); free(Zt); free(sc); free(sa); free(mc); free(ma); free(mu_new);
    free(prob); free(cumc); free(pk); free(perm); free(permc); free(Utc); free(Uta);
    return rc;
}
