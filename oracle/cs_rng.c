/*
 * cs_rng.c — base-R random number generation restated for the parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY (see cs_oracle.h).
 *
 * The reference (pure R) draws every random number through base R:
 *   set.seed      R/BayesNonparCluster.R:156
 *   runif         R/BayesNonparCluster.R:167
 *   sample.int    R/BayesNonparCluster.R:170,207
 *   rnorm         R/BayesNonparCluster.R:226
 *   sample        R/BayesNonparAlleleCluster.R:161
 * Base R is a third-party dependency that is not under /root/reference (no
 * version pinned in DESCRIPTION; the shipped fixtures were written by R 4.0.2
 * and 4.1.1).  What follows restates the published algorithms of R >= 3.6 with
 * the default RNGkind("Mersenne-Twister", "Inversion", "Rejection"), as
 * specified in SURVEY.md Appendix B.  It is pinned by reproducing the
 * reference's five shipped nonpara_clustering*.rds objects bit-for-bit.
 */
#include "cs_oracle.h"

#include <math.h>
#include <stdint.h>

/* ------------------------------------------------------------------ */
/* Mersenne Twister with R's seeding                                   */
/* ------------------------------------------------------------------ */

void cso_set_seed(cso_mt *g, uint32_t seed)
{
    uint32_t s = seed;
    for (int j = 0; j < 50; j++) s = 69069u * s + 1u;
    /* R fills 625 words: the first is the position slot, then mt[0..623] */
    s = 69069u * s + 1u; /* consumed by the position slot, then overwritten */
    for (int j = 0; j < 624; j++) {
        s = 69069u * s + 1u;
        g->mt[j] = s;
    }
    g->pos = 624; /* force a refill on the first draw */
    g->ndraws = 0;
}

static void mt_refill(cso_mt *g)
{
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
    uint32_t *mt = g->mt;
    int kk;
    for (kk = 0; kk < 624 - 397; kk++) {
        uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    for (; kk < 623; kk++) {
        uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    uint32_t y = (mt[623] & UPPER) | (mt[0] & LOWER);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    g->pos = 0;
}

double cso_unif_rand(cso_mt *g)
{
    if (g->pos >= 624) mt_refill(g);
    uint32_t y = g->mt[g->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    g->ndraws++;
    double v = (double)y * 2.3283064365386963e-10;
    /* R's fixup(): keep strictly inside (0,1) */
    if (v <= 0.0) return 0.5 * 2.328306437080797e-10;
    if (1.0 - v <= 0.0) return 1.0 - 0.5 * 2.328306437080797e-10;
    return v;
}

/* ------------------------------------------------------------------ */
/* Normal quantile: Wichura's AS 241 (PPND16), which R's qnorm uses    */
/* ------------------------------------------------------------------ */

double cso_qnorm(double p)
{
    static const double a[8] = {
        3.3871328727963666080e0, 1.3314166789178437745e+2, 1.9715909503065514427e+3,
        1.3731693765509461125e+4, 4.5921953931549871457e+4, 6.7265770927008700853e+4,
        3.3430575583588128105e+4, 2.5090809287301226727e+3};
    static const double b[8] = {
        1.0, 4.2313330701600911252e+1, 6.8718700749205790830e+2, 5.3941960214247511077e+3,
        2.1213794301586595867e+4, 3.9307895800092710610e+4, 2.8729085735721942674e+4,
        5.2264952788528545610e+3};
    static const double c[8] = {
        1.42343711074968357734e0, 4.63033784615654529590e0, 5.76949722146069140550e0,
        3.64784832476320460504e0, 1.27045825245236838258e0, 2.41780725177450611770e-1,
        2.27238449892691845833e-2, 7.74545014278341407640e-4};
    static const double d[8] = {
        1.0, 2.05319162663775882187e0, 1.67638483018380384940e0, 6.89767334985100004550e-1,
        1.48103976427480074590e-1, 1.51986665636164571966e-2, 5.47593808499534494600e-4,
        1.05075007164441684324e-9};
    static const double e[8] = {
        6.65790464350110377720e0, 5.46378491116411436990e0, 1.78482653991729133580e0,
        2.96560571828504891230e-1, 2.65321895265761230930e-2, 1.24266094738807843860e-3,
        2.71155556874348757815e-5, 2.01033439929228813265e-7};
    static const double f[8] = {
        1.0, 5.99832206555887937690e-1, 1.36929880922735805310e-1, 1.48753612908506148525e-2,
        7.86869131145613259100e-4, 1.84631831751005468180e-5, 1.42151175831644588870e-7,
        2.04426310338993978564e-15};

    if (p != p) return p;
    if (p <= 0.0) return -INFINITY;
    if (p >= 1.0) return INFINITY;
    double q = p - 0.5, r, num, den;
    if (fabs(q) <= 0.425) {
        r = 0.180625 - q * q;
        num = a[7]; den = b[7];
        for (int i = 6; i >= 0; i--) { num = num * r + a[i]; den = den * r + b[i]; }
        return q * num / den;
    }
    r = (q < 0) ? p : 1.0 - p;
    r = sqrt(-log(r));
    if (r <= 5.0) {
        r -= 1.6;
        num = c[7]; den = d[7];
        for (int i = 6; i >= 0; i--) { num = num * r + c[i]; den = den * r + d[i]; }
    } else {
        r -= 5.0;
        num = e[7]; den = f[7];
        for (int i = 6; i >= 0; i--) { num = num * r + e[i]; den = den * r + f[i]; }
    }
    double val = num / den;
    return (q < 0.0) ? -val : val;
}

/* rnorm(1, mean, sd) with the "Inversion" normal kind */
double cso_rnorm(cso_mt *g, double mean, double sd)
{
    if (sd == 0.0 || !isfinite(mean)) return mean; /* consumes no random numbers */
    double u = cso_unif_rand(g);
    u = (double)(int)(134217728.0 * u) + cso_unif_rand(g);
    return mean + sd * cso_qnorm(u / 134217728.0);
}

/* ------------------------------------------------------------------ */
/* sample.int(n, 1, prob = p): normalise, heap-sort descending, invert  */
/* ------------------------------------------------------------------ */

/* Descending, non-stable heap sort carrying an index vector (R's revsort).
 * Written with 1-based positions like the textbook sift-down it follows. */
void cso_revsort(double *a0, int *ib0, int n)
{
    if (n <= 1) return;
    double *a = a0 - 1;
    int *ib = ib0 - 1;
    int l = (n >> 1) + 1, ir = n;
    for (;;) {
        double ra;
        int ii;
        if (l > 1) {
            l--;
            ra = a[l];
            ii = ib[l];
        } else {
            ra = a[ir];
            ii = ib[ir];
            a[ir] = a[1];
            ib[ir] = ib[1];
            if (--ir == 1) {
                a[1] = ra;
                ib[1] = ii;
                return;
            }
        }
        int i = l, j = l << 1;
        while (j <= ir) {
            if (j < ir && a[j] > a[j + 1]) j++;
            if (ra > a[j]) {
                a[i] = a[j];
                ib[i] = ib[j];
                i = j;
                j += j;
            } else {
                j = ir + 1;
            }
        }
        a[i] = ra;
        ib[i] = ii;
    }
}

int cso_sample_int_prob(cso_mt *g, int n, double *p, int *perm)
{
    double sum = 0.0;
    for (int i = 0; i < n; i++)
        if (p[i] > 0) sum += p[i];
    for (int i = 0; i < n; i++) p[i] /= sum;
    for (int i = 0; i < n; i++) perm[i] = i + 1;
    cso_revsort(p, perm, n);
    for (int i = 1; i < n; i++) p[i] += p[i - 1];
    double u = cso_unif_rand(g);
    int j = 0;
    while (j < n - 1 && u > p[j]) j++;
    return perm[j];
}

/* R_unif_index(n) for sample.kind = "Rejection": returns 0..n-1 */
int cso_sample_unif_index(cso_mt *g, int n)
{
    if (n <= 0) return 0;
    int bits = (int)ceil(log2((double)n));
    for (;;) {
        int64_t v = 0;
        for (int k = 0; k <= bits; k += 16) {
            int v1 = (int)floor(cso_unif_rand(g) * 65536.0);
            v = 65536 * v + v1;
        }
        if (bits < 64) v &= (((int64_t)1 << bits) - 1);
        if ((double)v < (double)n) return (int)v;
    }
}

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11) — the product's fast RNG mode   */
/* ------------------------------------------------------------------ */

void cso_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
