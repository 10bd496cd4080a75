// cs_rstream.cuh — base R's random stream restated for the device (R-compatible modes).
// Mersenne-Twister MT19937 with R's seeding and output scaling, inversion rnorm (AS 241),
// and the descending heap sort `revsort` that sample.int's ProbSampleReplace applies
// (SURVEY.md Appendix B).  Translation units including this header are compiled with
// -fmad=false so that double arithmetic rounds like base R's C code.
#pragma once
#include "cs_common.cuh"

static __device__ void mt_refill(uint32_t *mt)
{
    const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, A = 0x9908b0dfu;
    int k = 0;
    for (; k < 624 - 397; k++) {
        const uint32_t y = (mt[k] & UP) | (mt[k + 1] & LO);
        mt[k] = mt[k + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    }
    for (; k < 623; k++) {
        const uint32_t y = (mt[k] & UP) | (mt[k + 1] & LO);
        mt[k] = mt[k - 227] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    }
    const uint32_t y = (mt[623] & UP) | (mt[0] & LO);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
}

__device__ __forceinline__ double unif_rand(uint32_t *mt, int &pos)
{
    if (pos >= 624) {
        mt_refill(mt);
        pos = 0;
    }
    uint32_t y = mt[pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    const double v = (double)y * 2.3283064365386963e-10;
    if (v <= 0.0) return 0.5 * 2.328306437080797e-10;
    if (1.0 - v <= 0.0) return 1.0 - 0.5 * 2.328306437080797e-10;
    return v;
}

static __device__ double rnorm_r(uint32_t *mt, int &pos, double mean, double sd)
{
    if (sd == 0.0 || !isfinite(mean)) return mean;
    double u = unif_rand(mt, pos);
    u = (double)(int)(134217728.0 * u) + unif_rand(mt, pos);
    return __dadd_rn(mean, __dmul_rn(sd, cs_qnorm(u / 134217728.0)));
}

// R's revsort: descending, unstable heap sort of a[] carrying ib[] (1-based sift-down)
static __device__ void revsort(double *a0, int *ib0, int n)
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

