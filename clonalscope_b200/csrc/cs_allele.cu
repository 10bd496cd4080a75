// cs_allele.cu — allele (coverage + major-allele-fraction) variant of the nested-CRP sampler.
//
// Replaces the loop body of BayesNonparAlleleCluster (reference
// R/BayesNonparAlleleCluster.R:101-290) and genotype_neighbor (R/genotype_neighbor.R:7-40):
// cluster rows hold integer ids into the (copies, major-allele fraction) genotype grid
// (:26-30), the likelihood is a coverage Gaussian at grid[state].rho plus an allele Gaussian
// at grid[state].theta (:179-183), proposals draw an unused grid state (:161), and the
// end-of-sweep redraw snaps continuous (coverage, allele) means to the nearest grid state
// (:215-226).  PARITY UNPINNED: the reference never calls this function and ships no
// fixture for it; the kernel is checked against the CPU restatement only.
//
// The update order is inherently serial and the inputs this path sees are small, so the
// whole chain runs in one persistent CTA (same organisation as the R-compatible coverage
// sampler): thread 0 owns the random stream and the sorts, all threads score clusters and
// do the end-of-sweep reductions.  rng_mode CS_RNG_R consumes base R's Mersenne-Twister
// stream in the reference's order; CS_RNG_PHILOX substitutes Philox4x32-10 words for that
// stream draw by draw.
#include "cs_rstream.cuh"

#include <string.h>
#include <vector>

template <typename T>
int cs_transpose(int rows, int cols, const T *d_in, T *d_out, cudaStream_t st);
void cs_r_set_seed(uint32_t seed, uint32_t mt[624]);

namespace {

constexpr int AT = 256;
constexpr int GMAX = 64; // grid states (maxcp <= 9 -> 54)

struct AlleleDev {
    int N, R, lcap, G, G6, rng_mode, niter;
    double alpha, beta;
    uint32_t seed;
    const double *Xc, *Xa; // N x R cell-major (coverage already capped at 3)
    int *Zt;               // N labels (1-based)
    int *Ut;               // lcap x R state ids
    double *Utc, *Uta;     // lcap x R continuous means (end of sweep)
    int *np0;              // lcap initial counts
    const double *sig0c, *sig0a;
    uint32_t *mt;
    // outputs
    int *Zall;
    double *sigc_all, *siga_all, *LL;
    int *kt_all;
    long long *u_off;
    int *u_labels, *u_state;
    double *u_cov, *u_allele;
    long long urec_cap;
    int *err;
    double *L0;
};

struct Stream {
    int mode;
    uint32_t *mt;
    int pos;
    uint32_t seed;
    unsigned long long n;
    __device__ double next()
    {
        if (mode == CS_RNG_R) return unif_rand(mt, pos);
        Philox4 o = cs_philox4x32((uint32_t)n, (uint32_t)(n >> 32), 0u, 0x05000000u, seed, 0x434C4F4Eu);
        n++;
        return cs_u53(o.v[0], o.v[1]);
    }
    // R_unif_index(n), sample.kind = "Rejection"
    __device__ int index(int nn)
    {
        if (nn <= 0) return 0;
        const int bits = (int)ceil(log2((double)nn));
        for (;;) {
            long long v = 0;
            for (int k = 0; k <= bits; k += 16) v = 65536 * v + (int)floor(next() * 65536.0);
            v &= ((1ll << bits) - 1);
            if (v < nn) return (int)v;
        }
    }
    __device__ double rnorm(double mean, double sd)
    {
        if (sd == 0.0 || !isfinite(mean)) return mean;
        double u = next();
        u = (double)(int)(134217728.0 * u) + next();
        return mean + sd * cs_qnorm(u / 134217728.0);
    }
};

__device__ __forceinline__ double dn_log(double x, double m, double s)
{
    const double z = fabs((x - m) / s);
    return -(CS_LN_SQRT_2PI + 0.5 * z * z + log(s));
}

__device__ int nearest_state(double cov, double allele, const double *gc, const double *ga, int G6)
{
    int best = 0;
    double bd = INFINITY;
    for (int i = 0; i < G6; i++) {
        const double a = cov - gc[i], b = allele - ga[i];
        const double d = sqrt(a * a + b * b);
        if (d < bd) {
            bd = d;
            best = i;
        }
    }
    return best + 1;
}

__global__ void __launch_bounds__(AT)
gibbs_allele_kernel(AlleleDev D)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lcap = D.lcap, R = D.R, N = D.N, G = D.G, tid = threadIdx.x;
    double *d = reinterpret_cast<double *>(s_raw);
    double *cumc = d; d += lcap + 1;
    double *prob = d; d += lcap + 1;
    double *xc = d; d += R;
    double *xa = d; d += R;
    double *sc = d; d += R;
    double *sa = d; d += R;
    double *gc = d; d += GMAX; // grid: coverage (maxcp of the call, then the maxcp = 6 grid for snapping)
    double *ga = d; d += GMAX;
    double *lsc = d; d += R;   // log(sigma) of both modalities (constant within a sweep)
    double *lsa = d; d += R;
    double *termc = d; d += (size_t)(AT / 32) * R; // one row of terms per warp and modality
    double *terma = d; d += (size_t)(AT / 32) * R;
    unsigned long long *s_used = reinterpret_cast<unsigned long long *>(d); d += R; // grid states in use, per segment
    int *p = reinterpret_cast<int *>(d);
    int *permc = p; p += lcap + 1;
    int *np = p; p += lcap;
    int *occ = p; p += lcap;
    int *mu_new = p; p += R;
    uint32_t *mt = reinterpret_cast<uint32_t *>(p);
    __shared__ int s_kt, s_nocc, s_valid, s_err, s_z, s_new;
    __shared__ long long s_uused;
    __shared__ double s_red[AT];
    __shared__ Stream S;

    for (int k = tid; k < lcap; k += AT) np[k] = D.np0[k];
    for (int r = tid; r < R; r += AT) {
        sc[r] = D.sig0c[r];
        sa[r] = D.sig0a[r];
        lsc[r] = log(sc[r]);
        lsa[r] = log(sa[r]);
    }
    for (int k = tid; k < 624; k += AT) mt[k] = D.mt[k];
    if (tid == 0) {
        int g = 0; // genotype grid (R/BayesNonparAlleleCluster.R:26-30); a maxcp = 6 grid is a prefix / extension
        for (int ii = 1; ii <= 9; ii++)
            for (int j = 0; j <= ii; j++) {
                gc[g] = 0.5 * ii;
                ga[g] = (j == 0) ? 0.0 : (double)j / (double)ii;
                g++;
            }
        s_kt = D.np0[lcap]; // K0 rides behind the counts
        S.mode = D.rng_mode;
        S.mt = mt;
        S.pos = 624;
        S.seed = D.seed;
        S.n = 0;
        s_valid = 0;
        s_err = 0;
        s_uused = 0;
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int k = 0; k < s_kt; k++)
            if (np[k] > 0) occ[n++] = k;
        s_nocc = n;
    }
    { // L0 at (U0, Z0, sigmas0) (:131-132); U0a has 0 -> 0.05 and 1 -> 0.95
        double acc = 0.0;
        for (int i = tid; i < N; i += AT) {
            const int *st = D.Ut + (size_t)(D.Zt[i] - 1) * R;
            const double *x1 = D.Xc + (size_t)i * R, *x2 = D.Xa + (size_t)i * R;
            for (int r = 0; r < R; r++) {
                double a = ga[st[r] - 1];
                a = (a == 0.0) ? 0.05 : ((a == 1.0) ? 0.95 : a);
                if (!cs_isna(x1[r])) acc += dn_log(x1[r], gc[st[r] - 1], sc[r]);
                if (!cs_isna(x2[r])) acc += dn_log(x2[r], a, sa[r]);
            }
        }
        s_red[tid] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int i = 0; i < AT; i++) t += s_red[i];
            *D.L0 = t;
        }
    }
    __syncthreads();

    for (int tt = 0; tt < D.niter && !s_err; tt++) {
        for (int ii = 0; ii < N; ii++) {
            const int kt = s_kt;
            const int zold = D.Zt[ii] - 1;
            for (int r = tid; r < R; r += AT) {
                xc[r] = D.Xc[(size_t)ii * R + r];
                xa[r] = D.Xa[(size_t)ii * R + r];
            }
            if (!s_valid) {
                const double sum = (double)N + (D.alpha > 0 ? D.alpha : 0.0);
                for (int k = tid; k <= kt; k += AT) {
                    cumc[k] = ((k < kt) ? (double)np[k] : D.alpha) / sum;
                    permc[k] = k + 1;
                }
            }
            // the grid states some cluster row holds, per segment (all threads; the draws below are thread 0's)
            for (int r = tid; r < R; r += AT) {
                unsigned long long used = 0;
                for (int k = 0; k < kt; k++) {
                    const int st = D.Ut[(size_t)k * R + r];
                    if (st >= 1 && st <= G) used |= 1ull << (st - 1);
                }
                s_used[r] = used;
            }
            __syncthreads();
            // ---- :158-166 proposal: an unused grid state per segment, or a live cluster's state
            if (tid == 0) {
                if (!s_valid) {
                    revsort(cumc, permc, kt + 1);
                    for (int k = 1; k <= kt; k++) cumc[k] += cumc[k - 1];
                    s_valid = 1;
                }
                for (int r = 0; r < R && !s_err; r++) {
                    const unsigned long long used = s_used[r];
                    const unsigned long long all = (G >= 64) ? ~0ull : ((1ull << G) - 1ull);
                    unsigned long long freeb = all & ~used;
                    const int nc = __popcll(freeb);
                    int tmpr = 0;
                    if (nc == 0) {
                        s_err = CS_ERR_STATE; // sample(integer(0), 1) is an error in the reference
                    } else if (nc == 1) {
                        tmpr = 1 + S.index(__ffsll((long long)freeb)); // sample(x, 1) == sample(1:x, 1)
                    } else {
                        int idx = S.index(nc);
                        while (idx--) freeb &= freeb - 1;
                        tmpr = __ffsll((long long)freeb);
                    }
                    const double u = S.next();
                    int j = 0;
                    while (j < kt && u > cumc[j]) j++;
                    const int indr = permc[j];
                    mu_new[r] = (indr <= kt) ? D.Ut[(size_t)(indr - 1) * R + r] : tmpr;
                }
                np[zold] -= 1; // :168
            }
            __syncthreads();
            if (s_err) break;
            // ---- :179-183 scores of the occupied clusters and of the new table
            const int nocc = s_nocc;
            { // a warp per candidate: the terms of all segments at once, then the two sums in segment
              // order by one lane (cell_ll_states' order of additions)
                const int lane = tid & 31, wid = tid >> 5;
                double *tc = termc + (size_t)wid * R, *ta = terma + (size_t)wid * R;
                for (int e = wid; e <= nocc; e += AT / 32) {
                    const int k = e < nocc ? occ[e] : -1;
                    const int n = e < nocc ? np[k] : 1;
                    if (n == 0) {
                        if (lane == 0) prob[e] = -INFINITY;
                        continue;
                    }
                    const int *st = e < nocc ? D.Ut + (size_t)k * R : mu_new;
                    for (int r = lane; r < R; r += 32) {
                        const int sidx = st[r];
                        const bool ok = sidx >= 1 && sidx <= G;
                        const double zc = fabs((xc[r] - (ok ? gc[sidx - 1] : NAN)) / sc[r]);
                        const double za = fabs((xa[r] - (ok ? ga[sidx - 1] : NAN)) / sa[r]);
                        tc[r] = -(CS_LN_SQRT_2PI + 0.5 * zc * zc + lsc[r]);
                        ta[r] = -(CS_LN_SQRT_2PI + 0.5 * za * za + lsa[r]);
                    }
                    __syncwarp();
                    if (lane == 0) {
                        double s1 = 0.0, s2 = 0.0;
                        for (int r = 0; r < R; r++) {
                            if (!cs_isna(xc[r])) s1 += tc[r];
                            if (!cs_isna(xa[r])) s2 += ta[r];
                        }
                        prob[e] = (s1 + s2) + log(e < nocc ? (double)n : D.beta);
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
            // ---- :185-190 categorical draw: sample.int(kt + 1, prob = c(p_k, p_new))
            { // the maximum and the exponentials by all threads; the sum in candidate order by thread 0
                double mx = -INFINITY;
                for (int e = tid; e <= nocc; e += AT)
                    if (e == nocc || np[occ[e]] != 0) mx = fmax(mx, prob[e]);
                s_red[tid] = mx;
                __syncthreads();
                for (int off = AT / 2; off >= 1; off >>= 1) {
                    if (tid < off) s_red[tid] = fmax(s_red[tid], s_red[tid + off]);
                    __syncthreads();
                }
                mx = s_red[0];
                __syncthreads();
                for (int e = tid; e <= nocc; e += AT) prob[e] = (e == nocc || np[occ[e]] != 0) ? exp(prob[e] - mx) : 0.0;
                __syncthreads();
            }
            if (tid == 0) {
                double sum = 0.0;
                for (int e = 0; e <= nocc; e++)
                    if (prob[e] > 0) sum += prob[e];
                // the full kt+1 vector (zeros for empty clusters) goes through revsort like in R
                for (int k = 0; k <= kt; k++) {
                    cumc[k] = 0.0;
                    permc[k] = k + 1;
                }
                for (int e = 0; e < nocc; e++) cumc[occ[e]] = prob[e] / sum;
                cumc[kt] = prob[nocc] / sum;
                revsort(cumc, permc, kt + 1);
                for (int k = 1; k <= kt; k++) cumc[k] += cumc[k - 1];
                const double u = S.next();
                int j = 0;
                while (j < kt && u > cumc[j]) j++;
                int z = permc[j];
                s_valid = 0; // the scratch held the proposal cache
                s_new = 0;
                if (z > kt) {
                    if (kt + 1 > lcap) {
                        s_err = CS_ERR_KCAP;
                        z = zold + 1;
                        np[zold] += 1;
                    } else {
                        np[kt] = 1;
                        occ[s_nocc++] = kt;
                        s_kt = kt + 1;
                        s_new = 1;
                    }
                } else {
                    np[z - 1] += 1;
                }
                if (z - 1 != zold && np[zold] == 0) {
                    int n = s_nocc, w = 0;
                    for (int e = 0; e < n; e++)
                        if (occ[e] != zold) occ[w++] = occ[e];
                    s_nocc = w;
                }
                D.Zt[ii] = z;
                s_z = z;
            }
            __syncthreads();
            if (s_new) {
                int *urow = D.Ut + (size_t)(s_z - 1) * R;
                for (int r = tid; r < R; r += AT) urow[r] = mu_new[r];
            }
            __syncthreads();
            if (s_err) break;
        }
        if (s_err) break;

        // ---- :203-231 end of sweep
        const int kt = s_kt, nocc = s_nocc;
        const int last = (tt == D.niter - 1);
        for (int pass = 0; pass < (last ? 2 : 1); pass++) {
            for (int e = tid; e < kt * R; e += AT) {
                D.Ut[e] = 0;
                D.Utc[e] = 0.0;
                D.Uta[e] = 0.0;
            }
            __syncthreads();
            // cluster means of both modalities: thread per (occupied cluster, segment)
            for (int e = tid; e < nocc * R; e += AT) {
                const int k = occ[e / R], r = e % R;
                double s1 = 0.0, s2 = 0.0;
                int c1 = 0, c2 = 0;
                for (int i = 0; i < N; i++) {
                    if (D.Zt[i] - 1 != k) continue;
                    const double a = D.Xc[(size_t)i * R + r], b = D.Xa[(size_t)i * R + r];
                    if (!cs_isna(a)) { s1 += a; c1++; }
                    if (!cs_isna(b)) { s2 += b; c2++; }
                }
                D.Utc[(size_t)k * R + r] = c1 ? s1 / (double)c1 : 0.0;
                D.Uta[(size_t)k * R + r] = c2 ? s2 / (double)c2 : 0.0;
            }
            __syncthreads();
            if (pass == 0 && tid == 0) {
                for (int e = 0; e < nocc; e++) {
                    const int k = occ[e];
                    double *uc = D.Utc + (size_t)k * R, *ua = D.Uta + (size_t)k * R;
                    bool allzero = true;
                    for (int r = 0; r < R; r++)
                        if (!(uc[r] == 0.0)) { allzero = false; break; }
                    if (allzero) {
                        for (int r = 0; r < R; r++) ua[r] = 0.0;
                        continue;
                    }
                    const double sq = sqrt((double)np[k]);
                    for (int r = 0; r < R; r++) {
                        const double v = S.rnorm(uc[r], sc[r] / sq);
                        uc[r] = v < 0.0 ? 0.0 : v;
                    }
                    for (int r = 0; r < R; r++) {
                        const double v = S.rnorm(ua[r], sa[r] / sq);
                        ua[r] = v < 0.0 ? 0.0 : v;
                    }
                }
            }
            __syncthreads();
            for (int e = tid; e < nocc * R; e += AT) { // snap to the nearest state of the maxcp = 6 grid
                const int k = occ[e / R], r = e % R;
                D.Ut[(size_t)k * R + r] = nearest_state(D.Utc[(size_t)k * R + r], D.Uta[(size_t)k * R + r], gc, ga, D.G6);
            }
            __syncthreads();
            if (pass == 0 && last) continue;
            for (int r = tid; r < R; r += AT) {
                double s1 = 0.0, s2 = 0.0;
                int c1 = 0, c2 = 0;
                for (int i = 0; i < N; i++) {
                    const double a = D.Xc[(size_t)i * R + r], b = D.Xa[(size_t)i * R + r];
                    const size_t o = (size_t)(D.Zt[i] - 1) * R + r;
                    if (!cs_isna(a)) { c1++; const double dd = a - D.Utc[o]; s1 += dd * dd; }
                    if (!cs_isna(b)) { c2++; const double dd = b - D.Uta[o]; s2 += dd * dd; }
                }
                double v1 = sqrt(1.0 / (double)c1 * s1), v2 = sqrt(1.0 / (double)c2 * s2);
                sc[r] = (v1 == 0.0) ? 0.01 : v1;
                sa[r] = (v2 == 0.0) ? 0.01 : v2;
                lsc[r] = log(sc[r]);
                lsa[r] = log(sa[r]);
            }
            __syncthreads();
        }
        // ---- record
        for (int i = tid; i < N; i += AT) D.Zall[(size_t)tt * N + i] = D.Zt[i];
        for (int r = tid; r < R; r += AT) {
            D.sigc_all[(size_t)tt * R + r] = sc[r];
            D.siga_all[(size_t)tt * R + r] = sa[r];
        }
        {
            double acc = 0.0;
            for (int i = tid; i < N; i += AT) {
                const double *x1 = D.Xc + (size_t)i * R, *x2 = D.Xa + (size_t)i * R;
                const double *m1 = D.Utc + (size_t)(D.Zt[i] - 1) * R, *m2 = D.Uta + (size_t)(D.Zt[i] - 1) * R;
                for (int r = 0; r < R; r++) {
                    if (!cs_isna(x1[r])) acc += dn_log(x1[r], m1[r], sc[r]);
                    if (!cs_isna(x2[r])) acc += dn_log(x2[r], m2[r], sa[r]);
                }
            }
            s_red[tid] = acc;
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int i = 0; i < AT; i++) s += s_red[i];
                D.LL[tt] = s;
                D.kt_all[tt] = kt;
                D.u_off[tt] = s_uused;
                D.u_off[tt + 1] = s_uused + nocc;
                if (s_uused + nocc > D.urec_cap) s_err = CS_ERR_UCAP;
            }
            __syncthreads();
        }
        if (!s_err) {
            for (int e = tid; e < nocc * R; e += AT) {
                const size_t src = (size_t)occ[e / R] * R + e % R, dst = (size_t)(s_uused + e / R) * R + e % R;
                D.u_state[dst] = D.Ut[src];
                D.u_cov[dst] = D.Utc[src];
                D.u_allele[dst] = D.Uta[src];
            }
            for (int e = tid; e < nocc; e += AT) D.u_labels[s_uused + e] = occ[e] + 1;
        }
        __syncthreads();
        if (tid == 0) {
            s_uused += nocc;
            s_valid = 0;
        }
        __syncthreads();
    }
    if (tid == 0 && s_err) *D.err = s_err;
}

__global__ void cap_kernel(size_t n, double *x, double cap)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && x[i] > cap) x[i] = cap;
}

int grid_size(int maxcp) { return maxcp * (maxcp + 3) / 2; }

} // namespace

extern "C" int cs_ncrp_gibbs_allele(int N, int R, const double *Xcov, const double *Xallele, int K0,
                                    const int *U0, const int *Z0, const double *sig0_cov,
                                    const double *sig0_allele, double alpha, double beta, int niter,
                                    uint32_t seed, int maxcp, const cs_gibbs_opts *opts, int *Zall,
                                    double *sig_cov_all, double *sig_allele_all, double *LL, double *L0,
                                    int *kt_all, int64_t ucap_rows, int64_t *u_off, int *u_labels, int *u_state,
                                    double *u_cov, double *u_allele)
{
    CS_REQUIRE(Xcov && Xallele && U0 && Z0 && sig0_cov && sig0_allele && Zall && LL && L0 && kt_all && u_off,
               "cs_ncrp_gibbs_allele: NULL argument");
    CS_REQUIRE(N >= 1 && R >= 1 && K0 >= 1 && niter >= 1, "cs_ncrp_gibbs_allele: bad dimensions");
    CS_REQUIRE(maxcp >= 1 && maxcp <= 9, "cs_ncrp_gibbs_allele: maxcp=%d outside 1..9", maxcp);
    cs_gibbs_opts o;
    memset(&o, 0, sizeof(o));
    o.rng_mode = CS_RNG_PHILOX;
    o.device = -1;
    if (opts) o = *opts;
    if (o.device >= 0) CS_CUDA(cudaSetDevice(o.device));
    const int lcap = o.kcap > 0 ? (o.kcap < K0 ? K0 : o.kcap) : 2048;
    const int G = grid_size(maxcp);
    for (int e = 0; e < K0 * R; e++)
        CS_REQUIRE(U0[e] >= 1 && U0[e] <= G, "U0 state id %d outside the genotype grid 1..%d", U0[e], G);
    CS_REQUIRE(K0 <= lcap, "K0=%d exceeds the label capacity %d", K0, lcap);
    std::vector<int> np0((size_t)lcap + 1, 0);
    for (int i = 0; i < N; i++) {
        if (Z0[i] < 1 || Z0[i] > K0) {
            cs_set_error("Z0[%d]=%d is not a row of U0 (1..%d)", i + 1, Z0[i], K0);
            return CS_ERR_STATE;
        }
        np0[Z0[i] - 1]++;
    }
    for (int k = 0; k < K0; k++) {
        if (np0[k] == 0) {
            cs_set_error("no cell starts in cluster %d: as.numeric(table(Zt)) would be shorter than nrow(U0) "
                         "and sample.int fails (R/BayesNonparAlleleCluster.R:148,163)", k + 1);
            return CS_ERR_STATE;
        }
    }
    np0[lcap] = K0;
    const size_t smem = sizeof(double) * (2 * ((size_t)lcap + 1) + 4 * (size_t)R + 2 * GMAX + 2 * (size_t)R + 2 * (size_t)(AT / 32) * R + (size_t)R) +
                        sizeof(int) * (((size_t)lcap + 1) + 2 * (size_t)lcap + R) + sizeof(uint32_t) * 624 + 64;
    if (smem > 227 * 1024) {
        cs_set_error("cs_ncrp_gibbs_allele: label capacity %d with R=%d needs %zu B of shared memory", lcap, R, smem);
        return CS_ERR_UNSUPPORTED;
    }
    const size_t nx = (size_t)N * R;
    const long long urec = (long long)niter * 128 + lcap;
    DevBuf<double> tmp, Xc, Xa, Utc, Uta, s0c, s0a, sigc, siga, dLL, ucov, uall, dL0;
    DevBuf<int> Zt, Ut, dnp, dZall, tZ, dkt, ulab, ust, derr;
    DevBuf<long long> uoff;
    DevBuf<uint32_t> dmt;
    CS_CUDA(tmp.alloc(nx)); CS_CUDA(Xc.alloc(nx)); CS_CUDA(Xa.alloc(nx));
    CS_CUDA(Utc.alloc((size_t)lcap * R)); CS_CUDA(Uta.alloc((size_t)lcap * R)); CS_CUDA(Ut.alloc((size_t)lcap * R));
    CS_CUDA(s0c.alloc(R)); CS_CUDA(s0a.alloc(R)); CS_CUDA(sigc.alloc((size_t)niter * R)); CS_CUDA(siga.alloc((size_t)niter * R));
    CS_CUDA(dLL.alloc(niter)); CS_CUDA(ucov.alloc((size_t)urec * R)); CS_CUDA(uall.alloc((size_t)urec * R));
    CS_CUDA(Zt.alloc(N)); CS_CUDA(dnp.alloc((size_t)lcap + 1)); CS_CUDA(dZall.alloc((size_t)niter * N)); CS_CUDA(tZ.alloc((size_t)niter * N));
    CS_CUDA(dkt.alloc(niter)); CS_CUDA(ulab.alloc((size_t)urec)); CS_CUDA(ust.alloc((size_t)urec * R)); CS_CUDA(derr.alloc(1));
    CS_CUDA(uoff.alloc((size_t)niter + 1)); CS_CUDA(dmt.alloc(624)); CS_CUDA(dL0.alloc(1));
    // inputs: column-major N x R -> cell-major; coverage capped at 3 (:37)
    CS_CUDA(cudaMemcpy(tmp.p, Xcov, sizeof(double) * nx, cudaMemcpyHostToDevice));
    int rc = cs_transpose<double>(R, N, tmp.p, Xc.p, 0);
    if (rc) return rc;
    cap_kernel<<<(unsigned)((nx + 255) / 256), 256>>>(nx, Xc.p, 3.0);
    CS_CUDA(cudaMemcpy(tmp.p, Xallele, sizeof(double) * nx, cudaMemcpyHostToDevice));
    if ((rc = cs_transpose<double>(R, N, tmp.p, Xa.p, 0))) return rc;
    std::vector<int> Urows((size_t)K0 * R);
    for (int k = 0; k < K0; k++)
        for (int r = 0; r < R; r++) Urows[(size_t)k * R + r] = U0[(size_t)r * K0 + k];
    CS_CUDA(cudaMemset(Ut.p, 0, sizeof(int) * (size_t)lcap * R));
    CS_CUDA(cudaMemcpy(Ut.p, Urows.data(), sizeof(int) * Urows.size(), cudaMemcpyHostToDevice));
    CS_CUDA(cudaMemcpy(Zt.p, Z0, sizeof(int) * N, cudaMemcpyHostToDevice));
    CS_CUDA(cudaMemcpy(dnp.p, np0.data(), sizeof(int) * np0.size(), cudaMemcpyHostToDevice));
    CS_CUDA(cudaMemcpy(s0c.p, sig0_cov, sizeof(double) * R, cudaMemcpyHostToDevice));
    CS_CUDA(cudaMemcpy(s0a.p, sig0_allele, sizeof(double) * R, cudaMemcpyHostToDevice));
    CS_CUDA(cudaMemset(derr.p, 0, sizeof(int)));
    uint32_t mt[624];
    cs_r_set_seed(seed, mt);
    CS_CUDA(cudaMemcpy(dmt.p, mt, sizeof(mt), cudaMemcpyHostToDevice));
    AlleleDev D;
    D.N = N; D.R = R; D.lcap = lcap; D.G = G; D.G6 = grid_size(6); D.rng_mode = o.rng_mode; D.niter = niter;
    D.alpha = alpha; D.beta = beta; D.seed = seed;
    D.Xc = Xc.p; D.Xa = Xa.p; D.Zt = Zt.p; D.Ut = Ut.p; D.Utc = Utc.p; D.Uta = Uta.p; D.np0 = dnp.p;
    D.sig0c = s0c.p; D.sig0a = s0a.p; D.mt = dmt.p; D.Zall = dZall.p; D.sigc_all = sigc.p; D.siga_all = siga.p;
    D.LL = dLL.p; D.kt_all = dkt.p; D.u_off = uoff.p; D.u_labels = ulab.p; D.u_state = ust.p; D.u_cov = ucov.p;
    D.u_allele = uall.p; D.urec_cap = urec; D.err = derr.p; D.L0 = dL0.p;
    CS_CUDA(cudaFuncSetAttribute(gibbs_allele_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gibbs_allele_kernel<<<1, AT, smem>>>(D);
    CS_CUDA(cudaGetLastError());
    CS_CUDA(cudaDeviceSynchronize());
    int herr = 0;
    CS_CUDA(cudaMemcpy(&herr, derr.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (herr == CS_ERR_KCAP) {
        cs_set_error("more clusters than the label capacity %d (raise cs_gibbs_opts.kcap)", lcap);
        return herr;
    }
    if (herr == CS_ERR_STATE) {
        cs_set_error("every genotype state is in use in some segment: sample(integer(0), 1) fails in the reference");
        return herr;
    }
    if (herr) {
        cs_set_error("device-side error %d", herr);
        return herr;
    }
    if ((rc = cs_transpose<int>(niter, N, dZall.p, tZ.p, 0))) return rc;
    CS_CUDA(cudaMemcpy(Zall, tZ.p, sizeof(int) * (size_t)niter * N, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(sig_cov_all, sigc.p, sizeof(double) * (size_t)niter * R, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(sig_allele_all, siga.p, sizeof(double) * (size_t)niter * R, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(LL, dLL.p, sizeof(double) * niter, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(L0, dL0.p, sizeof(double), cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(kt_all, dkt.p, sizeof(int) * niter, cudaMemcpyDeviceToHost));
    std::vector<long long> off((size_t)niter + 1);
    CS_CUDA(cudaMemcpy(off.data(), uoff.p, sizeof(long long) * off.size(), cudaMemcpyDeviceToHost));
    for (int t = 0; t <= niter; t++) u_off[t] = off[t];
    const long long need = off[niter];
    if (need > ucap_rows) {
        cs_set_error("ucap_rows=%lld too small: %lld rows recorded", (long long)ucap_rows, need);
        return CS_ERR_UCAP;
    }
    if (need) {
        CS_CUDA(cudaMemcpy(u_labels, ulab.p, sizeof(int) * (size_t)need, cudaMemcpyDeviceToHost));
        CS_CUDA(cudaMemcpy(u_state, ust.p, sizeof(int) * (size_t)need * R, cudaMemcpyDeviceToHost));
        CS_CUDA(cudaMemcpy(u_cov, ucov.p, sizeof(double) * (size_t)need * R, cudaMemcpyDeviceToHost));
        CS_CUDA(cudaMemcpy(u_allele, uall.p, sizeof(double) * (size_t)need * R, cudaMemcpyDeviceToHost));
    }
    return CS_OK;
}
