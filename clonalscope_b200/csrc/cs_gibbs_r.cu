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

__global__ void __launch_bounds__(RT)
gibbs_r_kernel(ChainDev D, int lcap, int t0, int t1, int last_sweep, double *__restrict__ scratch)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    RShared S;
    {
        double *d = reinterpret_cast<double *>(s_raw);
        S.cumc = d; d += lcap + 1;
        S.prob = d; d += lcap + 1;
        S.xrow = d; d += D.R;
        S.mu = d; d += D.R;
        S.sig = d; d += D.R;
        int *p = reinterpret_cast<int *>(d);
        S.permc = p; p += lcap + 1;
        S.perm = p; p += lcap + 1;
        S.np = p; p += lcap;
        S.occ = p; p += lcap;
        S.mt = reinterpret_cast<uint32_t *>(p);
    }
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
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int k = 0; k < s_kt; k++)
            if (S.np[k] > 0) S.occ[n++] = k;
        s_nocc = n;
    }
    __syncthreads();

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
            __syncthreads();
            // ---- :200-207 categorical draw (thread 0)
            if (tid == 0) {
                int pos = s_pos;
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
                const double u = unif_rand(S.mt, pos);
                s_pos = pos;
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
                        const double v = rnorm_r(S.mt, pos, m[r], S.sig[r] / sq);
                        u[r] = (v < 0.0) ? 0.0 : v;
                    }
                }
                s_pos = pos;
            }
            __syncthreads();
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
    for (int k = tid; k < 624; k += RT) D.mt[k] = S.mt[k];
    if (tid == 0) {
        D.scal->kt = s_kt;
        D.scal->mt_pos = s_pos;
        D.scal->n_reject += s_rej;
        D.scal->u_used = s_uused;
        if (s_err) D.scal->err = s_err;
    }
}

} // namespace

size_t cs_rmode_smem_bytes(int lcap, int R)
{
    return sizeof(double) * (2 * ((size_t)lcap + 1) + 3 * (size_t)R) + sizeof(int) * (2 * ((size_t)lcap + 1) + 2 * (size_t)lcap) +
           sizeof(uint32_t) * 624 + 64;
}

// sweeps [t0, t1) of the R-compatible chain; last_sweep = index of the run's final sweep
int cs_rmode_run(const ChainDev &D, int lcap, int t0, int t1, int last_sweep, double *scratch, cudaStream_t st)
{
    const size_t smem = cs_rmode_smem_bytes(lcap, D.R);
    if (smem > 227 * 1024) {
        cs_set_error("R-compatible sampler: label capacity %d with R=%d needs %zu B of shared memory", lcap, D.R, smem);
        return CS_ERR_UNSUPPORTED;
    }
    CS_CUDA(cudaFuncSetAttribute(gibbs_r_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gibbs_r_kernel<<<1, RT, smem, st>>>(D, lcap, t0, t1, last_sweep, scratch);
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
