// cs_gibbs_philox.cu — nested-CRP collapsed Gibbs sweep, Philox mode.
//
// Reference semantics: R/BayesNonparCluster.R:159-245 — cells are updated one after the
// other, every draw sees the counts / assignments / appended rows left by the previous
// cell.  That order is kept EXACTLY, but evaluated speculatively:
//
//   prepare_kernel (all SMs, one warp per cell)
//       scores every cell against all live clusters (Q matrix), draws its nested-CRP
//       proposal (:165-175) and its categorical draw (:193-207) against the state at
//       the START of the sweep.  As long as no earlier cell changed its label that state
//       is the state the sequential sampler would have seen (a cell that re-draws its
//       own label leaves counts and rows untouched), so all draws up to and including
//       the first label change ("event") are already the sequential result.
//   scan_kernel (one CTA, 32 warps)
//       resumes at the first event and walks the rest of the sweep in windows of 32
//       cells, one warp per cell, re-drawing against the committed state; a window is
//       cut at its first event, that event is committed, and the walk continues behind
//       it.  Q columns of clusters opened during the sweep are filled on demand.
//   end of sweep (:221-245): deterministic chunked reductions for cluster means and
//       sigmas, Philox/AS241 redraw of the means, total log-likelihood, compaction of
//       emptied clusters (labels keep growing, rows are never renumbered).
//
// Random numbers are counter-based (Philox4x32-10 keyed by seed; counter = cell, segment,
// sweep, purpose|attempt), so a draw does not depend on how many draws preceded it —
// that is what makes the speculation exact.  All sums that decide a draw are evaluated
// in one canonical order (lane-strided fma + butterfly), identical in every kernel and
// in the CPU oracle's Philox mode.
#include "cs_gibbs.cuh"

#include <limits.h>

namespace {

constexpr int PREP_THREADS = 256;
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

template <int NCH>
__device__ __forceinline__ void load_row(const double *__restrict__ xrow, int R, int lane, double (&x)[NCH])
{
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        const int r = lane + 32 * i;
        x[i] = (r < R) ? xrow[r] : 0.0;
    }
}

// canonical squared standardised distance of one cell to one row
template <int NCH>
__device__ __forceinline__ double q_row(const double (&x)[NCH], const double *u, const double *s_isig, int R,
                                        int lane)
{
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        const int r = lane + 32 * i;
        if (r < R) {
            const double xv = x[i];
            if (!cs_isna(xv)) {
                const double d = __dmul_rn(__dsub_rn(xv, u[r]), s_isig[r]);
                s = __fma_rn(d, d, s);
            }
        }
    }
    return cs_warp_sum_bfly(s);
}

template <int NCH>
__device__ __forceinline__ double q_regs(const double (&x)[NCH], const double (&mu)[NCH], const double *s_isig,
                                         int R, int lane)
{
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        const int r = lane + 32 * i;
        if (r < R) {
            const double xv = x[i];
            if (!cs_isna(xv)) {
                const double d = __dmul_rn(__dsub_rn(xv, mu[i]), s_isig[r]);
                s = __fma_rn(d, d, s);
            }
        }
    }
    return cs_warp_sum_bfly(s);
}

// nested-CRP proposal for one cell (R/BayesNonparCluster.R:166-173): per segment either a
// fresh Uniform(0.3, 2.5) state (prob alpha/(N+alpha)) or the state of the cluster of a
// uniformly chosen cell (== cluster chosen with prob npoints/(N+alpha), counts including
// the current cell as in the reference).
template <int NCH>
__device__ __forceinline__ void propose(const ChainDev &D, int c, int tt, uint32_t attempt, const int *Jrow,
                                        int lane, double (&mu)[NCH], int (&jj)[NCH])
{
    const double NpA = (double)D.N + D.alpha;
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        const int r = lane + 32 * i;
        if (r < D.R) {
            int j;
            double tm = 0.0;
            if (Jrow) {
                j = Jrow[r];
                if (j < 0) {
                    double ua, ub;
                    cs_philox_draw(D.seed, (uint32_t)c, (uint32_t)r, (uint32_t)tt, CS_PH_PROPOSAL, attempt, ua, ub);
                    tm = __fma_rn(2.5 - 0.3, ua, 0.3);
                }
            } else {
                double ua, ub;
                cs_philox_draw(D.seed, (uint32_t)c, (uint32_t)r, (uint32_t)tt, CS_PH_PROPOSAL, attempt, ua, ub);
                tm = __fma_rn(2.5 - 0.3, ua, 0.3);
                const double v = __dmul_rn(ub, NpA);
                if (v < (double)D.N) {
                    j = (int)v;
                    if (j > D.N - 1) j = D.N - 1;
                } else {
                    j = -1;
                }
            }
            jj[i] = j;
            mu[i] = (j >= 0) ? D.U[(size_t)D.Zs[j] * D.R + r] : tm;
        } else {
            jj[i] = -1;
            mu[i] = 0.0;
        }
    }
}

// any(apply(Ut, 1, identical, mu_new))  (R/BayesNonparCluster.R:174)
template <int NCH>
__device__ __forceinline__ bool is_duplicate(const ChainDev &D, int nlive, const double (&mu)[NCH], int lane)
{
    for (int s = 0; s < nlive; s++) {
        const double *u = D.U + (size_t)s * D.R;
        bool same = true;
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            const int r = lane + 32 * i;
            if (r < D.R) {
                const double uv = u[r];
                if (!(uv == mu[i]) && !(cs_isna(uv) && cs_isna(mu[i]))) same = false;
            }
        }
        if (__all_sync(CS_FULL, same)) return true;
    }
    return false;
}

// categorical draw over the live clusters (slot order == label order) and the new table
// (R/BayesNonparCluster.R:193-207).  Returns the slot, or nlive for "open a new cluster".
__device__ __forceinline__ int draw_cell(const double *Qrow, int nlive, const int *s_np, int zold, double q_new,
                                         double beta, double ua, int lane)
{
    double qmin = q_new;
    for (int s0 = 0; s0 < nlive; s0 += 32) {
        const int s = s0 + lane;
        if (s < nlive) {
            const int n = s_np[s] - (s == zold ? 1 : 0);
            if (n > 0) qmin = fmin(qmin, Qrow[s]);
        }
    }
    qmin = cs_warp_min(qmin);
    const double w_new = __dmul_rn(beta, exp(-0.5 * (q_new - qmin)));
    double w0 = 0.0; // weight of this lane's slot in the first group (kept for pass 2)
    double run = 0.0;
    for (int s0 = 0; s0 < nlive; s0 += 32) {
        const int s = s0 + lane;
        double w = 0.0;
        if (s < nlive) {
            const int n = s_np[s] - (s == zold ? 1 : 0);
            if (n > 0) w = __dmul_rn((double)n, exp(-0.5 * (Qrow[s] - qmin)));
        }
        if (s0 == 0) w0 = w;
        const int cnt = min(32, nlive - s0);
        for (int j = 0; j < cnt; j++) run = __dadd_rn(run, __shfl_sync(CS_FULL, w, j));
    }
    run = __dadd_rn(run, w_new);
    const double t = __dmul_rn(ua, run);
    run = 0.0;
    int z = -1, lastpos = -1;
    for (int s0 = 0; s0 < nlive && z < 0; s0 += 32) {
        const int s = s0 + lane;
        double w = w0;
        if (s0 != 0) {
            w = 0.0;
            if (s < nlive) {
                const int n = s_np[s] - (s == zold ? 1 : 0);
                if (n > 0) w = __dmul_rn((double)n, exp(-0.5 * (Qrow[s] - qmin)));
            }
        }
        const int cnt = min(32, nlive - s0);
        for (int j = 0; j < cnt; j++) {
            const double wj = __shfl_sync(CS_FULL, w, j);
            if (wj > 0.0) {
                run = __dadd_rn(run, wj);
                lastpos = s0 + j;
                if (run > t) {
                    z = s0 + j;
                    break;
                }
            }
        }
    }
    if (z < 0) {
        if (w_new > 0.0) {
            run = __dadd_rn(run, w_new);
            lastpos = nlive;
            if (run > t) z = nlive;
        }
        if (z < 0) z = lastpos;
    }
    return z;
}

// ---------------------------------------------------------------- prepare (all SMs)
template <int NCH>
__global__ void __launch_bounds__(PREP_THREADS)
prepare_kernel(ChainDev D, int tt, int *__restrict__ rej)
{
    extern __shared__ double s_dyn[];
    double *s_isig = s_dyn;
    int *s_np = reinterpret_cast<int *>(s_dyn + D.R);
    const int nlive = D.scal->nlive;
    for (int r = threadIdx.x; r < D.R; r += PREP_THREADS) s_isig[r] = D.isig[r];
    for (int s = threadIdx.x; s < nlive; s += PREP_THREADS) s_np[s] = D.np[s];
    __syncthreads();
    const bool reject_live = D.single && tt == 0;
    const int lane = threadIdx.x & 31;
    const int wpb = PREP_THREADS / 32;
    int first = INT_MAX;
    for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < D.N; c += gridDim.x * wpb) {
        double x[NCH], mu[NCH];
        int jj[NCH];
        load_row<NCH>(D.X + (size_t)c * D.R, D.R, lane, x);
        const int zold = D.Zs[c];
        double *Qrow = D.Q + (size_t)c * D.kcap;
        for (int s0 = 0; s0 < nlive; s0 += 32) {
            double myq = 0.0;
            const int cnt = min(32, nlive - s0);
            for (int j = 0; j < cnt; j++) {
                const double q = q_row<NCH>(x, D.U + (size_t)(s0 + j) * D.R, s_isig, D.R, lane);
                if (lane == j) myq = q;
            }
            if (lane < cnt) Qrow[s0 + lane] = myq;
        }
        uint32_t attempt = 0;
        for (;;) {
            propose<NCH>(D, c, tt, attempt, nullptr, lane, mu, jj);
            if (!reject_live || !is_duplicate<NCH>(D, nlive, mu, lane)) break;
            attempt++;
        }
        if (reject_live && lane == 0) rej[c] = (int)attempt;
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            const int r = lane + 32 * i;
            if (r < D.R) D.J[(size_t)c * D.R + r] = jj[i];
        }
        const double q_new = q_regs<NCH>(x, mu, s_isig, D.R, lane);
        double ua, ub;
        cs_philox_draw(D.seed, (uint32_t)c, 0u, (uint32_t)tt, CS_PH_ASSIGN, 0u, ua, ub);
        __syncwarp();
        const int z = draw_cell(Qrow, nlive, s_np, zold, q_new, D.beta, ua, lane);
        if (lane == 0) {
            D.zspec[c] = z;
            D.qcols[c] = nlive;
            if (z != zold && c < first) first = c;
        }
    }
    if (lane == 0 && first != INT_MAX) atomicMin(&D.scal->first_event, first);
}

// ---------------------------------------------------------------- scan (one CTA)
template <int NCH>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(ChainDev D, int tt, int force_all, int *__restrict__ rej)
{
    extern __shared__ double s_dyn[];
    double *s_isig = s_dyn;
    int *s_np = reinterpret_cast<int *>(s_dyn + D.R);
    int *s_label = s_np + D.kcap;
    __shared__ int s_nlive, s_kt, s_fmin[2], s_err;
    __shared__ long long s_scanned, s_events;

    const int f0 = force_all ? 0 : D.scal->first_event;
    if (f0 >= D.N) return;
    for (int r = threadIdx.x; r < D.R; r += SCAN_THREADS) s_isig[r] = D.isig[r];
    for (int s = threadIdx.x; s < D.kcap; s += SCAN_THREADS) {
        s_np[s] = D.np[s];
        s_label[s] = D.slot_label[s];
    }
    if (threadIdx.x == 0) {
        s_nlive = D.scal->nlive;
        s_kt = D.scal->kt;
        s_fmin[0] = INT_MAX;
        s_fmin[1] = INT_MAX;
        s_err = 0;
        s_scanned = 0;
        s_events = 0;
    }
    __syncthreads();
    const bool reject_live = D.single && tt == 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool use_J = !reject_live && !force_all;
    int a = f0, round = 0;
    while (a < D.N) {
        const int c = a + warp;
        const int nlive = s_nlive;
        double x[NCH], mu[NCH];
        int jj[NCH];
        int z = -1, zold = -1;
        if (c < D.N) {
            load_row<NCH>(D.X + (size_t)c * D.R, D.R, lane, x);
            zold = D.Zs[c];
            double *Qrow = D.Q + (size_t)c * D.kcap;
            uint32_t attempt = 0;
            for (;;) {
                propose<NCH>(D, c, tt, attempt, use_J ? D.J + (size_t)c * D.R : nullptr, lane, mu, jj);
                if (!reject_live || !is_duplicate<NCH>(D, nlive, mu, lane)) break;
                attempt++;
            }
            if (reject_live && lane == 0) rej[c] = (int)attempt;
            // Q columns of the clusters opened since this row was last scored
            for (int s = D.qcols[c]; s < nlive; s++) {
                const double q = q_row<NCH>(x, D.U + (size_t)s * D.R, s_isig, D.R, lane);
                if (lane == 0) Qrow[s] = q;
            }
            if (lane == 0) D.qcols[c] = nlive;
            const double q_new = q_regs<NCH>(x, mu, s_isig, D.R, lane);
            double ua, ub;
            cs_philox_draw(D.seed, (uint32_t)c, 0u, (uint32_t)tt, CS_PH_ASSIGN, 0u, ua, ub);
            __syncwarp();
            z = draw_cell(Qrow, nlive, s_np, zold, q_new, D.beta, ua, lane);
            if (z != zold && lane == 0) atomicMin(&s_fmin[round & 1], c);
        }
        __syncthreads();
        const int f = s_fmin[round & 1];
        if (f != INT_MAX && c == f) { // this warp owns the first event of the window: commit it
            if (z == nlive) {
                if (nlive >= D.kcap) {
                    if (lane == 0) s_err = CS_ERR_KCAP;
                } else {
                    double *urow = D.U + (size_t)nlive * D.R;
#pragma unroll
                    for (int i = 0; i < NCH; i++) {
                        const int r = lane + 32 * i;
                        if (r < D.R) urow[r] = mu[i];
                    }
                    if (lane == 0) {
                        s_np[zold] -= 1;
                        s_np[nlive] = 1;
                        s_label[nlive] = ++s_kt;
                        s_nlive = nlive + 1;
                        D.Zs[c] = nlive;
                    }
                }
            } else if (lane == 0) {
                s_np[zold] -= 1;
                s_np[z] += 1;
                D.Zs[c] = z;
            }
            if (lane == 0) s_events += 1;
        }
        if (threadIdx.x == 0) {
            s_scanned += (f != INT_MAX ? (f - a + 1) : min(SCAN_WARPS, D.N - a));
            s_fmin[(round + 1) & 1] = INT_MAX; // next round's slot; this round's was read above
        }
        __syncthreads();
        if (s_err) break;
        a = (f != INT_MAX) ? f + 1 : a + SCAN_WARPS;
        round++;
    }
    for (int s = threadIdx.x; s < D.kcap; s += SCAN_THREADS) {
        D.np[s] = s_np[s];
        D.slot_label[s] = s_label[s];
    }
    if (threadIdx.x == 0) {
        D.scal->nlive = s_nlive;
        D.scal->kt = s_kt;
        D.scal->cells_scanned += s_scanned;
        D.scal->n_events += s_events;
        if (s_err) D.scal->err = s_err;
    }
}

// ---------------------------------------------------------------- end of sweep
// per-chunk per-cluster sums of X and non-NA counts (R/BayesNonparCluster.R:223)
__global__ void __launch_bounds__(256)
stats_partial_kernel(ChainDev D, int smem_doubles)
{
    extern __shared__ double s_dyn[];
    const int nlive = D.scal->nlive;
    const int R = D.R;
    const int ST = max(1, smem_doubles / (R + (R + 1) / 2)); // slots per pass (double acc + int cnt)
    double *acc = s_dyn;
    int *cnt = reinterpret_cast<int *>(s_dyn + (size_t)ST * R);
    const int i0 = blockIdx.x * CS_STAT_CHUNK, i1 = min(D.N, i0 + CS_STAT_CHUNK);
    for (int t0 = 0; t0 < nlive; t0 += ST) {
        const int ns = min(ST, nlive - t0);
        for (int e = threadIdx.x; e < ns * R; e += blockDim.x) {
            acc[e] = 0.0;
            cnt[e] = 0;
        }
        __syncthreads();
        for (int i = i0; i < i1; i++) {
            const int s = D.Zs[i] - t0;
            if (s < 0 || s >= ns) continue;
            const double *x = D.X + (size_t)i * R;
            for (int r = threadIdx.x; r < R; r += blockDim.x) {
                const double v = x[r];
                if (!cs_isna(v)) {
                    acc[s * R + r] += v;
                    cnt[s * R + r] += 1;
                }
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < ns * R; e += blockDim.x) {
            const size_t o = ((size_t)blockIdx.x * D.kcap + t0) * R + e;
            D.psum[o] = acc[e];
            D.pcnt[o] = cnt[e];
        }
        __syncthreads();
    }
}

__global__ void stats_final_kernel(ChainDev D, int nchunks)
{
    const int nlive = D.scal->nlive;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nlive * D.R) return;
    double s = 0.0;
    long long c = 0;
    for (int ch = 0; ch < nchunks; ch++) {
        const size_t o = (size_t)ch * D.kcap * D.R + e;
        s += D.psum[o];
        c += D.pcnt[o];
    }
    D.mean[e] = c ? s / (double)c : 0.0;
}

// redraw of the cluster means (R/BayesNonparCluster.R:221-232; :260-266 on the last sweep)
__global__ void __launch_bounds__(128)
redraw_kernel(ChainDev D, int tt, int last)
{
    const int s = blockIdx.x;
    if (s >= D.scal->nlive) return;
    const int n = D.np[s];
    if (n == 0) return; // emptied cluster: its row becomes NA in Uall and is dropped at compaction
    const double *m = D.mean + (size_t)s * D.R;
    double *u = D.U + (size_t)s * D.R;
    int nz = 0;
    for (int r = threadIdx.x; r < D.R; r += blockDim.x) nz |= !(m[r] == 0.0);
    const int anynz = __syncthreads_or(nz);
    const double sq = sqrt((double)n);
    const uint32_t label = (uint32_t)D.slot_label[s];
    for (int r = threadIdx.x; r < D.R; r += blockDim.x) {
        double v;
        if (last) {
            v = m[r];
        } else if (!anynz) {
            v = 0.0;
        } else {
            const double sd = D.sig[r] / sq;
            if (sd == 0.0) {
                v = m[r];
            } else {
                double ua, ub;
                cs_philox_draw(D.seed, label, (uint32_t)r, (uint32_t)tt, CS_PH_RNORM, 0u, ua, ub);
                v = __dadd_rn(m[r], __dmul_rn(sd, cs_qnorm(ua)));
            }
            if (v < 0.0) v = 0.0;
        }
        u[r] = v;
    }
}

// per-chunk per-segment squared residuals (R/BayesNonparCluster.R:234)
__global__ void __launch_bounds__(256)
sigma_partial_kernel(ChainDev D)
{
    const int i0 = blockIdx.x * CS_STAT_CHUNK, i1 = min(D.N, i0 + CS_STAT_CHUNK);
    for (int r = threadIdx.x; r < D.R; r += blockDim.x) {
        double s = 0.0;
        int c = 0;
        for (int i = i0; i < i1; i++) {
            const double x = D.X[(size_t)i * D.R + r];
            if (cs_isna(x)) continue;
            c++;
            const double d = x - D.U[(size_t)D.Zs[i] * D.R + r];
            if (!cs_isna(d)) s = __fma_rn(d, d, s);
        }
        D.psum[(size_t)blockIdx.x * D.R + r] = s;
        D.pcnt[(size_t)blockIdx.x * D.R + r] = c;
    }
}

__global__ void sigma_final_kernel(ChainDev D, int nchunks, int tt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= D.R) return;
    double s = 0.0;
    long long c = 0;
    for (int ch = 0; ch < nchunks; ch++) {
        s += D.psum[(size_t)ch * D.R + r];
        c += D.pcnt[(size_t)ch * D.R + r];
    }
    const double sg = sqrt((1.0 / (double)c) * s);
    D.sig[r] = sg;
    D.isig[r] = 1.0 / sg;
    if (tt >= 0) D.sigma_all[(size_t)tt * D.R + r] = sg;
}

// total fixed-state log-likelihood (R/BayesNonparCluster.R:133,245): per-chunk partials
__global__ void __launch_bounds__(256)
ll_partial_kernel(ChainDev D, const double *__restrict__ U, const int *__restrict__ Zs, const double *__restrict__ sig)
{
    __shared__ double s_w[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * CS_STAT_CHUNK, i1 = min(D.N, i0 + CS_STAT_CHUNK);
    double acc = 0.0;
    for (int i = i0 + warp; i < i1; i += 8) {
        const double *x = D.X + (size_t)i * D.R;
        const double *u = U + (size_t)Zs[i] * D.R;
        double t = 0.0;
        for (int r = lane; r < D.R; r += 32) {
            const double xv = x[r];
            if (!cs_isna(xv)) {
                const double sg = sig[r];
                const double zz = (xv - u[r]) / sg;
                const double term = -(CS_LN_SQRT_2PI + 0.5 * zz * zz + log(sg));
                if (!cs_isna(term)) t += term;
            }
        }
        acc += cs_warp_sum_bfly(t);
    }
    if (lane == 0) s_w[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += s_w[w];
        D.psum[blockIdx.x] = s;
    }
}

__global__ void ll_final_kernel(ChainDev D, int nchunks, double *__restrict__ out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int ch = 0; ch < nchunks; ch++) s += D.psum[ch];
        *out = s;
    }
}

// record Uall[[tt]] (non-NA rows), kt; compute the compaction map
__global__ void __launch_bounds__(256)
record_kernel(ChainDev D, int tt)
{
    __shared__ int s_nocc;
    __shared__ long long s_off;
    const int nlive = D.scal->nlive;
    if (threadIdx.x == 0) {
        int k = 0;
        for (int s = 0; s < nlive; s++) D.newslot[s] = (D.np[s] > 0) ? k++ : -1;
        s_nocc = k;
        s_off = D.scal->u_used;
        D.kt_all[tt] = D.scal->kt;
        D.u_off[tt] = s_off;
        D.u_off[tt + 1] = s_off + k;
        if (s_off + k > D.urec_cap) D.scal->err = CS_ERR_UCAP;
        D.scal->u_used = s_off + k;
    }
    __syncthreads();
    const bool fits = s_off + s_nocc <= D.urec_cap;
    for (int s = 0; s < nlive; s++) {
        const int k = D.newslot[s];
        if (k < 0) continue;
        const double *u = D.U + (size_t)s * D.R;
        double *un = D.Unext + (size_t)k * D.R;
        for (int r = threadIdx.x; r < D.R; r += blockDim.x) {
            const double v = u[r];
            un[r] = v;
            if (fits) D.u_rows[(size_t)(s_off + k) * D.R + r] = v;
        }
        if (fits && threadIdx.x == 0) D.u_labels[s_off + k] = D.slot_label[s];
    }
}

__global__ void relabel_kernel(ChainDev D, int tt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.N) return;
    const int s = D.Zs[i];
    D.Zall[(size_t)tt * D.N + i] = D.slot_label[s];
    D.Zs[i] = D.newslot[s];
}

__global__ void compact_kernel(ChainDev D, int *__restrict__ rej, int reject_live)
{
    __shared__ long long s_rej[256];
    if (reject_live) { // rejected proposals of the first sweep (diagnostic parity with the oracle)
        long long t = 0;
        for (int i = threadIdx.x; i < D.N; i += blockDim.x) t += rej[i];
        s_rej[threadIdx.x] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long tot = 0;
            for (int i = 0; i < (int)blockDim.x; i++) tot += s_rej[i];
            D.scal->n_reject += tot;
        }
    }
    if (threadIdx.x == 0) {
        const int nlive = D.scal->nlive;
        int k = 0;
        for (int s = 0; s < nlive; s++) {
            if (D.newslot[s] >= 0) {
                D.np[k] = D.np[s];
                D.slot_label[k] = D.slot_label[s];
                k++;
            }
        }
        D.scal->nlive = k;
        D.scal->first_event = INT_MAX;
    }
}

template <int NCH>
int launch_sweep(const ChainDev &D, int tt, int last, int force_all, int *rej, int nsm, cudaStream_t st,
                 cudaEvent_t *ev)
{
    const size_t smem_prep = sizeof(double) * D.R + sizeof(int) * D.kcap;
    const size_t smem_scan = sizeof(double) * D.R + sizeof(int) * 2 * D.kcap;
    if (smem_scan > 48 * 1024) {
        CS_CUDA(cudaFuncSetAttribute(prepare_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep));
        CS_CUDA(cudaFuncSetAttribute(scan_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scan));
    }
    if (ev) CS_CUDA(cudaEventRecord(ev[0], st));
    if (!force_all) {
        int grid = (D.N + PREP_THREADS / 32 - 1) / (PREP_THREADS / 32);
        if (grid > nsm * 8) grid = nsm * 8;
        prepare_kernel<NCH><<<grid, PREP_THREADS, smem_prep, st>>>(D, tt, rej);
    } else {
        CS_CUDA(cudaMemsetAsync(D.qcols, 0, sizeof(int) * (size_t)D.N, st));
    }
    if (ev) CS_CUDA(cudaEventRecord(ev[1], st));
    scan_kernel<NCH><<<1, SCAN_THREADS, smem_scan, st>>>(D, tt, force_all, rej);
    if (ev) CS_CUDA(cudaEventRecord(ev[2], st));
    (void)last;
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

} // namespace

int cs_philox_max_R() { return 1024; }

// one full sweep (asynchronous).  D.U / D.Unext are swapped by the caller afterwards.
int cs_philox_sweep(const ChainDev &D, int tt, int last, int force_all, int *rej, cudaStream_t st,
                    cudaEvent_t *ev)
{
    const int nsm = cs_num_sms();
    int rc;
    if (D.R <= 32) rc = launch_sweep<1>(D, tt, last, force_all, rej, nsm, st, ev);
    else if (D.R <= 64) rc = launch_sweep<2>(D, tt, last, force_all, rej, nsm, st, ev);
    else if (D.R <= 128) rc = launch_sweep<4>(D, tt, last, force_all, rej, nsm, st, ev);
    else if (D.R <= 320) rc = launch_sweep<10>(D, tt, last, force_all, rej, nsm, st, ev);
    else if (D.R <= 1024) rc = launch_sweep<32>(D, tt, last, force_all, rej, nsm, st, ev);
    else {
        cs_set_error("Philox sampler supports up to 1024 segments (R=%d)", D.R);
        return CS_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
    const int nchunks = (D.N + CS_STAT_CHUNK - 1) / CS_STAT_CHUNK;
    const int stat_smem = 160 * 1024;
    CS_CUDA(cudaFuncSetAttribute(stats_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, stat_smem));
    stats_partial_kernel<<<nchunks, 256, stat_smem, st>>>(D, stat_smem / 8);
    stats_final_kernel<<<(D.kcap * D.R + 255) / 256, 256, 0, st>>>(D, nchunks);
    redraw_kernel<<<D.kcap, 128, 0, st>>>(D, tt, last);
    sigma_partial_kernel<<<nchunks, 256, 0, st>>>(D);
    sigma_final_kernel<<<(D.R + 127) / 128, 128, 0, st>>>(D, nchunks, tt);
    ll_partial_kernel<<<nchunks, 256, 0, st>>>(D, D.U, D.Zs, D.sig);
    ll_final_kernel<<<1, 32, 0, st>>>(D, nchunks, D.LL + tt);
    record_kernel<<<1, 256, 0, st>>>(D, tt);
    relabel_kernel<<<(D.N + 255) / 256, 256, 0, st>>>(D, tt);
    compact_kernel<<<1, 256, 0, st>>>(D, rej, D.single && tt == 0);
    if (ev) CS_CUDA(cudaEventRecord(ev[3], st));
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// L0 = total log-likelihood at (U0, Z0 (0-based rows), sigmas0)   (R/BayesNonparCluster.R:133)
int cs_total_loglik_async(const ChainDev &D, const double *U, const int *Zrows, const double *sig, double *d_out,
                          cudaStream_t st)
{
    const int nchunks = (D.N + CS_STAT_CHUNK - 1) / CS_STAT_CHUNK;
    ll_partial_kernel<<<nchunks, 256, 0, st>>>(D, U, Zrows, sig);
    ll_final_kernel<<<1, 32, 0, st>>>(D, nchunks, d_out);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
