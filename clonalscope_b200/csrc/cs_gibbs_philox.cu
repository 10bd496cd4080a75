// cs_gibbs_philox.cu — nested-CRP collapsed Gibbs sweep, Philox mode.
//
// Reference semantics: R/BayesNonparCluster.R:159-245 — cells are updated one after the
// other, every draw sees the counts / assignments / appended rows left by the previous
// cell.  That order is kept EXACTLY, but evaluated speculatively:
//
//   prepare_kernel (all SMs, one warp per cell)
//       scores every cell against all live clusters (Q), draws its nested-CRP proposal
//       (:165-175) and its categorical draw (:193-207) against the state at the START of
//       the sweep, and caches what a later re-draw needs (the exponentials E, their
//       reference value qmin, the draw's uniform, the proposal sources J).  As long as no
//       earlier cell changed its label that state is the state the sequential sampler would
//       have seen (a cell that re-draws its own label leaves counts and rows untouched), so
//       every draw up to and including the first label change ("event") is already the
//       sequential result.
//   index kernels (all SMs; skipped when the sweep has no event)
//       inverse source index: for every cell f the (cell, segment) pairs whose proposal
//       copies f's cluster state — what an event at f invalidates.
//   scan_ring_kernel / scan_kernel (one CTA walks, thread per cell)
//       resumes at the first event and walks the rest of the sweep in windows of 256
//       cells: each thread re-draws its cell against the committed state from the cached
//       exponentials (no exp, no gathers: counts are the only thing an event changes for
//       almost every cell) and keeps a margin — how far its sums may move before the draw
//       could change; the window's events are committed in order for as long as every draw
//       in front of the next one provably still holds, their dependents' proposal
//       distances are corrected in place (fixed-point, so the correction is exact), and
//       the walk continues behind the last one.  scan_ring_kernel (<= 32 live clusters)
//       keeps what a re-draw needs in a shared-memory ring filled by asynchronous copies
//       and runs as a cluster of CTAs whose other members score the column of a newly
//       opened cluster; scan_kernel is the generic walk against global memory.
//   seq_scan_kernel (one CTA, warp per cell): the plain sequential evaluation, used for the
//       first sweep of a single-cluster start (rejection loop live, :174) and as the
//       test hook that proves the speculative path draws the same chain.
//   end of sweep (:221-245): deterministic chunked reductions for cluster means and
//       sigmas, Philox/AS241 redraw of the means, total log-likelihood, compaction of
//       emptied clusters (labels keep growing, rows are never renumbered).
//
// Random numbers are counter-based (Philox4x32-10 keyed by seed; counter = cell, segment,
// sweep, purpose|attempt), so a draw does not depend on how many draws preceded it — that
// is what makes the speculation exact.  Squared distances are accumulated in 64-bit fixed
// point (2^-32), so they do not depend on summation order and can be corrected term by term;
// the CPU oracle's Philox mode uses the same definition.
#include "cs_gibbs.cuh"

#include <cooperative_groups.h>
#include <limits.h>
#include <stdlib.h>
#include <math_constants.h>

namespace cg = cooperative_groups;

namespace {

constexpr int PREP_THREADS = 256;
constexpr int SCAN_THREADS = 1024;  // sequential kernel: one warp per cell
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int WIN_THREADS = 512;    // windowed kernel: one thread per cell of the window (the upper
                                    // half only helps with corrections and new columns);
constexpr int WIN_MIN = 256;        // cells re-drawn per round.  (The width can adapt between WIN_MIN and WIN_CELLS
constexpr int WIN_CELLS = 256;      // to how far rounds get; measured: 64-128 cell windows lose more to fewer events
                                    // per round than they gain from cheaper rounds, so both are 256.)
constexpr double Q_SCALE = 1.0 / 4294967296.0;

template <int NCH>
__device__ __forceinline__ void load_row(const double *__restrict__ xrow, int R, int lane, double (&x)[NCH])
{
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        const int r = lane + 32 * i;
        x[i] = (r < R) ? xrow[r] : 0.0;
    }
}

// one segment's contribution, 2^-32 fixed point: llrint(((x - u) * 65536 / sigma)^2)
__device__ __forceinline__ long long tint(double x, double u, double isig16)
{
    const double d = __dmul_rn(__dsub_rn(x, u), isig16);
    double t = __dmul_rn(d, d);
    if (!(t < 4611686018427387904.0)) t = 4611686018427387904.0;
    return __double2ll_rn(t);
}

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(CS_FULL, v, off);
    return v;
}

// fixed-point squared standardised distance of one cell (registers) to one row
template <int NCH>
__device__ __forceinline__ long long q_row(const double (&x)[NCH], const double *u, const double *s_isig16, int R,
                                           int lane)
{
    long long s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        const int r = lane + 32 * i;
        if (r < R) {
            const double xv = x[i];
            if (!cs_isna(xv)) s += tint(xv, u[r], s_isig16[r]);
        }
    }
    return warp_sum_ll(s);
}

template <int NCH>
__device__ __forceinline__ long long q_regs(const double (&x)[NCH], const double (&mu)[NCH], const double *s_isig16,
                                            int R, int lane)
{
    long long s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        const int r = lane + 32 * i;
        if (r < R) {
            const double xv = x[i];
            if (!cs_isna(xv)) s += tint(xv, mu[i], s_isig16[r]);
        }
    }
    return warp_sum_ll(s);
}

__device__ __forceinline__ double fresh_state(const ChainDev &D, int c, int r, int tt, uint32_t attempt)
{
    double ua, ub;
    cs_philox_draw(D.seed, (uint32_t)c, (uint32_t)r, (uint32_t)tt, CS_PH_PROPOSAL, attempt, ua, ub);
    return __fma_rn(2.5 - 0.3, ua, 0.3);
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
                if (j < 0) tm = fresh_state(D, c, r, tt, attempt);
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
// (R/BayesNonparCluster.R:193-207), one warp.  Qc / Ec point at this cell's element of slot
// 0 (slot stride = N).  Returns the slot, or nlive for "open a new cluster".  When Ec is
// given the exponentials are stored for later re-draws; *qmin_out receives their reference.
__device__ __forceinline__ int draw_cell_warp(const double *Qc, double *Ec, size_t stride, int nlive,
                                              const int *s_np, int zold, double q_new, double beta, double ua,
                                              int lane, double *qmin_out)
{
    double qmin = q_new;
    for (int s0 = 0; s0 < nlive; s0 += 32) {
        const int s = s0 + lane;
        if (s < nlive) {
            const int n = s_np[s] - (s == zold ? 1 : 0);
            if (n > 0) qmin = fmin(qmin, Qc[(size_t)s * stride]);
        }
    }
    qmin = cs_warp_min(qmin);
    if (qmin_out) *qmin_out = qmin;
    const double w_new = __dmul_rn(beta, exp(-0.5 * (q_new - qmin)));
    double run = 0.0;
    for (int s0 = 0; s0 < nlive; s0 += 32) {
        const int s = s0 + lane;
        double w = 0.0;
        if (s < nlive) {
            const double e = exp(-0.5 * (Qc[(size_t)s * stride] - qmin));
            if (Ec) Ec[(size_t)s * stride] = e;
            const int n = s_np[s] - (s == zold ? 1 : 0);
            if (n > 0) w = __dmul_rn((double)n, e);
        }
        const int cnt = min(32, nlive - s0);
        for (int j = 0; j < cnt; j++) run = __dadd_rn(run, __shfl_sync(CS_FULL, w, j));
    }
    run = __dadd_rn(run, w_new);
    const double t = __dmul_rn(ua, run);
    run = 0.0;
    int z = -1, lastpos = -1;
    for (int s0 = 0; s0 < nlive && z < 0; s0 += 32) {
        const int s = s0 + lane;
        double w = 0.0;
        if (s < nlive) {
            const int n = s_np[s] - (s == zold ? 1 : 0);
            if (n > 0) w = __dmul_rn((double)n, exp(-0.5 * (Qc[(size_t)s * stride] - qmin)));
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

// the same draw by ONE thread from cached exponentials: slots below ncached use Ec when the
// cached reference value still is the minimum over today's candidates, everything else is
// re-exponentiated.  Same operations in the same order as draw_cell_warp.
__device__ __forceinline__ int draw_cell_thread(const double *Qc, const double *Ec, size_t stride, int ncached,
                                                int nlive, const int *s_np, int zold, double q_new,
                                                double qmin_cached, double beta, double ua)
{
    constexpr int CH = 8; // loads are issued CH at a time, unconditionally, so they overlap
    double qmin = q_new;
    for (int s0 = 0; s0 < nlive; s0 += CH) {
        double q[CH];
#pragma unroll
        for (int i = 0; i < CH; i++) q[i] = (s0 + i < nlive) ? Qc[(size_t)(s0 + i) * stride] : 0.0;
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int s = s0 + i;
            if (s < nlive && s_np[s] - (s == zold ? 1 : 0) > 0) qmin = fmin(qmin, q[i]);
        }
    }
    const int nfast = (qmin == qmin_cached) ? ncached : 0;
    const double w_new = __dmul_rn(beta, exp(-0.5 * (q_new - qmin)));
    double run = 0.0;
    for (int s0 = 0; s0 < nlive; s0 += CH) {
        double e[CH];
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int s = s0 + i;
            e[i] = (s < nfast) ? Ec[(size_t)s * stride] : ((s < nlive) ? Qc[(size_t)s * stride] : 0.0);
        }
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int s = s0 + i;
            if (s < nlive) {
                const int n = s_np[s] - (s == zold ? 1 : 0);
                if (n > 0) {
                    const double ee = (s < nfast) ? e[i] : exp(-0.5 * (e[i] - qmin));
                    run = __dadd_rn(run, __dmul_rn((double)n, ee));
                }
            }
        }
    }
    run = __dadd_rn(run, w_new);
    const double t = __dmul_rn(ua, run);
    run = 0.0;
    int lastpos = -1;
    for (int s0 = 0; s0 < nlive; s0 += CH) {
        double e[CH];
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int s = s0 + i;
            e[i] = (s < nfast) ? Ec[(size_t)s * stride] : ((s < nlive) ? Qc[(size_t)s * stride] : 0.0);
        }
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int s = s0 + i;
            if (s < nlive) {
                const int n = s_np[s] - (s == zold ? 1 : 0);
                if (n > 0) {
                    const double ee = (s < nfast) ? e[i] : exp(-0.5 * (e[i] - qmin));
                    const double w = __dmul_rn((double)n, ee);
                    if (w > 0.0) {
                        run = __dadd_rn(run, w);
                        lastpos = s;
                        if (run > t) return s;
                    }
                }
            }
        }
    }
    if (w_new > 0.0) lastpos = nlive;
    return lastpos;
}

// the common case of the windowed scan: at most 32 live clusters, all exponentials cached,
// candidate set as at the start of the sweep.  One pass; the cumulative sums also go to shared
// memory (s_cum[s * WIN_CELLS + t]) so that the two that bracket the draw can be fetched by
// index.  s_npd[s] / s_npd1[s] = (double) np[s] / (double)(np[s] - 1), 0 when not positive.
//
// *margin receives how far every cumulative sum and the threshold may move before this draw
// could change: half the distance from t = u * total to the nearer of the two cumulative sums
// that bracket it, minus a rounding allowance (the sums have at most 33 non-negative terms:
// their floating-point error is below 1e-14 * total; 1e-12 is charged).  An event that moves
// one cell from slot a to slot b moves every sum and t by at most E[a] + E[b].
template <int KMAX>
__device__ __forceinline__ int draw_cell_fast(const double *Ec, size_t stride, int nlive, const double *s_npd,
                                              const double *s_npd1, int zold, double w_new, double ua,
                                              double *s_cum, double *s_e, double *margin)
{
    unsigned pos = 0;
    double run = 0.0;
#pragma unroll
    for (int s = 0; s < KMAX; s++) {
        if (s < nlive) {
            const double nd = (s == zold) ? s_npd1[s] : s_npd[s];
            const double e = Ec[(size_t)s * stride];
            s_e[s * WIN_CELLS] = e; // the margins are charged with these while events are committed
            const double w = __dmul_rn(nd, e);
            run = __dadd_rn(run, w);
            s_cum[s * WIN_CELLS] = run;
            if (w > 0.0) pos |= 1u << s;
        }
    }
    const double tot = __dadd_rn(run, w_new);
    const double t = __dmul_rn(ua, tot);
    unsigned le = 0; // cumulative sums are non-decreasing: those <= t form a prefix
#pragma unroll
    for (int s = 0; s < KMAX; s++)
        if (s < nlive && !(s_cum[s * WIN_CELLS] > t)) le |= 1u << s;
    const int k = __popc(le); // first slot whose cumulative sum exceeds t (it has positive weight)
    int z;
    double gap;
    const double below = t - (k > 0 ? s_cum[(k - 1) * WIN_CELLS] : 0.0);
    if (k < nlive) {
        z = k;
        gap = fmin(below, s_cum[k * WIN_CELLS] - t);
    } else if (w_new > 0.0) {
        z = nlive;
        gap = below;
    } else {
        z = pos ? 31 - __clz(pos) : -1;
        gap = 0.0;
    }
    *margin = 0.5 * gap - 1e-12 * tot;
    return z;
}

// ---------------------------------------------------------------- prepare (all SMs)
template <int NCH>
__global__ void __launch_bounds__(PREP_THREADS, 4)
prepare_kernel(ChainDev D, int tt, int *__restrict__ rej)
{
    extern __shared__ double s_dyn[];
    double *s_isig16 = s_dyn;
    int *s_np = reinterpret_cast<int *>(s_dyn + D.R);
    const int nlive = D.scal->nlive;
    for (int r = threadIdx.x; r < D.R; r += PREP_THREADS) s_isig16[r] = D.isig[r] * 65536.0;
    for (int s = threadIdx.x; s < nlive; s += PREP_THREADS) s_np[s] = D.np[s];
    if (blockIdx.x == 0) // the counts the speculative draws were made against (the walks consult them)
        for (int s = threadIdx.x; s < D.kcap; s += PREP_THREADS) D.np0[s] = s < nlive ? D.np[s] : 0;
    __syncthreads();
    const bool reject_live = D.single && tt == 0;
    const int lane = threadIdx.x & 31;
    const int wpb = PREP_THREADS / 32;
    const size_t N = (size_t)D.N;
    int first = INT_MAX;
    for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < D.N; c += gridDim.x * wpb) {
        double x[NCH], mu[NCH];
        int jj[NCH];
        load_row<NCH>(D.X + (size_t)c * D.R, D.R, lane, x);
        const int zold = D.Zs[c];
        double *Qc = D.Q + c, *Ec = D.E + c;
        for (int s0 = 0; s0 < nlive; s0 += 32) {
            double myq = 0.0;
            const int cnt = min(32, nlive - s0);
            for (int j = 0; j < cnt; j++) {
                const long long q = q_row<NCH>(x, D.U + (size_t)(s0 + j) * D.R, s_isig16, D.R, lane);
                if (lane == j) myq = (double)q * Q_SCALE;
            }
            if (lane < cnt) Qc[(size_t)(s0 + lane) * N] = myq;
        }
        uint32_t attempt = 0;
        for (;;) {
            propose<NCH>(D, c, tt, attempt, nullptr, lane, mu, jj);
            if (!reject_live || !is_duplicate<NCH>(D, nlive, mu, lane)) break;
            attempt++;
        }
        if (reject_live && lane == 0) rej[c] = (int)attempt;
        int nnear = 0; // sources at most CS_NEAR_W cells in front of this one: what the walk re-checks within a round
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            const int r = lane + 32 * i;
            if (r < D.R) D.J[(size_t)c * D.R + r] = jj[i];
            const bool nr = r < D.R && jj[i] >= 0 && jj[i] < c && c - jj[i] <= CS_NEAR_W;
            const unsigned nb = __ballot_sync(CS_FULL, nr);
            if (nr) {
                const int slot = nnear + __popc(nb & ((1u << lane) - 1u));
                if (slot < CS_NEAR_CAP) D.near[(size_t)c * CS_NEAR_CAP + slot] = ((unsigned)jj[i] << 10) | (unsigned)r;
            }
            nnear += __popc(nb);
        }
        if (lane < CS_NEAR_CAP && lane >= nnear) D.near[(size_t)c * CS_NEAR_CAP + lane] = CS_NEAR_EMPTY;
        if (nnear > CS_NEAR_CAP && lane == 0) D.near[(size_t)c * CS_NEAR_CAP + CS_NEAR_CAP - 1] = CS_NEAR_OVER;
        const long long qn = q_regs<NCH>(x, mu, s_isig16, D.R, lane);
        double ua, ub;
        cs_philox_draw(D.seed, (uint32_t)c, 0u, (uint32_t)tt, CS_PH_ASSIGN, 0u, ua, ub);
        __syncwarp();
        double qmin;
        const double q_new = (double)qn * Q_SCALE;
        const int z = draw_cell_warp(Qc, Ec, N, nlive, s_np, zold, q_new, D.beta, ua, lane, &qmin);
        // a live candidate that attains qmin (255: only the proposal does).  While it stays a
        // candidate, and nothing smaller joins, the cached exponentials stay exact.
        int km = 255;
        for (int s = lane; s < nlive && s < 255; s += 32)
            if (km == 255 && s_np[s] - (s == zold ? 1 : 0) > 0 && Qc[(size_t)s * N] == qmin) km = s;
        km = cs_warp_min_i(km);
        if (lane == 0) {
            D.zspec[c] = z;
            D.qcols[c] = nlive;
            D.qnew_i[c] = qn;
            D.qmin[c] = qmin;
            D.ua[c] = ua;
            D.kmin[c] = km;
            D.wn_q[c] = qn;
            D.wn_w[c] = __dmul_rn(D.beta, exp(-0.5 * (q_new - qmin)));
            if (z != zold && c < first) first = c;
        }
    }
    if (lane == 0 && first != INT_MAX) atomicMin(&D.scal->first_event, first);
}

// ---------------------------------------------------------------- inverse source index
// icnt[j] = number of (cell, segment) proposals that copy cell j's cluster state
__global__ void index_count_kernel(ChainDev D)
{
    if (D.scal->first_event >= D.N) return;
    const size_t n = (size_t)D.N * D.R;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const int j = D.J[e];
        if (j >= 0 && (int)(e / (size_t)D.R) > j) atomicAdd(&D.icnt[j], 1); // only later cells are ever corrected
    }
}

// exclusive scan of icnt[0..N) in place (icnt[N] = total); one CTA, N is at most a few 10^6
__global__ void __launch_bounds__(1024) index_scan_kernel(ChainDev D)
{
    if (D.scal->first_event >= D.N) return;
    __shared__ int s_w[33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int carry = 0;
    for (int base = 0; base < D.N; base += 1024 * 4) {
        int v[4], sum = 0;
        const int i0 = base + threadIdx.x * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            v[k] = (i0 + k < D.N) ? D.icnt[i0 + k] : 0;
            sum += v[k];
        }
        int inc = sum;
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(CS_FULL, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const int w = s_w[lane];
            int winc = w;
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(CS_FULL, winc, off);
                if (lane >= off) winc += t;
            }
            s_w[lane] = winc - w;
            if (lane == 31) s_w[32] = winc;
        }
        __syncthreads();
        int ex = carry + s_w[wid] + inc - sum;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < D.N) {
                D.icnt[i0 + k] = ex;
                D.icur[i0 + k] = ex;
            }
            ex += v[k];
        }
        carry += s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) D.icnt[D.N] = carry;
}

__global__ void index_fill_kernel(ChainDev D)
{
    if (D.scal->first_event >= D.N) return;
    const size_t n = (size_t)D.N * D.R;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const int j = D.J[e];
        const int cc = (int)(e / (size_t)D.R);
        if (j >= 0 && cc > j) {
            const int p = atomicAdd(&D.icur[j], 1);
            InvEntry en; // (the value the correction needs rides along: one dependent load less)
            en.cr = ((unsigned)cc << 10) | (unsigned)(e % (size_t)D.R);
            en.pad = 0u;
            en.x = D.X[e];
            *reinterpret_cast<int4 *>(D.inv + p) = *reinterpret_cast<const int4 *>(&en);
        }
    }
}


// ---------------------------------------------------------------- near sources inside a round
// A round of the walks commits several events; a cell behind some of them keeps its draw when
// (i) the counts' changes stay inside its margin and (ii) so does the change of its new-table
// weight, if its proposal copied the cluster state of a cell that moves in the same round.
// This is (ii): the cell's near sources (prepare_kernel) are looked up among the round's
// candidates in front of it (s_cidx: candidate number of a window cell, 255 = none), the
// proposal distance is corrected exactly as the commit will correct it, and the margin is
// charged with the change of the weight (only the threshold u * total moves with it: by at
// most that change).  Returns the margin that is left (<= 0: the cell has to be re-drawn).
__device__ __forceinline__ double charge_near_sources(const ChainDev &D, const unsigned *nearc, int c, int pos, int K,
                                                      const unsigned char *s_cidx, const int *s_ev_a, const int *s_ev_b,
                                                      const double *s_isig16, long long qi, double qmin, double w_new,
                                                      double m)
{
    long long dq = 0;
    bool dep = false;
#pragma unroll 1
    for (int k = 0; k < CS_NEAR_CAP; k++) {
        const unsigned en = nearc[k];
        if (en == CS_NEAR_EMPTY) break;
        if (en == CS_NEAR_OVER) return -1.0; // (more near sources than the list holds: re-drawn next round)
        const int j = (int)(en >> 10);
        if (j < pos) continue;
        const int ci = s_cidx[j - pos];
        if (ci >= K) continue;               // (not a candidate, or one this round cannot reach)
        const int r = (int)(en & 1023u);
        const double xv = D.X[(size_t)c * D.R + r];
        dep = true;
        if (!cs_isna(xv))
            dq += tint(xv, D.U[(size_t)s_ev_b[ci] * D.R + r], s_isig16[r]) - tint(xv, D.U[(size_t)s_ev_a[ci] * D.R + r], s_isig16[r]);
    }
    if (dep && dq != 0) {
        const double q2 = (double)(qi + dq) * Q_SCALE;
        if (!(q2 >= qmin)) return -1.0;      // (the proposal would become the reference of the cached exponentials)
        const double w2 = __dmul_rn(D.beta, exp(-0.5 * (q2 - qmin)));
        m -= fabs(w2 - w_new) * 1.000000001 + 1e-300;
    }
    return m;
}

// ---------------------------------------------------------------- scan (one CTA, thread per cell)
constexpr int MAX_EV = 32; // events committed per window round

template <int NCH>
__global__ void __launch_bounds__(WIN_THREADS)
scan_kernel(ChainDev D, int tt)
{
    extern __shared__ double s_dyn[];
    double *s_isig16 = s_dyn;                 // R
    double *s_npd = s_dyn + D.R;              // kcap : (double) np[s]      (0 when np <= 0)
    double *s_npd1 = s_npd + D.kcap;          // kcap : (double)(np[s] - 1) (0 when np <= 1)
    double *s_cum = s_npd1 + D.kcap;          // 32 x WIN_CELLS cumulative sums of the fast draw
    double *s_e = s_cum + 32 * WIN_CELLS;     // 32 x WIN_CELLS exponentials of the window's cells
    int *s_np = reinterpret_cast<int *>(s_e + 32 * WIN_CELLS);
    int *s_label = s_np + D.kcap;
    int *s_np0 = s_label + D.kcap;            // counts at the start of the sweep
    __shared__ int s_nlive, s_kt, s_err, s_block, s_wcnt[WIN_THREADS / 32];
    // the round's event candidates in cell order: cell, old slot, new slot, dependents' range, nearest dependent
    __shared__ int s_ev_f[MAX_EV], s_ev_a[MAX_EV], s_ev_b[MAX_EV], s_ev_e0[MAX_EV], s_ev_e1[MAX_EV];
    __shared__ unsigned char s_cidx[WIN_CELLS]; // candidate number of every window cell (255: not a candidate)
    __shared__ long long s_scanned, s_events;
    __shared__ int s_diag[4]; // rounds, rounds cut short by a margin / an opened cluster / the event list

    const int f0 = D.scal->first_event;
    if (f0 >= D.N) return;
    for (int r = threadIdx.x; r < D.R; r += WIN_THREADS) s_isig16[r] = D.isig[r] * 65536.0;
    if (threadIdx.x < 4) s_diag[threadIdx.x] = 0;
    for (int s = threadIdx.x; s < D.kcap; s += WIN_THREADS) {
        const int n = D.np[s];
        s_np[s] = n;
        s_npd[s] = n > 0 ? (double)n : 0.0;
        s_npd1[s] = n > 1 ? (double)(n - 1) : 0.0;
        s_np0[s] = D.np0[s];
        s_label[s] = D.slot_label[s];
    }
    if (threadIdx.x == 0) {
        s_nlive = D.scal->nlive;
        s_kt = D.scal->kt;
        s_err = 0;
        s_scanned = 0;
        s_events = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t N = (size_t)D.N;
    int pos = f0, W = WIN_MIN;
    while (pos < D.N) {
        // ---- every cell of the window is re-drawn exactly against the committed state
        const int c = pos + threadIdx.x;
        const bool active = threadIdx.x < W && c < D.N;
        const int nlive = s_nlive;
        int z = -1, zold = -1;
        double margin = 0.0; // what this draw tolerates before it has to be redone
        if (active) {
            zold = D.Zs[c];
            const long long qi = __ldcg(D.qnew_i + c);
            const double q_new = (double)qi * Q_SCALE;
            const double qmin = D.qmin[c];
            const double ua = D.ua[c];
            // the cached exponentials are exactly today's when the cached minimum still is the
            // minimum over today's candidates: the cluster that attained it still is a candidate,
            // the proposal is not below it, and no cluster below it became a candidate (clusters
            // opened this sweep clear kmin themselves; the cell's own cluster joins the candidates
            // when it stops being a singleton)
            const int km = D.kmin[c];
            // (when only the proposal attained the cached minimum, it must still sit exactly there)
            bool fast = nlive <= 32 &&
                        (km != 255 ? (q_new >= qmin && s_np[km] - (km == zold ? 1 : 0) > 0) : q_new == qmin);
            if (fast && s_np0[zold] < 2 && s_np[zold] >= 2 && D.Q[(size_t)zold * N + c] < qmin) fast = false;
            if (fast) {
                double w_new = D.wn_w[c];
                if (D.wn_q[c] != qi) { // the proposal was corrected since the weight was cached
                    w_new = __dmul_rn(D.beta, exp(-0.5 * (q_new - qmin)));
                    D.wn_q[c] = qi;
                    D.wn_w[c] = w_new;
                }
                double *cum = s_cum + threadIdx.x;
                if (nlive <= 8) z = draw_cell_fast<8>(D.E + c, N, nlive, s_npd, s_npd1, zold, w_new, ua, cum, s_e + threadIdx.x, &margin);
                else if (nlive <= 16) z = draw_cell_fast<16>(D.E + c, N, nlive, s_npd, s_npd1, zold, w_new, ua, cum, s_e + threadIdx.x, &margin);
                else if (nlive <= 24) z = draw_cell_fast<24>(D.E + c, N, nlive, s_npd, s_npd1, zold, w_new, ua, cum, s_e + threadIdx.x, &margin);
                else z = draw_cell_fast<32>(D.E + c, N, nlive, s_npd, s_npd1, zold, w_new, ua, cum, s_e + threadIdx.x, &margin);
            } else {
                z = draw_cell_thread(D.Q + c, D.E + c, N, nlive, nlive, s_np, zold, q_new, qmin, D.beta, ua);
            }
        }
        // ---- the window's event candidates, in cell order (ballot + prefix over the warps)
        const bool cand = active && z != zold;
        const unsigned bal = __ballot_sync(CS_FULL, cand);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int jc = __popc(bal & ((1u << lane) - 1u)), ncand = 0; // candidates in front of this cell
        for (int w = 0; w < WIN_THREADS / 32; w++) {
            const int cw = s_wcnt[w];
            if (w < warp) jc += cw;
            ncand += cw;
        }
        if (cand && jc < MAX_EV) { // what a commit needs, where all threads can read it
            s_ev_f[jc] = c;
            s_ev_a[jc] = zold;
            s_ev_b[jc] = z;
            s_ev_e0[jc] = D.icnt[c];     // range of its dependents in the inverse index
            s_ev_e1[jc] = D.icnt[c + 1];
        }
        if (threadIdx.x < WIN_CELLS) s_cidx[threadIdx.x] = (cand && jc < MAX_EV) ? (unsigned char)jc : (unsigned char)255;
        if (threadIdx.x == 0) s_block = INT_MAX;
        __syncthreads();
        // ---- how many of them can be committed in this round, in order, with every draw in front
        // of each provably unchanged.  Candidate j is out when it lies behind a candidate that
        // opens a cluster (a column the window has not scored) ...
        int K = min(ncand, MAX_EV);
        for (int j = 0; j < K; j++) {
            if (s_ev_b[j] == nlive) {
                K = j + 1;
                break;
            }
        }
        // ... or when a cell between the previous candidate and itself (itself included) runs out
        // of margin: each of the jc events in front of a cell moves its sums and threshold by at
        // most E[a] + E[b].  (A cluster that stops or starts being a candidate for some cell needs
        // no special case: its weight n * E[.] changes by at most E[.], and re-referencing the
        // exponentials to another minimum rescales all weights of a cell by one factor — relative
        // effect ~1e-15, inside the allowance.)
        // ... or when the proposal of such a cell copied the cluster state of a cell that moves in this
        // round and its new-table weight changes by more than what is left of the margin.
        if (active && jc >= 1 && jc < K) {
            double m = margin;
            for (int i = 0; i < jc && m > 0.0; i++)
                m -= (s_e[s_ev_a[i] * WIN_CELLS + threadIdx.x] + s_e[s_ev_b[i] * WIN_CELLS + threadIdx.x]) * 1.000000001;
            if (m > 0.0)
                m = charge_near_sources(D, D.near + (size_t)c * CS_NEAR_CAP, c, pos, K, s_cidx, s_ev_a, s_ev_b, s_isig16,
                                        __ldcg(D.qnew_i + c), D.qmin[c], D.wn_w[c], m);
            if (!(m > 0.0)) atomicMin(&s_block, jc);
        }
        __syncthreads();
        const int K0 = K;
        K = min(K, s_block);
        const int nev = K;
        const bool opened = nev > 0 && s_ev_b[nev - 1] == nlive;
        const int lo = nev ? s_ev_f[nev - 1] + 1 : pos;
        // ---- commit them: counts (order-free integer updates), assignments
        if (threadIdx.x < nev) {
            atomicAdd(&s_np[s_ev_a[threadIdx.x]], -1);
            if (s_ev_b[threadIdx.x] == nlive) s_np[nlive] = 1; // (a slot past the live ones may hold a stale count)
            else atomicAdd(&s_np[s_ev_b[threadIdx.x]], 1);
            D.Zs[s_ev_f[threadIdx.x]] = s_ev_b[threadIdx.x];
        }
        __syncthreads();
        if (opened) { // :209-212 the last event opens a cluster with its proposed per-segment states
            const int f = s_ev_f[nev - 1], b = nlive;
            if (b >= D.kcap) {
                if (threadIdx.x == 0) s_err = CS_ERR_KCAP;
            } else {
                // (its own assignment is not a source of its proposal: restore the old slot for the gather)
                double *urow = D.U + (size_t)b * D.R;
                const int *Jrow = D.J + (size_t)f * D.R;
                const int a_f = s_ev_a[nev - 1];
                for (int r = threadIdx.x; r < D.R; r += WIN_THREADS) {
                    const int j = Jrow[r];
                    const int zs = (j == f) ? a_f : ((j >= 0) ? D.Zs[j] : 0);
                    urow[r] = (j >= 0) ? D.U[(size_t)zs * D.R + r] : fresh_state(D, f, r, tt, 0u);
                }
                if (threadIdx.x == 0) {
                    s_label[b] = ++s_kt;
                    s_nlive = b + 1;
                }
            }
        }
        for (int sl = threadIdx.x; sl <= nlive && sl < D.kcap; sl += WIN_THREADS) {
            const int n = s_np[sl];
            s_npd[sl] = n > 0 ? (double)n : 0.0;
            s_npd1[sl] = n > 1 ? (double)(n - 1) : 0.0;
        }
        if (threadIdx.x == 0) {
            s_events += nev;
            s_diag[0]++;
            if (nev < ncand) s_diag[nev < K0 ? 1 : (opened ? 2 : 3)]++;
        }
        __syncthreads();
        if (s_err) break;
        // ---- proposals that copied a moved cell's cluster state: correct their distance exactly
        // (the lists of all events of the round are walked as one range, so their loads overlap)
        {
            int tot_e = 0;
            for (int i = 0; i < nev; i++) tot_e += s_ev_e1[i] - s_ev_e0[i];
            for (int g = threadIdx.x; g < tot_e; g += WIN_THREADS) {
                int i = 0, e = g;
                while (e >= s_ev_e1[i] - s_ev_e0[i]) {
                    e -= s_ev_e1[i] - s_ev_e0[i];
                    i++;
                }
                e += s_ev_e0[i];
                const int f = s_ev_f[i];
                const unsigned cr = D.inv[e].cr;
                const double xv = D.inv[e].x;
                const int cc = (int)(cr >> 10), r = (int)(cr & 1023u);
                if (cc > f && !cs_isna(xv)) {
                    const long long d = tint(xv, D.U[(size_t)s_ev_b[i] * D.R + r], s_isig16[r]) -
                                        tint(xv, D.U[(size_t)s_ev_a[i] * D.R + r], s_isig16[r]);
                    if (d) atomicAdd(reinterpret_cast<unsigned long long *>(D.qnew_i + cc), (unsigned long long)d);
                }
            }
        }
        // ---- a new cluster is a new candidate for every later cell: score them now
        if (opened) {
            const int f = s_ev_f[nev - 1], b = s_ev_b[nev - 1];
            double x[NCH];
            const double *urow = D.U + (size_t)b * D.R;
            for (int cc = f + 1 + warp; cc < D.N; cc += WIN_THREADS / 32) {
                load_row<NCH>(D.X + (size_t)cc * D.R, D.R, lane, x);
                const long long q = q_row<NCH>(x, urow, s_isig16, D.R, lane);
                const double qd = (double)q * Q_SCALE, qm = D.qmin[cc];
                if (qd < qm) {
                    // the new cluster is this cell's best candidate: it becomes the reference of the
                    // cached exponentials (slots 0..b re-exponentiated, one lane each)
                    for (int s = lane; s <= b; s += 32) {
                        const double qs = (s == b) ? qd : D.Q[(size_t)s * N + cc];
                        D.E[(size_t)s * N + cc] = exp(-0.5 * (qs - qd));
                    }
                    if (lane == 0) {
                        D.Q[(size_t)b * N + cc] = qd;
                        D.qmin[cc] = qd;
                        D.kmin[cc] = b < 255 ? b : 255;
                        D.wn_q[cc] = LLONG_MIN; // the cached new-table weight used the old reference
                    }
                } else if (lane == 0) {
                    D.Q[(size_t)b * N + cc] = qd;
                    D.E[(size_t)b * N + cc] = exp(-0.5 * (qd - qm)); // cached like the other slots'
                }
            }
        }
        if (threadIdx.x == 0) s_scanned += (nev ? lo - pos : min(W, D.N - pos));
        __syncthreads(); // orders the corrections (L2 atomics) before the next window's .cg loads
        const int adv = nev ? lo - pos : W;
        pos += adv;
        // sparse events: wider windows (fewer rounds); dense events: narrower ones (cheaper rounds)
        if (nev == 0) W = min(WIN_CELLS, W * 2);
        else if (adv < W / 4 && nev < MAX_EV) W = max(WIN_MIN, W / 2);
    }
    __syncthreads();
    for (int s = threadIdx.x; s < D.kcap; s += WIN_THREADS) {
        D.np[s] = s_np[s];
        D.slot_label[s] = s_label[s];
    }
    if (threadIdx.x == 0) {
        D.scal->nlive = s_nlive;
        D.scal->kt = s_kt;
        D.scal->cells_scanned += s_scanned;
        D.scal->n_events += s_events;
        D.scal->n_rounds += s_diag[0];
        D.scal->cut_margin += s_diag[1];
        D.scal->cut_open += s_diag[2];
        D.scal->cut_maxev += s_diag[3];
        if (s_err) D.scal->err = s_err;
    }
}

// ---------------------------------------------------------------- scan, resident window (one cluster of CTAs)
// The same walk as scan_kernel for sweeps with at most 32 live clusters, with what a re-draw
// needs kept in shared memory: a ring of RING cells (cached exponentials, proposal distance,
// reference, uniform, old slot, dependents' range) is filled by asynchronous copies well ahead
// of the window, so a round touches global memory only for the corrections it commits.
//   * copies are issued at the end of a round and waited for in the middle of the next one;
//     a window never reaches past what has landed (RING - WIN_CELLS cells of slack);
//   * a correction of a proposal distance goes to global memory AND, for cells already
//     copied, to the ring.  No copy is in flight while corrections are applied, and copies
//     go past L1 (16 bytes, .cg: cells are copied four at a time), so a later copy sees
//     every correction;
//   * opening a cluster rewrites cached columns of all later cells: the other CTAs of the
//     cluster (parked at the cluster barrier in between) score the new column together with
//     this one, then the ring is filled again (cluster.sync invalidates L1);
//   * when a cluster opens that the ring has no slot for, the walk is handed on at that cell
//     (to the other shape of this kernel, then to scan_kernel).
// two shapes of the resident window: cells in the ring x live clusters it holds x widest window
template <int RING_, int SLOTS_, int WINC_>
struct RingCfg {
    static constexpr int RING = RING_, RMASK = RING_ - 1, SLOTS = SLOTS_, WINC = WINC_;
};
using RingWide = RingCfg<1024, 12, 512>; // few clusters (the later, sparser sweeps): 512-cell windows, half the rounds
using RingDeep = RingCfg<512, 32, 256>;  // up to 32 clusters

__device__ __forceinline__ void cp_async_8(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16_cg(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NLEFT>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(NLEFT) : "memory"); }

constexpr int DEP_STAGE = 2 * WIN_THREADS; // dependents' index entries staged per round
// fixed part of the kernel's shared memory (compile-time offsets); behind it: np, label, np0
// (kcap ints each) and isig16 (R doubles)
template <typename Cfg>
struct RingFixed {
    long long qi[Cfg::RING], wq[Cfg::RING];        // proposal distance (fixed point); the one the cached weight is for
    double w[Cfg::RING], qmin[Cfg::RING], ua[Cfg::RING]; // cached new-table weight, reference, uniform
    double e[Cfg::SLOTS * Cfg::RING];         // cached exponentials [slot][cell]
    int z[Cfg::RING], km[Cfg::RING], i0[Cfg::RING]; // old slot, cluster at the reference, dependents' offset
    unsigned near[Cfg::RING * CS_NEAR_CAP];         // near sources of the proposal (see charge_near_sources)
    InvEntry dep[DEP_STAGE];             // staged index entries of the round's dependents
    double npd[Cfg::SLOTS + 2], npd1[Cfg::SLOTS + 2];
};
template <typename Cfg>
inline size_t ring_smem_bytes(int R, int kcap)
{
    return sizeof(RingFixed<Cfg>) + sizeof(int) * 3 * (size_t)kcap + sizeof(double) * ((size_t)R + 1) + 16;
}

// asynchronous copies of cells [lo, hi) (multiples of 4) into the ring, 16 bytes at a time: a
// warp takes a field, its lanes two cells (8-byte fields) or four (4-byte fields) each.  The
// arrays are padded, so whole groups can be read up to the cell behind the last one (whose
// dependents' offset closes the last cell's range).
template <typename Cfg>
__device__ __forceinline__ void ring_issue(const ChainDev &D, RingFixed<Cfg> &F, int lo, int hi, int nlive)
{
    const int ncell = hi - lo;
    if (ncell <= 0) return;
    const size_t N = (size_t)D.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < nlive + 9; k += WIN_THREADS / 32) {
        if (k == nlive + 8) { // near sources: CS_NEAR_CAP words = two 16-byte groups per cell
            for (int q = lane; q < ncell * 2; q += 32) {
                const int c = lo + (q >> 1), h = (q & 1) * 4;
                cp_async_16_cg(F.near + (size_t)(c & Cfg::RMASK) * CS_NEAR_CAP + h, D.near + (size_t)c * CS_NEAR_CAP + h);
            }
            continue;
        }
        if (k < nlive + 5) {
            const double *src;
            double *dst;
            bool aligned = true;
            if (k < nlive) {
                src = D.E + (size_t)k * N;
                dst = F.e + (size_t)k * Cfg::RING;
                aligned = (((size_t)k * N) & 1) == 0; // (odd N: every other column starts at an odd element)
            } else if (k == nlive) {
                src = reinterpret_cast<const double *>(D.wn_q);
                dst = reinterpret_cast<double *>(F.wq);
            } else if (k == nlive + 1) {
                src = D.wn_w;
                dst = F.w;
            } else if (k == nlive + 2) {
                src = D.qmin;
                dst = F.qmin;
            } else if (k == nlive + 3) {
                src = D.ua;
                dst = F.ua;
            } else {
                src = reinterpret_cast<const double *>(D.qnew_i);
                dst = reinterpret_cast<double *>(F.qi);
            }
            for (int p = lane; p < ncell / 2; p += 32) {
                const int c = lo + 2 * p;
                if (aligned) {
                    cp_async_16_cg(dst + (c & Cfg::RMASK), src + c);
                } else {
                    cp_async_8(dst + (c & Cfg::RMASK), src + c);
                    cp_async_8(dst + ((c + 1) & Cfg::RMASK), src + c + 1);
                }
            }
        } else {
            const int *src;
            int *dst;
            if (k == nlive + 5) {
                src = D.Zs;
                dst = F.z;
            } else if (k == nlive + 6) {
                src = D.kmin;
                dst = F.km;
            } else {
                src = D.icnt;
                dst = F.i0;
            }
            for (int q = lane; q < ncell / 4; q += 32) {
                const int c = lo + 4 * q;
                cp_async_16_cg(dst + (c & Cfg::RMASK), src + c);
            }
        }
    }
}

// draw of one cell from the ring: draw_cell_fast's arithmetic with few registers.  Slots are
// taken eight at a time with unconditional loads (a slot past the live ones weighs exactly +0,
// which leaves every sum as it is); the first pass keeps the running sum after each group, the
// second one re-adds only the group in which the threshold falls — the same additions in the
// same order, so the same sums.
template <typename Cfg>
__device__ __forceinline__ int draw_cell_ring(const double *ecol, int nlive, const double *s_npd, const double *s_npd1,
                                              int zold, double w_new, double ua, double *margin)
{
    constexpr int NG = (Cfg::SLOTS + 7) / 8;
    double cg[NG];
    double run = 0.0;
#pragma unroll
    for (int g = 0; g < NG; g++) {
        if (8 * g < nlive) {
            double e[8];
#pragma unroll
            for (int i = 0; i < 8; i++) e[i] = (8 * g + i < Cfg::SLOTS) ? ecol[(size_t)(8 * g + i) * Cfg::RING] : 0.0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int s = 8 * g + i;
                const double nd = (s == zold) ? s_npd1[s] : s_npd[s];
                run = __dadd_rn(run, (s < nlive) ? __dmul_rn(nd, e[i]) : 0.0);
            }
        }
        cg[g] = run;
    }
    const double tot = __dadd_rn(run, w_new);
    const double t = __dmul_rn(ua, tot);
    // cumulative sums are non-decreasing: those <= t form a prefix.  Whole groups first ...
    int gsel = 0;
    double start = 0.0;
#pragma unroll
    for (int g = 0; g < NG; g++) {
        if (8 * g < nlive && !(cg[g] > t)) {
            gsel = g + 1;
            start = cg[g];
        }
    }
    int z;
    double gap;
    if (8 * gsel < nlive) { // ... then the slots of the group the threshold falls in
        double cum = start, cum_below = start, cum_at = CUDART_INF;
        int k = 8 * gsel;
        const double *eg = ecol + (size_t)(8 * gsel) * Cfg::RING;
        double e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) e[i] = (8 * gsel + i < Cfg::SLOTS) ? eg[(size_t)i * Cfg::RING] : 0.0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int s = 8 * gsel + i;
            const double nd = (s == zold) ? s_npd1[s] : s_npd[s];
            cum = __dadd_rn(cum, (s < nlive) ? __dmul_rn(nd, e[i]) : 0.0);
            const bool over = cum > t;
            cum_below = over ? cum_below : cum;
            cum_at = over ? fmin(cum_at, cum) : cum_at;
            k += (!over && s < nlive) ? 1 : 0;
        }
        z = k; // (the group's last sum exceeds t, so k < nlive)
        gap = fmin(t - cum_below, cum_at - t);
    } else if (w_new > 0.0) {
        z = nlive;
        gap = t - start;
    } else {
        z = -1; // (all weights zero, or the threshold at the very top: the last positive weight)
        for (int s = 0; s < nlive; s++) {
            const double nd = (s == zold) ? s_npd1[s] : s_npd[s];
            if (__dmul_rn(nd, ecol[(size_t)s * Cfg::RING]) > 0.0) z = s;
        }
        gap = 0.0;
    }
    *margin = 0.5 * gap - 1e-12 * tot;
    return z;
}

// a new cluster (slot b, opened at cell f) is a new candidate for every later cell: this CTA's
// share of their scores.  Cells are dealt to the warps of the cluster round-robin.
template <int NCH>
__device__ __noinline__ void score_new_column(const ChainDev &D, int f, int b, int rank, int csize,
                                                 const double *s_isig16, int lane, int warp)
{
    const size_t N = (size_t)D.N;
    const int stride = (WIN_THREADS / 32) * csize;
    double ur[NCH], x[NCH], xn[NCH];
    load_row<NCH>(D.U + (size_t)b * D.R, D.R, lane, ur);
    int cc = f + 1 + rank * (WIN_THREADS / 32) + warp;
    if (cc < D.N) load_row<NCH>(D.X + (size_t)cc * D.R, D.R, lane, x);
    for (; cc < D.N; cc += stride) {
        const int nx = cc + stride;
        if (nx < D.N) load_row<NCH>(D.X + (size_t)nx * D.R, D.R, lane, xn); // next row in flight
        const double qm = D.qmin[cc];
        const long long q = q_regs<NCH>(x, ur, s_isig16, D.R, lane);
        const double qd = (double)q * Q_SCALE;
        if (qd < qm) {
            // the new cluster is this cell's best candidate: it becomes the reference of the
            // cached exponentials (slots 0..b re-exponentiated, one lane each)
            for (int s = lane; s <= b; s += 32) {
                const double qs = (s == b) ? qd : D.Q[(size_t)s * N + cc];
                D.E[(size_t)s * N + cc] = exp(-0.5 * (qs - qd));
            }
            if (lane == 0) {
                D.Q[(size_t)b * N + cc] = qd;
                D.qmin[cc] = qd;
                D.kmin[cc] = b < 255 ? b : 255;
                D.wn_q[cc] = LLONG_MIN; // the cached new-table weight used the old reference
            }
        } else if (lane == 0) {
            D.Q[(size_t)b * N + cc] = qd;
            D.E[(size_t)b * N + cc] = exp(-0.5 * (qd - qm)); // cached like the other slots'
        }
#pragma unroll
        for (int i = 0; i < NCH; i++) x[i] = xn[i];
    }
}

template <int NCH, typename Cfg>
__global__ void __launch_bounds__(WIN_THREADS, 1)
scan_ring_kernel(ChainDev D, int tt)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
    extern __shared__ __align__(16) unsigned char s_ring_raw[];
    RingFixed<Cfg> &F = *reinterpret_cast<RingFixed<Cfg> *>(s_ring_raw);
    int *s_np = reinterpret_cast<int *>(s_ring_raw + sizeof(RingFixed<Cfg>));
    int *s_label = s_np + D.kcap, *s_np0 = s_label + D.kcap;
    double *s_isig16 = reinterpret_cast<double *>(s_ring_raw + sizeof(RingFixed<Cfg>) + sizeof(int) * ((3 * (size_t)D.kcap + 1) & ~(size_t)1));
    __shared__ int s_nlive, s_kt, s_err, s_block, s_wcnt[WIN_THREADS / 32];
    __shared__ int s_ev_f[MAX_EV], s_ev_a[MAX_EV], s_ev_b[MAX_EV], s_ev_e0[MAX_EV], s_ev_e1[MAX_EV];
    __shared__ unsigned char s_cidx[Cfg::WINC]; // candidate number of every window cell (255: not a candidate)
    __shared__ int s_ev_off[MAX_EV + 1]; // dependents in front of each candidate's own (prefix of e1 - e0)
    __shared__ long long s_scanned, s_events;
    __shared__ int s_diag[4]; // rounds, rounds cut short by a margin / an opened cluster / the event list

    // (uniform over the cluster: the scalars are written back only after the last cluster barrier)
    const int f0 = D.scal->first_event;
    if (f0 >= D.N || D.scal->nlive > Cfg::SLOTS) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = threadIdx.x; r < D.R; r += WIN_THREADS) s_isig16[r] = D.isig[r] * 65536.0;
    volatile int *cmd_cell = &D.scal->open_cell, *cmd_slot = &D.scal->open_slot;
    if (rank != 0) {
        // helper CTAs: parked at the cluster barrier until a cluster opens or the walk ends
        __syncthreads();
        for (;;) {
            cluster.sync();
            const int f = *cmd_cell, b = *cmd_slot;
            if (f < 0) break;
            score_new_column<NCH>(D, f, b, rank, csize, s_isig16, lane, warp);
            cluster.sync();
        }
        return;
    }
    for (int s = threadIdx.x; s < D.kcap; s += WIN_THREADS) {
        const int n = D.np[s];
        s_np[s] = n;
        s_np0[s] = D.np0[s];
        s_label[s] = D.slot_label[s];
        if (s <= Cfg::SLOTS) {
            F.npd[s] = n > 0 ? (double)n : 0.0;
            F.npd1[s] = n > 1 ? (double)(n - 1) : 0.0;
        }
    }
    if (threadIdx.x == 0) {
        s_nlive = D.scal->nlive;
        s_kt = D.scal->kt;
        s_err = 0;
        s_scanned = 0;
        s_events = 0;
    }
    if (threadIdx.x < 4) s_diag[threadIdx.x] = 0;
    const size_t N = (size_t)D.N;
    const int Npad = (D.N + 4) & ~3; // cells 0..N, in whole groups of four
    int pos = f0, W = WIN_MIN;
    int head = min(Npad, (pos & ~3) + Cfg::RING), landed;
    ring_issue<Cfg>(D, F, pos & ~3, head, D.scal->nlive);
    cp_async_wait_all();
    __syncthreads();
    landed = head;
    int resume = D.N; // where scan_kernel takes over (N: nowhere)
    while (pos < D.N) {
        // ---- every cell of the window is re-drawn exactly against the committed state
        // (a cell's range of dependents ends at its successor's offset: one cell short of what has landed)
        const int We = min(W, landed - 1 - pos);
        const int c = pos + threadIdx.x;
        const bool active = threadIdx.x < We && c < D.N;
        const int ri = c & Cfg::RMASK;
        const int nlive = s_nlive;
        int z = -1, zold = -1;
        double margin = 0.0; // what this draw tolerates before it has to be redone
        if (active) {
            zold = F.z[ri];
            const long long qi = F.qi[ri];
            const double q_new = (double)qi * Q_SCALE;
            const double qmin = F.qmin[ri];
            const double ua = F.ua[ri];
            const int km = F.km[ri];
            // (the conditions under which the cached exponentials are today's: see scan_kernel)
            bool fast = (km != 255 ? (q_new >= qmin && s_np[km] - (km == zold ? 1 : 0) > 0) : q_new == qmin);
            if (fast && s_np0[zold] < 2 && s_np[zold] >= 2 && D.Q[(size_t)zold * N + c] < qmin) fast = false;
            if (fast) {
                double w_new = F.w[ri];
                if (F.wq[ri] != qi) { // the proposal was corrected since the weight was cached
                    w_new = __dmul_rn(D.beta, exp(-0.5 * (q_new - qmin)));
                    F.wq[ri] = qi;
                    F.w[ri] = w_new;
                }
                const double *ecol = F.e + ri;
                z = draw_cell_ring<Cfg>(ecol, nlive, F.npd, F.npd1, zold, w_new, ua, &margin);
            } else {
                z = draw_cell_thread(D.Q + c, D.E + c, N, nlive, nlive, s_np, zold, q_new, qmin, D.beta, ua);
            }
        }
        // ---- the window's event candidates, in cell order (ballot + prefix over the warps)
        const bool cand = active && z != zold;
        const unsigned bal = __ballot_sync(CS_FULL, cand);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int jc = __popc(bal & ((1u << lane) - 1u)), ncand = 0; // candidates in front of this cell
#pragma unroll
        for (int w = 0; w < WIN_THREADS / 32; w++) {
            const int cw = s_wcnt[w];
            if (w < warp) jc += cw;
            ncand += cw;
        }
        if (cand && jc < MAX_EV) {
            s_ev_f[jc] = c;
            s_ev_a[jc] = zold;
            s_ev_b[jc] = z;
            s_ev_e0[jc] = F.i0[ri];
            s_ev_e1[jc] = F.i0[(c + 1) & Cfg::RMASK];
        }
        if (threadIdx.x < Cfg::WINC) s_cidx[threadIdx.x] = (cand && jc < MAX_EV) ? (unsigned char)jc : (unsigned char)255;
        if (threadIdx.x == 0) s_block = INT_MAX;
        __syncthreads();
        // ---- how many can be committed in this round (scan_kernel's rules; every warp works it
        // out for itself, one candidate per lane): candidate j is out when it lies behind one that
        // opens a cluster
        int K = min(ncand, MAX_EV);
        {
            const bool mine = lane < K;
            const bool opens = mine && s_ev_b[lane] == nlive;
            const unsigned opn = __ballot_sync(CS_FULL, opens);
            if (opn) K = min(K, __ffs(opn));
            if (warp == 0) { // where each candidate's dependents start in the round's flattened list
                int len = lane < K ? s_ev_e1[lane] - s_ev_e0[lane] : 0, inc = len;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int o = __shfl_up_sync(CS_FULL, inc, off);
                    if (lane >= off) inc += o;
                }
                s_ev_off[lane] = inc - len;
                if (lane == 31) s_ev_off[32] = inc;
            }
        }
        // ... or when a cell between the previous candidate and itself (itself included) runs out
        // of margin: each of the jc events in front of a cell moves its sums and threshold by at
        // most E[a] + E[b]
        // most E[a] + E[b]; so does its new-table weight when its proposal copied the cluster state of
        // a cell that moves in this round (charge_near_sources)
        if (active && jc >= 1 && jc < K) {
            double m = margin;
            const double *ecol = F.e + ri;
            for (int i = 0; i < jc; i++)
                m -= (ecol[(size_t)s_ev_a[i] * Cfg::RING] + ecol[(size_t)s_ev_b[i] * Cfg::RING]) * 1.000000001;
            if (m > 0.0)
                m = charge_near_sources(D, F.near + (size_t)ri * CS_NEAR_CAP, c, pos, K, s_cidx, s_ev_a, s_ev_b, s_isig16,
                                        F.qi[ri], F.qmin[ri], F.w[ri], m);
            if (!(m > 0.0)) atomicMin(&s_block, jc);
        }
        __syncthreads();
        const int K0 = K;
        K = min(K, s_block);
        const int nev = K;
        const bool opened = nev > 0 && s_ev_b[nev - 1] == nlive;
        const int lo = nev ? s_ev_f[nev - 1] + 1 : pos;
        // ---- the index entries of the committed events' dependents are staged now (asynchronous
        // copies, two per thread); they are read after the commit, when their latency has passed
        const int tot_e = s_ev_off[nev];
        int dep_i[2] = {-1, -1};
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int g = threadIdx.x + u * WIN_THREADS;
            if (g < tot_e) {
                int i = 0;
                for (int j = 1; j < nev; j++) i += (s_ev_off[j] <= g) ? 1 : 0;
                const int e = s_ev_e0[i] + (g - s_ev_off[i]);
                cp_async_16_cg(&F.dep[g], D.inv + e);
                dep_i[u] = i;
            }
        }
        // ---- commit them: counts (order-free integer updates), assignments
        if (threadIdx.x < nev) {
            atomicAdd(&s_np[s_ev_a[threadIdx.x]], -1);
            if (s_ev_b[threadIdx.x] == nlive) s_np[nlive] = 1; // (a slot past the live ones may hold a stale count)
            else atomicAdd(&s_np[s_ev_b[threadIdx.x]], 1);
            D.Zs[s_ev_f[threadIdx.x]] = s_ev_b[threadIdx.x];
        }
        cp_async_wait_all(); // the copies issued at the end of the previous round
        __syncthreads();
        landed = head;
        if (opened) { // :209-212 the last event opens a cluster with its proposed per-segment states
            const int f = s_ev_f[nev - 1], b = nlive;
            if (b >= D.kcap) {
                if (threadIdx.x == 0) s_err = CS_ERR_KCAP;
            } else {
                double *urow = D.U + (size_t)b * D.R;
                const int *Jrow = D.J + (size_t)f * D.R;
                const int a_f = s_ev_a[nev - 1]; // (its own assignment was already moved)
                for (int r = threadIdx.x; r < D.R; r += WIN_THREADS) {
                    const int j = Jrow[r];
                    const int zs = (j == f) ? a_f : ((j >= 0) ? D.Zs[j] : 0);
                    urow[r] = (j >= 0) ? D.U[(size_t)zs * D.R + r] : fresh_state(D, f, r, tt, 0u);
                }
                if (threadIdx.x == 0) {
                    s_label[b] = ++s_kt;
                    s_nlive = b + 1;
                }
            }
        }
        if (threadIdx.x <= nlive && threadIdx.x <= Cfg::SLOTS && threadIdx.x < D.kcap) {
            const int n = s_np[threadIdx.x];
            F.npd[threadIdx.x] = n > 0 ? (double)n : 0.0;
            F.npd1[threadIdx.x] = n > 1 ? (double)(n - 1) : 0.0;
        }
        if (threadIdx.x == 0) {
            s_events += nev;
            s_scanned += (nev ? lo - pos : min(We, D.N - pos));
            s_diag[0]++;
            if (nev < ncand) s_diag[nev < K0 ? 1 : (opened ? 2 : 3)]++;
        }
        __syncthreads();
        if (s_err) break;
        // ---- proposals that copied a moved cell's cluster state: correct their distance exactly,
        // in global memory and, for cells already copied, in the ring
        {
            auto correct = [&](int i, int cc, int r, double xv) {
                if (cc > s_ev_f[i] && !cs_isna(xv)) {
                    const long long d = tint(xv, D.U[(size_t)s_ev_b[i] * D.R + r], s_isig16[r]) -
                                        tint(xv, D.U[(size_t)s_ev_a[i] * D.R + r], s_isig16[r]);
                    if (d) {
                        atomicAdd(reinterpret_cast<unsigned long long *>(D.qnew_i + cc), (unsigned long long)d);
                        if (cc < head)
                            atomicAdd(reinterpret_cast<unsigned long long *>(F.qi + (cc & Cfg::RMASK)), (unsigned long long)d);
                    }
                }
            };
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (dep_i[u] >= 0) {
                    const unsigned cr = F.dep[threadIdx.x + u * WIN_THREADS].cr;
                    correct(dep_i[u], (int)(cr >> 10), (int)(cr & 1023u), F.dep[threadIdx.x + u * WIN_THREADS].x);
                }
            for (int g = threadIdx.x + 2 * WIN_THREADS; g < tot_e; g += WIN_THREADS) { // (rarely: long lists)
                int i = 0;
                for (int j = 1; j < nev; j++) i += (s_ev_off[j] <= g) ? 1 : 0;
                const int e = s_ev_e0[i] + (g - s_ev_off[i]);
                const unsigned cr = D.inv[e].cr;
                correct(i, (int)(cr >> 10), (int)(cr & 1023u), D.inv[e].x);
            }
        }
        __syncthreads(); // corrections done (global ones ordered before the copies issued below)
        const int adv = nev ? lo - pos : We;
        pos += adv;
        if (nev == 0 || adv > W / 2) W = min(Cfg::WINC, W * 2); // (rounds get far: wider windows, fewer rounds)
        else if (adv < W / 4 && nev < MAX_EV) W = max(WIN_MIN, W / 2);
        if (opened) {
            // ---- the new column, scored by the whole cluster; then the ring is filled again
            const int f = s_ev_f[nev - 1], b = s_ev_b[nev - 1];
            if (threadIdx.x == 0) {
                *cmd_cell = f;
                *cmd_slot = b;
            }
            cluster.sync();
            score_new_column<NCH>(D, f, b, 0, csize, s_isig16, lane, warp);
            cluster.sync();
            if (b + 1 > Cfg::SLOTS) { // one cluster more than the ring holds: the next walk takes over here
                resume = pos;
                break;
            }
            ring_issue<Cfg>(D, F, pos & ~3, head, b + 1);
            cp_async_wait_all();
            __syncthreads();
        }
        // ---- copies for the cells the window will reach after the next round
        {
            const int nh = min(Npad, (pos & ~3) + Cfg::RING);
            ring_issue<Cfg>(D, F, head, nh, s_nlive);
            head = max(head, nh);
        }
    }
    cp_async_wait_all();
    if (threadIdx.x == 0) *cmd_cell = -1;
    cluster.sync();
    for (int s = threadIdx.x; s < D.kcap; s += WIN_THREADS) {
        D.np[s] = s_np[s];
        D.slot_label[s] = s_label[s];
    }
    if (threadIdx.x == 0) {
        D.scal->nlive = s_nlive;
        D.scal->kt = s_kt;
        D.scal->cells_scanned += s_scanned;
        D.scal->n_events += s_events;
        D.scal->n_rounds += s_diag[0];
        D.scal->cut_margin += s_diag[1];
        D.scal->cut_open += s_diag[2];
        D.scal->cut_maxev += s_diag[3];
        D.scal->first_event = s_err ? D.N : resume; // what is left for scan_kernel
        if (s_err) D.scal->err = s_err;
    }
}

// ---------------------------------------------------------------- sequential scan (one CTA, warp per cell)
// every cell from `start` on is re-proposed, re-scored and re-drawn against the committed
// state, 32 cells at a time (a window is cut at its first event).
template <int NCH>
__global__ void __launch_bounds__(SCAN_THREADS)
seq_scan_kernel(ChainDev D, int tt, int force_all, int *__restrict__ rej)
{
    extern __shared__ double s_dyn[];
    double *s_isig16 = s_dyn;
    int *s_np = reinterpret_cast<int *>(s_dyn + D.R);
    int *s_label = s_np + D.kcap;
    __shared__ int s_nlive, s_kt, s_fmin[2], s_err;
    __shared__ long long s_scanned, s_events;

    const int f0 = force_all ? 0 : D.scal->first_event;
    if (f0 >= D.N) return;
    for (int r = threadIdx.x; r < D.R; r += SCAN_THREADS) s_isig16[r] = D.isig[r] * 65536.0;
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
    const bool use_J = !reject_live && !force_all;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t N = (size_t)D.N;
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
            double *Qc = D.Q + c;
            uint32_t attempt = 0;
            for (;;) {
                propose<NCH>(D, c, tt, attempt, use_J ? D.J + (size_t)c * D.R : nullptr, lane, mu, jj);
                if (!reject_live || !is_duplicate<NCH>(D, nlive, mu, lane)) break;
                attempt++;
            }
            if (reject_live && lane == 0) rej[c] = (int)attempt;
            // Q entries of the clusters opened since this cell was last scored
            for (int s = D.qcols[c]; s < nlive; s++) {
                const long long q = q_row<NCH>(x, D.U + (size_t)s * D.R, s_isig16, D.R, lane);
                if (lane == 0) Qc[(size_t)s * N] = (double)q * Q_SCALE;
            }
            if (lane == 0) D.qcols[c] = nlive;
            const long long qn = q_regs<NCH>(x, mu, s_isig16, D.R, lane);
            double ua, ub;
            cs_philox_draw(D.seed, (uint32_t)c, 0u, (uint32_t)tt, CS_PH_ASSIGN, 0u, ua, ub);
            __syncwarp();
            z = draw_cell_warp(Qc, nullptr, N, nlive, s_np, zold, (double)qn * Q_SCALE, D.beta, ua, lane, nullptr);
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

// ---------------------------------------------------------------- fixed-point walk (all SMs, cooperative)
// The sequential sweep as the fixed point of a parallel map.  Let h be a hypothesis for the
// outcome of every cell of the sweep (the slot it ends in).  F(h)[c] = the draw of cell c
// against the state that h[0..c) induces: counts = counts at the start of the sweep + the moves
// h makes in front of c, proposal = the cluster states of its source cells (h for sources in
// front of c, the sweep's start for the others), same uniform.  The sequential chain is the
// only h with F(h) = h (induction over c: cell 0 sees the start state; if h is right in front
// of c, F(h)[c] is the sequential draw), and from ANY h the iteration h <- F(h) is right up to
// and including the first cell it changes — so it ends after at most N steps, in practice after
// a handful: a draw flips only when the moves in front of it shift its cumulative weights past
// its uniform, and each step's flips are a small fraction of the previous step's.  Every step
// is one pass of ALL SMs over the cells behind the last settled one; there is no sequential
// walk, no margin bookkeeping and no inverse index.  Per step and 256-cell tile:
//   pull   (warp per cell) the proposal distance under h: the cell's sources (J) are looked up
//          in a bitmap of the cells h moves; for a hit the segment's term is exchanged exactly
//          (fixed point): tint(x, U[h[j]]) - tint(x, U[Zs[j]]);
//   draw   (thread per cell) counts in front of the cell = prefix over the tiles (a small scan
//          between two grid barriers) + prefix over the tile's warps + ballots inside the warp;
//          weights from the cached exponentials while their reference still is the minimum
//          over today's candidates, re-exponentiated otherwise — the arithmetic of
//          draw_cell_warp in the same order;
//   publish the tile's net moves per slot, its bits of the bitmap, the first changed cell, the
//          first cell that opens a cluster.
// A cell that opens a cluster stops the prefix: when everything up to it has settled, its
// proposal becomes the new row (:209-212), the new column is scored for all later cells by the
// whole grid, and the iteration goes on behind it with one slot more.
constexpr int FX_THREADS = 256, FX_WARPS = FX_THREADS / 32;
constexpr int FX_STEP = 128;      // source entries per warp and step of the pull (16 bytes per lane)
constexpr int FX_NS = 6;          // steps in the ring of asynchronous copies
constexpr int FX_HCAP = 256;      // collected hits per warp
constexpr int FX_INF = 0x7f7f7f7f; // (what cudaMemsetAsync(0x7f) leaves)

struct FixScal {
    int fc[2], fo[2]; // per step parity: first changed cell, first cell that opens a cluster
};

__host__ __device__ inline int fx_tiles(int N) { return (N + FX_THREADS - 1) / FX_THREADS; }
__host__ __device__ inline int fx_tpad(int N) { return (fx_tiles(N) + 31) & ~31; }
__host__ __device__ inline int fx_words(int N) { return ((N + FX_THREADS - 1) / FX_THREADS) * FX_WARPS; }
inline size_t fix_smem_bytes(int N, int R, int kcap, bool bits_in_smem)
{
    return sizeof(double) * (size_t)R + sizeof(long long) * FX_THREADS + sizeof(int) * (size_t)(kcap + 1) * (FX_WARPS + 1) +
           sizeof(int) * FX_WARPS * FX_NS * FX_STEP + sizeof(int2) * FX_WARPS * FX_HCAP +
           (bits_in_smem ? sizeof(unsigned) * (size_t)fx_words(N) : 0) + 48;
}

// the tile's net moves per slot (tdelta[s][tile]), its 256 bits of the moved-cell bitmap, the
// first cell whose hypothesis opens a cluster
__device__ __forceinline__ void fx_publish(const ChainDev &D, int tile, int c, bool active, bool counts_open, int z, int zo,
                                           int nmat, int *s_td, unsigned *bits_a, unsigned *bits_b, int *fo)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Tp = fx_tpad(D.N);
    for (int s = threadIdx.x; s <= nmat; s += FX_THREADS) s_td[s] = 0;
    __syncthreads();
    const bool moved = active && z != zo;
    if (moved) {
        atomicAdd(&s_td[z], 1);
        atomicAdd(&s_td[zo], -1);
    }
    const unsigned bal = __ballot_sync(CS_FULL, moved);
    if (lane == 0) {
        bits_a[tile * FX_WARPS + warp] = bal;
        if (bits_b) bits_b[tile * FX_WARPS + warp] = bal;
    }
    const unsigned opn = __ballot_sync(CS_FULL, counts_open && z == nmat);
    if (opn && lane == 0) atomicMin(fo, tile * FX_THREADS + warp * 32 + __ffs(opn) - 1);
    __syncthreads();
    for (int s = threadIdx.x; s <= nmat; s += FX_THREADS) D.fx_td[(size_t)s * Tp + tile] = s_td[s];
    (void)c;
}

// exclusive prefix of tdelta over the tiles, per slot (a warp per slot, warps of the whole grid)
__device__ __forceinline__ void fx_scan_tiles(const ChainDev &D, int nmat)
{
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * FX_WARPS + (threadIdx.x >> 5), nw = gridDim.x * FX_WARPS;
    const int T = fx_tiles(D.N), Tp = fx_tpad(D.N);
    for (int s = gw; s <= nmat; s += nw) {
        int run = 0;
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int t = t0 + lane;
            const int v = t < T ? D.fx_td[(size_t)s * Tp + t] : 0;
            int inc = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int o = __shfl_up_sync(CS_FULL, inc, off);
                if (lane >= off) inc += o;
            }
            if (t < T) D.fx_tp[(size_t)s * Tp + t] = run + inc - v;
            run += __shfl_sync(CS_FULL, inc, 31);
        }
    }
}

// the column of a cluster opened at cell f (slot b) for every later cell: score_new_column over
// the warps of the whole grid
template <int NCH>
__device__ __forceinline__ void fx_score_column(const ChainDev &D, int f, int b, const double *s_isig16)
{
    const size_t N = (size_t)D.N;
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * FX_WARPS + (threadIdx.x >> 5), stride = gridDim.x * FX_WARPS;
    double ur[NCH], x[NCH];
    load_row<NCH>(D.U + (size_t)b * D.R, D.R, lane, ur);
    for (int cc = f + 1 + gw; cc < D.N; cc += stride) {
        load_row<NCH>(D.X + (size_t)cc * D.R, D.R, lane, x);
        const double qm = D.qmin[cc];
        const long long q = q_regs<NCH>(x, ur, s_isig16, D.R, lane);
        const double qd = (double)q * Q_SCALE;
        if (qd < qm) {
            // the new cluster is this cell's best candidate: it becomes the reference of the
            // cached exponentials (slots 0..b re-exponentiated, one lane each)
            for (int s = lane; s <= b; s += 32) {
                const double qs = (s == b) ? qd : D.Q[(size_t)s * N + cc];
                D.E[(size_t)s * N + cc] = exp(-0.5 * (qs - qd));
            }
            if (lane == 0) {
                D.Q[(size_t)b * N + cc] = qd;
                D.qmin[cc] = qd;
                D.kmin[cc] = b < 255 ? b : 255;
            }
        } else if (lane == 0) {
            D.Q[(size_t)b * N + cc] = qd;
            D.E[(size_t)b * N + cc] = exp(-0.5 * (qd - qm)); // cached like the other slots'
        }
    }
}

template <int NCH>
__global__ void __launch_bounds__(FX_THREADS, 3)
fix_walk_kernel(ChainDev D, int tt, int bits_in_smem)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char s_fx_raw[];
    double *s_isig16 = reinterpret_cast<double *>(s_fx_raw);            // R
    long long *s_qh = reinterpret_cast<long long *>(s_isig16 + D.R);    // what h changes in the proposal distances of the tile's cells
    int *s_base = reinterpret_cast<int *>(s_qh + FX_THREADS);           // kcap + 1: counts in front of the tile
    int *s_wd = s_base + (D.kcap + 1);                                  // FX_WARPS x (kcap + 1): moves in front of each warp
    const int KS = D.kcap + 1;
    int *s_ring = reinterpret_cast<int *>((reinterpret_cast<uintptr_t>(s_wd + FX_WARPS * KS) + 15) & ~(uintptr_t)15); // FX_WARPS rings of FX_NS x FX_STEP source entries (16-byte copies)
    int2 *s_hits = reinterpret_cast<int2 *>(s_ring + FX_WARPS * FX_NS * FX_STEP);
    unsigned *s_bits = reinterpret_cast<unsigned *>(s_hits + FX_WARPS * FX_HCAP);

    // (uniform over the grid: the scalars are rewritten only behind the last grid barrier)
    const int f0 = D.scal->first_event;
    if (f0 >= D.N) return;
    int nmat = D.scal->nlive, kt = D.scal->kt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned ltmask = (1u << lane) - 1u;
    const size_t N = (size_t)D.N;
    const int T = fx_tiles(D.N), Tp = fx_tpad(D.N), NW = fx_words(D.N);
    for (int r = threadIdx.x; r < D.R; r += FX_THREADS) s_isig16[r] = D.isig[r] * 65536.0;
    int *zh[2] = {D.zspec, D.zh1};
    unsigned *bits[2] = {D.fx_bits, D.fx_bits + NW};
    FixScal *fxs = reinterpret_cast<FixScal *>(D.fx_scal);

    // ---- step -1: the hypothesis is prepare_kernel's draws against the start of the sweep
    for (int tile = blockIdx.x; tile < T; tile += gridDim.x) {
        const int c = tile * FX_THREADS + threadIdx.x;
        const bool active = c < D.N;
        const int z = active ? D.zspec[c] : -1, zo = active ? D.Zs[c] : -1;
        if (active) D.zh1[c] = z;
        // (rows of slots that open later in the sweep: nothing of an earlier sweep may be left in them)
        for (int s2 = nmat + 1 + threadIdx.x; s2 <= D.kcap; s2 += FX_THREADS) D.fx_td[(size_t)s2 * Tp + tile] = 0;
        fx_publish(D, tile, c, active, active, z, zo, nmat, s_wd, bits[0], bits[1], &fxs->fo[1]);
        __syncthreads();
    }
    grid.sync();
    int open_cur = *reinterpret_cast<volatile int *>(&fxs->fo[1]);
    fx_scan_tiles(D, nmat);
    grid.sync();

    int lo = f0, err = 0, steps = 0, opened = 0;
    long long evaluated = 0;
    for (int it = 0;; it++) {
        const int cur = it & 1, nxt = cur ^ 1;
        if (blockIdx.x == 0 && threadIdx.x == 0) { // the other parity's minima: read by everyone before the last barrier
            fxs->fc[nxt] = FX_INF;
            fxs->fo[nxt] = FX_INF;
        }
        const unsigned *bcur = bits[cur];
        if (bits_in_smem) {
            for (int w = threadIdx.x; w < NW; w += FX_THREADS) s_bits[w] = bcur[w];
            bcur = s_bits;
        }
        const int *zcur = zh[cur];
        steps++;
        for (int tile = lo / FX_THREADS + blockIdx.x; tile < T; tile += gridDim.x) {
            __syncthreads();
            const int c = tile * FX_THREADS + threadIdx.x;
            const bool active = c < D.N, evalc = active && c >= lo;
            const int zo = active ? D.Zs[c] : -1, zc = active ? zcur[c] : -1;
            // ---- counts in front of the tile, moves in front of each warp
            for (int e = threadIdx.x; e < FX_WARPS * KS; e += FX_THREADS) s_wd[e] = 0;
            s_qh[threadIdx.x] = 0;
            for (int s = threadIdx.x; s <= nmat; s += FX_THREADS) s_base[s] = (s < D.kcap ? D.np0[s] : 0) + D.fx_tp[(size_t)s * Tp + tile];
            __syncthreads();
            const bool moved = active && zc != zo;
            if (moved) {
                atomicAdd(&s_wd[warp * KS + zc], 1);
                atomicAdd(&s_wd[warp * KS + zo], -1);
            }
            const bool warp_moves = __any_sync(CS_FULL, moved);
            __syncthreads();
            for (int s = threadIdx.x; s <= nmat; s += FX_THREADS) {
                int run = 0;
#pragma unroll
                for (int w = 0; w < FX_WARPS; w++) {
                    const int t = s_wd[w * KS + s];
                    s_wd[w * KS + s] = run;
                    run += t;
                }
            }
            __syncthreads();
            // ---- pull: what h changes in the proposal distances of this warp's 32 cells.  Their
            // source lists are one contiguous stretch of J (and of X).  The warp streams over it
            // through a ring of asynchronous copies (FX_NS steps of 128 entries; nothing waits on
            // DRAM latency), looks every source up in the bitmap of moved cells and collects the
            // hits; the collected hits are then worked off a lane each — both slots and the value
            // loaded together, the term exchanged exactly, the difference added to the cell's sum.
            {
                const int cbase = tile * FX_THREADS + warp * 32;
                const int k0 = max(0, lo - cbase), k1 = min(32, D.N - cbase);
                if (k0 < k1) {
                    const size_t ebase = (size_t)cbase * D.R;
                    const int *Jw = D.J + ebase;
                    const double *Xw = D.X + ebase;
                    const int le0 = k0 * D.R, le1 = k1 * D.R; // entries relative to the warp's first cell
                    const unsigned long long rdiv = (0x100000000ull + (unsigned)D.R - 1) / (unsigned)D.R; // le / R for le < 2^16
                    int *ring = s_ring + warp * (FX_NS * FX_STEP);
                    int2 *hits = s_hits + warp * FX_HCAP;
                    int nh = 0;
                    auto flush = [&]() {
                        __syncwarp();
#pragma unroll 2
                        for (int idx = lane; idx < nh; idx += 32) {
                            const int2 h = hits[idx];
                            const int le = h.x, j = h.y;
                            const int zj = zcur[j], zs = D.Zs[j];
                            const double xv = Xw[le];
                            const int kc = (int)(((unsigned long long)le * rdiv) >> 32), r = le - kc * D.R;
                            // (a source that opens a cluster lies behind the settled prefix: no row yet)
                            const bool ok = zj < nmat && !cs_isna(xv);
                            const double u1 = ok ? D.U[(size_t)zj * D.R + r] : 0.0, u0 = ok ? D.U[(size_t)zs * D.R + r] : 0.0;
                            if (ok) {
                                const double is = s_isig16[r];
                                const long long d = tint(xv, u1, is) - tint(xv, u0, is);
                                if (d) atomicAdd(reinterpret_cast<unsigned long long *>(s_qh + warp * 32 + kc), (unsigned long long)d);
                            }
                        }
                        nh = 0;
                        __syncwarp();
                    };
                    const int sbeg = le0 & ~(FX_STEP - 1);
                    const int nsteps = (le1 - sbeg + FX_STEP - 1) / FX_STEP;
#pragma unroll
                    for (int p = 0; p < FX_NS - 1; p++) { // (the list is padded: a step may read past the stretch)
                        if (p < nsteps) cp_async_16_cg(ring + p * FX_STEP + 4 * lane, Jw + sbeg + p * FX_STEP + 4 * lane);
                        cp_async_commit();
                    }
                    for (int st = 0; st < nsteps; st++) {
                        cp_async_wait_group<FX_NS - 2>();
                        __syncwarp();
                        const int *stage = ring + (st % FX_NS) * FX_STEP;
                        int jj[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) jj[u] = stage[u * 32 + lane];
                        __syncwarp();
                        {
                            const int p = st + FX_NS - 1;
                            if (p < nsteps) cp_async_16_cg(ring + (p % FX_NS) * FX_STEP + 4 * lane, Jw + sbeg + p * FX_STEP + 4 * lane);
                            cp_async_commit();
                        }
                        const int s0 = sbeg + st * FX_STEP;
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int le = s0 + u * 32 + lane;
                            const int j = jj[u];
                            const int kc = (int)(((unsigned long long)le * rdiv) >> 32);
                            const bool h = le >= le0 && le < le1 && j >= 0 && j < cbase + kc && ((bcur[j >> 5] >> (j & 31)) & 1u);
                            const unsigned hb = __ballot_sync(CS_FULL, h);
                            if (h) hits[nh + __popc(hb & ltmask)] = make_int2(le, j);
                            nh += __popc(hb);
                        }
                        if (nh > FX_HCAP - FX_STEP) flush();
                    }
                    flush();
                    cp_async_wait_all();
                }
                __syncwarp();
            }
            // ---- draw (the slot loops are uniform over the warp: the counts come from ballots)
            int z = zc;
            {
                const long long qi = evalc ? D.qnew_i[c] + s_qh[threadIdx.x] : 0;
                const double q_new = (double)qi * Q_SCALE;
                const double qmin_c = evalc ? D.qmin[c] : 0.0, ua = evalc ? D.ua[c] : 0.0;
                const int km = evalc ? D.kmin[c] : 255;
                const int *wd = s_wd + warp * KS;
                auto count_at = [&](int s) -> int { // cells in slot s in front of this cell (itself included while it has not moved)
                    int n = s_base[s] + wd[s];
                    if (warp_moves) {
                        const unsigned bn = __ballot_sync(CS_FULL, zc == s), bo = __ballot_sync(CS_FULL, zo == s);
                        n += __popc(bn & ltmask) - __popc(bo & ltmask);
                    }
                    return n;
                };
                int nkm = 1, nzo = 0;
                for (int s = 0; s < nmat; s++) {
                    const int n = count_at(s);
                    if (s == km) nkm = n;
                    if (s == zo) nzo = n;
                }
                // (the conditions under which the cached exponentials are today's: their reference
                //  still is the minimum over the candidates)
                bool fast = (km != 255 ? (q_new >= qmin_c && nkm - (km == zo ? 1 : 0) > 0) : q_new == qmin_c);
                if (evalc && fast && D.np0[zo] < 2 && nzo >= 2 && D.Q[(size_t)zo * N + c] < qmin_c) fast = false;
                double qref = qmin_c;
                if (__any_sync(CS_FULL, evalc && !fast)) {
                    double qm = q_new;
                    for (int s = 0; s < nmat; s++) {
                        const int n = count_at(s) - (s == zo ? 1 : 0);
                        if (evalc && !fast && n > 0) qm = fmin(qm, D.Q[(size_t)s * N + c]);
                    }
                    if (!fast) {
                        fast = qm == qmin_c;
                        qref = qm;
                    }
                }
                const double w_new = __dmul_rn(D.beta, exp(-0.5 * (q_new - qref)));
                auto weight = [&](int s) -> double {
                    const int n = count_at(s) - (s == zo ? 1 : 0);
                    double w = 0.0;
                    if (evalc && n > 0) {
                        const double e = fast ? D.E[(size_t)s * N + c] : exp(-0.5 * (D.Q[(size_t)s * N + c] - qref));
                        w = __dmul_rn((double)n, e);
                    }
                    return w;
                };
                double run = 0.0;
                for (int s = 0; s < nmat; s++) run = __dadd_rn(run, weight(s));
                const double t = __dmul_rn(ua, __dadd_rn(run, w_new));
                run = 0.0;
                int zz = -1, lastpos = -1;
                for (int s = 0; s < nmat; s++) {
                    const double w = weight(s);
                    if (w > 0.0) {
                        run = __dadd_rn(run, w);
                        lastpos = s;
                        if (zz < 0 && run > t) zz = s;
                    }
                }
                if (zz < 0) zz = (w_new > 0.0) ? nmat : lastpos;
                if (evalc) z = zz;
            }
            if (evalc) {
                zh[nxt][c] = z;
                if (z != zc) atomicMin(&fxs->fc[cur], c);
            }
            __syncthreads();
            fx_publish(D, tile, c, active, evalc, z, zo, nmat, s_wd, bits[nxt], nullptr, &fxs->fo[cur]);
            if (threadIdx.x == 0) evaluated += min(D.N, (tile + 1) * FX_THREADS) - max(lo, tile * FX_THREADS);
        }
        grid.sync();
        const int changed_min = *reinterpret_cast<volatile int *>(&fxs->fc[cur]);
        const int open_next = *reinterpret_cast<volatile int *>(&fxs->fo[cur]);
        if (changed_min >= FX_INF && open_cur >= FX_INF) break; // nothing changed, nothing opens: h is the chain
        if (open_cur < changed_min) {
            // ---- everything up to the cell that opens a cluster has settled: its proposal becomes
            // the row of slot nmat (:209-212), scored for every later cell by the whole grid
            const int f = open_cur, b = nmat;
            if (b >= D.kcap) {
                err = CS_ERR_KCAP;
                break;
            }
            if (blockIdx.x == 0) {
                const int *Jrow = D.J + (size_t)f * D.R;
                for (int r = threadIdx.x; r < D.R; r += FX_THREADS) {
                    const int j = Jrow[r];
                    double v;
                    if (j < 0) v = fresh_state(D, f, r, tt, 0u);
                    else v = D.U[(size_t)(j < f ? zh[nxt][j] : D.Zs[j]) * D.R + r];
                    D.U[(size_t)b * D.R + r] = v;
                }
                if (threadIdx.x == 0) D.slot_label[b] = kt + 1;
            }
            kt++;
            nmat++;
            opened++;
            grid.sync();
            fx_score_column<NCH>(D, f, b, s_isig16);
            lo = f + 1;
            open_cur = FX_INF;
        } else {
            // (that cell's new draw saw a settled prefix, so it is settled itself; it is evaluated
            //  once more so that both copies of h receive it)
            lo = changed_min;
            open_cur = open_next;
        }
        fx_scan_tiles(D, nmat);
        grid.sync();
    }
    // ---- h is the sweep's outcome (both copies agree): assignments, counts, scalars
    {
        const int *zf = zh[0];
        long long nev = 0;
        for (int c = blockIdx.x * FX_THREADS + threadIdx.x; c < D.N; c += gridDim.x * FX_THREADS) {
            const int z = zf[c];
            if (z != D.Zs[c]) {
                D.Zs[c] = z;
                nev++;
            }
        }
        nev = warp_sum_ll(nev);
        if (lane == 0 && nev) atomicAdd(reinterpret_cast<unsigned long long *>(&D.scal->n_events), (unsigned long long)nev);
        if (threadIdx.x == 0 && evaluated) atomicAdd(reinterpret_cast<unsigned long long *>(&D.scal->cells_scanned), (unsigned long long)evaluated);
        if (blockIdx.x == 0) {
            for (int s = threadIdx.x; s < nmat; s += FX_THREADS) {
                int tot = 0;
                for (int t = 0; t < T; t++) tot += D.fx_td[(size_t)s * Tp + t];
                D.np[s] = D.np0[s] + tot;
            }
            if (threadIdx.x == 0) {
                D.scal->nlive = nmat;
                D.scal->kt = kt;
                D.scal->n_rounds += steps;
                D.scal->cut_open += opened;
                if (err) D.scal->err = err;
            }
        }
    }
}

// ---------------------------------------------------------------- end of sweep
// per-chunk per-cluster sums of X and non-NA counts (R/BayesNonparCluster.R:223)
__global__ void __launch_bounds__(1024)
stats_partial_kernel(ChainDev D, int smem_doubles)
{
    extern __shared__ double s_dyn[];
    const int nlive = D.scal->nlive;
    const int R = D.R;
    const int ST = max(1, smem_doubles / (R + (R + 1) / 2)); // slots per pass (double acc + int cnt)
    double *acc = s_dyn;
    int *cnt = reinterpret_cast<int *>(s_dyn + (size_t)ST * R);
    const int i0 = blockIdx.x * D.chunk, i1 = min(D.N, i0 + D.chunk);
    for (int t0 = 0; t0 < nlive; t0 += ST) {
        const int ns = min(ST, nlive - t0);
        for (int e = threadIdx.x; e < ns * R; e += blockDim.x) {
            acc[e] = 0.0;
            cnt[e] = 0;
        }
        __syncthreads();
        // a thread owns segment r: cells in order (fixed summation order), 16 loads in flight
        for (int r = threadIdx.x; r < R; r += blockDim.x) {
            for (int i = i0; i < i1; i += 16) {
                int sl[16];
                double v[16];
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    sl[u] = (i + u < i1) ? D.Zs[i + u] - t0 : -1;
                    v[u] = (i + u < i1) ? D.X[(size_t)(i + u) * R + r] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    if (sl[u] >= 0 && sl[u] < ns && !cs_isna(v[u])) {
                        acc[sl[u] * R + r] += v[u];
                        cnt[sl[u] * R + r] += 1;
                    }
                }
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < ns * R; e += blockDim.x) {
            const size_t o = ((size_t)blockIdx.x * nlive + t0) * R + e; // partials are packed by nlive
            D.psum[o] = acc[e];
            D.pcnt[o] = cnt[e];
        }
        __syncthreads();
    }
}

// (the chunk partials of an element are summed by FIN_PARTS threads, each over a contiguous
// quarter of the chunks in order, and the quarters are added in order: a fixed order again,
// at a quarter of the latency)
constexpr int FIN_ELEMS = 64, FIN_PARTS = 4;
__global__ void __launch_bounds__(FIN_ELEMS * FIN_PARTS) stats_final_kernel(ChainDev D, int nchunks)
{
    __shared__ double s_s[FIN_PARTS][FIN_ELEMS];
    __shared__ long long s_c[FIN_PARTS][FIN_ELEMS];
    const int nlive = D.scal->nlive;
    const int le = threadIdx.x % FIN_ELEMS, part = threadIdx.x / FIN_ELEMS;
    const int e = blockIdx.x * FIN_ELEMS + le;
    const int per = (nchunks + FIN_PARTS - 1) / FIN_PARTS;
    const int c0 = part * per, c1 = min(nchunks, c0 + per);
    double s = 0.0;
    long long c = 0;
    if (e < nlive * D.R) {
        for (int ch = c0; ch < c1; ch += 16) { // chunks in order, 16 loads in flight
            double ps[16];
            int pc[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const size_t o = (size_t)(ch + k) * nlive * D.R + e;
                ps[k] = (ch + k < c1) ? D.psum[o] : 0.0;
                pc[k] = (ch + k < c1) ? D.pcnt[o] : 0;
            }
#pragma unroll
            for (int k = 0; k < 16; k++) {
                s += ps[k];
                c += pc[k];
            }
        }
    }
    s_s[part][le] = s;
    s_c[part][le] = c;
    __syncthreads();
    if (part == 0 && e < nlive * D.R) {
        for (int p2 = 1; p2 < FIN_PARTS; p2++) {
            s += s_s[p2][le];
            c += s_c[p2][le];
        }
        D.mean[e] = c ? s / (double)c : 0.0;
    }
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
__global__ void __launch_bounds__(1024)
sigma_partial_kernel(ChainDev D)
{
    const int i0 = blockIdx.x * D.chunk, i1 = min(D.N, i0 + D.chunk);
    for (int r = threadIdx.x; r < D.R; r += blockDim.x) {
        double s = 0.0;
        int c = 0;
        for (int i = i0; i < i1; i += 16) { // cells in order, 16 (x, u) pairs in flight
            double x[16], u[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const bool ok = i + k < i1;
                x[k] = ok ? D.X[(size_t)(i + k) * D.R + r] : NAN;
                u[k] = ok ? D.U[(size_t)D.Zs[i + k] * D.R + r] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (cs_isna(x[k])) continue;
                c++;
                const double d = x[k] - u[k];
                if (!cs_isna(d)) s = __fma_rn(d, d, s);
            }
        }
        D.psum[(size_t)blockIdx.x * D.R + r] = s;
        D.pcnt[(size_t)blockIdx.x * D.R + r] = c;
    }
}

// sigma_r = sqrt(S_r / c_r) with S_r the squared residuals and c_r the non-NA count of segment
// r (:234).  The total log-likelihood at the state just recorded (:245) follows from the same
// sums: sum_i sum_r dnorm(x; u, sigma_r, log) = -sum_r [c_r (ln sqrt(2 pi) + ln sigma_r) +
// S_r / (2 sigma_r^2)] — no second pass over the data.  A thread per segment (chunks in
// order); the per-segment terms go to D.mean[0..R) (free once the means are redrawn) and are
// summed in segment order by record_kernel.
__global__ void __launch_bounds__(FIN_ELEMS * FIN_PARTS) sigma_final_kernel(ChainDev D, int nchunks, int tt)
{
    __shared__ double s_s[FIN_PARTS][FIN_ELEMS];
    __shared__ long long s_c[FIN_PARTS][FIN_ELEMS];
    const int le = threadIdx.x % FIN_ELEMS, part = threadIdx.x / FIN_ELEMS;
    const int r = blockIdx.x * FIN_ELEMS + le;
    const int per = (nchunks + FIN_PARTS - 1) / FIN_PARTS;
    const int c0 = part * per, c1 = min(nchunks, c0 + per);
    double s = 0.0;
    long long c = 0;
    if (r < D.R) {
        for (int ch = c0; ch < c1; ch += 16) { // chunks in order, 16 loads in flight
            double ps[16];
            int pc[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                ps[k] = (ch + k < c1) ? D.psum[(size_t)(ch + k) * D.R + r] : 0.0;
                pc[k] = (ch + k < c1) ? D.pcnt[(size_t)(ch + k) * D.R + r] : 0;
            }
#pragma unroll
            for (int k = 0; k < 16; k++) {
                s += ps[k];
                c += pc[k];
            }
        }
    }
    s_s[part][le] = s;
    s_c[part][le] = c;
    __syncthreads();
    if (part != 0 || r >= D.R) return;
    for (int p2 = 1; p2 < FIN_PARTS; p2++) {
        s += s_s[p2][le];
        c += s_c[p2][le];
    }
    const double sg = sqrt((1.0 / (double)c) * s);
    D.sig[r] = sg;
    D.isig[r] = 1.0 / sg;
    D.sigma_all[(size_t)tt * D.R + r] = sg;
    D.mean[r] = (c > 0) ? (double)c * (CS_LN_SQRT_2PI + log(sg)) + 0.5 * s / (sg * sg) : 0.0;
}

// total fixed-state log-likelihood (R/BayesNonparCluster.R:133,245): per-chunk partials
__global__ void __launch_bounds__(256)
ll_partial_kernel(ChainDev D, const double *__restrict__ U, const int *__restrict__ Zs, const double *__restrict__ sig)
{
    __shared__ double s_w[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * D.chunk, i1 = min(D.N, i0 + D.chunk);
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
    if (threadIdx.x == 32) { // total log-likelihood: sigma_final_kernel's per-segment terms, in segment order
        double ll = 0.0;
        for (int r0 = 0; r0 < D.R; r0 += 16) {
            double t[16];
#pragma unroll
            for (int k = 0; k < 16; k++) t[k] = (r0 + k < D.R) ? D.mean[r0 + k] : 0.0;
#pragma unroll
            for (int k = 0; k < 16; k++) ll -= t[k];
        }
        D.LL[tt] = ll;
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

// CTAs in the resident-window walk's cluster (0: the kernel cannot run with this much shared
// memory): the largest of 16, 8, 4, 2, 1 the device co-schedules.  Worked out once per shape.
// (function attributes and what a device co-schedules are per device: the caches below are
//  indexed by the current device; host threads may race to fill an entry with the same value)
constexpr int MAX_DEVICES = 64;
inline int current_device()
{
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) d = 0;
    return d;
}

template <int NCH, typename Cfg>
int ring_cluster_size(size_t smem)
{
    static size_t known_smem_d[MAX_DEVICES] = {};
    static int known_d[MAX_DEVICES] = {};
    static bool have_d[MAX_DEVICES] = {};
    const int dev = current_device();
    if (have_d[dev] && known_smem_d[dev] == smem) return known_d[dev];
    size_t &known_smem = known_smem_d[dev];
    int &known = known_d[dev];
    have_d[dev] = true;
    known_smem = smem;
    known = 0;
    if (smem > 227 * 1024) return 0;
    if (cudaFuncSetAttribute(scan_ring_kernel<NCH, Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(scan_ring_kernel<NCH, Cfg>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    // (CS_RING_CLUSTER caps the cluster: smaller clusters co-schedule more easily when several
    //  chains share the GPU; the helpers only matter in sweeps that open clusters)
    int cap = 16;
    if (const char *e = getenv("CS_RING_CLUSTER")) cap = atoi(e) < 1 ? 1 : atoi(e);
    for (int cs = 16; cs >= 1; cs >>= 1) {
        if (cs > cap) continue;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs);
        cfg.blockDim = dim3(WIN_THREADS);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = cs;
        at.val.clusterDim.y = 1;
        at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, scan_ring_kernel<NCH, Cfg>, &cfg) == cudaSuccess && n >= 1) {
            known = cs;
            break;
        }
        (void)cudaGetLastError();
    }
    return known;
}

template <int NCH, typename Cfg>
int launch_ring(const ChainDev &D, int tt, cudaStream_t st)
{
    const size_t smem_ring = ring_smem_bytes<Cfg>(D.R, D.kcap);
    const int csize = ring_cluster_size<NCH, Cfg>(smem_ring);
    if (csize <= 0) return CS_OK;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize);
    cfg.blockDim = dim3(WIN_THREADS);
    cfg.dynamicSmemBytes = smem_ring;
    cfg.stream = st;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = csize;
    at.val.clusterDim.y = 1;
    at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    CS_CUDA(cudaLaunchKernelEx(&cfg, scan_ring_kernel<NCH, Cfg>, D, tt));
    return CS_OK;
}

// the fixed-point walk: a cooperative launch of as many CTAs as the device keeps resident (at
// most one per tile)
template <int NCH>
int launch_fix(const ChainDev &D, int tt, int nsm, cudaStream_t st)
{
    bool bits_in_smem = (size_t)fx_words(D.N) * sizeof(unsigned) <= 64 * 1024;
    const size_t smem = fix_smem_bytes(D.N, D.R, D.kcap, bits_in_smem);
    static size_t attr_d[MAX_DEVICES] = {};
    static int occ_d[MAX_DEVICES] = {};
    const int dev = current_device();
    if (attr_d[dev] != smem) {
        CS_CUDA(cudaFuncSetAttribute(fix_walk_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        CS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fix_walk_kernel<NCH>, FX_THREADS, smem));
        if (occ < 1) {
            cs_set_error("fix_walk_kernel cannot be resident with %zu bytes of shared memory", smem);
            return CS_ERR_UNSUPPORTED;
        }
        occ_d[dev] = occ;
        attr_d[dev] = smem;
    }
    int grid = occ_d[dev] * nsm;
    if (const char *e = getenv("CS_FIX_CTAS_PER_SM")) grid = (atoi(e) < 1 ? 1 : (atoi(e) > occ_d[dev] ? occ_d[dev] : atoi(e))) * nsm;
    const int tiles = fx_tiles(D.N);
    if (grid > tiles) grid = tiles;
    CS_CUDA(cudaMemsetAsync(D.fx_scal, 0x7f, sizeof(FixScal), st));
    ChainDev Dc = D;
    int tt_c = tt, bis = bits_in_smem ? 1 : 0;
    void *args[] = {&Dc, &tt_c, &bis};
    CS_CUDA(cudaLaunchCooperativeKernel((const void *)fix_walk_kernel<NCH>, dim3(grid), dim3(FX_THREADS), args, smem, st));
    return CS_OK;
}

template <int NCH>
int launch_sweep(const ChainDev &D, int tt, int flags, int *rej, int nsm, cudaStream_t st, cudaEvent_t *ev)
{
    const int force_all = flags & 1;
    const size_t smem_prep = sizeof(double) * D.R + sizeof(int) * D.kcap;
    const size_t smem_scan = sizeof(double) * ((size_t)D.R + 2 * D.kcap + 64 * WIN_CELLS) + sizeof(int) * 3 * D.kcap + 16;
    // (function attributes are set when the sizes change, not every sweep)
    static size_t attr_scan_d[MAX_DEVICES] = {}, attr_prep_d[MAX_DEVICES] = {};
    size_t &attr_scan = attr_scan_d[current_device()], &attr_prep = attr_prep_d[current_device()];
    if (attr_scan != smem_scan || attr_prep != smem_prep) {
        CS_CUDA(cudaFuncSetAttribute(scan_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scan));
        if (smem_scan > 48 * 1024) {
            CS_CUDA(cudaFuncSetAttribute(prepare_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep));
            CS_CUDA(cudaFuncSetAttribute(seq_scan_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scan));
        }
        attr_scan = smem_scan;
        attr_prep = smem_prep;
    }
    const bool reject_live = D.single && tt == 0;
    if (ev) CS_CUDA(cudaEventRecord(ev[0], st));
    if (force_all) {
        CS_CUDA(cudaMemsetAsync(D.qcols, 0, sizeof(int) * (size_t)D.N, st));
        if (ev) CS_CUDA(cudaEventRecord(ev[1], st));
        seq_scan_kernel<NCH><<<1, SCAN_THREADS, smem_scan, st>>>(D, tt, 1, rej);
    } else {
        int grid = (D.N + PREP_THREADS / 32 - 1) / (PREP_THREADS / 32);
        if (grid > nsm * 8) grid = nsm * 8;
        prepare_kernel<NCH><<<grid, PREP_THREADS, smem_prep, st>>>(D, tt, rej);
        if (ev) CS_CUDA(cudaEventRecord(ev[1], st));
        if (reject_live) {
            seq_scan_kernel<NCH><<<1, SCAN_THREADS, smem_scan, st>>>(D, tt, 0, rej);
        } else if (!(flags & 4)) {
            if (ev) CS_CUDA(cudaEventRecord(ev[4], st));
            { int rc2 = launch_fix<NCH>(D, tt, nsm, st); if (rc2) return rc2; }
        } else { // (flags bit 2: the round-1 walk — inverse index + one-CTA ring walk)
            CS_CUDA(cudaMemsetAsync(D.icnt, 0, sizeof(int) * ((size_t)D.N + 1), st));
            index_count_kernel<<<nsm * 8, 256, 0, st>>>(D);
            index_scan_kernel<<<1, 1024, 0, st>>>(D);
            index_fill_kernel<<<nsm * 8, 256, 0, st>>>(D);
            if (ev) CS_CUDA(cudaEventRecord(ev[4], st));
            // the resident-window walks — the wide shape while there are few clusters, the deep one
            // for up to 32 — then the generic one for whatever they left (nothing, unless a 33rd
            // cluster opened); each starts where the one before stopped
            { int rc2 = launch_ring<NCH, RingWide>(D, tt, st); if (rc2) return rc2; }
            { int rc2 = launch_ring<NCH, RingDeep>(D, tt, st); if (rc2) return rc2; }
            scan_kernel<NCH><<<1, WIN_THREADS, smem_scan, st>>>(D, tt);
        }
    }
    if (ev) CS_CUDA(cudaEventRecord(ev[2], st));
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

} // namespace

int cs_philox_max_R() { return 1024; }

// one full sweep (asynchronous).  D.U / D.Unext are swapped by the caller afterwards.
int cs_philox_sweep(const ChainDev &D, int tt, int last, int flags, int *rej, cudaStream_t st,
                    cudaEvent_t *ev)
{
    const int nsm = cs_num_sms();
    int rc;
    if (D.R <= 32) rc = launch_sweep<1>(D, tt, flags, rej, nsm, st, ev);
    else if (D.R <= 64) rc = launch_sweep<2>(D, tt, flags, rej, nsm, st, ev);
    else if (D.R <= 128) rc = launch_sweep<4>(D, tt, flags, rej, nsm, st, ev);
    else if (D.R <= 320) rc = launch_sweep<10>(D, tt, flags, rej, nsm, st, ev);
    else if (D.R <= 1024) rc = launch_sweep<32>(D, tt, flags, rej, nsm, st, ev);
    else {
        cs_set_error("Philox sampler supports up to 1024 segments (R=%d)", D.R);
        return CS_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
    const int nchunks = (D.N + D.chunk - 1) / D.chunk;
    // (a thread per segment; 72 KB of accumulators = 20 clusters of 300 segments per pass, three CTAs per SM)
    const int stat_smem = 72 * 1024;
    const int seg_threads = D.R >= 1024 ? 1024 : (D.R < 64 ? 64 : ((D.R + 31) / 32) * 32);
    static bool attr_stats_d[MAX_DEVICES] = {};
    bool &attr_stats = attr_stats_d[current_device()];
    if (!attr_stats) {
        CS_CUDA(cudaFuncSetAttribute(stats_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, stat_smem));
        attr_stats = true;
    }
    stats_partial_kernel<<<nchunks, seg_threads, stat_smem, st>>>(D, stat_smem / 8);
    stats_final_kernel<<<(D.kcap * D.R + FIN_ELEMS - 1) / FIN_ELEMS, FIN_ELEMS * FIN_PARTS, 0, st>>>(D, nchunks);
    redraw_kernel<<<D.kcap, 128, 0, st>>>(D, tt, last);
    sigma_partial_kernel<<<nchunks, seg_threads, 0, st>>>(D);
    sigma_final_kernel<<<(D.R + FIN_ELEMS - 1) / FIN_ELEMS, FIN_ELEMS * FIN_PARTS, 0, st>>>(D, nchunks, tt);
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
    const int nchunks = (D.N + D.chunk - 1) / D.chunk;
    ll_partial_kernel<<<nchunks, 256, 0, st>>>(D, U, Zrows, sig);
    ll_final_kernel<<<1, 32, 0, st>>>(D, nchunks, d_out);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
