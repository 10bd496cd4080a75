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
//   fix_walk_kernel (cooperative, all SMs; skipped when the sweep has no event)
//       the rest of the sweep as the fixed point of a parallel map: a hypothesis for the outcome
//       of every cell is re-drawn against the state it induces itself, for all cells behind the
//       last settled one at once, until nothing changes (a handful of steps); a cell that opens
//       a cluster stops the prefix, its row is appended and scored by the whole grid, and the
//       iteration goes on behind it.  See the comment above the kernel.
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
constexpr double Q_SCALE = 1.0 / 4294967296.0;

// the same with streaming loads (evict-first): a pass over all of X must not push what the walk
// reads next (J, E, Q) out of L2
template <int NCH>
__device__ __forceinline__ void load_row_stream(const double *__restrict__ xrow, int R, int lane, double (&x)[NCH])
{
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        const int r = lane + 32 * i;
        x[i] = (r < R) ? __ldcs(xrow + r) : 0.0;
    }
}

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
    // llrint(min(d*d, 2^62)), NaN -> 2^62.  The clamp is taken on the integer: the conversion
    // saturates above 2^63 and gives 0x8000000000000000 for a NaN, so "high word >= 0x40000000
    // as unsigned" is exactly "d*d >= 2^62 or NaN" (three integer instructions instead of the
    // seven of a double-precision minimum)
    const double d = __dmul_rn(__dsub_rn(x, u), isig16);
    const long long v = __double2ll_rn(__dmul_rn(d, d));
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    const bool big = hi >= 0x40000000u;
    return (long long)(((unsigned long long)(big ? 0x40000000u : hi) << 32) | (big ? 0u : lo));
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

// ---------------------------------------------------------------- prepare (all SMs)
template <int NCH, int CTAS = 4>
__global__ void __launch_bounds__(PREP_THREADS, CTAS)
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
        load_row_stream<NCH>(D.X + (size_t)c * D.R, D.R, lane, x);
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
        // the sources IN FRONT of this cell, packed (source << 10 | segment) at the start of its row
        // of J: the only part of the proposal that the sweep can still change (fix_walk_kernel)
        int nsrc = 0;
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            const int r = lane + 32 * i;
            const bool fr = r < D.R && jj[i] >= 0 && jj[i] < c;
            const unsigned fb = __ballot_sync(CS_FULL, fr);
            if (fr) D.J[(size_t)c * D.R + nsrc + __popc(fb & ((1u << lane) - 1u))] = (jj[i] << 10) | r;
            nsrc += __popc(fb);
        }
        if (lane == 0) D.jcnt[c] = nsrc;
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
            if (z != zold && c < first) first = c;
        }
    }
    if (lane == 0 && first != INT_MAX) atomicMin(&D.scal->first_event, first);
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
                propose<NCH>(D, c, tt, attempt, nullptr, lane, mu, jj);
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
constexpr int FX_FU = 4;          // hits per lane worked off together
constexpr int FX_PFD = 12;        // source rows prefetched into L2 ahead of the one being tested
constexpr int FX_INF = 0x7f7f7f7f; // (what cudaMemsetAsync(0x7f) leaves)

__device__ __forceinline__ unsigned long long fx_gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct FixScal {
    int fc[3], fo[3]; // per step (mod 3): first changed cell, first cell that opens a cluster
};

__host__ __device__ inline int fx_tiles(int N) { return (N + FX_THREADS - 1) / FX_THREADS; }
__host__ __device__ inline int fx_tpad(int N) { return (fx_tiles(N) + 31) & ~31; }
__host__ __device__ inline int fx_words(int N) { return ((N + FX_THREADS - 1) / FX_THREADS) * FX_WARPS; }
// collected hits per warp: a cell's worth above the flush threshold
__host__ __device__ constexpr int fx_hcap(int NCH) { return 32 * NCH + 192 > 512 ? 32 * NCH + 192 : 512; } // (>= the draw's staging area: 16 slots x 32 lanes)
inline size_t fix_smem_bytes(int N, int R, int kcap, bool bits_in_smem, int NCH)
{
    return sizeof(double) * (size_t)R + sizeof(long long) * FX_THREADS + sizeof(int) * (size_t)(kcap + 1) * (FX_WARPS + 1) +
           sizeof(int2) * FX_WARPS * fx_hcap(NCH) + sizeof(int) * FX_WARPS * 34 +
           (bits_in_smem ? sizeof(unsigned) * (size_t)fx_words(N) : 0) + 48;
}

// the tile's net moves per slot (tdelta[s][tile]), its 256 bits of the moved-cell bitmap, the
// first cell whose hypothesis opens a cluster
__device__ __noinline__ void fx_publish(const ChainDev &D, int tile, int c, bool active, bool counts_open, int z, int zo,
                                           int nmat, int *s_td, unsigned *bits_a, unsigned *bits_b, int *td_a, int *td_b, int *fo)
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
    for (int s = threadIdx.x; s <= nmat; s += FX_THREADS) {
        td_a[(size_t)s * Tp + tile] = s_td[s];
        if (td_b) td_b[(size_t)s * Tp + tile] = s_td[s];
    }
    (void)c;
}

// the column of a cluster opened at cell f (slot b) for every later cell: score_new_column over
// the warps of the whole grid
template <int NCH>
__device__ __noinline__ void fx_score_column(const ChainDev &D, int f, int b, const double *s_isig16)
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

__device__ __forceinline__ void cp_async_8(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// what a warp of the walk works with (one tile, one step)
struct FxWarp {
    const int *zcur;         // the hypothesis h
    const unsigned *bcur;    // bitmap of the cells h moves
    const double *s_isig16;
    long long *qh;           // this warp's 32 sums: what h changes in its cells' proposal distances
    int2 *hits;              // this warp's list of collected hits (also the draw's staging area)
    int *hoff;               // [k]: where cell k's hits start in the list
    int nmat, lo, dbg;
};

// the collected hits of a warp, FX_FU x 32 at a time: the loads of a level are issued together
// (one round costs about one trip to DRAM), the term of every hit is exchanged exactly, the
// differences go back into the list and are summed per cell by the lane of that cell
__device__ __noinline__ void fx_flush(const ChainDev &D, const FxWarp &W, const double *Xw, int nh, int kcur)
{
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (W.dbg & 1) nh = 0;
    for (int base = 0; base < nh; base += 32 * FX_FU) {
        int kr[FX_FU], zj[FX_FU], zs[FX_FU];
        double xv[FX_FU], u1[FX_FU], u0[FX_FU];
#pragma unroll
        for (int u = 0; u < FX_FU; u++) {
            const int idx = base + 32 * u + lane;
            const bool v = idx < nh;
            const int2 h = v ? W.hits[idx] : make_int2(0, 0);
            kr[u] = h.x;
            zj[u] = v ? W.zcur[h.y] : INT_MAX;
            zs[u] = v ? D.Zs[h.y] : 0;
            xv[u] = v ? ((W.dbg & 8) ? 1.0 : Xw[(h.x >> 10) * D.R + (h.x & 1023)]) : CUDART_NAN;
        }
#pragma unroll
        for (int u = 0; u < FX_FU; u++) {
            // (a source that opens a cluster lies behind the settled prefix: no row yet)
            const bool ok = zj[u] < W.nmat && !cs_isna(xv[u]);
            if (!ok) kr[u] = -1;
            u1[u] = ok ? D.U[(size_t)zj[u] * D.R + (kr[u] & 1023)] : 0.0;
            u0[u] = ok ? D.U[(size_t)zs[u] * D.R + (kr[u] & 1023)] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < FX_FU; u++) {
            const int idx = base + 32 * u + lane;
            long long d = 0;
            if (kr[u] >= 0) {
                const double is = W.s_isig16[kr[u] & 1023];
                d = tint(xv[u], u1[u], is) - tint(xv[u], u0[u], is);
            }
            if (idx < nh) *reinterpret_cast<long long *>(W.hits + idx) = d;
        }
    }
    __syncwarp();
    {
        const int i0 = lane <= kcur ? W.hoff[lane] : nh, i1 = lane < kcur ? W.hoff[lane + 1] : nh;
        long long sum = 0;
        for (int idx = i0; idx < i1; idx++) sum += *reinterpret_cast<const long long *>(W.hits + idx);
        if (sum) W.qh[lane] += sum;
    }
    __syncwarp();
    W.hoff[lane] = 0;
    __syncwarp();
}

// pull: what h changes in the proposal distances of a warp's 32 cells.  A cell's sources in
// front of it (packed by prepare_kernel at the start of its row of J; rows are fetched two
// cells ahead) are looked up in the bitmap of moved cells; the hits are appended to the warp's
// list (one prefix sum over the lanes per cell) and worked off in batches (fx_flush) — hits are
// latency, not work: batching them is what overlaps it.
template <int NCH>
__device__ __forceinline__ void fx_pull(const ChainDev &D, const FxWarp &W, int cbase)
{
    constexpr int HCAP = fx_hcap(NCH);
    const int lane = threadIdx.x & 31;
    const int k0 = max(0, W.lo - cbase), k1 = min(32, D.N - cbase);
    if (k0 >= k1) return;
    const size_t ebase = (size_t)cbase * D.R;
    const double *Xw = D.X + ebase;
    const int mycnt = (lane >= k0 && lane < k1) ? D.jcnt[cbase + lane] : 0; // sources in front of cell cbase + lane
    W.hoff[lane] = 0;
    __syncwarp();
    int nh = 0;
    auto load_sources = [&](int k, int (&dst)[NCH]) {
        const int cnt = __shfl_sync(CS_FULL, mycnt, k & 31);
        const int *Jrow = D.J + ebase + (size_t)k * D.R;
#pragma unroll
        for (int i = 0; i < NCH; i++) dst[i] = (k < k1 && lane + 32 * i < cnt) ? __ldg(Jrow + lane + 32 * i) : -1;
    };
    auto test_sources = [&](int k, const int (&src)[NCH]) {
        if (lane == 0) W.hoff[k] = nh;
        unsigned m = 0;
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            const int e = src[i];
            if (e >= 0 && ((W.bcur[e >> 15] >> ((e >> 10) & 31)) & 1u)) m |= 1u << i;
        }
        if (!(W.dbg & 16) && __any_sync(CS_FULL, m != 0)) {
            const int mine = __popc(m);
            int inc = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int o = __shfl_up_sync(CS_FULL, inc, off);
                if (lane >= off) inc += o;
            }
            int pos = nh + inc - mine;
#pragma unroll
            for (int i = 0; i < NCH; i++)
                if ((m >> i) & 1u) W.hits[pos++] = make_int2((k << 10) | (src[i] & 1023), src[i] >> 10);
            nh += __shfl_sync(CS_FULL, inc, 31);
            if (nh > HCAP - 32 * NCH) {
                fx_flush(D, W, Xw, nh, k);
                nh = 0;
            }
        }
    };
    auto prefetch_row = [&](int k) { // row k of the warp's cells into L2: a 128-byte line per lane
        if (k < k1) {
            const int cnt = __shfl_sync(CS_FULL, mycnt, k & 31);
            if (lane * 32 < cnt) asm volatile("prefetch.global.L2 [%0];" ::"l"(D.J + ebase + (size_t)k * D.R + lane * 32));
        }
    };
    for (int k = k0 + 2; k < k0 + FX_PFD && k < k1; k++) prefetch_row(k);
    int ja[NCH], jb[NCH];
    load_sources(k0, ja);
    load_sources(k0 + 1, jb);
    for (int k = k0; k < k1; k += 2) {
        prefetch_row(k + FX_PFD);
        prefetch_row(k + FX_PFD + 1);
        test_sources(k, ja);
        load_sources(k + 2, ja);
        if (k + 1 < k1) test_sources(k + 1, jb);
        load_sources(k + 3, jb);
    }
    fx_flush(D, W, Xw, nh, k1 - 1);
}

// draw of one cell (a thread) against the state h induces in front of it.  The slot loops are
// uniform over the warp (the counts come from ballots); the cached exponentials of up to FX_G
// slots are staged in the warp's list area by asynchronous copies while the counts are worked
// out, and the cumulative weights replace them there, so the draw reads them once.  The
// arithmetic is draw_cell_warp's, in the same order.
constexpr int FX_G = 16;
__device__ __noinline__ int fx_draw(const ChainDev &D, const FxWarp &W, int c, bool evalc, int zc, int zo, bool warp_moves,
                                    const int *s_base, const int *wd)
{
    const int lane = threadIdx.x & 31;
    const unsigned ltmask = (1u << lane) - 1u;
    const size_t N = (size_t)D.N;
    const int nmat = W.nmat;
    double *stage = reinterpret_cast<double *>(W.hits); // [slot in group][lane]
    const int ng = (nmat + FX_G - 1) / FX_G;
    auto stage_group = [&](int g) {
        const int s1 = min(nmat, (g + 1) * FX_G);
        if (evalc)
            for (int s = g * FX_G; s < s1; s++) cp_async_8(stage + (s - g * FX_G) * 32 + lane, D.E + (size_t)s * N + c);
    };
    __syncwarp();
    stage_group(0);
    auto count_at = [&](int s) -> int { // cells in slot s in front of this cell (itself included while it has not moved)
        int n = s_base[s] + wd[s];
        if (warp_moves) {
            const unsigned bn = __ballot_sync(CS_FULL, zc == s), bo = __ballot_sync(CS_FULL, zo == s);
            n += __popc(bn & ltmask) - __popc(bo & ltmask);
        }
        return n;
    };
    const long long qi = evalc ? D.qnew_i[c] + W.qh[lane] : 0;
    const double q_new = (double)qi * Q_SCALE;
    const double qmin_c = evalc ? D.qmin[c] : 0.0, ua = evalc ? D.ua[c] : 0.0;
    const int km = evalc ? D.kmin[c] : 255;
    int nkm = 1, nzo = 0;
    for (int s = 0; s < nmat; s++) {
        const int n = count_at(s);
        if (s == km) nkm = n;
        if (s == zo) nzo = n;
    }
    // (the conditions under which the cached exponentials are today's: their reference still is
    //  the minimum over the candidates)
    bool cached = (km != 255 ? (q_new >= qmin_c && nkm - (km == zo ? 1 : 0) > 0) : q_new == qmin_c);
    if (evalc && cached && D.np0[zo] < 2 && nzo >= 2 && D.Q[(size_t)zo * N + c] < qmin_c) cached = false;
    double qref = qmin_c;
    if (__any_sync(CS_FULL, evalc && !cached)) { // (rare: the minimum over today's candidates, from the distances)
        double qm = q_new;
        for (int s = 0; s < nmat; s++) {
            const int n = count_at(s) - (s == zo ? 1 : 0);
            if (evalc && !cached && n > 0) qm = fmin(qm, D.Q[(size_t)s * N + c]);
        }
        if (!cached) {
            cached = qm == qmin_c;
            qref = qm;
        }
    }
    const double w_new = __dmul_rn(D.beta, exp(-0.5 * (q_new - qref)));
    auto weight = [&](int s, double e) -> double {
        const int n = count_at(s) - (s == zo ? 1 : 0);
        double w = 0.0;
        if (evalc && n > 0) w = __dmul_rn((double)n, cached ? e : exp(-0.5 * (D.Q[(size_t)s * N + c] - qref)));
        return w;
    };
    double run = 0.0;
    int lastpos = -1;
    for (int g = 0; g < ng; g++) {
        if (g) {
            __syncwarp();
            stage_group(g);
        }
        cp_async_wait_all();
        __syncwarp();
        const int s1 = min(nmat, (g + 1) * FX_G);
        for (int s = g * FX_G; s < s1; s++) {
            double *slot = stage + (s - g * FX_G) * 32 + lane;
            const double w = weight(s, *slot);
            run = __dadd_rn(run, w);
            if (w > 0.0) lastpos = s;
            *slot = run; // (cumulative weights: what the second pass reads when one group holds all slots)
        }
    }
    const double t = __dmul_rn(ua, __dadd_rn(run, w_new));
    int zz = -1;
    if (ng == 1) {
        for (int s = 0; s < nmat; s++)
            if (zz < 0 && stage[s * 32 + lane] > t) zz = s;
    } else {
        run = 0.0;
        for (int g = 0; g < ng; g++) {
            __syncwarp();
            stage_group(g);
            cp_async_wait_all();
            __syncwarp();
            const int s1 = min(nmat, (g + 1) * FX_G);
            for (int s = g * FX_G; s < s1; s++) {
                const double w = weight(s, stage[(s - g * FX_G) * 32 + lane]);
                if (w > 0.0) {
                    run = __dadd_rn(run, w);
                    if (zz < 0 && run > t) zz = s;
                }
            }
        }
    }
    if (zz < 0) zz = (w_new > 0.0) ? nmat : lastpos;
    __syncwarp();
    return evalc ? zz : zc;
}

template <int NCH>
__global__ void __launch_bounds__(FX_THREADS, 3)
fix_walk_kernel(ChainDev Darg, int tt, int bits_in_smem, unsigned long long *prof, int dbg)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char s_fx_raw[];
    double *s_isig16 = reinterpret_cast<double *>(s_fx_raw);            // R
    long long *s_qh = reinterpret_cast<long long *>(s_isig16 + Darg.R);    // what h changes in the proposal distances of the tile's cells
    int *s_base = reinterpret_cast<int *>(s_qh + FX_THREADS);           // kcap + 1: counts in front of the tile
    int *s_wd = s_base + (Darg.kcap + 1);                                  // FX_WARPS x (kcap + 1): moves in front of each warp
    const int KS = Darg.kcap + 1;
    int2 *s_hits = reinterpret_cast<int2 *>((reinterpret_cast<uintptr_t>(s_wd + FX_WARPS * KS) + 15) & ~(uintptr_t)15); // FX_WARPS lists of collected hits
    constexpr int HCAP = fx_hcap(NCH);
    int *s_hoff = reinterpret_cast<int *>(s_hits + FX_WARPS * HCAP);
    unsigned *s_bits = reinterpret_cast<unsigned *>(s_hoff + FX_WARPS * 34);

    // (uniform over the grid: the scalars are rewritten only behind the last grid barrier)
    const int f0 = Darg.scal->first_event;
    if (f0 >= Darg.N) return;
    // (the out-of-line helpers take the descriptor by reference: a copy in shared memory, not a
    //  local-memory copy of the kernel parameter)
    __shared__ ChainDev s_D;
    if (threadIdx.x == 0) s_D = Darg;
    __syncthreads();
    const ChainDev &D = s_D;
    int nmat = D.scal->nlive, kt = D.scal->kt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = fx_tiles(D.N), Tp = fx_tpad(D.N), NW = fx_words(D.N);
    for (int r = threadIdx.x; r < D.R; r += FX_THREADS) s_isig16[r] = D.isig[r] * 65536.0;
    int *zh[2] = {D.zspec, D.zh1};
    unsigned *bits[2] = {D.fx_bits, D.fx_bits + NW};
    int *td[2] = {D.fx_td, D.fx_tp}; // the tiles' net moves per slot, one copy per step parity
    FixScal *fxs = reinterpret_cast<FixScal *>(D.fx_scal);

    // ---- step -1: the hypothesis is prepare_kernel's draws against the start of the sweep
    for (int tile = blockIdx.x; tile < T; tile += gridDim.x) {
        const int c = tile * FX_THREADS + threadIdx.x;
        const bool active = c < D.N;
        const int z = active ? D.zspec[c] : -1, zo = active ? D.Zs[c] : -1;
        if (active) D.zh1[c] = z;
        // (rows of slots that open later in the sweep: nothing of an earlier sweep may be left in them)
        for (int s2 = nmat + 1 + threadIdx.x; s2 <= D.kcap; s2 += FX_THREADS) {
            td[0][(size_t)s2 * Tp + tile] = 0;
            td[1][(size_t)s2 * Tp + tile] = 0;
        }
        fx_publish(D, tile, c, active, active, z, zo, nmat, s_wd, bits[0], bits[1], td[0], td[1], &fxs->fo[2]);
        __syncthreads();
    }
    grid.sync();
    int open_cur = *reinterpret_cast<volatile int *>(&fxs->fo[2]);

    int lo = f0, err = 0, steps = 0, opened = 0;
    long long evaluated = 0;
    for (int it = 0;; it++) {
        const int cur = it & 1, nxt = cur ^ 1, set = it % 3;
        if (blockIdx.x == 0 && threadIdx.x == 0) { // next step's minima (last read behind the barrier before the previous one)
            fxs->fc[(it + 1) % 3] = FX_INF;
            fxs->fo[(it + 1) % 3] = FX_INF;
        }
        const unsigned *bcur = bits[cur];
        if (bits_in_smem) {
            for (int w = threadIdx.x; w < NW; w += FX_THREADS) s_bits[w] = bcur[w];
            bcur = s_bits;
        }
        const int *zcur = zh[cur];
        steps++;
        unsigned long long pt0 = 0, pt1 = 0, pt2 = 0, pt3 = 0;
        if (prof && threadIdx.x == 0) pt0 = fx_gtime();
        for (int tile = lo / FX_THREADS + blockIdx.x; tile < T; tile += gridDim.x) {
            __syncthreads();
            const int c = tile * FX_THREADS + threadIdx.x;
            const bool active = c < D.N, evalc = active && c >= lo;
            const int zo = active ? D.Zs[c] : -1, zc = active ? zcur[c] : -1;
            // ---- counts in front of the tile, moves in front of each warp
            for (int e = threadIdx.x; e < FX_WARPS * KS; e += FX_THREADS) s_wd[e] = 0;
            s_qh[threadIdx.x] = 0;
            for (int s = warp; s <= nmat; s += FX_WARPS) { // (a warp per slot: the net moves of all tiles in front of this one)
                const int *row = td[cur] + (size_t)s * Tp;
                int part = 0;
                for (int t = lane; t < tile; t += 32) part += row[t];
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) part += __shfl_xor_sync(CS_FULL, part, off);
                if (lane == 0) s_base[s] = (s < D.kcap ? D.np0[s] : 0) + part;
            }
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
            unsigned long long ph0 = 0, ph1 = 0, ph2 = 0;
            if (prof && threadIdx.x == 0) ph0 = fx_gtime();
            FxWarp W;
            W.zcur = zcur;
            W.bcur = bcur;
            W.s_isig16 = s_isig16;
            W.qh = s_qh + warp * 32;
            W.hits = s_hits + warp * HCAP;
            W.hoff = s_hoff + warp * 34;
            W.nmat = nmat;
            W.lo = lo;
            W.dbg = dbg;
            if (!(dbg & 4)) fx_pull<NCH>(D, W, tile * FX_THREADS + warp * 32);
            if (prof) {
                __syncthreads();
                if (threadIdx.x == 0) ph1 = fx_gtime();
            }
            const int z = fx_draw(D, W, c, evalc, zc, zo, warp_moves, s_base, s_wd + warp * KS);
            if (evalc) {
                zh[nxt][c] = z;
                if (z != zc) atomicMin(&fxs->fc[set], c);
            }
            __syncthreads();
            if (prof && threadIdx.x == 0 && it < 64) {
                ph2 = fx_gtime();
                atomicAdd(prof + 512 + 8 * it + 0, ph0 - pt0); // (one tile per CTA assumed: since the start of the step)
                atomicAdd(prof + 512 + 8 * it + 1, ph1 - ph0);
                atomicAdd(prof + 512 + 8 * it + 2, ph2 - ph1);
            }
            fx_publish(D, tile, c, active, evalc, z, zo, nmat, s_wd, bits[nxt], nullptr, td[nxt], nullptr, &fxs->fo[set]);
            if (threadIdx.x == 0) evaluated += min(D.N, (tile + 1) * FX_THREADS) - max(lo, tile * FX_THREADS);
        }
        if (prof && threadIdx.x == 0) pt1 = fx_gtime();
        grid.sync();
        if (prof && threadIdx.x == 0) pt2 = fx_gtime();
        const int changed_min = *reinterpret_cast<volatile int *>(&fxs->fc[set]);
        const int open_next = *reinterpret_cast<volatile int *>(&fxs->fo[set]);
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
                const double NpA = (double)D.N + D.alpha;
                for (int r = threadIdx.x; r < D.R; r += FX_THREADS) { // (propose(): the same draws, the sources' slots under h)
                    double ua, ub;
                    cs_philox_draw(D.seed, (uint32_t)f, (uint32_t)r, (uint32_t)tt, CS_PH_PROPOSAL, 0u, ua, ub);
                    const double v = __dmul_rn(ub, NpA);
                    double val;
                    if (v < (double)D.N) {
                        int j = (int)v;
                        if (j > D.N - 1) j = D.N - 1;
                        val = D.U[(size_t)(j < f ? zh[nxt][j] : D.Zs[j]) * D.R + r];
                    } else {
                        val = __fma_rn(2.5 - 0.3, ua, 0.3);
                    }
                    D.U[(size_t)b * D.R + r] = val;
                }
                if (threadIdx.x == 0) D.slot_label[b] = kt + 1;
            }
            kt++;
            nmat++;
            opened++;
            grid.sync();
            fx_score_column<NCH>(D, f, b, s_isig16);
            grid.sync();
            lo = f + 1;
            open_cur = FX_INF;
        } else {
            // (that cell's new draw saw a settled prefix, so it is settled itself; it is evaluated
            //  once more so that both copies of h receive it)
            lo = changed_min;
            open_cur = open_next;
        }
        if (prof && threadIdx.x == 0) pt3 = fx_gtime();
        if (prof && threadIdx.x == 0 && it < 64) { // per step: max / sum of the CTAs' evaluation time, CTA 0's waits
            unsigned long long *pr = prof + 8 * it;
            atomicMax(pr + 0, pt1 - pt0);
            atomicAdd(pr + 1, pt1 - pt0);
            if (pt1 - pt0 > 2000) atomicAdd(pr + 6, 1ull);
            if (blockIdx.x == 0) {
                pr[2] = pt2 - pt1;          // first barrier (waits for the slowest CTA)
                pr[3] = pt3 - pt2;          // opening a cluster (two more barriers)
                pr[4] = 0;
                pr[5] = (unsigned long long)lo;
            }
        }
    }
    // ---- h is the sweep's outcome (both copies agree): assignments, counts, scalars
    {
        const int *zf = zh[0];
        const int *tdf = td[steps & 1]; // (the copy the last step wrote)
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
                for (int t = 0; t < T; t++) tot += tdf[(size_t)s * Tp + t];
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

// (function attributes and what a device co-schedules are per device: the caches below are
//  indexed by the current device; host threads may race to fill an entry with the same value)
constexpr int MAX_DEVICES = 64;
inline int current_device()
{
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) d = 0;
    return d;
}

// the fixed-point walk: a cooperative launch of as many CTAs as the device keeps resident (at
// most one per tile)
template <int NCH>
int launch_fix(const ChainDev &D, int tt, int nsm, cudaStream_t st)
{
    bool bits_in_smem = (size_t)fx_words(D.N) * sizeof(unsigned) <= 64 * 1024;
    const size_t smem = fix_smem_bytes(D.N, D.R, D.kcap, bits_in_smem, NCH);
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
    unsigned long long *prof = nullptr; // (CS_FIX_PROF=1: per-step timings of the first 64 steps, read by tools/walk_diag.py)
    if (getenv("CS_FIX_PROF")) {
        prof = reinterpret_cast<unsigned long long *>(D.tile_sums);
        CS_CUDA(cudaMemsetAsync(prof, 0, sizeof(unsigned long long) * 8 * 128, st));
    }
    int dbg = getenv("CS_FIX_DBG") ? atoi(getenv("CS_FIX_DBG")) : 0;
    void *args[] = {&Dc, &tt_c, &bis, &prof, &dbg};
    CS_CUDA(cudaLaunchCooperativeKernel((const void *)fix_walk_kernel<NCH>, dim3(grid), dim3(FX_THREADS), args, smem, st));
    return CS_OK;
}

template <int NCH>
int launch_sweep(const ChainDev &D, int tt, int flags, int *rej, int nsm, cudaStream_t st, cudaEvent_t *ev)
{
    const int force_all = flags & 1;
    const size_t smem_prep = sizeof(double) * D.R + sizeof(int) * D.kcap;
    const size_t smem_scan = sizeof(double) * (size_t)D.R + sizeof(int) * 2 * D.kcap + 16;
    // (function attributes are set when the sizes change, not every sweep)
    static size_t attr_scan_d[MAX_DEVICES] = {}, attr_prep_d[MAX_DEVICES] = {};
    size_t &attr_scan = attr_scan_d[current_device()], &attr_prep = attr_prep_d[current_device()];
    if (attr_scan != smem_scan || attr_prep != smem_prep) {
        if (smem_scan > 48 * 1024) CS_CUDA(cudaFuncSetAttribute(seq_scan_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scan));
        if (smem_prep > 48 * 1024) CS_CUDA(cudaFuncSetAttribute(prepare_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep));
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
        static const int prep_ctas = getenv("CS_PREP_CTAS") ? atoi(getenv("CS_PREP_CTAS")) : 4;
        if (prep_ctas == 3) {
            if (grid > nsm * 6) grid = nsm * 6;
            prepare_kernel<NCH, 3><<<grid, PREP_THREADS, smem_prep, st>>>(D, tt, rej);
        } else if (prep_ctas == 2) {
            if (grid > nsm * 4) grid = nsm * 4;
            prepare_kernel<NCH, 2><<<grid, PREP_THREADS, smem_prep, st>>>(D, tt, rej);
        } else
            prepare_kernel<NCH><<<grid, PREP_THREADS, smem_prep, st>>>(D, tt, rej);
        if (ev) CS_CUDA(cudaEventRecord(ev[1], st));
        if (reject_live) {
            seq_scan_kernel<NCH><<<1, SCAN_THREADS, smem_scan, st>>>(D, tt, 0, rej);
        } else {
            if (ev) CS_CUDA(cudaEventRecord(ev[4], st));
            { int rc2 = launch_fix<NCH>(D, tt, nsm, st); if (rc2) return rc2; }
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
