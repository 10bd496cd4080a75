// cs_gibbs.cuh — device state of one nested-CRP Gibbs chain (shared by the Philox
// path, the R-compatible path and the host API).
#pragma once
#include "cs_common.cuh"

// device-resident scalars of a chain
struct ChainScal {
    int nlive;        // Philox: live slots (rows of U in use) ; R mode: unused
    int kt;           // labels ever created == nrow(Uall[[t]])
    int first_event;  // Philox: first cell of the sweep whose speculative draw moved it
    int err;          // CS_ERR_* raised on the device
    int mt_pos;       // R mode: Mersenne-Twister position
    int prop_valid;   // R mode scratch
    int open_cell;    // Philox ring scan: command to the helper CTAs (cell that opened a cluster, -1 = leave)
    int open_slot;    //                    the slot it opened
    long long cells_scanned, n_events, n_reject, u_used;
    long long n_rounds, cut_margin, cut_open, cut_maxev; // walk diagnostics: rounds, and what ended a round's commits early
};

// one entry of the inverse source index: a (cell, segment) whose proposal copies the source
// cell's cluster state, with the value the correction needs riding along (one 16-byte record:
// one store when the index is built, one 16-byte copy when it is read)
struct __align__(16) InvEntry {
    unsigned cr;      // cell << 10 | segment
    unsigned pad;
    double x;         // X[cell][segment]
};

struct ChainDev {
    int N, R, kcap;
    int chunk;        // cells per end-of-sweep reduction chunk
    double alpha, beta;
    uint32_t seed;
    int single;       // initial assignment had one distinct label (rejection live in sweep 0)
    double *X;        // N x R cell-major
    int *Zs;          // Philox: slot of every cell ; R mode: 1-based label
    double *U;        // kcap x R rows
    double *Unext;    // compaction target (Philox)
    int *np;          // kcap counts
    int *np0;         // kcap counts at the start of the sweep (Philox; written by prepare_kernel)
    int *slot_label;  // kcap (Philox)
    double *sig;      // R
    double *isig;     // R : 1/sig
    // Philox path (all "columns" are slot-major: element (slot s, cell c) at [s*N + c])
    double *Q;        // kcap x N   squared standardised distance of cell c to cluster s (2^-32 fixed point)
    double *E;        // kcap x N   exp(-0.5 (Q - qmin[c])) cached by the prepare kernel
    long long *qnew_i; // N         fixed-point distance of cell c to ITS proposed new cluster
    double *qmin;     // N          the reference value the cached exponentials were taken against
    double *ua;       // N          uniform of the categorical draw of this sweep
    long long *wn_q;  // N          proposal distance the cached new-table weight was computed at
    double *wn_w;     // N          beta * exp(-0.5 (q_new - qmin)) at wn_q
    int *kmin;        // N          a live cluster attaining qmin (255: only the proposal does)
    int *J;           // N x R      row c: the proposal's sources in front of cell c, packed (source << 10 | segment)
    int *jcnt;        // N          how many
    int *zspec;       // N          speculative draws
    int *qcols;       // N          (sequential kernel) valid columns of Q
    int *icnt;        // N+1        inverse source index: offsets (after the scan)
    int *icur;        // N          fill cursors
    InvEntry *inv;    // N x R      inverse source index, entries grouped by source cell
    unsigned *near;   // N x CS_NEAR_CAP  proposal sources at most CS_NEAR_W cells in front of the cell
                      //            (source << 10 | segment; CS_NEAR_EMPTY behind the last; CS_NEAR_OVER: more than fit)
    long long *tile_sums; // scan scratch
    // fixed-point walk (fix_walk_kernel): second copy of the hypothesis (the first is zspec), two
    // copies of the moved-cell bitmap, the tiles' net moves per slot and their prefix over the
    // tiles ((kcap + 1) rows of fx_tpad(N) tiles), the per-step minima
    int *zh1;
    unsigned *fx_bits;
    int *fx_td, *fx_tp;
    int *fx_scal;
    double *mean;     // kcap x R cluster means
    double *psum;     // partial sums workspace
    int *pcnt;        // partial counts workspace
    int *newslot;     // kcap
    uint32_t *mt;     // 624 (R mode)
    ChainScal *scal;
    // outputs (device)
    int *Zall;        // niter x N sweep-major
    double *sigma_all; // niter x R
    double *LL;       // niter
    int *kt_all;      // niter
    long long *u_off; // niter+1
    int *u_labels;    // urec_cap
    double *u_rows;   // urec_cap x R
    long long urec_cap;
};

// near sources of a cell's proposal: the cells within CS_NEAR_W in front of it whose cluster state
// it copies (what an event inside the same window of the walk can invalidate)
constexpr int CS_NEAR_W = 512, CS_NEAR_CAP = 8;
constexpr unsigned CS_NEAR_EMPTY = 0xFFFFFFFFu, CS_NEAR_OVER = 0xFFFFFFFEu;

// cells per deterministic reduction chunk: at most 512 chunks, at least 256 cells each
inline int cs_stat_chunk(int N) { const int c = (N + 511) / 512; return c < 256 ? 256 : ((c + 31) / 32) * 32; }
