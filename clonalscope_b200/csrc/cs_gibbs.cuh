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
    int pad0, pad1;
    long long cells_scanned, n_events, n_reject, u_used;
};

struct ChainDev {
    int N, R, kcap;
    double alpha, beta;
    uint32_t seed;
    int single;       // initial assignment had one distinct label (rejection live in sweep 0)
    double *X;        // N x R cell-major
    int *Zs;          // Philox: slot of every cell ; R mode: 1-based label
    double *U;        // kcap x R rows
    double *Unext;    // compaction target (Philox)
    int *np;          // kcap counts
    int *slot_label;  // kcap (Philox)
    double *sig;      // R
    double *isig;     // R : 1/sig
    double *Q;        // N x kcap (Philox)
    int *J;           // N x R proposal sources (Philox)
    int *zspec;       // N speculative draws (Philox)
    int *qcols;       // N valid columns of Q (Philox)
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

constexpr int CS_STAT_CHUNK = 1024; // cells per deterministic reduction chunk
