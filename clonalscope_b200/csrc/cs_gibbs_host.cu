// cs_gibbs_host.cu — host side of the nested-CRP Gibbs chain: device-resident chain
// object (cs_chain_*) and the R-layout one-shot entry point cs_ncrp_gibbs.
// Reference boundary: BayesNonparCluster, R/BayesNonparCluster.R:21-289.
#include "cs_gibbs.cuh"

#include <chrono>
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

template <typename T>
int cs_transpose(int rows, int cols, const T *d_in, T *d_out, cudaStream_t st);
int cs_philox_sweep(const ChainDev &D, int tt, int last, int flags, int *rej, cudaStream_t st,
                    cudaEvent_t *ev);
int cs_total_loglik_async(const ChainDev &D, const double *U, const int *Zrows, const double *sig, double *d_out,
                          cudaStream_t st);
int cs_rmode_run(const ChainDev &D, int lcap, int t0, int t1, int last_sweep, double *scratch, cudaStream_t st);
void cs_r_set_seed(uint32_t seed, uint32_t mt[624]);

struct cs_chain {
    int N = 0, R = 0, kcap = 0, niter_cap = 0, rng_mode = 0, flags = 0;
    int sweeps_done = 0;
    bool inited = false;
    cudaStream_t st = nullptr;
    ChainDev D{};
    DevBuf<double> X, Xc, Ua, Ub, sig, isig, Q, E, qmin, ua, mean, psum, sigma_all, LL, u_rows, L0;
    DevBuf<long long> qnew_i, tile_sums, wn_q;
    DevBuf<double> wn_w;
    DevBuf<int> kmin, icnt, icur;
    DevBuf<unsigned> near, fx_bits;
    DevBuf<int> zh1, fx_td, fx_tp, fx_scal, jcnt;
    DevBuf<InvEntry> inv;
    DevBuf<int> Zs, np, np0, slot_label, J, zspec, qcols, pcnt, newslot, rej, Zall, kt_all, u_labels;
    DevBuf<long long> u_off;
    DevBuf<uint32_t> mt;
    DevBuf<ChainScal> scal;
    DevBuf<int> zt;                 // cs_chain_fetch: Zall transposed into R's layout (kept between calls)
    // timing of the last cs_chain_run: whole run, and (flags bit 1) per-sweep kernel groups
    cudaEvent_t run_ev[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> sweep_ev;
    int prof_sweeps = 0;
    ~cs_chain()
    {
        for (auto e : sweep_ev) cudaEventDestroy(e);
        for (auto e : run_ev)
            if (e) cudaEventDestroy(e);
        if (st) cudaStreamDestroy(st);
    }
};

static const int kDefaultKcapPhilox = 256, kDefaultLcapR = 4096;
static thread_local cs_chain *g_cached = nullptr; // see cs_ncrp_gibbs
static thread_local int g_cached_dev = -1;

extern "C" int cs_chain_create(int N, int R, int kcap, int niter_cap, int rng_mode, cs_chain **out)
{
    CS_REQUIRE(out != nullptr, "cs_chain_create: NULL output");
    *out = nullptr;
    CS_REQUIRE(N >= 1 && R >= 1 && niter_cap >= 1, "cs_chain_create: bad dimensions N=%d R=%d niter=%d", N, R, niter_cap);
    CS_REQUIRE(rng_mode != CS_RNG_PHILOX || N < (1 << 22), "cs_chain_create: the Philox sampler takes up to 4,194,303 cells per chain (N=%d)", N);
    CS_REQUIRE(rng_mode == CS_RNG_R || rng_mode == CS_RNG_PHILOX, "cs_chain_create: unknown rng_mode %d", rng_mode);
    if (kcap <= 0) kcap = rng_mode == CS_RNG_PHILOX ? kDefaultKcapPhilox : kDefaultLcapR;
    if (rng_mode == CS_RNG_PHILOX && R > 1024) {
        cs_set_error("Philox sampler supports up to 1024 segments (R=%d)", R);
        return CS_ERR_UNSUPPORTED;
    }
    cs_chain *c = new cs_chain();
    struct Guard {
        cs_chain *c;
        ~Guard() { delete c; }
    } guard{c};
    c->N = N;
    c->R = R;
    c->kcap = kcap;
    c->niter_cap = niter_cap;
    c->rng_mode = rng_mode;
    CS_CUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    CS_CUDA(cudaEventCreate(&c->run_ev[0]));
    CS_CUDA(cudaEventCreate(&c->run_ev[1]));
    const size_t nx = (size_t)N * R;
    const int chunk = cs_stat_chunk(N);
    const int nchunks = (N + chunk - 1) / chunk;
    const bool ph = rng_mode == CS_RNG_PHILOX;
    // rows of Uall recorded on the device: a sweep records at most min(kcap, N) occupied clusters
    // (R mode: kcap is the label capacity, far above what a sweep holds; 512 per sweep on average)
    const long long per_sweep = ph ? (kcap < N ? kcap : N) : (N < 512 ? N : 512);
    const long long urec_cap = (long long)niter_cap * per_sweep + kcap;
    CS_CUDA(c->X.alloc(nx));
    CS_CUDA(c->Xc.alloc(nx));
    CS_CUDA(c->Ua.alloc((size_t)kcap * R));
    CS_CUDA(c->Ub.alloc((size_t)kcap * R));
    CS_CUDA(c->sig.alloc(R));
    CS_CUDA(c->isig.alloc(R));
    CS_CUDA(c->mean.alloc((size_t)kcap * R));
    CS_CUDA(c->Zs.alloc((size_t)N + 8)); // (+8 here and below: the ring scan copies whole groups of four cells)
    CS_CUDA(c->np.alloc(kcap));
    CS_CUDA(c->np0.alloc(kcap));
    CS_CUDA(cudaMemset(c->np0.p, 0, sizeof(int) * (size_t)kcap));
    CS_CUDA(c->slot_label.alloc(kcap));
    CS_CUDA(c->newslot.alloc(kcap));
    CS_CUDA(c->scal.alloc(1));
    CS_CUDA(c->mt.alloc(624));
    CS_CUDA(c->L0.alloc(1));
    CS_CUDA(c->psum.alloc(ph ? (size_t)nchunks * kcap * R : (size_t)nchunks + 16));
    CS_CUDA(c->pcnt.alloc(ph ? (size_t)nchunks * kcap * R : 16));
    if (ph) {
        CS_CUDA(c->Q.alloc((size_t)N * kcap));
        CS_CUDA(c->E.alloc((size_t)N * kcap + 8));
        CS_CUDA(c->qmin.alloc((size_t)N + 8));
        CS_CUDA(c->ua.alloc((size_t)N + 8));
        CS_CUDA(c->qnew_i.alloc((size_t)N + 8));
        CS_CUDA(c->wn_q.alloc((size_t)N + 8));
        CS_CUDA(c->wn_w.alloc((size_t)N + 8));
        CS_CUDA(c->kmin.alloc((size_t)N + 8));
        CS_CUDA(c->icnt.alloc((size_t)N + 8));
        CS_CUDA(c->icur.alloc(N));
        CS_CUDA(c->near.alloc(((size_t)N + 8) * CS_NEAR_CAP));
        CS_CUDA(c->inv.alloc(nx));
        CS_CUDA(c->tile_sums.alloc((size_t)N / 1024 + 8 + 1024)); // (also the fixed-point walk's optional per-step timings)
        CS_CUDA(c->J.alloc(nx + 256)); // (the fixed-point walk copies whole 128-entry steps)
        CS_CUDA(c->zspec.alloc(N));
        CS_CUDA(c->qcols.alloc(N));
        CS_CUDA(c->rej.alloc(N));
        const size_t tiles = ((size_t)N + 255) / 256, tpad = (tiles + 31) & ~(size_t)31;
        CS_CUDA(c->zh1.alloc(N));
        CS_CUDA(c->jcnt.alloc((size_t)N + 32));
        CS_CUDA(c->fx_bits.alloc(2 * tiles * 8));
        CS_CUDA(c->fx_td.alloc(((size_t)kcap + 1) * tpad));
        CS_CUDA(c->fx_tp.alloc(((size_t)kcap + 1) * tpad));
        CS_CUDA(c->fx_scal.alloc(8));
        CS_CUDA(cudaMemset(c->fx_td.p, 0, sizeof(int) * ((size_t)kcap + 1) * tpad));
        CS_CUDA(cudaMemset(c->fx_tp.p, 0, sizeof(int) * ((size_t)kcap + 1) * tpad));
    }
    CS_CUDA(c->Zall.alloc((size_t)niter_cap * N));
    CS_CUDA(c->sigma_all.alloc((size_t)niter_cap * R));
    CS_CUDA(c->LL.alloc(niter_cap));
    CS_CUDA(c->kt_all.alloc(niter_cap));
    CS_CUDA(c->u_off.alloc((size_t)niter_cap + 1));
    CS_CUDA(c->u_labels.alloc((size_t)urec_cap));
    CS_CUDA(c->u_rows.alloc((size_t)urec_cap * R));
    ChainDev &D = c->D;
    D.N = N; D.R = R; D.kcap = kcap; D.chunk = chunk;
    D.X = c->X.p; D.Zs = c->Zs.p; D.U = c->Ua.p; D.Unext = c->Ub.p; D.np = c->np.p; D.np0 = c->np0.p;
    D.slot_label = c->slot_label.p; D.sig = c->sig.p; D.isig = c->isig.p; D.Q = c->Q.p; D.J = c->J.p;
    D.E = c->E.p; D.qmin = c->qmin.p; D.ua = c->ua.p; D.qnew_i = c->qnew_i.p; D.icnt = c->icnt.p;
    D.wn_q = c->wn_q.p; D.wn_w = c->wn_w.p; D.kmin = c->kmin.p;
    D.icur = c->icur.p; D.inv = c->inv.p; D.near = c->near.p; D.tile_sums = c->tile_sums.p;
    D.jcnt = c->jcnt.p; D.zh1 = c->zh1.p; D.fx_bits = c->fx_bits.p; D.fx_td = c->fx_td.p; D.fx_tp = c->fx_tp.p; D.fx_scal = c->fx_scal.p;
    D.zspec = c->zspec.p; D.qcols = c->qcols.p; D.mean = c->mean.p; D.psum = c->psum.p; D.pcnt = c->pcnt.p;
    D.newslot = c->newslot.p; D.mt = c->mt.p; D.scal = c->scal.p;
    D.Zall = c->Zall.p; D.sigma_all = c->sigma_all.p; D.LL = c->LL.p; D.kt_all = c->kt_all.p;
    D.u_off = c->u_off.p; D.u_labels = c->u_labels.p; D.u_rows = c->u_rows.p; D.urec_cap = urec_cap;
    guard.c = nullptr;
    *out = c;
    return CS_OK;
}

extern "C" void cs_chain_destroy(cs_chain *c) { delete c; }

static __global__ void isig_kernel(int R, const double *sig, double *isig)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) isig[r] = 1.0 / sig[r];
}
static __global__ void fill_kernel(size_t n, double *p, double v)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" int cs_chain_init(cs_chain *c, const double *X, int K0, const double *U0, const int *Z0,
                             const double *sigmas0, double alpha, double beta, uint32_t seed, int flags)
{
    CS_REQUIRE(c && X && U0 && Z0 && sigmas0, "cs_chain_init: NULL argument");
    const int N = c->N, R = c->R;
    const bool ph = c->rng_mode == CS_RNG_PHILOX;
    CS_REQUIRE(K0 >= 1, "cs_chain_init: K0=%d", K0);
    // NA rows of U0 (first column NA, as :93 tests) keep their label and hold no cell: the
    // reference never scores (:193 which(npoints != 0)) or proposes (:170 prob 0) them.  The
    // Philox path gives them no slot; the R-compatible path keeps them as empty labels.
    std::vector<int> slot_of(K0, -1);
    int nrows_live = 0;
    for (int k = 0; k < K0; k++)
        if (!(U0[k] != U0[k])) slot_of[k] = nrows_live++;
    CS_REQUIRE(nrows_live >= 1, "cs_chain_init: every row of U0 is NA");
    CS_REQUIRE((ph ? nrows_live : K0) <= c->kcap, "cs_chain_init: %d rows of U0 exceed the capacity kcap=%d",
               ph ? nrows_live : K0, c->kcap);
    // the reference indexes U0[Z0,] (:133) and table(Zt)[1:nrow(U0)] (:150): labels outside
    // 1..K0 give NA counts and an error further down
    for (int i = 0; i < N; i++) {
        if (Z0[i] < 1 || Z0[i] > K0) {
            cs_set_error("Z0[%d]=%d is not a row of U0 (1..%d)", i + 1, Z0[i], K0);
            return CS_ERR_STATE;
        }
        if (slot_of[Z0[i] - 1] < 0) {
            cs_set_error("Z0[%d]=%d names an NA row of U0", i + 1, Z0[i]);
            return CS_ERR_STATE;
        }
    }
    for (int r = 0; r < R; r++) CS_REQUIRE(sigmas0[r] > 0.0, "sigmas0[%d]=%g must be positive", r + 1, sigmas0[r]);
    bool single = true;
    for (int i = 1; i < N && single; i++) single = Z0[i] == Z0[0];
    if (single && Z0[0] != 1) {
        cs_set_error("all cells start in cluster %d != 1: the reference's npoints would be NA (R/BayesNonparCluster.R:141-151)", Z0[0]);
        return CS_ERR_STATE;
    }
    cudaStream_t st = c->st;
    ChainDev &D = c->D;
    D.U = c->Ua.p;
    D.Unext = c->Ub.p;
    D.alpha = alpha;
    D.beta = beta;
    D.seed = seed;
    D.single = single ? 1 : 0;
    c->flags = flags;
    c->sweeps_done = 0;
    // X: column-major N x R (== row-major R x N) -> cell-major
    CS_CUDA(cudaStreamSynchronize(st));
    int rc = cs_h2d_staged(c->Xc.p, X, sizeof(double) * (size_t)N * R);
    if (rc) return rc;
    rc = cs_transpose<double>(R, N, c->Xc.p, c->X.p, st);
    if (rc) return rc;
    // U0: column-major K0 x R -> rows (Philox: the non-NA rows, in label order, one slot each)
    const int nrows_up = ph ? nrows_live : K0;
    std::vector<double> Urows((size_t)nrows_up * R);
    for (int k = 0; k < K0; k++) {
        const int row = ph ? slot_of[k] : k;
        if (row < 0) continue;
        for (int r = 0; r < R; r++) Urows[(size_t)row * R + r] = U0[(size_t)r * K0 + k];
    }
    CS_CUDA(cudaMemcpyAsync(D.U, Urows.data(), sizeof(double) * Urows.size(), cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaMemcpyAsync(D.sig, sigmas0, sizeof(double) * R, cudaMemcpyHostToDevice, st));
    isig_kernel<<<(R + 127) / 128, 128, 0, st>>>(R, D.sig, D.isig);
    // L0 at (U0, Z0, sigmas0)  (:133) — before the single-cluster replacement of U0
    std::vector<int> zrows(N);
    for (int i = 0; i < N; i++) zrows[i] = ph ? slot_of[Z0[i] - 1] : Z0[i] - 1;
    CS_CUDA(cudaMemcpyAsync(D.Zs, zrows.data(), sizeof(int) * N, cudaMemcpyHostToDevice, st));
    rc = cs_total_loglik_async(D, D.U, D.Zs, D.sig, c->L0.p, st);
    if (rc) return rc;
    CS_CUDA(cudaStreamSynchronize(st)); // Urows / zrows are reused below
    // state (:141-151)
    std::vector<int> np(c->kcap, 0), lab(c->kcap, 0);
    int kt, nlive;
    for (int k = 0; k < c->kcap; k++) lab[k] = k + 1;
    if (single) {
        kt = nlive = 1;
        np[0] = N;
        fill_kernel<<<(R + 127) / 128, 128, 0, st>>>((size_t)R, D.U, 1.0);
    } else {
        kt = K0;
        nlive = ph ? nrows_live : K0;
        for (int i = 0; i < N; i++) np[zrows[i]]++;
        if (ph) { // slots past the initial ones are labelled as they open (label = ++kt)
            for (int k = 0; k < K0; k++)
                if (slot_of[k] >= 0) lab[slot_of[k]] = k + 1;
        }
    }
    std::vector<int> zinit(N);
    for (int i = 0; i < N; i++) zinit[i] = ph ? (single ? 0 : zrows[i]) : Z0[i];
    ChainScal h{};
    h.nlive = nlive;
    h.kt = kt;
    h.first_event = INT_MAX;
    h.mt_pos = 624;
    CS_CUDA(cudaMemcpyAsync(D.Zs, zinit.data(), sizeof(int) * N, cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaMemcpyAsync(D.np, np.data(), sizeof(int) * c->kcap, cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaMemcpyAsync(D.slot_label, lab.data(), sizeof(int) * c->kcap, cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaMemcpyAsync(D.scal, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    uint32_t mt[624];
    cs_r_set_seed(seed, mt);
    CS_CUDA(cudaMemcpyAsync(D.mt, mt, sizeof(mt), cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaStreamSynchronize(st));
    CS_CUDA(cudaGetLastError());
    c->inited = true;
    return CS_OK;
}

extern "C" int cs_chain_run(cs_chain *c, int nsweeps, int is_last, int sync)
{
    CS_REQUIRE(c && c->inited, "cs_chain_run: chain not initialised");
    CS_REQUIRE(nsweeps >= 0 && c->sweeps_done + nsweeps <= c->niter_cap,
               "cs_chain_run: %d more sweeps exceed niter_cap=%d (done %d)", nsweeps, c->niter_cap, c->sweeps_done);
    const int t0 = c->sweeps_done, t1 = t0 + nsweeps;
    int rc = CS_OK;
    const bool prof = (c->flags & 2) != 0 && c->rng_mode == CS_RNG_PHILOX;
    if (prof) {
        while ((int)c->sweep_ev.size() < 5 * nsweeps) {
            cudaEvent_t e;
            CS_CUDA(cudaEventCreate(&e));
            c->sweep_ev.push_back(e);
        }
    }
    c->prof_sweeps = prof ? nsweeps : 0;
    CS_CUDA(cudaEventRecord(c->run_ev[0], c->st));
    if (c->rng_mode == CS_RNG_R) {
        if (nsweeps > 0) rc = cs_rmode_run(c->D, c->kcap, t0, t1, is_last ? t1 - 1 : -1, c->mean.p, c->st);
    } else {
        for (int tt = t0; tt < t1 && rc == CS_OK; tt++) {
            rc = cs_philox_sweep(c->D, tt, is_last && tt == t1 - 1, c->flags, c->rej.p, c->st,
                                 prof ? &c->sweep_ev[5 * (size_t)(tt - t0)] : nullptr);
            double *t = c->D.U;
            c->D.U = c->D.Unext;
            c->D.Unext = t;
        }
    }
    if (rc) return rc;
    CS_CUDA(cudaEventRecord(c->run_ev[1], c->st));
    c->sweeps_done = t1;
    if (sync) return cs_chain_sync(c);
    return CS_OK;
}

extern "C" int cs_chain_sync(cs_chain *c)
{
    CS_REQUIRE(c != nullptr, "cs_chain_sync: NULL chain");
    CS_CUDA(cudaStreamSynchronize(c->st));
    ChainScal h;
    CS_CUDA(cudaMemcpy(&h, c->D.scal, sizeof(h), cudaMemcpyDeviceToHost));
    if (h.err == CS_ERR_KCAP) {
        cs_set_error("more clusters than capacity kcap=%d (raise cs_gibbs_opts.kcap)", c->kcap);
        return CS_ERR_KCAP;
    }
    if (h.err == CS_ERR_UCAP) {
        cs_set_error("device Uall record capacity exceeded (%lld rows; a chain that keeps more than ~512 occupied "
                     "clusters per sweep in R-compatible mode: run it in shorter pieces with cs_chain_run / cs_chain_fetch)",
                     c->D.urec_cap);
        return CS_ERR_UCAP;
    }
    if (h.err) {
        cs_set_error("device-side error %d", h.err);
        return h.err;
    }
    return CS_OK;
}

extern "C" int cs_chain_sweeps_done(cs_chain *c) { return c ? c->sweeps_done : -1; }

extern "C" int cs_chain_timing(cs_chain *c, double *run_ms, double *prepare_ms, double *index_ms, double *scan_ms,
                               double *eos_ms)
{
    CS_REQUIRE(c != nullptr, "cs_chain_timing: NULL chain");
    CS_CUDA(cudaStreamSynchronize(c->st));
    float ms = 0.f;
    if (run_ms) {
        CS_CUDA(cudaEventElapsedTime(&ms, c->run_ev[0], c->run_ev[1]));
        *run_ms = ms;
    }
    double a = 0, ix = 0, b = 0, d = 0;
    for (int s = 0; s < c->prof_sweeps; s++) {
        cudaEvent_t *e = &c->sweep_ev[5 * (size_t)s];
        CS_CUDA(cudaEventElapsedTime(&ms, e[0], e[1]));
        a += ms;
        // the sequential kernel (test hook, first sweep of a single-cluster start) builds no index
        const bool has_index = !(c->flags & 1) && !(c->D.single && c->sweeps_done - c->prof_sweeps + s == 0);
        if (has_index) {
            CS_CUDA(cudaEventElapsedTime(&ms, e[1], e[4]));
            ix += ms;
            CS_CUDA(cudaEventElapsedTime(&ms, e[4], e[2]));
            b += ms;
        } else {
            CS_CUDA(cudaEventElapsedTime(&ms, e[1], e[2]));
            b += ms;
        }
        CS_CUDA(cudaEventElapsedTime(&ms, e[2], e[3]));
        d += ms;
    }
    if (prepare_ms) *prepare_ms = a;
    if (index_ms) *index_ms = ix;
    if (scan_ms) *scan_ms = b;
    if (eos_ms) *eos_ms = d;
    return CS_OK;
}

extern "C" int cs_chain_stats(cs_chain *c, int64_t *cells_scanned, int64_t *n_events, int64_t *n_reject)
{
    CS_REQUIRE(c != nullptr, "cs_chain_stats: NULL chain");
    CS_CUDA(cudaStreamSynchronize(c->st));
    ChainScal h;
    CS_CUDA(cudaMemcpy(&h, c->D.scal, sizeof(h), cudaMemcpyDeviceToHost));
    if (cells_scanned) *cells_scanned = h.cells_scanned;
    if (n_events) *n_events = h.n_events;
    if (n_reject) *n_reject = h.n_reject;
    return CS_OK;
}

extern "C" int cs_chain_walk_stats(cs_chain *c, int64_t *rounds, int64_t *cut_margin, int64_t *cut_open, int64_t *cut_list)
{
    CS_REQUIRE(c != nullptr, "cs_chain_walk_stats: NULL chain");
    CS_CUDA(cudaStreamSynchronize(c->st));
    ChainScal h;
    CS_CUDA(cudaMemcpy(&h, c->D.scal, sizeof(h), cudaMemcpyDeviceToHost));
    if (rounds) *rounds = h.n_rounds;
    if (cut_margin) *cut_margin = h.cut_margin;
    if (cut_open) *cut_open = h.cut_open;
    if (cut_list) *cut_list = h.cut_maxev;
    return CS_OK;
}

extern "C" int cs_chain_fix_prof(cs_chain *c, uint64_t *out1024)
{
    CS_REQUIRE(c && out1024, "cs_chain_fix_prof: NULL argument");
    CS_CUDA(cudaStreamSynchronize(c->st));
    CS_CUDA(cudaMemcpy(out1024, c->tile_sums.p, sizeof(uint64_t) * 1024, cudaMemcpyDeviceToHost));
    return CS_OK;
}

extern "C" int cs_chain_fetch(cs_chain *c, int *Zall, double *sigma_all, double *LL, double *L0, int *kt_all,
                              int64_t ucap_rows, int64_t *u_off, int *u_labels, double *u_rows)
{
    CS_REQUIRE(c && c->inited, "cs_chain_fetch: chain not initialised");
    int rc = cs_chain_sync(c);
    if (rc) return rc;
    const int T = c->sweeps_done, N = c->N, R = c->R;
    if (L0) CS_CUDA(cudaMemcpy(L0, c->L0.p, sizeof(double), cudaMemcpyDeviceToHost));
    if (T == 0) return CS_OK;
    if (Zall) { // device sweep-major T x N (row-major) -> R's column-major T x N (== row-major N x T)
        CS_CUDA(c->zt.reserve((size_t)c->niter_cap * N));
        rc = cs_transpose<int>(T, N, c->Zall.p, c->zt.p, c->st);
        if (rc) return rc;
        CS_CUDA(cudaStreamSynchronize(c->st));
        if ((rc = cs_d2h_staged(Zall, c->zt.p, sizeof(int) * (size_t)T * N))) return rc;
    }
    // sigma_all: device T x R row-major == column-major R x T: column t = sweep t
    if (sigma_all) CS_CUDA(cudaMemcpy(sigma_all, c->sigma_all.p, sizeof(double) * (size_t)T * R, cudaMemcpyDeviceToHost));
    if (LL) CS_CUDA(cudaMemcpy(LL, c->LL.p, sizeof(double) * T, cudaMemcpyDeviceToHost));
    if (kt_all) CS_CUDA(cudaMemcpy(kt_all, c->kt_all.p, sizeof(int) * T, cudaMemcpyDeviceToHost));
    if (u_off) {
        std::vector<long long> off((size_t)T + 1);
        CS_CUDA(cudaMemcpy(off.data(), c->u_off.p, sizeof(long long) * off.size(), cudaMemcpyDeviceToHost));
        for (int t = 0; t <= T; t++) u_off[t] = off[t];
        const long long need = off[T];
        if (need > ucap_rows) {
            cs_set_error("ucap_rows=%lld too small: %lld rows recorded", (long long)ucap_rows, need);
            return CS_ERR_UCAP;
        }
        if (u_labels && need) CS_CUDA(cudaMemcpy(u_labels, c->u_labels.p, sizeof(int) * (size_t)need, cudaMemcpyDeviceToHost));
        if (u_rows && need && (rc = cs_d2h_staged(u_rows, c->u_rows.p, sizeof(double) * (size_t)need * R))) return rc;
    }
    return CS_OK;
}

extern "C" int cs_ncrp_gibbs(int N, int R, const double *X, int K0, const double *U0, const int *Z0,
                             const double *sigmas0, double alpha, double beta, int niter, uint32_t seed,
                             const cs_gibbs_opts *opts, int *Zall, double *sigma_all, double *LL, double *L0,
                             int *kt_all, int64_t ucap_rows, int64_t *u_off, int *u_labels, double *u_rows,
                             int64_t *n_reject, int32_t *rng_state_out)
{
    cs_gibbs_opts o;
    memset(&o, 0, sizeof(o));
    o.rng_mode = CS_RNG_PHILOX;
    o.device = -1;
    if (opts) o = *opts;
    CS_REQUIRE(niter >= 1, "niter=%d must be >= 1", niter);
    struct DeviceRestore { // the caller's current device is left as it was found
        int prev = -1;
        ~DeviceRestore()
        {
            if (prev >= 0) cudaSetDevice(prev);
        }
    } restore;
    if (o.device >= 0) {
        int cur = -1;
        CS_CUDA(cudaGetDevice(&cur));
        if (cur != o.device) {
            CS_CUDA(cudaSetDevice(o.device));
            restore.prev = cur;
        }
    }
    int kcap = o.kcap;
    if (kcap > 0 && kcap < K0) kcap = K0;
    if (kcap <= 0) kcap = o.rng_mode == CS_RNG_PHILOX ? kDefaultKcapPhilox : kDefaultLcapR;
    int dev = 0;
    CS_CUDA(cudaGetDevice(&dev));
    // the device buffers of one chain (~1.5 GB at 100k x 300) are kept between calls of the same
    // shape: RunCovCluster calls the sampler up to ten times on same-sized inputs
    cs_chain *c = g_cached;
    g_cached = nullptr;
    int rc;
    if (!(c && g_cached_dev == dev && c->N == N && c->R == R && c->kcap == kcap && c->niter_cap == niter &&
          c->rng_mode == o.rng_mode)) {
        cs_chain_destroy(c);
        c = nullptr;
        if ((rc = cs_chain_create(N, R, kcap, niter, o.rng_mode, &c))) return rc;
    }
    struct Guard { // an error drops the chain; success hands it back to the cache
        cs_chain *c;
        int dev;
        bool ok = false;
        ~Guard()
        {
            if (ok) {
                g_cached = c;
                g_cached_dev = dev;
            } else {
                cs_chain_destroy(c);
            }
        }
    } guard{c, dev};
    const bool e2e_prof = getenv("CS_E2E_PROF") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
    auto t0 = now();
    if ((rc = cs_chain_init(c, X, K0, U0, Z0, sigmas0, alpha, beta, seed, o.flags))) return rc;
    const double t_init = ms_since(t0);
    t0 = now();
    if ((rc = cs_chain_run(c, niter, 1, 1))) return rc;
    const double t_run = ms_since(t0);
    t0 = now();
    if ((rc = cs_chain_fetch(c, Zall, sigma_all, LL, L0, kt_all, ucap_rows, u_off, u_labels, u_rows))) return rc;
    if (e2e_prof) fprintf(stderr, "cs_ncrp_gibbs: init %.1f ms  run %.1f ms  fetch %.1f ms\n", t_init, t_run, ms_since(t0));
    if (n_reject) {
        int64_t sc, ev;
        if ((rc = cs_chain_stats(c, &sc, &ev, n_reject))) return rc;
    }
    if (rng_state_out && o.rng_mode == CS_RNG_R) { // .Random.seed[-1]: position, then the 624 state words
        ChainScal h;
        CS_CUDA(cudaMemcpy(&h, c->D.scal, sizeof(h), cudaMemcpyDeviceToHost));
        rng_state_out[0] = h.mt_pos;
        CS_CUDA(cudaMemcpy(rng_state_out + 1, c->D.mt, sizeof(uint32_t) * 624, cudaMemcpyDeviceToHost));
    }
    guard.ok = true;
    return CS_OK;
}

extern "C" void cs_release_cached(void)
{
    cs_chain_destroy(g_cached);
    g_cached = nullptr;
}
