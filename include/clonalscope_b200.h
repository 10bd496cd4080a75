/*
 * clonalscope_b200.h — C ABI of libclonalscope_b200.so
 *
 * B200-native (sm_100a) implementation of Clonalscope's cell-clustering hot
 * path.  This header is the drop-in boundary: the functions below are what an
 * R `.Call` shim (or cgo/ctypes/JNI — plain pointers and sizes only) binds in
 * place of the reference's pure-R bodies.  Citations are file:line inside the
 * reference repository (seasoncloud/Clonalscope).
 *
 * Conventions
 *   - Every entry point returns an int status (CS_OK == 0).  No C++ exception
 *     crosses the boundary; cs_last_error() returns a thread-local message.
 *   - Host-pointer entry points take R's own memory layout: matrices are
 *     COLUMN-major doubles/ints, NA_real_ is any NaN, labels are 1-based.
 *     They copy to the device, run the CUDA path and copy back; there is no
 *     CPU fallback — without a CUDA device they fail with CS_ERR_CUDA.
 *   - `_dev` entry points take device pointers in the library's device layout
 *     (cell-major, documented per function) and run on the given stream
 *     (a cudaStream_t passed as void*; NULL = default stream).
 */
#ifndef CLONALSCOPE_B200_H
#define CLONALSCOPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    CS_OK = 0,
    CS_ERR_ARG = 1,         /* invalid argument (message says which) */
    CS_ERR_CUDA = 2,        /* CUDA runtime error / no device */
    CS_ERR_KCAP = 3,        /* more live clusters than opts.kcap / label capacity */
    CS_ERR_UCAP = 4,        /* ucap_rows too small; u_off[niter] holds the need */
    CS_ERR_UNSUPPORTED = 5, /* shape outside what the kernels support */
    CS_ERR_STATE = 6        /* reference would have errored (e.g. NA counts) */
};

enum { CS_RNG_R = 0, CS_RNG_PHILOX = 1 };

int cs_version(void);
const char *cs_last_error(void);
int cs_device_count(void);          /* >= 0, or -1 on CUDA failure */
int cs_set_device(int device);

/* ------------------------------------------------------------------------
 * (1) gene -> bin aggregation
 * replaces the hot loop of Gen_bin_cell_rna_filtered
 *   R/Gene_bin_cell_rna_filtered.R:63-81  (findOverlaps table -> per-bin column sums
 *   -> Matrix(sparse=TRUE));  caller R/CreateSegtableNoWGS.R:53-58.
 * Input  : gene x cell counts as dgCMatrix slots (@p colptr, @i rowidx, @x).
 * Map    : the (bin, gene) overlap table as CSR over genes (map_ptr[G+1], map_bin),
 *          bins 0-based in the sorted order of :30-37; built host-side (integer
 *          interval logic, see clonalscope_b200/intervals.py).
 * Output : bin x cell dgCMatrix slots, ascending rows per column, zero sums dropped.
 * Two-call protocol because R must allocate the result: begin() runs the kernels
 * and reports nnz; fetch() copies into caller-allocated slots and frees the job.
 * ---------------------------------------------------------------------- */
typedef struct cs_bin_job cs_bin_job;
int cs_bin_counts_begin(int G, int N, int B,
                        const int32_t *colptr, const int32_t *rowidx, const double *x,
                        const int32_t *map_ptr, const int32_t *map_bin,
                        cs_bin_job **job, int64_t *nnz_out);
int cs_bin_counts_fetch(cs_bin_job *job, int32_t *out_colptr, int32_t *out_rowidx,
                        double *out_x);
void cs_bin_counts_free(cs_bin_job *job);

/* Device-resident form.  A plan owns the uploaded gene->bin map and scratch. */
typedef struct cs_bin_plan cs_bin_plan;
int cs_bin_plan_create(int G, int B, const int32_t *map_ptr, const int32_t *map_bin,
                       cs_bin_plan **plan);
void cs_bin_plan_destroy(cs_bin_plan *plan);
/* device time (CUDA events on the call's stream) of the count, scan and fill kernels of the
 * last completed cs_bin_counts_dev call that produced output */
int cs_bin_plan_timing(cs_bin_plan *plan, double *count_ms, double *scan_ms, double *fill_ms);
/* d_colptr/d_out_colptr: int64[N+1]; d_rowidx/d_out_rowidx: int32; d_x/d_out_x: double.
 * Pass d_out_rowidx == NULL to count only (fills d_out_colptr, *nnz_out).
 * With outputs given, out_cap is their capacity in entries; CS_ERR_UCAP if too small
 * (d_out_colptr and *nnz_out are still valid).  Synchronises the stream once. */
int cs_bin_counts_dev(cs_bin_plan *plan, int N,
                      const int64_t *d_colptr, const int32_t *d_rowidx, const double *d_x,
                      int64_t *d_out_colptr, int32_t *d_out_rowidx, double *d_out_x,
                      int64_t out_cap, int64_t *nnz_out, void *stream);

/* ------------------------------------------------------------------------
 * (2a) fixed-state cell x cluster Gaussian log-likelihood matrix
 *   LL[i,k] = sum_r dnorm(X[i,r], U[k,r], sigma[r], log=TRUE), NA segments skipped
 * replaces R/BayesNonparCluster.R:91-96 (Z0 initialisation), :193 (per-sweep
 * scoring) and R/AssignCluster.R:58 (re-scoring).  X: N x R, U: K x R,
 * LL: N x K, all column-major (R matrices).
 * ---------------------------------------------------------------------- */
int cs_loglik_matrix(int N, int R, int K, const double *X, const double *U,
                     const double *sigma, double *LL);
/* device layout: d_X[i*R+r], d_U[k*R+r], d_LL[i*K+k] */
int cs_loglik_matrix_dev(int N, int R, int K, const double *d_X, const double *d_U,
                         const double *d_sigma, double *d_LL, void *stream);

/* ------------------------------------------------------------------------
 * (2b) nested-CRP collapsed Gibbs sampler — replaces the loop body of
 *   BayesNonparCluster  R/BayesNonparCluster.R:133-277
 * (argument defaulting, U0 truncation, dimnames and messages stay in the shim).
 * The reference's sequential within-sweep order and its quirks (proposal counts
 * include the current cell :170 vs :188; first-sweep rejection :174; last sweep
 * overwritten noise-free :260-277; clusters never renumbered) are preserved.
 * rng_mode CS_RNG_R consumes base R's Mersenne-Twister stream in the reference's
 * order (reproduces the shipped fixtures); CS_RNG_PHILOX is the fast counter-based
 * stream (same model and order, speculative-parallel evaluation with exact
 * sequential semantics).
 * ---------------------------------------------------------------------- */
typedef struct {
    int rng_mode;   /* CS_RNG_R | CS_RNG_PHILOX */
    int device;     /* CUDA ordinal, -1 = current */
    int kcap;       /* Philox: max live clusters within a sweep (0 = 256);
                       R mode: max labels ever created (0 = 4096) */
    int flags;      /* bit 0: disable speculative evaluation (scan every cell);
                       bit 1: record per-kernel-group CUDA events (cs_chain_timing) */
    int reserved[4];
} cs_gibbs_opts;

int cs_ncrp_gibbs(int N, int R, const double *X /* N x R col-major */,
                  int K0, const double *U0 /* K0 x R col-major */, const int *Z0,
                  const double *sigmas0 /* R */, double alpha, double beta,
                  int niter, uint32_t seed, const cs_gibbs_opts *opts,
                  int *Zall,          /* niter x N col-major (results$Zall) */
                  double *sigma_all,  /* R x niter col-major: column t = sigma_all[[t]] */
                  double *LL,         /* niter (results$Likelihood) */
                  double *L0,         /* priors$L0 */
                  int *kt_all,        /* niter: nrow(Uall[[t]]) */
                  int64_t ucap_rows,  /* capacity (rows) of u_labels / u_rows */
                  int64_t *u_off,     /* niter+1 */
                  int *u_labels,      /* labels of the non-NA rows of Uall[[t]] */
                  double *u_rows,     /* those rows, row-major rows x R */
                  int64_t *n_reject,  /* rejected proposals (may be NULL) */
                  int32_t *rng_state_out /* R mode: 625 ints = .Random.seed[-1] after the
                                            run; may be NULL */);

/* cs_ncrp_gibbs keeps the device buffers of its last chain for a same-shaped next call
 * (per host thread); this frees them. */
void cs_release_cached(void);

/* Device-resident chain (what bench.py times as `value`). */
typedef struct cs_chain cs_chain;
int cs_chain_create(int N, int R, int kcap, int niter_cap, int rng_mode, cs_chain **c);
/* X col-major host; uploads, transposes to cell-major, computes L0, initialises state */
int cs_chain_init(cs_chain *c, const double *X, int K0, const double *U0, const int *Z0,
                  const double *sigmas0, double alpha, double beta, uint32_t seed, int flags);
/* run `nsweeps` more sweeps; is_last marks the final sweep of the run (:260-277).
 * Asynchronous on the chain's stream unless sync != 0. */
int cs_chain_run(cs_chain *c, int nsweeps, int is_last, int sync);
int cs_chain_sync(cs_chain *c);
int cs_chain_sweeps_done(cs_chain *c);
/* device time of the last cs_chain_run (CUDA events on the chain's stream); with init flag
 * bit 1 set also the per-sweep totals of the prepare kernel, the inverse-index kernels, the
 * scan kernel and the end-of-sweep kernels.  Synchronises the chain. */
int cs_chain_timing(cs_chain *c, double *run_ms, double *prepare_ms, double *index_ms, double *scan_ms,
                    double *eos_ms);
/* number of cells the sequential scan kernel had to process (diagnostics) */
int cs_chain_stats(cs_chain *c, int64_t *cells_scanned, int64_t *n_events, int64_t *n_reject);
/* the walk since cs_chain_init: steps of the fixed-point iteration (rounds), clusters opened
 * inside sweeps (cut_open); the other two count what cut the rounds of the round-1 ring walk
 * (init flag bit 2) short */
int cs_chain_walk_stats(cs_chain *c, int64_t *rounds, int64_t *cut_margin, int64_t *cut_open, int64_t *cut_list);
/* diagnostics (environment CS_FIX_PROF=1): per step of the last sweep's fixed-point walk, eight
 * words — max and sum over the CTAs of the evaluation time, CTA 0's wait at the first grid
 * barrier, its scan, its wait at the second barrier (ns), the first unsettled cell, the number of
 * CTAs that had a tile, unused; 64 steps; then, per step, the sums over the CTAs of the time
 * before the pull, in the pull and in the draw (8 words per step again) */
int cs_chain_fix_prof(cs_chain *c, uint64_t *out1024);
int cs_chain_fetch(cs_chain *c, int *Zall, double *sigma_all, double *LL, double *L0,
                   int *kt_all, int64_t ucap_rows, int64_t *u_off, int *u_labels,
                   double *u_rows);
void cs_chain_destroy(cs_chain *c);

/* ------------------------------------------------------------------------
 * (2c) allele variant — replaces the loop body of
 *   BayesNonparAlleleCluster  R/BayesNonparAlleleCluster.R:101-290
 *   genotype_neighbor         R/genotype_neighbor.R:7-40
 * Parity of this path is UNPINNED by the reference (never called, no fixture).
 * ---------------------------------------------------------------------- */
int cs_ncrp_gibbs_allele(int N, int R, const double *Xcov, const double *Xallele,
                         int K0, const int *U0 /* K0 x R state ids, col-major */,
                         const int *Z0, const double *sig0_cov, const double *sig0_allele,
                         double alpha, double beta, int niter, uint32_t seed, int maxcp,
                         const cs_gibbs_opts *opts,
                         int *Zall, double *sig_cov_all, double *sig_allele_all,
                         double *LL, double *L0, int *kt_all,
                         int64_t ucap_rows, int64_t *u_off, int *u_labels,
                         int *u_state, double *u_cov, double *u_allele);

/* ------------------------------------------------------------------------
 * (3) posterior assignment tally — replaces R/AssignCluster.R:37 (and :232)
 *   maj_vote = apply(Zall, 2, function(x) which.max(table(x)[as.character(1:max(Zall))]))
 * on the MCMCtrim'ed Zall (R/MCMCtrim.R:27-36).  Zall: T x N col-major.
 * Ties go to the smallest label.
 * ---------------------------------------------------------------------- */
int cs_majority_vote(int T, int N, const int *Zall, int *zmaj);
/* element (t, i) at d_Zall[t*stride_t + i*stride_i] */
int cs_majority_vote_dev(int T, int N, const int *d_Zall, int64_t stride_t, int64_t stride_i,
                         int *d_zmaj, void *stream);

/* ------------------------------------------------------------------------
 * cell-sharded sufficient statistics (multi-GPU M-step):
 * per-cluster sums / non-NA counts / cluster sizes and per-segment squared
 * residuals for the cells this rank holds; the caller all-reduces the buffers.
 *   sum_x[k*R+r], cnt_x[k*R+r] (as double), n_k[k];  sq[r], cnt[r]
 * (R/BayesNonparCluster.R:221-234).  d_Z: 1-based labels in 1..K.
 * ---------------------------------------------------------------------- */
int cs_cluster_stats_dev(int N, int R, int K, const double *d_X, const int *d_Z,
                         double *d_sum_x, double *d_cnt_x, double *d_nk, void *stream);
int cs_sigma_stats_dev(int N, int R, int K, const double *d_X, const int *d_Z,
                       const double *d_U, double *d_sq, double *d_cnt, void *stream);
/* Host (R layout) forms — what AssignCluster's `.Call` stubs bind for its means and sigmas
 *   R/AssignCluster.R:46-53, :67-77 (colMeans per cluster, sqrt(1/N * colSums((X - U[Z,])^2))).
 * X: N x R, sum_x / cnt_x / U: K x R, all column-major; labels outside 1..K are skipped
 * (the cells AssignCluster has set aside); cnt[r] counts the non-NA entries of ALL cells. */
int cs_cluster_stats(int N, int R, int K, const double *X, const int *Z,
                     double *sum_x, double *cnt_x, double *n_k);
int cs_sigma_stats(int N, int R, int K, const double *X, const int *Z, const double *U,
                   double *sq, double *cnt);

/* ------------------------------------------------------------------------
 * per-gene sums and sums of squares over the cells of a G x N dgCMatrix (slots @i @x, nnz
 * entries) — what the variance filter of EstRegionCov needs (replaces the dense
 * `apply(mtx, 1, var)` / `rowSums` of R/EstRegionCov.R:42-43; the per-segment and library-size
 * column sums of :88-159 go through cs_bin_counts_* with a gene -> segment map).
 * ---------------------------------------------------------------------- */
int cs_gene_moments(int G, int64_t nnz, const int32_t *rowidx, const double *x, double *sum, double *sumsq);

/* ------------------------------------------------------------------------
 * EstRegionCov on a device-resident copy of the gene x cell matrix (uploaded once)
 *
 * cs_csc_upload: dgCMatrix slots (G genes x N cells, 0-based) -> handle.  With col_sel != NULL the
 *   handle holds the Nsel columns col_sel[0..Nsel) in that order (gathered on the device): the
 *   mtx[, match(celltype0[, 1], barcodes[, 1])] of R/EstRegionCov.R:40.
 * cs_csc_gene_moments: per-gene sum and sum of squares over the handle's cells (:42-48).
 * cs_csc_region_cov: R/EstRegionCov.R:88-159 for S segments at once.  map_ptr[G+1] / map_seg: the
 *   genes of every segment as a CSR map gene -> segments (findOverlaps on the host); lib_flag[G]:
 *   1 for the genes whose counts make up the library size (:131-139); celltype[N]: 0 normal,
 *   1 tumor, other = neither; include: 0 tumor, 1 control, 2 all.  Outputs, N x S column-major:
 *   sums (may be NULL) = per-segment column sums; deltas = the coverage change of the cells the
 *   segment reports (positive sum and of the included type), NaN elsewhere; alpha[N] = library
 *   sizes; nsel[S] = cells with a positive sum.  The medians (:141,:153) are exact.
 * ---------------------------------------------------------------------- */
typedef struct cs_csc cs_csc;
int cs_csc_upload(int G, int N, const int32_t *colptr, const int32_t *rowidx, const double *x, int Nsel,
                  const int32_t *col_sel, cs_csc **out);
void cs_csc_free(cs_csc *m);
int cs_csc_gene_moments(cs_csc *m, double *sum, double *sumsq);
int cs_csc_region_cov(cs_csc *m, int S, const int32_t *map_ptr, const int32_t *map_seg,
                      const unsigned char *lib_flag, const int *celltype, int include, double *deltas,
                      double *sums, double *alpha, int *nsel);
/* The two halves of cs_csc_region_cov with device-resident results, for cell-sharded runs: the
 * medians of :141,:153 need every cell, so the shards' sums and library sizes are all-gathered in
 * between (clonalscope_b200/distributed.py region_cov_sharded).  d_sums: S x N row-major (row s =
 * segment s), d_alpha: N; map_ptr / map_seg / lib_flag are host pointers as above. */
int cs_csc_region_sums_dev(cs_csc *m, int S, const int32_t *map_ptr, const int32_t *map_seg,
                           const unsigned char *lib_flag, double *d_sums, double *d_alpha, void *stream);
int cs_region_ratios_dev(int S, int N, const double *d_sums, const double *d_alpha, const int *d_celltype,
                         int include, double *d_deltas, int *d_nsel, void *stream);

/* ------------------------------------------------------------------------
 * the caller side of the gene -> bin aggregation (CreateSegtableNoWGS + Segmentation_bulk)
 *
 * cs_rowsum_by_label: bin x cluster pooling of a bin x cell dgCMatrix (slots as above, B bins,
 *   N cells): out[b + k*B] = sum over the cells with label[c] == k of counts[b, c]; labels
 *   outside [0, K) are skipped.  Replaces bin_by_cluster[, k] = rowSums(bin_mtx[, Zest == k])
 *   of R/CreateSegtableNoWGS.R:62-69 and ref_counts = rowSums(bin_mtx[, normal]) of :77.
 *   _dev: device pointers (int64 colptr), e.g. the output of cs_bin_counts_dev as it lies.
 * cs_segment_hmm: nseq sequences stored back to back (sequence s = val[off[s] .. off[s+1])), one
 *   per chromosome and pooled profile: centred running mean of width nmean with shrinking ends
 *   (caTools::runmean(x, k), R/Segmentation_bulk.R:152) and the Viterbi path of the 4-state
 *   Gaussian HMM (HiddenMarkov::dthmm + Viterbi, :154-166; means / sds per state, initial
 *   distribution delta, transition matrix Pi row-major).  state[i] is 0-based.  Host pointers.
 * ---------------------------------------------------------------------- */
int cs_rowsum_by_label(int B, int N, const int32_t *colptr, const int32_t *rowidx, const double *x,
                       const int *label, int K, double *out);
int cs_rowsum_by_label_dev(int B, int N, const int64_t *d_colptr, const int32_t *d_rowidx, const double *d_x,
                           const int *d_label, int K, double *d_out, void *stream);
/* the same pooling on the device-resident result of cs_bin_counts_begin (before or instead of
 * cs_bin_counts_fetch; the job stays valid and is released by fetch or free as before) */
int cs_bin_job_rowsum_by_label(cs_bin_job *job, const int *label, int K, double *out);
int cs_segment_hmm(int nseq, const int64_t *off, const double *val, int nmean, const double *means,
                   const double *sds, const double *delta, const double *Pi, double *smooth, int *state);

/* ------------------------------------------------------------------------
 * MatrixMarket ingest: Matrix::readMM + as(., "dgCMatrix") of the tutorials
 * (samples/P5931/scRNA/README.md:72; R/Gene_bin_cell_rna_filtered.R:23).  `path`: a coordinate
 * real / integer / pattern general file, plain or .gz.  begin() parses the text on host threads,
 * builds the compressed columns on the device (duplicates summed, rows ascending) and reports
 * dims = {rows, cols, nnz}; fetch() copies into caller-allocated dgCMatrix slots and frees the job.
 * ---------------------------------------------------------------------- */
typedef struct cs_mm_job cs_mm_job;
int cs_mm_read_begin(const char *path, cs_mm_job **job, int64_t *dims);
int cs_mm_read_fetch(cs_mm_job *job, int32_t *colptr, int32_t *rowidx, double *x);
void cs_mm_read_free(cs_mm_job *job);

#ifdef __cplusplus
}
#endif
#endif /* CLONALSCOPE_B200_H */
