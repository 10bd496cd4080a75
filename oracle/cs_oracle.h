/*
 * cs_oracle.h — CPU restatement of Clonalscope's cell-clustering hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is the parity checker for the CUDA
 * product path in clonalscope_b200/.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * package never links, imports or calls anything in this directory.
 *
 * Parity status: PINNED.  The R-compatible mode (rng_mode = CSO_RNG_R) is
 * checked bit-for-bit against the five nonpara_clustering*.rds objects the
 * reference ships (tests/test_oracle_golden.py; vectors in tests/golden/).
 * The allele sampler is NOT pinned (no caller / fixture in the reference).
 *
 * All matrices are ROW-major here (numpy C order): X[i*R + r], U[k*R + r].
 * NA is any NaN.  Labels are 1-based as in R.
 */
#ifndef CS_ORACLE_H
#define CS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { CSO_RNG_R = 0, CSO_RNG_PHILOX = 1 };

/* flags for cso_gibbs */
enum {
    CSO_FLAG_SORT_ONCE = 1 /* sort the proposal probability vector once per cell
                              instead of once per sample.int call; identical
                              results, used only to time a kinder CPU baseline */
};

/* ---- base-R RNG restatement ---- */
typedef struct {
    uint32_t mt[624];
    int pos;
    uint64_t ndraws;
} cso_mt;

void cso_set_seed(cso_mt *g, uint32_t seed);
double cso_unif_rand(cso_mt *g);
double cso_qnorm(double p);
double cso_rnorm(cso_mt *g, double mean, double sd);
/* sample.int(n, 1, prob = p): p is overwritten (normalised, sorted, cumulated) */
int cso_sample_int_prob(cso_mt *g, int n, double *p, int *perm);
/* sample(x, 1) with uniform probability, R >= 3.6 "Rejection" sampling */
int cso_sample_unif_index(cso_mt *g, int n);
void cso_revsort(double *a, int *ib, int n);

/* ---- Philox4x32-10 ---- */
void cso_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* ---- fixed-state log-likelihood matrix  (R/BayesNonparCluster.R:91-96,193) ---- */
void cso_loglik_matrix(int N, int R, int K, const double *X, const double *U,
                       const double *sigma, double *LL /* N*K row-major */);

/* total fixed-state log-likelihood  (R/BayesNonparCluster.R:133,245) */
double cso_total_loglik(int N, int R, const double *X, const double *U, int K,
                        const int *Z, const double *sigma);

/* ---- nested-CRP Gibbs sampler  (R/BayesNonparCluster.R:21-289) ---- */
int cso_gibbs(int N, int R, const double *X,
              int K0, const double *U0, const int *Z0,
              const double *sigmas0, double alpha, double beta,
              int niter, uint32_t seed, int rng_mode, int flags,
              int *Zall,          /* niter*N, row t = sweep t */
              double *sigma_all,  /* niter*R */
              double *LL,         /* niter */
              double *L0,         /* 1 */
              int *kt_all,        /* niter: nrow(Uall[[t]]) */
              int64_t ucap_rows,  /* capacity of u_labels/u_rows (rows) */
              int64_t *u_off,     /* niter+1 offsets into u_labels/u_rows */
              int *u_labels,      /* labels of the non-NA rows of Uall[[t]] */
              double *u_rows,     /* their values, ucap_rows*R */
              int64_t *n_reject); /* rejected proposals (first sweep) */

/* ---- allele variant  (R/BayesNonparAlleleCluster.R:20-307; parity UNPINNED) ---- */
int cso_genotype_grid(int maxcp, double *mu /* n*2 row-major */);
int cso_genotype_neighbor(double cov, double allele, int maxcp);
int cso_gibbs_allele(int N, int R, const double *Xcov, const double *Xallele,
                     int K0, const int *U0 /* state ids, K0*R */, const int *Z0,
                     const double *sig0_cov, const double *sig0_allele,
                     double alpha, double beta, int niter, uint32_t seed,
                     int maxcp, int rng_mode,
                     int *Zall, double *sig_cov_all, double *sig_allele_all,
                     double *LL, double *L0, int *kt_all,
                     int64_t ucap_rows, int64_t *u_off, int *u_labels,
                     int *u_state, double *u_cov, double *u_allele);

/* ---- posterior tally  (R/MCMCtrim.R:16-36, R/AssignCluster.R:37) ---- */
void cso_majority_vote(int T, int N, const int *Zall /* T*N row t = sweep */,
                       int *zmaj);

/* ---- gene -> bin aggregation  (R/Gene_bin_cell_rna_filtered.R:63-81) ---- */
/* sparse restatement: CSC in (G x N), gene->bin map as CSR over genes, CSC out.
 * out_rowidx/out_x may be NULL to count only.  Returns nnz_out. */
int64_t cso_bin_counts(int G, int N, int B,
                       const int64_t *colptr, const int32_t *rowidx, const double *x,
                       const int32_t *map_ptr, const int32_t *map_bin,
                       int64_t *out_colptr, int32_t *out_rowidx, double *out_x);
/* R-faithful dense algorithm (densify, loop over bins) for small N */
void cso_bin_counts_dense(int G, int N, int B,
                          const int64_t *colptr, const int32_t *rowidx, const double *x,
                          const int32_t *map_ptr, const int32_t *map_bin,
                          double *out /* B*N col-major */);

#ifdef __cplusplus
}
#endif
#endif
