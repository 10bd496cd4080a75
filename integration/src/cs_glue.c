/*
 * cs_glue.c — `.Call` glue between R and libclonalscope_b200.so.
 *
 * Written against the documented R C API and kept deliberately thin — unpack SEXPs, call the
 * C ABI of include/clonalscope_b200.h, pack the results.  No arithmetic happens here.
 * R is not installed in this repository's image: the file is compiled in the CPU test suite
 * against a declarations-only copy of the R API (integration/rstub/) so that any drift against
 * include/clonalscope_b200.h is a compile error; it is linked and run only in an R installation.
 *
 * Conventions: R owns all memory (allocVector + PROTECT); inputs are read
 * only; errors are raised with Rf_error() from this outermost C frame only, after every
 * library call has returned (no C++ scope or CUDA call is live beneath the longjmp).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>

#include "clonalscope_b200.h"

static void check(int rc)
{
    if (rc != CS_OK) Rf_error("clonalscope_b200: %s", cs_last_error());
}

/* rows of Uall the one-shot samplers may record: niter x min(N, kcap) + slack (kcap: the
 * library's default capacities, include/clonalscope_b200.h cs_gibbs_opts.kcap) */
static int64_t ucap_rows(int niter, int N, int rng_mode)
{
    const int per = rng_mode == CS_RNG_R ? 512 : 256;
    return (int64_t)niter * (N < per ? N : per) + 4096;
}

/* .Call("cs_R_loglik_matrix", X, U, sigma): N x K matrix
 * (R/BayesNonparCluster.R:91-96 Z0 default, R/AssignCluster.R:58 orphans' scores) */
SEXP cs_R_loglik_matrix(SEXP X, SEXP U, SEXP sigma)
{
    const int N = Rf_nrows(X), R = Rf_ncols(X), K = Rf_nrows(U);
    SEXP LL = PROTECT(Rf_allocMatrix(REALSXP, N, K));
    int rc = cs_loglik_matrix(N, R, K, REAL(X), REAL(U), REAL(sigma), REAL(LL));
    UNPROTECT(1);
    check(rc);
    return LL;
}

/* .Call("cs_R_bin_counts", p, i, x, G, map_ptr, map_bin, B): list(p, i, x) of the bin x cell dgCMatrix
 * (R/Gene_bin_cell_rna_filtered.R:69-81; R/EstRegionCov.R:88-93,131-137 with segments as bins) */
SEXP cs_R_bin_counts(SEXP p, SEXP i, SEXP x, SEXP G, SEXP map_ptr, SEXP map_bin, SEXP B)
{
    const int N = LENGTH(p) - 1;
    cs_bin_job *job = NULL;
    int64_t nnz = 0;
    check(cs_bin_counts_begin(Rf_asInteger(G), N, Rf_asInteger(B), INTEGER(p), INTEGER(i), REAL(x),
                              INTEGER(map_ptr), INTEGER(map_bin), &job, &nnz));
    SEXP op = PROTECT(Rf_allocVector(INTSXP, N + 1));
    SEXP oi = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)nnz));
    SEXP ox = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)nnz));
    int rc = cs_bin_counts_fetch(job, INTEGER(op), INTEGER(oi), REAL(ox)); /* frees the job */
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, op);
    SET_VECTOR_ELT(out, 1, oi);
    SET_VECTOR_ELT(out, 2, ox);
    UNPROTECT(4);
    check(rc);
    return out;
}

/* .Call("cs_R_ncrp_gibbs", X, U0, Z0, sigmas0, alpha, beta, niter, seed, rng_mode, device)
 * -> list(Zall, sigma_all, Likelihood, L0, kt, u_off, u_labels, u_rows, rng_state)
 * (R/BayesNonparCluster.R:133-277) */
SEXP cs_R_ncrp_gibbs(SEXP X, SEXP U0, SEXP Z0, SEXP sigmas0, SEXP alpha, SEXP beta, SEXP niter_, SEXP seed,
                     SEXP rng_mode, SEXP device)
{
    const int N = Rf_nrows(X), R = Rf_ncols(X), K0 = Rf_nrows(U0), niter = Rf_asInteger(niter_);
    cs_gibbs_opts o;
    memset(&o, 0, sizeof(o));
    o.rng_mode = Rf_asInteger(rng_mode);
    o.device = Rf_asInteger(device);
    const int64_t ucap = ucap_rows(niter, N, o.rng_mode);
    SEXP Zall = PROTECT(Rf_allocMatrix(INTSXP, niter, N));
    SEXP sig = PROTECT(Rf_allocMatrix(REALSXP, R, niter));
    SEXP LL = PROTECT(Rf_allocVector(REALSXP, niter));
    SEXP L0 = PROTECT(Rf_allocVector(REALSXP, 1));
    SEXP kt = PROTECT(Rf_allocVector(INTSXP, niter));
    SEXP uoff = PROTECT(Rf_allocVector(REALSXP, niter + 1));
    SEXP ulab = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)ucap));
    SEXP urows = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)ucap * R));
    SEXP rst = PROTECT(Rf_allocVector(INTSXP, 625));
    int64_t *off = (int64_t *)R_alloc((size_t)niter + 1, sizeof(int64_t));
    int64_t nrej = 0;
    int rc = cs_ncrp_gibbs(N, R, REAL(X), K0, REAL(U0), INTEGER(Z0), REAL(sigmas0), Rf_asReal(alpha),
                           Rf_asReal(beta), niter, (uint32_t)Rf_asInteger(seed), &o, INTEGER(Zall), REAL(sig),
                           REAL(LL), REAL(L0), INTEGER(kt), ucap, off, INTEGER(ulab), REAL(urows), &nrej,
                           INTEGER(rst));
    for (int t = 0; t <= niter; t++) REAL(uoff)[t] = (double)off[t];
    const char *names[] = {"Zall", "sigma_all", "Likelihood", "L0", "kt", "u_off", "u_labels", "u_rows", "rng_state", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SEXP parts[] = {Zall, sig, LL, L0, kt, uoff, ulab, urows, rst};
    for (int k = 0; k < 9; k++) SET_VECTOR_ELT(out, k, parts[k]);
    UNPROTECT(10);
    check(rc);
    return out;
}

/* .Call("cs_R_ncrp_gibbs_allele", Xcov, Xallele, U0, Z0, sig0_cov, sig0_allele, alpha, beta, niter, seed,
 *       maxcp, rng_mode, device)
 * -> list(Zall, sigma_cov_all, sigma_allele_all, Likelihood, L0, kt, u_off, u_labels, u_state, u_cov, u_allele)
 * (R/BayesNonparAlleleCluster.R:101-290, R/genotype_neighbor.R:7-40); U0: integer state ids */
SEXP cs_R_ncrp_gibbs_allele(SEXP Xcov, SEXP Xallele, SEXP U0, SEXP Z0, SEXP sig0_cov, SEXP sig0_allele, SEXP alpha,
                            SEXP beta, SEXP niter_, SEXP seed, SEXP maxcp, SEXP rng_mode, SEXP device)
{
    const int N = Rf_nrows(Xcov), R = Rf_ncols(Xcov), K0 = Rf_nrows(U0), niter = Rf_asInteger(niter_);
    cs_gibbs_opts o;
    memset(&o, 0, sizeof(o));
    o.rng_mode = Rf_asInteger(rng_mode);
    o.device = Rf_asInteger(device);
    const int64_t ucap = ucap_rows(niter, N, o.rng_mode);
    SEXP Zall = PROTECT(Rf_allocMatrix(INTSXP, niter, N));
    SEXP sigc = PROTECT(Rf_allocMatrix(REALSXP, R, niter));
    SEXP siga = PROTECT(Rf_allocMatrix(REALSXP, R, niter));
    SEXP LL = PROTECT(Rf_allocVector(REALSXP, niter));
    SEXP L0 = PROTECT(Rf_allocVector(REALSXP, 1));
    SEXP kt = PROTECT(Rf_allocVector(INTSXP, niter));
    SEXP uoff = PROTECT(Rf_allocVector(REALSXP, niter + 1));
    SEXP ulab = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)ucap));
    SEXP ustate = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)ucap * R));
    SEXP ucov = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)ucap * R));
    SEXP uall = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)ucap * R));
    int64_t *off = (int64_t *)R_alloc((size_t)niter + 1, sizeof(int64_t));
    int rc = cs_ncrp_gibbs_allele(N, R, REAL(Xcov), REAL(Xallele), K0, INTEGER(U0), INTEGER(Z0), REAL(sig0_cov),
                                  REAL(sig0_allele), Rf_asReal(alpha), Rf_asReal(beta), niter,
                                  (uint32_t)Rf_asInteger(seed), Rf_asInteger(maxcp), &o, INTEGER(Zall), REAL(sigc),
                                  REAL(siga), REAL(LL), REAL(L0), INTEGER(kt), ucap, off, INTEGER(ulab),
                                  INTEGER(ustate), REAL(ucov), REAL(uall));
    for (int t = 0; t <= niter; t++) REAL(uoff)[t] = (double)off[t];
    const char *names[] = {"Zall", "sigma_cov_all", "sigma_allele_all", "Likelihood", "L0", "kt", "u_off", "u_labels",
                           "u_state", "u_cov", "u_allele", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SEXP parts[] = {Zall, sigc, siga, LL, L0, kt, uoff, ulab, ustate, ucov, uall};
    for (int k = 0; k < 11; k++) SET_VECTOR_ELT(out, k, parts[k]);
    UNPROTECT(12);
    check(rc);
    return out;
}

/* .Call("cs_R_majority_vote", Zall): integer vector, one label per cell (R/AssignCluster.R:37, :232) */
SEXP cs_R_majority_vote(SEXP Zall)
{
    const int T = Rf_nrows(Zall), N = Rf_ncols(Zall);
    SEXP z = PROTECT(Rf_allocVector(INTSXP, N));
    int rc = cs_majority_vote(T, N, INTEGER(Zall), INTEGER(z));
    UNPROTECT(1);
    check(rc);
    return z;
}

/* .Call("cs_R_cluster_stats", X, Z, K): list(sum, count, n) — K x R sums and non-NA counts of the
 * cells of each label, and the cluster sizes (R/AssignCluster.R:46-51, :67-72; labels NA / outside
 * 1..K are skipped: NA_integer_ is INT_MIN) */
SEXP cs_R_cluster_stats(SEXP X, SEXP Z, SEXP K_)
{
    const int N = Rf_nrows(X), R = Rf_ncols(X), K = Rf_asInteger(K_);
    SEXP s = PROTECT(Rf_allocMatrix(REALSXP, K, R)), c = PROTECT(Rf_allocMatrix(REALSXP, K, R));
    SEXP n = PROTECT(Rf_allocVector(REALSXP, K));
    int rc = cs_cluster_stats(N, R, K, REAL(X), INTEGER(Z), REAL(s), REAL(c), REAL(n));
    const char *names[] = {"sum", "count", "n", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, s);
    SET_VECTOR_ELT(out, 1, c);
    SET_VECTOR_ELT(out, 2, n);
    UNPROTECT(4);
    check(rc);
    return out;
}

/* .Call("cs_R_sigma_stats", X, Z, U): list(sq, count) — per-segment squared residuals against the
 * cells' own cluster rows and non-NA counts (R/AssignCluster.R:53, :77) */
SEXP cs_R_sigma_stats(SEXP X, SEXP Z, SEXP U)
{
    const int N = Rf_nrows(X), R = Rf_ncols(X), K = Rf_nrows(U);
    SEXP q = PROTECT(Rf_allocVector(REALSXP, R)), c = PROTECT(Rf_allocVector(REALSXP, R));
    int rc = cs_sigma_stats(N, R, K, REAL(X), INTEGER(Z), REAL(U), REAL(q), REAL(c));
    const char *names[] = {"sq", "count", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, q);
    SET_VECTOR_ELT(out, 1, c);
    UNPROTECT(3);
    check(rc);
    return out;
}

/* ---- EstRegionCov on a device-resident matrix (R/EstRegionCov.R:36-48, :80-159) */
static void csc_finalizer(SEXP ptr)
{
    cs_csc *m = (cs_csc *)R_ExternalPtrAddr(ptr);
    if (m) cs_csc_free(m);
    R_ClearExternalPtr(ptr);
}

/* .Call("cs_R_csc_upload", p, i, x, G, col_sel): external pointer to the uploaded matrix; col_sel =
 * match(celltype0[,1], barcodes[,1]) - 1L (or integer(0) for all columns in order) */
SEXP cs_R_csc_upload(SEXP p, SEXP i, SEXP x, SEXP G, SEXP col_sel)
{
    cs_csc *m = NULL;
    const int nsel = LENGTH(col_sel);
    check(cs_csc_upload(Rf_asInteger(G), LENGTH(p) - 1, INTEGER(p), INTEGER(i), REAL(x), nsel,
                        nsel ? INTEGER(col_sel) : NULL, &m));
    SEXP ptr = PROTECT(R_MakeExternalPtr(m, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, csc_finalizer, 1);
    UNPROTECT(1);
    return ptr;
}

/* .Call("cs_R_csc_gene_moments", handle, G): list(sum, sumsq) (:42-43) */
SEXP cs_R_csc_gene_moments(SEXP handle, SEXP G)
{
    const int g = Rf_asInteger(G);
    SEXP s = PROTECT(Rf_allocVector(REALSXP, g)), q = PROTECT(Rf_allocVector(REALSXP, g));
    int rc = cs_csc_gene_moments((cs_csc *)R_ExternalPtrAddr(handle), REAL(s), REAL(q));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, s);
    SET_VECTOR_ELT(out, 1, q);
    UNPROTECT(3);
    check(rc);
    return out;
}

/* .Call("cs_R_csc_region_cov", handle, N, S, map_ptr, map_seg, lib_flag (raw), celltype, include)
 * -> list(deltas N x S, sums N x S, alpha, nsel) (:88-159 for every segment at once) */
SEXP cs_R_csc_region_cov(SEXP handle, SEXP N_, SEXP S_, SEXP map_ptr, SEXP map_seg, SEXP lib_flag, SEXP celltype,
                         SEXP include)
{
    const int N = Rf_asInteger(N_), S = Rf_asInteger(S_);
    SEXP d = PROTECT(Rf_allocMatrix(REALSXP, N, S)), s = PROTECT(Rf_allocMatrix(REALSXP, N, S));
    SEXP a = PROTECT(Rf_allocVector(REALSXP, N)), n = PROTECT(Rf_allocVector(INTSXP, S));
    int rc = cs_csc_region_cov((cs_csc *)R_ExternalPtrAddr(handle), S, INTEGER(map_ptr), INTEGER(map_seg), RAW(lib_flag),
                               INTEGER(celltype), Rf_asInteger(include), REAL(d), REAL(s), REAL(a), INTEGER(n));
    const char *names[] = {"deltas", "sums", "alpha", "nsel", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, d);
    SET_VECTOR_ELT(out, 1, s);
    SET_VECTOR_ELT(out, 2, a);
    SET_VECTOR_ELT(out, 3, n);
    UNPROTECT(5);
    check(rc);
    return out;
}

/* .Call("cs_R_rowsum_by_label", p, i, x, B, label, K): B x K matrix, column k = rowSums over the
 * cells with label k (0-based; others skipped) (R/CreateSegtableNoWGS.R:62-69, :77) */
SEXP cs_R_rowsum_by_label(SEXP p, SEXP i, SEXP x, SEXP B_, SEXP label, SEXP K_)
{
    const int B = Rf_asInteger(B_), K = Rf_asInteger(K_);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, B, K));
    int rc = cs_rowsum_by_label(B, LENGTH(p) - 1, INTEGER(p), INTEGER(i), REAL(x), INTEGER(label), K, REAL(out));
    UNPROTECT(1);
    check(rc);
    return out;
}

/* .Call("cs_R_segment_hmm", off, val, nmean, means, sds, delta, Pi): list(smooth, state) for the
 * sequences val[off[s] .. off[s+1]) (R/Segmentation_bulk.R:152-166: runmean + dthmm + Viterbi);
 * off is a double vector (R has no 64-bit integers), Pi row-major, state 1-based like Viterbi() */
SEXP cs_R_segment_hmm(SEXP off, SEXP val, SEXP nmean, SEXP means, SEXP sds, SEXP delta, SEXP Pi)
{
    const int nseq = LENGTH(off) - 1;
    const R_xlen_t total = XLENGTH(val);
    int64_t off64[1024];
    if (nseq > 1023) Rf_error("cs_R_segment_hmm: at most 1023 sequences per call");
    for (int s = 0; s <= nseq; s++) off64[s] = (int64_t)REAL(off)[s];
    SEXP sm = PROTECT(Rf_allocVector(REALSXP, total)), st = PROTECT(Rf_allocVector(INTSXP, total));
    int rc = cs_segment_hmm(nseq, off64, REAL(val), Rf_asInteger(nmean), REAL(means), REAL(sds), REAL(delta), REAL(Pi),
                            REAL(sm), INTEGER(st));
    if (rc == CS_OK)
        for (R_xlen_t e = 0; e < total; e++) INTEGER(st)[e] += 1;
    const char *names[] = {"smooth", "state", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, sm);
    SET_VECTOR_ELT(out, 1, st);
    UNPROTECT(3);
    check(rc);
    return out;
}

/* .Call("cs_R_read_mm", path): list(Dim, p, i, x) of the dgCMatrix (Matrix::readMM + as(., "dgCMatrix"),
 * samples/P5931/scRNA/README.md:72) */
SEXP cs_R_read_mm(SEXP path)
{
    cs_mm_job *job = NULL;
    int64_t dims[3] = {0, 0, 0};
    check(cs_mm_read_begin(CHAR(STRING_ELT(path, 0)), &job, dims));
    SEXP dim = PROTECT(Rf_allocVector(INTSXP, 2));
    INTEGER(dim)[0] = (int)dims[0];
    INTEGER(dim)[1] = (int)dims[1];
    SEXP op = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)dims[1] + 1));
    SEXP oi = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)dims[2]));
    SEXP ox = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)dims[2]));
    int rc = cs_mm_read_fetch(job, INTEGER(op), INTEGER(oi), REAL(ox)); /* frees the job */
    const char *names[] = {"Dim", "p", "i", "x", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, dim);
    SET_VECTOR_ELT(out, 1, op);
    SET_VECTOR_ELT(out, 2, oi);
    SET_VECTOR_ELT(out, 3, ox);
    UNPROTECT(5);
    check(rc);
    return out;
}

static const R_CallMethodDef calls[] = {
    {"cs_R_loglik_matrix", (DL_FUNC)&cs_R_loglik_matrix, 3},
    {"cs_R_bin_counts", (DL_FUNC)&cs_R_bin_counts, 7},
    {"cs_R_ncrp_gibbs", (DL_FUNC)&cs_R_ncrp_gibbs, 10},
    {"cs_R_ncrp_gibbs_allele", (DL_FUNC)&cs_R_ncrp_gibbs_allele, 13},
    {"cs_R_majority_vote", (DL_FUNC)&cs_R_majority_vote, 1},
    {"cs_R_cluster_stats", (DL_FUNC)&cs_R_cluster_stats, 3},
    {"cs_R_sigma_stats", (DL_FUNC)&cs_R_sigma_stats, 3},
    {"cs_R_csc_upload", (DL_FUNC)&cs_R_csc_upload, 5},
    {"cs_R_csc_gene_moments", (DL_FUNC)&cs_R_csc_gene_moments, 2},
    {"cs_R_csc_region_cov", (DL_FUNC)&cs_R_csc_region_cov, 8},
    {"cs_R_rowsum_by_label", (DL_FUNC)&cs_R_rowsum_by_label, 6},
    {"cs_R_segment_hmm", (DL_FUNC)&cs_R_segment_hmm, 7},
    {"cs_R_read_mm", (DL_FUNC)&cs_R_read_mm, 1},
    {NULL, NULL, 0}};

void R_init_Clonalscope(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, calls, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
