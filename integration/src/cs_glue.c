/*
 * cs_glue.c — `.Call` glue between R and libclonalscope_b200.so.
 *
 * NOT compiled in this repository's image (no R headers there): written against the
 * documented R C API and kept deliberately thin — unpack SEXPs, call the C ABI of
 * include/clonalscope_b200.h, pack the results.  No arithmetic happens here.
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

/* .Call("cs_R_loglik_matrix", X, U, sigma): N x K matrix */
SEXP cs_R_loglik_matrix(SEXP X, SEXP U, SEXP sigma)
{
    const int N = Rf_nrows(X), R = Rf_ncols(X), K = Rf_nrows(U);
    SEXP LL = PROTECT(Rf_allocMatrix(REALSXP, N, K));
    int rc = cs_loglik_matrix(N, R, K, REAL(X), REAL(U), REAL(sigma), REAL(LL));
    UNPROTECT(1);
    check(rc);
    return LL;
}

/* .Call("cs_R_bin_counts", p, i, x, G, map_ptr, map_bin, B): list(p, i, x) of the bin x cell dgCMatrix */
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
 * -> list(Zall, sigma_all, Likelihood, L0, kt, u_off, u_labels, u_rows, rng_state) */
SEXP cs_R_ncrp_gibbs(SEXP X, SEXP U0, SEXP Z0, SEXP sigmas0, SEXP alpha, SEXP beta, SEXP niter_, SEXP seed,
                     SEXP rng_mode, SEXP device)
{
    const int N = Rf_nrows(X), R = Rf_ncols(X), K0 = Rf_nrows(U0), niter = Rf_asInteger(niter_);
    const int64_t ucap = (int64_t)niter * 160 + 4096;
    cs_gibbs_opts o;
    memset(&o, 0, sizeof(o));
    o.rng_mode = Rf_asInteger(rng_mode);
    o.device = Rf_asInteger(device);
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

/* .Call("cs_R_majority_vote", Zall): integer vector, one label per cell */
SEXP cs_R_majority_vote(SEXP Zall)
{
    const int T = Rf_nrows(Zall), N = Rf_ncols(Zall);
    SEXP z = PROTECT(Rf_allocVector(INTSXP, N));
    int rc = cs_majority_vote(T, N, INTEGER(Zall), INTEGER(z));
    UNPROTECT(1);
    check(rc);
    return z;
}

/* .Call("cs_R_gene_moments", i, x, G): list(sum, sumsq), one entry per gene (EstRegionCov :42-43) */
SEXP cs_R_gene_moments(SEXP i, SEXP x, SEXP G)
{
    const int g = Rf_asInteger(G);
    SEXP s = PROTECT(Rf_allocVector(REALSXP, g)), q = PROTECT(Rf_allocVector(REALSXP, g));
    int rc = cs_gene_moments(g, (int64_t)XLENGTH(x), INTEGER(i), REAL(x), REAL(s), REAL(q));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, s);
    SET_VECTOR_ELT(out, 1, q);
    UNPROTECT(3);
    check(rc);
    return out;
}

static const R_CallMethodDef calls[] = {
    {"cs_R_loglik_matrix", (DL_FUNC)&cs_R_loglik_matrix, 3},
    {"cs_R_bin_counts", (DL_FUNC)&cs_R_bin_counts, 7},
    {"cs_R_ncrp_gibbs", (DL_FUNC)&cs_R_ncrp_gibbs, 10},
    {"cs_R_majority_vote", (DL_FUNC)&cs_R_majority_vote, 1},
    {"cs_R_gene_moments", (DL_FUNC)&cs_R_gene_moments, 3},
    {NULL, NULL, 0}};

void R_init_Clonalscope(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, calls, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
