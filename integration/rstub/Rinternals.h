/* Declarations-only stand-in for R's <Rinternals.h> — TEST INFRASTRUCTURE (see R.h here). */
#ifndef CS_RSTUB_RINTERNALS_H
#define CS_RSTUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_mkNamed(SEXPTYPE, const char **);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int LENGTH(SEXP);
R_xlen_t XLENGTH(SEXP);
int *INTEGER(SEXP);
double *REAL(SEXP);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP R_do_slot(SEXP, SEXP);
SEXP Rf_install(const char *);
#define RAWSXP 24
typedef unsigned char Rbyte;
Rbyte *RAW(SEXP);
extern SEXP R_NilValue;
SEXP STRING_ELT(SEXP, R_xlen_t);
const char *R_CHAR(SEXP);
#define CHAR(x) R_CHAR(x)
typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void *, SEXP, SEXP);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, int);
SEXP Rf_ScalarInteger(int);
#endif
