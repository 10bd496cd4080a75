/* Declarations-only stand-in for R's <R_ext/Rdynload.h> — TEST INFRASTRUCTURE (see ../R.h). */
#ifndef CS_RSTUB_RDYNLOAD_H
#define CS_RSTUB_RDYNLOAD_H
typedef void *(*DL_FUNC)(void);
typedef struct {
    const char *name;
    DL_FUNC fun;
    int numArgs;
} R_CallMethodDef;
typedef R_CallMethodDef R_CMethodDef;
typedef R_CallMethodDef R_FortranMethodDef;
typedef R_CallMethodDef R_ExternalMethodDef;
typedef struct _DllInfo DllInfo;
typedef int Rboolean;
#define FALSE 0
#define TRUE 1
int R_registerRoutines(DllInfo *, const R_CMethodDef *, const R_CallMethodDef *, const R_FortranMethodDef *,
                       const R_ExternalMethodDef *);
Rboolean R_useDynamicSymbols(DllInfo *, Rboolean);
#endif
