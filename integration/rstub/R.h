/* Declarations-only stand-in for R's <R.h> — TEST INFRASTRUCTURE.
 * R is not installed in the build image; tests/test_host_logic.py compiles
 * integration/src/cs_glue.c against these prototypes (gcc -fsyntax-only semantics: nothing is
 * linked) so that a change of include/clonalscope_b200.h that the glue does not follow is a
 * compile error.  Signatures follow the documented R C API ("Writing R Extensions", section 6). */
#ifndef CS_RSTUB_R_H
#define CS_RSTUB_R_H
#include <stddef.h>
char *R_alloc(size_t n, int size);
void Rf_error(const char *fmt, ...) __attribute__((noreturn));
void Rf_warning(const char *fmt, ...);
void R_CheckUserInterrupt(void);
#endif
