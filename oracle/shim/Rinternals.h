/* TEST INFRASTRUCTURE ONLY (oracle): stand-in for R's <Rinternals.h>.
 *
 * The solver core touches SEXP only in the verbose callback
 * (src/dopri.c:838-871), which is unreachable with verbose = VERBOSE_QUIET.
 * The symbols exist so the file links; calling any of them aborts.
 */
#ifndef DDE_ORACLE_SHIM_RINTERNALS_H
#define DDE_ORACLE_SHIM_RINTERNALS_H

#include "R.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dde_shim_sexprec *SEXP;
extern SEXP R_NilValue;

#define REALSXP 14
#define PROTECT(x) (x)
#define UNPROTECT(n) ((void) (n))

SEXP VECTOR_ELT(SEXP x, long i);
SEXP Rf_ScalarReal(double x);
SEXP Rf_allocVector(unsigned int type, long n);
double *REAL(SEXP x);
SEXP Rf_lang4(SEXP a, SEXP b, SEXP c, SEXP d);
SEXP Rf_eval(SEXP call, SEXP env);

#ifdef __cplusplus
}
#endif

#endif
