/* TEST INFRASTRUCTURE ONLY (oracle): stand-in for R's <R.h>.
 *
 * R is not installed in this image, but the reference solver core
 * (/root/reference/src/{dopri,dopri_5,dopri_853,difeq}.c) only needs a handful
 * of R API entry points.  This header supplies them so those files compile
 * unmodified, from where they lie, into oracle/_ref/ (see oracle/Makefile).
 *
 * Semantics kept:
 *   R_Calloc zero-fills, R_Free releases           (src/dopri.c:13,38-49,222-234)
 *   R_alloc is transient storage                    (src/difeq.c:183)
 *   Rf_error does not return                        (src/dopri.c:333, src/difeq.c:86,91,268)
 *   NA_REAL is R's NA: a quiet NaN, low word 1954   (src/dopri.c:782, src/difeq.c:333)
 *   R_FINITE(x) is false for NA/NaN/Inf             (src/dopri.c:332)
 */
#ifndef DDE_ORACLE_SHIM_R_H
#define DDE_ORACLE_SHIM_R_H

#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

void *dde_shim_calloc(size_t n, size_t size);
void dde_shim_free(void *p);
char *R_alloc(size_t n, int size);
void Rf_error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
void Rprintf(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
double dde_shim_na_real(void);

#define R_Calloc(n, type) ((type *) dde_shim_calloc((size_t) (n), sizeof(type)))
#define R_Free(p) (dde_shim_free((void *) (p)), (p) = NULL)
#define R_FINITE(x) isfinite(x)
#define NA_REAL dde_shim_na_real()

#ifdef __cplusplus
}
#endif

#endif
