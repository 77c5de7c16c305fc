/* <dde/dde.h> -- what a compiled model includes.
 *
 * Source-compatible with the reference's model plugin header
 * (/root/reference/inst/include/dde/dde.h:3-13): the same eight history
 * functions with the same signatures.  A model source written against it --
 *
 *     void deriv(size_t n, double t, const double *y, double *dydt, const void *data);
 *     void output(size_t n, double t, const double *y, size_t n_out, double *out, const void *data);
 *     void event(size_t n, double t, double *y, void *data);
 *     void target(size_t n, size_t step, const double *y, double *y_next,
 *                 size_t n_out, double *out, const void *data);
 *
 * (src/dopri.h:49-53, src/difeq.h:19-21) -- compiles unchanged either
 *   * with a C compiler against the reference solver (the oracle), where the
 *     functions below are the reference's (src/dopri.c:776-828,
 *     src/difeq.c:274-324), or
 *   * with nvcc inside a dde_b200 model translation unit, where they are
 *     __device__ functions reading the calling trajectory's ring buffer in
 *     HBM through an implicit per-thread context, and libm `pow`/`exp` are
 *     routed to the bit-reproducible dde_pow/dde_exp.
 *
 * The only addition is the DDE_MODEL_FN decoration on the model's own
 * functions (empty for C, `__device__` for nvcc).  Where the reference asks
 * for `#include <dde/dde.c>` once per shared library (inst/include/dde/dde.c),
 * a dde_b200 model file ends with DDE_REGISTER_* (dde_b200/csrc/dde_model.cuh).
 */
#ifndef DDE_DDE_H
#define DDE_DDE_H

#include <math.h>
#include <stddef.h>

#if defined(__CUDACC__)

#define DDE_MODEL_FN static __device__ __forceinline__

/* dopri: */
static __device__ double ylag_1(double t, size_t i);
static __device__ void ylag_all(double t, double *y);
static __device__ void ylag_vec(double t, const size_t *idx, size_t nidx, double *y);
static __device__ void ylag_vec_int(double t, const int *idx, size_t nidx, double *y);

/* difeq: */
static __device__ double yprev_1(int step, size_t i);
static __device__ void yprev_all(int step, double *y);
static __device__ void yprev_vec(int step, const size_t *idx, size_t nidx, double *y);
static __device__ void yprev_vec_int(int step, const int *idx, size_t nidx, double *y);

#else /* plain C: the reference's declarations */

#define DDE_MODEL_FN

#ifdef __cplusplus
extern "C" {
#endif

/* dopri: */
double ylag_1(double t, size_t i);
void ylag_all(double t, double *y);
void ylag_vec(double t, const size_t *idx, size_t nidx, double *y);
void ylag_vec_int(double t, const int *idx, size_t nidx, double *y);

/* difeq: */
double yprev_1(int step, size_t i);
void yprev_all(int step, double *y);
void yprev_vec(int step, const size_t *idx, size_t nidx, double *y);
void yprev_vec_int(int step, const int *idx, size_t nidx, double *y);

#ifdef __cplusplus
}
#endif

#endif

#endif /* DDE_DDE_H */
