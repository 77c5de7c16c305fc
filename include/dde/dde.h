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
 * a dde_b200 model file ends with DDE_REGISTER_* (include/dde_b200/dde_model.cuh).
 */
#ifndef DDE_DDE_H
#define DDE_DDE_H

#include <math.h>
#include <stddef.h>

/* ---- large models: cooperative evaluation (an extension, optional) --------
 *
 * A model with hundreds of equations is integrated by a whole thread block
 * per trajectory (include/dde_b200/dopri_coop_kernel.cuh), and its deriv() is
 * then called by all threads of the block at once.  Written with the three
 * macros below the same source still compiles as plain sequential C (for the
 * reference solver / the oracle) and as a thread-per-trajectory device
 * function:
 *
 *   DDE_FOR(i, lo, hi) { ... }   independent iterations i = lo .. hi-1; the
 *                                block's threads take them in strides
 *   DDE_SCRATCH_BEGIN();          once, before the first DDE_SCRATCH
 *   DDE_SCRATCH(type, name, len); an array all threads share (a local array
 *                                in sequential code)
 *   DDE_SYNC();                   all writes before it are visible to all
 *                                threads after it (nothing in sequential code)
 *
 * y, dydt and data are shared by the block; ylag_vec / ylag_vec_int / ylag_all
 * are collective (every thread calls them with the same arguments, index and
 * result arrays are shared, the results are visible on return).
 */
#if defined(__CUDACC__) && defined(DDE_COOPERATIVE)
#define DDE_FOR(i, lo, hi) \
  for (int i = (int) (lo) + (int) threadIdx.x; i < (int) (hi); i += (int) blockDim.x)
#define DDE_SYNC() __syncthreads()
#define DDE_SCRATCH_BEGIN() size_t dde_scratch_offset_ = 0
#define DDE_SCRATCH(type, name, len)                                       \
  type *name = (type *) (dde_coop_scratch() + dde_scratch_offset_);        \
  dde_scratch_offset_ += (sizeof(type) * (size_t) (len) + 15) & ~(size_t) 15
static __device__ char *dde_coop_scratch(void);
#else
#define DDE_FOR(i, lo, hi) for (int i = (int) (lo); i < (int) (hi); ++i)
#define DDE_SYNC() ((void) 0)
#define DDE_SCRATCH_BEGIN() ((void) 0)
#define DDE_SCRATCH(type, name, len) type name[len]
#endif

#if defined(__CUDACC__) || defined(DDE_HOST_EMU) /* DDE_HOST_EMU: tests/emu, the device code compiled by g++ */

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
