/* dde_model.cuh -- what a dde_b200 model translation unit includes.
 *
 * A model file looks like the reference's (compare
 * /root/reference/inst/examples/seir.c): model functions written against
 * <dde/dde.h>, then -- where the reference has `#include <dde/dde.c>` once
 * per shared library -- one DDE_REGISTER_* line per model, which
 * instantiates the solver kernels around the (inlined) model functions and
 * records a launcher under the model's name:
 *
 *     #include "dde_model.cuh"
 *     #include "my_model.h"     // DDE_MODEL_FN void my_rhs(size_t n, double t, ...)
 *     DDE_REGISTER_DOPRI(my_model, my_rhs, 3, 2)              // n = 3, 2 parms
 *     DDE_REGISTER_DOPRI_OUTPUT(lorenz, lorenz, 3, 3, lorenz_output, 2)
 *     DDE_REGISTER_DIFEQ(logistic, logistic, 0, -1, 0)        // any n
 *
 * n > 0 fixes the number of equations at compile time (state in registers);
 * n = 0 compiles a runtime-n variant (n <= DDE_DYN_MAXN, state in local
 * memory).  n_parms / n_out: >= 0 fixed, -1 runtime.
 *
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (no FMA
 * contraction -- bit-parity with the reference depends on it).
 */
#ifndef DDE_MODEL_CUH
#define DDE_MODEL_CUH

#include <math.h>
#include <stddef.h>

#include <type_traits>

#include <dde/dde.h>

#include "difeq_kernels.cuh"
#include "dopri_kernels.cuh"
#include "dopri_dde1_kernel.cuh"
#include "registry.h"
#include "dde_model_def.h"

/* Model sources call libm by name; give them the bit-reproducible versions
   (see dde_math.h). */
#define pow(x, y) dde_pow((x), (y))
#define exp(x) dde_exp((x))

/* shared memory a block may spend on an on-chip difeq history */
#ifndef DDE_DIFEQ_STAGE_BUDGET
#define DDE_DIFEQ_STAGE_BUDGET (56 * 1024)
#endif

namespace dde {

template <class Kernel>
inline int persistent_grid(Kernel kernel, long long n_items, int device_sms, size_t dyn_smem = 0) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, DDE_TPB, dyn_smem) !=
          cudaSuccess ||
      per_sm < 1) {
    per_sm = 1;
    (void) cudaGetLastError();
  }
  /* one resident wave: a multiple of the SM count, never more blocks than
     there is work */
  const long long wanted = (n_items + DDE_TPB - 1) / DDE_TPB;
  const long long resident = (long long) device_sms * per_sm;
  const long long grid = wanted < resident ? wanted : resident;
  return (int) (grid < 1 ? 1 : grid);
}

/* stands in for a model dopri_dde1_kernel cannot take, so that the template is only
   instantiated for models it can (the branch that would launch it is never taken) */
struct Dde1Placeholder {
  static constexpr int N = 1, NP = 0, N_OUT = 0, N_EVENTS = 0, N_LAGS = 1;
  typedef Dde1Scope Scoped;
  static __device__ __forceinline__ void deriv_scoped(Scoped &, size_t, double, const double *,
                                                      double *dydt, const void *) {
    dydt[0] = 0.0;
  }
  static __device__ __forceinline__ void lags(double t, const void *, double *tlag) { tlag[0] = t; }
  static __device__ __forceinline__ void deriv(size_t, double, const double *, double *dydt,
                                               const void *) {
    dydt[0] = 0.0;
  }
  static __device__ __forceinline__ void output(size_t, double, const double *, size_t, double *,
                                                const void *) {}
  static __device__ __forceinline__ void event(int, size_t, double, double *, void *) {}
};

template <class M, int METHOD>
cudaError_t launch_dopri_method(const DopriArgs &args_in, int device_sms, cudaStream_t stream) {
  DopriArgs args = args_in;
  /* stage the lag window's coefficients in shared memory when they fit */
  const size_t order = METHOD == 0 ? 5 : 8;
  size_t dyn = DDE_CTX_BYTES;
  args.win_staged = 0;
  if (args.n_history > 0) {
    const size_t bytes = sizeof(double) * DDE_WIN * order * (size_t) args.n * DDE_TPB;
    if (bytes <= DDE_WIN_COEF_BUDGET) {
      dyn += bytes;
      args.win_staged = 1;
    }
  }
  /* start-up of every trajectory (row 0, k1, first step size) */
  const long long init_grid = (args.n_traj + DDE_TPB - 1) / DDE_TPB;
  cudaError_t err = cudaFuncSetAttribute(dopri_init_kernel<M, METHOD>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int) DDE_CTX_BYTES);
  if (err != cudaSuccess) {
    return err;
  }
  dopri_init_kernel<M, METHOD><<<(unsigned) init_grid, DDE_TPB, DDE_CTX_BYTES, stream>>>(args);
  err = cudaGetLastError();
  if (err != cudaSuccess) {
    return err;
  }
  /* the stepping kernel: one resident wave of persistent threads */
  /* a small model without history: the unrolled stage loop (dopri_kernels.cuh) */
  constexpr bool CAN_UNROLL = M::N >= 1 && M::N <= 4;
  auto kernel = dopri_tpt_kernel<M, METHOD, 0>;
  if (CAN_UNROLL && args.n_history == 0) {
    kernel = dopri_tpt_kernel<M, METHOD, CAN_UNROLL ? 1 : 0>;
    /* the threads' row buffers, then the block's copy of the pow tables */
    dyn = DDE_CTX_BYTES + sizeof(double) * (DDE_STAGE_ROW_DOUBLES + 1) * DDE_TPB + 16 +
          DDE_TAB_SHARED_BYTES;
  }
#ifndef DDE_NO_CALL_DERIV
  /* one equation, a few parameters, a history: unrolled stages around ONE
     out-of-line copy of the model (the Mackey-Glass batch: 15.7 M -> 16.6 M
     trajectories/s) */
  constexpr bool CAN_CALL = M::N == 1 && M::NP >= 0 && M::NP <= 8;
  if (CAN_CALL && args.n_history > 0) {
    kernel = dopri_tpt_kernel<M, METHOD, CAN_CALL ? 2 : 0>;
  }
#endif
#ifndef DDE_NO_DDE1_KERNEL
  /* dop853 of one delay equation with a declared lag, the plain case (forward, fresh
     start, no critical times, output function or stiffness test): dopri_dde1_kernel.cuh */
  constexpr bool CAN_DDE1 =
      M::N == 1 && M::NP >= 0 && M::NP <= 8 && M::N_LAGS == 1 && METHOD == 1;
  if (CAN_DDE1 && args.n_history > 0 && args.sign > 0 && !args.restart && args.n_tcrit == 0 &&
      args.n_out == 0 && args.stiff_check == 0) {
    kernel = dopri_dde1_kernel<typename std::conditional<CAN_DDE1, M, Dde1Placeholder>::type>;
    dyn = DDE1_SMEM_BYTES;
  }
#endif
  err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn);
  if (err != cudaSuccess) {
    return err;
  }
  const int grid = persistent_grid(kernel, args.n_traj, device_sms, dyn);
  kernel<<<grid, DDE_TPB, dyn, stream>>>(args);
  return cudaGetLastError();
}

template <class M>
cudaError_t launch_dopri_model(int method, const DopriArgs &args, int device_sms,
                               cudaStream_t stream, int *n_launches) {
  *n_launches = 2;
  return method == 0 ? launch_dopri_method<M, 0>(args, device_sms, stream)
                     : launch_dopri_method<M, 1>(args, device_sms, stream);
}

template <class M>
cudaError_t launch_difeq_model(const DifeqArgs &args_in, int device_sms, cudaStream_t stream,
                               int *n_launches) {
  DifeqArgs a = args_in;
  /* keep the history on chip when it fits: [slot][component][thread] doubles */
  size_t dyn = 0; /* the difeq context is static shared memory, lag_ctx.cuh */
  a.staged = 0;
  if (a.n_history > 0) {
    const size_t bytes = sizeof(double) * (size_t) a.n_history * (size_t) a.n * DDE_TPB;
    if (bytes <= DDE_DIFEQ_STAGE_BUDGET) {
      dyn += bytes;
      a.staged = 1;
    }
  }
  const bool keep = a.restart || a.restartable;
  auto kernel = a.staged ? (keep ? difeq_tpt_kernel<M, true, true> : difeq_tpt_kernel<M, false, true>)
                         : (keep ? difeq_tpt_kernel<M, true, false>
                                 : difeq_tpt_kernel<M, false, false>);
  cudaError_t err =
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn);
  if (err != cudaSuccess) {
    return err;
  }
  const int grid = persistent_grid(kernel, a.n_rep, device_sms, dyn);
  kernel<<<grid, DDE_TPB, dyn, stream>>>(a);
  *n_launches = 1;
  return cudaGetLastError();
}

} /* namespace dde */

#define DDE_REGISTER_DOPRI_EVENTS(NAME, DERIV, NEQ, NPARMS, OUTPUT, NOUT, EVENTS, NEVENTS) \
  DDE_DEFINE_DOPRI_MODEL(NAME, DERIV, NEQ, NPARMS, OUTPUT, NOUT, EVENTS, NEVENTS)          \
  DDE_REGISTER_DOPRI_DESC(NAME, NEQ, NPARMS, NOUT, NEVENTS)

/* ... with the model's lags declared (dde_model_def.h: M::lags), e.g.
 *     DDE_MODEL_FN void mackey_glass_lags(double t, const void *data, double *tlag) {
 *       tlag[0] = t - 17.0;
 *     }
 * and the model's source compiled a second time inside a struct (dde_model_scoped.inc):
 *     #define DDE_SCOPED_MODEL mackey_glass
 *     #define DDE_SCOPED_SOURCE "mackey_glass.h"
 *     #include "dde_model_scoped.inc"
 *     DDE_REGISTER_DOPRI_LAGS(mackey_glass, mackey_glass, 1, 3, mackey_glass_lags, 1) */
#define DDE_REGISTER_DOPRI_LAGS(NAME, DERIV, NEQ, NPARMS, LAGS, NLAGS)                          \
  DDE_DEFINE_DOPRI_MODEL_LAGS(NAME, DERIV, NEQ, NPARMS, DDE_NO_OUTPUT, 0, DDE_NO_EVENTS, 0, LAGS, \
                              NLAGS)                                                           \
  DDE_REGISTER_DOPRI_DESC(NAME, NEQ, NPARMS, 0, 0)

#define DDE_REGISTER_DOPRI_DESC(NAME, NEQ, NPARMS, NOUT, NEVENTS)                              \
  static dde::ModelDesc dde_dopri_desc_##NAME = {#NAME,                                        \
                                                 1,                                            \
                                                 (NEQ),                                        \
                                                 (NPARMS),                                     \
                                                 (NOUT),                                       \
                                                 DDE_DYN_MAXN,                                 \
                                                 DDE_DYN_MAXP,                                 \
                                                 DDE_DYN_MAXOUT,                               \
                                                 (NEVENTS),                                    \
                                                 &dde::launch_dopri_model<dde_dopri_model_##NAME>, \
                                                 nullptr,                                      \
                                                 nullptr};                                     \
  static const int dde_dopri_reg_##NAME = (dde_register_model(&dde_dopri_desc_##NAME), 0);

#define DDE_REGISTER_DOPRI_OUTPUT(NAME, DERIV, NEQ, NPARMS, OUTPUT, NOUT) \
  DDE_REGISTER_DOPRI_EVENTS(NAME, DERIV, NEQ, NPARMS, OUTPUT, NOUT, DDE_NO_EVENTS, 0)

#define DDE_REGISTER_DOPRI(NAME, DERIV, NEQ, NPARMS) \
  DDE_REGISTER_DOPRI_OUTPUT(NAME, DERIV, NEQ, NPARMS, DDE_NO_OUTPUT, 0)

#define DDE_REGISTER_DIFEQ(NAME, TARGET, NEQ, NPARMS, NOUT) \
  DDE_REGISTER_DIFEQ_FULL(NAME, TARGET, NEQ, NPARMS, NOUT, false, DDE_NO_DIFEQ_SCOPE_TYPE, \
                          DDE_NO_SCOPE_CALL)

/* ... with the model's source compiled a second time inside a struct (dde_model_scoped.inc,
 *     #define DDE_SCOPED_BASE ::dde::DifeqScope): the history state yprev_* read travels in
 *     registers of the iterating thread instead of shared memory (difeq_kernels.cuh) */
#define DDE_REGISTER_DIFEQ_SCOPED(NAME, TARGET, NEQ, NPARMS, NOUT) \
  DDE_REGISTER_DIFEQ_FULL(NAME, TARGET, NEQ, NPARMS, NOUT, true, DDE_SCOPE_TYPE, DDE_SCOPE_CALL)

#define DDE_NO_DIFEQ_SCOPE_TYPE(NAME) ::dde::DifeqScope

#define DDE_REGISTER_DIFEQ_FULL(NAME, TARGET, NEQ, NPARMS, NOUT, IS_SCOPED, SCOPE_TYPE,        \
                                SCOPE_CALL)                                                    \
  struct dde_difeq_model_##NAME {                                                              \
    static constexpr int N = (NEQ), NP = (NPARMS), N_OUT = (NOUT);                             \
    static constexpr bool SCOPED = (IS_SCOPED);                                                \
    typedef SCOPE_TYPE(NAME) Scoped;                                                           \
    static __device__ __forceinline__ void target_scoped(Scoped &sc, size_t n, size_t step,    \
                                                         const double *y, double *y_next,      \
                                                         size_t n_out, double *out,            \
                                                         const void *data) {                   \
      SCOPE_CALL(sc, TARGET)(n, step, y, y_next, n_out, out, data);                            \
    }                                                                                          \
    static __device__ __forceinline__ void target(size_t n, size_t step, const double *y,      \
                                                  double *y_next, size_t n_out, double *out,   \
                                                  const void *data) {                          \
      TARGET(n, step, y, y_next, n_out, out, data);                                            \
    }                                                                                          \
  };                                                                                           \
  static dde::ModelDesc dde_difeq_desc_##NAME = {#NAME,                                        \
                                                 2,                                            \
                                                 (NEQ),                                        \
                                                 (NPARMS),                                     \
                                                 (NOUT),                                       \
                                                 DDE_DYN_MAXN,                                 \
                                                 DDE_DYN_MAXP,                                 \
                                                 DDE_DYN_MAXOUT,                               \
                                                 0,                                            \
                                                 nullptr,                                      \
                                                 &dde::launch_difeq_model<dde_difeq_model_##NAME>, \
                                                 nullptr};                                     \
  static const int dde_difeq_reg_##NAME = (dde_register_model(&dde_difeq_desc_##NAME), 0);

#endif /* DDE_MODEL_CUH */
