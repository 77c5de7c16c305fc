/* dde_model_coop.cuh -- what a translation unit of LARGE models includes: the
 * block-per-trajectory counterpart of dde_model.cuh.  Model functions are
 * written with the cooperative macros of <dde/dde.h> (DDE_FOR, DDE_SCRATCH,
 * DDE_SYNC); one DDE_REGISTER_DOPRI_COOP line per model instantiates
 * dopri_coop_kernel around the inlined deriv():
 *
 *     #include "dde_model_coop.cuh"
 *     #include "seir_age.h"
 *     DDE_REGISTER_DOPRI_COOP(seir_age, seir_age, 128, 4, 3)
 *         // name, deriv, threads per trajectory, components per thread, n_parms
 *
 * The model then accepts any n <= threads x components.  A model is
 * registered either here or in a thread-per-trajectory unit, not both.
 */
#ifndef DDE_MODEL_COOP_CUH
#define DDE_MODEL_COOP_CUH

#define DDE_COOPERATIVE 1

#include <math.h>
#include <stddef.h>

#include <dde/dde.h>

#include "dopri_coop_kernel.cuh"
#include "registry.h"

#define pow(x, y) dde_pow((x), (y))
#define exp(x) dde_exp((x))

namespace dde {

template <class M, int METHOD, int TPT, int CMAX>
cudaError_t launch_dopri_coop_method(const DopriArgs &args, int device_sms, cudaStream_t stream) {
  auto kernel = args.n == TPT * CMAX ? dopri_coop_kernel<M, METHOD, TPT, CMAX, TPT * CMAX>
                                     : dopri_coop_kernel<M, METHOD, TPT, CMAX, 0>;
  const size_t table =
      args.n_history > 0 && args.n_history <= DDE_COOP_MAX_TABLE ? 2 * (size_t) args.n_history : 0;
  const size_t dyn = sizeof(double) * (DDE_COOP_NVEC * (size_t) args.n + table) + DDE_COOP_SCRATCH_BYTES;
  cudaError_t err =
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn);
  if (err != cudaSuccess) {
    return err;
  }
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TPT, dyn) != cudaSuccess ||
      per_sm < 1) {
    per_sm = 1;
    (void) cudaGetLastError();
  }
  /* persistent blocks: one resident wave, never more blocks than trajectories */
  const long long resident = (long long) device_sms * per_sm;
  const long long grid = args.n_traj < resident ? args.n_traj : resident;
  kernel<<<(unsigned) (grid < 1 ? 1 : grid), TPT, dyn, stream>>>(args);
  return cudaGetLastError();
}

template <class M, int TPT, int CMAX>
cudaError_t launch_dopri_coop_model(int method, const DopriArgs &args, int device_sms,
                                    cudaStream_t stream, int *n_launches) {
  *n_launches = 1;
  if (args.n_out > 0) {
    return cudaErrorNotSupported; /* output functions: thread-per-trajectory models only */
  }
  return method == 0 ? launch_dopri_coop_method<M, 0, TPT, CMAX>(args, device_sms, stream)
                     : launch_dopri_coop_method<M, 1, TPT, CMAX>(args, device_sms, stream);
}

} /* namespace dde */

#define DDE_REGISTER_DOPRI_COOP(NAME, DERIV, THREADS, COMPONENTS, NPARMS)                      \
  struct dde_coop_model_##NAME {                                                               \
    static __device__ __forceinline__ void deriv(size_t n, double t, const double *y,          \
                                                 double *dydt, const void *data) {             \
      DERIV(n, t, y, dydt, data);                                                              \
    }                                                                                          \
  };                                                                                           \
  static dde::ModelDesc dde_coop_desc_##NAME = {                                               \
      #NAME,                                                                                   \
      1,                                                                                       \
      0,                                                                                       \
      (NPARMS),                                                                                \
      0,                                                                                       \
      (THREADS) * (COMPONENTS),                                                                \
      (NPARMS) >= 0 ? (NPARMS) : (THREADS) * (COMPONENTS),                                     \
      0,                                                                                       \
      0,                                                                                       \
      &dde::launch_dopri_coop_model<dde_coop_model_##NAME, (THREADS), (COMPONENTS)>,           \
      nullptr,                                                                                 \
      nullptr};                                                                                \
  static const int dde_coop_reg_##NAME = (dde_register_model(&dde_coop_desc_##NAME), 0);

#endif /* DDE_MODEL_COOP_CUH */
