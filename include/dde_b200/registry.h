/* registry.h -- host-side table of compiled models.
 *
 * The reference resolves a model by symbol name in a dyn.load'ed shared
 * object (/root/reference/R/common.R:89-98) and calls it through a function
 * pointer (src/dopri.h:56-57).  A device function cannot be resolved that
 * way without giving up inlining, so a dde_b200 model translation unit
 * instantiates the kernels for its model and registers a launcher here
 * under the model's name (DDE_REGISTER_* in dde_model.cuh).  Any shared
 * library built that way registers itself when it is loaded.
 */
#ifndef DDE_REGISTRY_H
#define DDE_REGISTRY_H

#include <cuda_runtime.h>

namespace dde {

struct DopriArgs;
struct DifeqArgs;

struct ModelDesc {
  const char *name;
  int kind;     /* bit 0: dopri (deriv), bit 1: difeq (target) */
  int n;        /* compiled number of equations; 0 = any n up to DDE_DYN_MAXN */
  int n_parms;  /* >= 0 fixed; -1 = any up to DDE_DYN_MAXP */
  int n_out;    /* >= 0 fixed; -1 = any up to DDE_DYN_MAXOUT */
  int dyn_max_n, dyn_max_parms, dyn_max_out;
  int n_events; /* event functions the model registered (src/dopri.h:53) */
  /* launchers: enqueue the kernel(s) on `stream`, return the CUDA status and
     the number of kernels launched */
  cudaError_t (*launch_dopri)(int method, const DopriArgs &args, int device_sms,
                              cudaStream_t stream, int *n_launches);
  cudaError_t (*launch_difeq)(const DifeqArgs &args, int device_sms, cudaStream_t stream,
                              int *n_launches);
  ModelDesc *next;
};

} /* namespace dde */

/* exported by libdde_b200.so; model libraries call it from a static
   initialiser */
extern "C" void dde_register_model(dde::ModelDesc *desc);

#endif
