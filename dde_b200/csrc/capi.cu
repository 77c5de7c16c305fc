/* capi.cu -- host side of the C-ABI declared in include/dde_b200.h.
 *
 * Owns device memory, streams and the model registry; validates arguments
 * the way the reference's reset functions do (/root/reference/src/dopri.c:
 * 138-219, src/difeq.c:74-105) and launches the model's kernels.  No solver
 * arithmetic happens on the host: if the CUDA path is unavailable every
 * entry point fails loudly.
 */
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "dde_b200.h"
#include "dde_math.h"
#include "difeq_kernels.cuh"
#include "dopri_kernels.cuh"
#include "registry.h"

/* trajectories per chunk of a streamed download: 2^15 (a few MB of results) */
#ifndef DDE_CHUNK_SHIFT
#define DDE_CHUNK_SHIFT 15
#endif

namespace {

thread_local std::string g_last_error;

int fail(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return 1;
}

#define DDE_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t err__ = (call);                                                 \
    if (err__ != cudaSuccess) {                                                 \
      return fail("CUDA error %s at %s:%d (%s)", cudaGetErrorName(err__), __FILE__, __LINE__, \
                  cudaGetErrorString(err__));                                   \
    }                                                                           \
  } while (0)

dde::ModelDesc *g_models = nullptr;
std::mutex g_models_mutex;

/* exact-n registration wins over a runtime-n one of the same name */
dde::ModelDesc *find_model(const char *name, int kind, size_t n) {
  std::lock_guard<std::mutex> lock(g_models_mutex);
  dde::ModelDesc *wild = nullptr;
  for (dde::ModelDesc *m = g_models; m != nullptr; m = m->next) {
    if (strcmp(m->name, name) != 0 || (m->kind & kind) == 0) {
      continue;
    }
    if (m->n == (int) n) {
      return m;
    }
    if (m->n == 0) {
      wild = m;
    }
  }
  return wild;
}

bool model_name_known(const char *name) {
  std::lock_guard<std::mutex> lock(g_models_mutex);
  for (dde::ModelDesc *m = g_models; m != nullptr; m = m->next) {
    if (strcmp(m->name, name) == 0) {
      return true;
    }
  }
  return false;
}

int check_model_shape(const dde::ModelDesc *m, size_t n, size_t n_parms, size_t n_out) {
  if (m->n > 0 ? n != (size_t) m->n : (n < 1 || n > (size_t) m->dyn_max_n)) {
    return fail("model '%s' is compiled for n = %d%s, got n = %zu", m->name,
                m->n > 0 ? m->n : m->dyn_max_n, m->n > 0 ? "" : " at most", n);
  }
  if (m->n_parms >= 0 ? n_parms != (size_t) m->n_parms : n_parms > (size_t) m->dyn_max_parms) {
    return fail("model '%s' takes %d%s parameters, got %zu", m->name,
                m->n_parms >= 0 ? m->n_parms : m->dyn_max_parms, m->n_parms >= 0 ? "" : " at most",
                n_parms);
  }
  if (n_out != 0 &&
      (m->n_out >= 0 ? n_out != (size_t) m->n_out : n_out > (size_t) m->dyn_max_out)) {
    return fail("model '%s' has %d output variables, got n_out = %zu", m->name,
                m->n_out >= 0 ? m->n_out : m->dyn_max_out, n_out);
  }
  return 0;
}

__global__ void max_uint_kernel(const unsigned *x, long long n, unsigned *out) {
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  unsigned v = i < n ? x[i] : 0u;
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned o = __shfl_down_sync(0xffffffffu, v, off);
    v = o > v ? o : v;
  }
  if ((threadIdx.x & 31) == 0 && v > 0) {
    atomicMax(out, v);
  }
}

/* grow_history on a restart: lay the surviving entries of every trajectory's ring out
 * again in a ring of larger capacity (logical entry idx lives in slot idx % capacity) */
__global__ void ring_regrow_kernel(const double *old_ring, double *new_ring,
                                   const unsigned *count, long long n_traj, unsigned old_cap,
                                   unsigned new_cap, int hl) {
  const long long gid = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if (gid >= n_traj * old_cap) {
    return;
  }
  const long long traj = gid / old_cap;
  const unsigned e = (unsigned) (gid % old_cap);
  const unsigned c = count[traj];
  const unsigned used = c < old_cap ? c : old_cap;
  if (e >= used) {
    return;
  }
  const unsigned idx = c - used + e;
  const double *src = old_ring + ((size_t) traj * old_cap + idx % old_cap) * (size_t) hl;
  double *dst = new_ring + ((size_t) traj * new_cap + idx % new_cap) * (size_t) hl;
  for (int i = 0; i < hl; ++i) {
    dst[i] = src[i];
  }
}

/* hand-out order by the last run's step counts (dde_dopri_batch_order_by_last_run): a
   counting sort, most attempts first.  Bin = min(attempts, DDE_ORDER_BINS - 1). */
#define DDE_ORDER_BINS 4096
__global__ void order_hist_kernel(const int *stats, long long n, unsigned *hist) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int k = stats[4 * i + 1];
    k = k < 0 ? 0 : (k > DDE_ORDER_BINS - 1 ? DDE_ORDER_BINS - 1 : k);
    atomicAdd(&hist[k], 1u);
  }
}
/* one block: hist[k] becomes the number of trajectories in bins above k */
__global__ void order_scan_kernel(unsigned *hist) {
  __shared__ unsigned part[256];
  constexpr int PER = DDE_ORDER_BINS / 256;
  const int t = threadIdx.x; /* thread t owns bins [BINS - (t + 1) PER, BINS - t PER), top first */
  unsigned sum = 0;
  for (int j = 0; j < PER; ++j) {
    sum += hist[DDE_ORDER_BINS - 1 - (t * PER + j)];
  }
  part[t] = sum;
  __syncthreads();
  unsigned before = 0;
  for (int u = 0; u < t; ++u) {
    before += part[u];
  }
  for (int j = 0; j < PER; ++j) {
    const int k = DDE_ORDER_BINS - 1 - (t * PER + j);
    const unsigned c = hist[k];
    hist[k] = before;
    before += c;
  }
}
__global__ void order_scatter_kernel(const int *stats, long long n, unsigned *offset,
                                     unsigned *order) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int k = stats[4 * i + 1];
    k = k < 0 ? 0 : (k > DDE_ORDER_BINS - 1 ? DDE_ORDER_BINS - 1 : k);
    order[atomicAdd(&offset[k], 1u)] = (unsigned) i;
  }
}

__global__ void fill_int_kernel(int *x, long long n, int value) {
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if (i < n) {
    x[i] = value;
  }
}

template <class T>
void free_dev(T *&p) {
  if (p != nullptr) {
    cudaFree(p);
    p = nullptr;
  }
}

/* FP64 issue-rate probe: 8 independent register chains per thread, no memory
 * traffic.  fma = 1: DFMA (2 flop/instr); fma = 0: alternating DMUL / DADD
 * (1 flop/instr), the ceiling that applies to the FMA-free solver code. */
template <int FMA>
__global__ void fp64_peak_kernel(double *sink, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    if (FMA) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    } else {
      x0 = x0 * a; x1 = x1 * a; x2 = x2 * a; x3 = x3 * a;
      x4 = x4 * a; x5 = x5 * a; x6 = x6 * a; x7 = x7 * a;
      x0 = x0 + b; x1 = x1 + b; x2 = x2 + b; x3 = x3 + b;
      x4 = x4 + b; x5 = x5 + b; x6 = x6 + b; x7 = x7 + b;
    }
  }
  const double r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (r == 123.456) {
    sink[0] = r;
  }
}

} /* namespace */

extern "C" void dde_register_model(dde::ModelDesc *desc) {
  std::lock_guard<std::mutex> lock(g_models_mutex);
  desc->next = g_models;
  g_models = desc;
}

/* ---- handles ------------------------------------------------------------- */

struct dde_dopri_batch_s {
  dde::ModelDesc *model = nullptr;
  dde_dopri_control ctl;
  size_t n = 0, n_out = 0, n_parms = 0, n_traj = 0;
  size_t parms_stride = 0;
  size_t n_times = 0, n_tcrit = 0, nt = 0;
  size_t cap_times = 0, cap_tcrit = 0, cap_y_out = 0, cap_out = 0;
  int device = 0, device_sms = 0;
  int order = 5;
  int reset_code = 0; /* ERR_ZERO_TIME_DIFFERENCE / ERR_INCONSISTENT_TIME from reset */
  double sign = 1.0;
  int tcrit_idx0 = 0;
  bool have_times = false, have_y0 = false, ran = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int last_launches = 0;
  /* device memory */
  double *d_y0 = nullptr, *d_parms = nullptr, *d_times = nullptr, *d_tcrit = nullptr;
  double *d_y_out = nullptr, *d_out = nullptr, *d_t_last = nullptr, *d_step_size = nullptr;
  int *d_stats = nullptr, *d_code = nullptr;
  double *d_ring = nullptr;
  double *d_init_k1 = nullptr, *d_init_h = nullptr, *d_y_final = nullptr;
  unsigned *d_order = nullptr; /* [n_traj] the order trajectories are handed out in, when set */
  bool have_order = false;
  int *d_event = nullptr; /* [cap_event] event index per critical time, or none set */
  size_t cap_event = 0;
  bool have_events = false;
  unsigned *d_ring_count = nullptr;
  unsigned long long *d_next = nullptr;
  /* streamed download (dde_dopri_batch_run_download), allocated on first use */
  cudaStream_t copy_stream = nullptr;
  unsigned *d_chunk_done = nullptr;
  int *h_chunk_flag = nullptr; /* pinned + mapped */
  int *d_chunk_flag = nullptr; /* its device address */
  size_t n_chunks = 0;
  bool streaming = false;      /* the run in flight raises chunk flags */
};

struct dde_difeq_batch_s {
  dde::ModelDesc *model = nullptr;
  size_t n = 0, n_out = 0, n_parms = 0, n_rep = 0, n_history = 0;
  size_t parms_stride = 0;
  size_t n_steps = 0, nt = 0, cap_steps = 0, cap_y_out = 0, cap_out = 0;
  int return_initial = 1;
  int device = 0, device_sms = 0;
  bool have_steps = false, have_y0 = false, ran = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double *d_y0 = nullptr, *d_parms = nullptr, *d_y_out = nullptr, *d_out = nullptr;
  double *d_y_final = nullptr, *d_ring = nullptr;
  unsigned long long *d_steps = nullptr, *d_next = nullptr;
  int *d_code = nullptr;
  unsigned *d_ring_count = nullptr;
  bool restartable = false, kept_history = false;
  unsigned long long steps_first = 0, steps_last = 0; /* of the steps set last */
  std::vector<unsigned long long> h_steps;            /* ... and the steps themselves */
  bool continued = false; /* the history holds more than one run (difeq_continue) */
  unsigned long long step_origin = 0; /* first step of the handle's first run */
  unsigned long long last_step = 0;   /* obj->step after the last run */
};

extern "C" {

const char *dde_last_error(void) { return g_last_error.c_str(); }

const char *dde_version(void) { return "dde_b200 0.1.0 (sm_100a)"; }

void dde_dopri_control_default(dde_dopri_control *ctl, int32_t method) {
  memset(ctl, 0, sizeof *ctl);
  ctl->atol = 1e-6;
  ctl->rtol = 1e-6;
  ctl->step_size_min = 0.0;       /* R/dopri.R:296; clamped to DBL_EPSILON at use */
  ctl->step_size_max = INFINITY;  /* R/dopri.R:296; clamped to DBL_MAX at use */
  ctl->step_size_initial = 0.0;
  ctl->step_max_n = 100000;
  ctl->stiff_check = 0;
  ctl->n_history = 0;
  ctl->method = method;
  ctl->step_size_min_allow = 0;
  ctl->grow_history = 0;
  ctl->return_initial = 1;
}

size_t dde_model_count(void) {
  std::lock_guard<std::mutex> lock(g_models_mutex);
  size_t k = 0;
  for (dde::ModelDesc *m = g_models; m != nullptr; m = m->next) {
    ++k;
  }
  return k;
}

const char *dde_model_name(size_t i) {
  std::lock_guard<std::mutex> lock(g_models_mutex);
  size_t k = 0;
  for (dde::ModelDesc *m = g_models; m != nullptr; m = m->next, ++k) {
    if (k == i) {
      return m->name;
    }
  }
  return nullptr;
}

int dde_model_info(const char *model, int32_t *kind, size_t *n, size_t *n_parms, size_t *n_out) {
  std::lock_guard<std::mutex> lock(g_models_mutex);
  for (dde::ModelDesc *m = g_models; m != nullptr; m = m->next) {
    if (strcmp(m->name, model) == 0) {
      if (kind) *kind = m->kind;
      if (n) *n = (size_t) m->n;
      if (n_parms) *n_parms = m->n_parms >= 0 ? (size_t) m->n_parms : (size_t) -1;
      if (n_out) *n_out = m->n_out >= 0 ? (size_t) m->n_out : (size_t) -1;
      return 0;
    }
  }
  return fail("unknown model '%s'", model);
}

/* r_dopri_error, src/r_dopri.c:264-297 */
int dde_dopri_strerror(int32_t code, double t, uint64_t n_history, char *buf, size_t len) {
  switch (code) {
    case DDE_ERR_ZERO_TIME_DIFFERENCE:
      snprintf(buf, len, "Initialisation failure: Beginning and end times are the same");
      break;
    case DDE_ERR_INCONSISTENT_TIME:
      snprintf(buf, len, "Initialisation failure: Times have inconsistent sign");
      break;
    case DDE_ERR_TOO_MANY_STEPS:
      snprintf(buf, len, "Integration failure: too many steps (at t = %2.5f)", t);
      break;
    case DDE_ERR_STEP_SIZE_TOO_SMALL:
      snprintf(buf, len, "Integration failure: step size too small (at t = %2.5f)", t);
      break;
    case DDE_ERR_STEP_SIZE_VANISHED:
      snprintf(buf, len, "Integration failure: step size vanished (at t = %2.5f)", t);
      break;
    case DDE_ERR_YLAG_FAIL:
      if (n_history == 0) {
        snprintf(buf, len, "Integration failure: can't use ylag in model with no history");
      } else {
        snprintf(buf, len,
                 "Integration failure: did not find time in history (at t = %2.5f)", t);
      }
      break;
    case DDE_ERR_STIFF:
      snprintf(buf, len, "Integration failure: problem became stiff (at t = %2.5f)", t);
      break;
    case DDE_ERR_NONFINITE_DERIV: /* src/dopri.c:333 */
      snprintf(buf, len, "non-finite derivative at initial time");
      break;
    case DDE_ERR_DIFEQ_HISTORY: /* src/difeq.c:268 */
      snprintf(buf, len, "difeq failure: did not find step in history");
      break;
    case DDE_ERR_RESTART: /* src/r_dopri.c:216 */
      snprintf(buf, len, "Incorrect initial time on integration restart");
      break;
    case DDE_ERR_LAG_UNDECLARED: /* an extension: include/dde_b200/dde_model_def.h, M::lags */
      snprintf(buf, len,
               "Integration failure: the model made a history look-up it did not declare "
               "(at t = %2.5f)", t);
      break;
    case DDE_OK_COMPLETE:
    case DDE_OK_INTERRUPTED:
      snprintf(buf, len, "OK");
      return 0;
    default:
      snprintf(buf, len, "Integration failure: (code %d) [dde bug]", (int) code);
      break;
  }
  return 1;
}

double dde_pow_host(double x, double y) { return dde_pow(x, y); }
double dde_exp_host(double x) { return dde_exp(x); }
void dde_pow2_host(double x, double y1, double y2, double *out) {
  const DdePow2 r = dde_pow2(x, y1, y2);
  out[0] = r.a;
  out[1] = r.b;
}

/* TFLOP/s of the FP64 pipe on the current device (see fp64_peak_kernel). */
double dde_fp64_peak(int fma) {
  cudaDeviceProp prop;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    (void) cudaGetLastError();
    return -1.0;
  }
  double *sink = nullptr;
  if (cudaMalloc(&sink, sizeof(double)) != cudaSuccess) {
    (void) cudaGetLastError();
    return -1.0;
  }
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    if (fma) {
      fp64_peak_kernel<1><<<blocks, threads>>>(sink, iters, 1.0000001, 1e-9);
    } else {
      fp64_peak_kernel<0><<<blocks, threads>>>(sink, iters, 1.0000001, 1e-9);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) {
      best = ms;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  if (cudaGetLastError() != cudaSuccess) {
    return -1.0;
  }
  const double instr = (double) blocks * threads * (double) iters * 16.0;
  return instr * (fma ? 2.0 : 1.0) / (best * 1e-3) / 1e12;
}

void *dde_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    fail("cudaHostAlloc(%zu) failed", bytes);
    (void) cudaGetLastError();
    return nullptr;
  }
  return p;
}

void dde_host_free(void *p) {
  if (p != nullptr) {
    cudaFreeHost(p);
  }
}

/* ---- dopri ----------------------------------------------------------------- */

void dde_dopri_batch_free(dde_dopri_batch_t *b) {
  if (b == nullptr) {
    return;
  }
  cudaSetDevice(b->device);
  free_dev(b->d_y0);
  free_dev(b->d_parms);
  free_dev(b->d_times);
  free_dev(b->d_tcrit);
  free_dev(b->d_y_out);
  free_dev(b->d_out);
  free_dev(b->d_t_last);
  free_dev(b->d_step_size);
  free_dev(b->d_stats);
  free_dev(b->d_code);
  free_dev(b->d_ring);
  free_dev(b->d_init_k1);
  free_dev(b->d_init_h);
  free_dev(b->d_y_final);
  free_dev(b->d_event);
  free_dev(b->d_order);
  free_dev(b->d_ring_count);
  free_dev(b->d_next);
  free_dev(b->d_chunk_done);
  if (b->h_chunk_flag) cudaFreeHost(b->h_chunk_flag);
  if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

dde_dopri_batch_t *dde_dopri_batch_alloc(const char *model, const dde_dopri_control *ctl,
                                         size_t n, size_t n_out, size_t n_parms,
                                         size_t n_traj) {
  if (model == nullptr || ctl == nullptr) {
    fail("model and ctl must not be NULL");
    return nullptr;
  }
  dde::ModelDesc *m = find_model(model, 1, n);
  if (m == nullptr) {
    if (model_name_known(model)) {
      fail("model '%s' is not compiled for dopri with n = %zu", model, n);
    } else {
      fail("unknown model '%s'", model);
    }
    return nullptr;
  }
  if (check_model_shape(m, n, n_parms, n_out) != 0) {
    return nullptr;
  }
  if (ctl->method != DDE_DOPRI_5 && ctl->method != DDE_DOPRI_853) {
    fail("unknown method %d", (int) ctl->method);
    return nullptr;
  }
  if (n_traj == 0) {
    fail("n_traj must be positive");
    return nullptr;
  }
  if (ctl->n_history > 0x7fffffffULL) {
    fail("n_history too large");
    return nullptr;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    fail("no CUDA device available: dde_b200 has no CPU fallback");
    (void) cudaGetLastError();
    return nullptr;
  }
  dde_dopri_batch_t *b = new dde_dopri_batch_t();
  b->model = m;
  b->ctl = *ctl;
  /* src/r_dopri.c:160-161 */
  b->ctl.step_size_min = fmax(fabs(ctl->step_size_min), DBL_EPSILON);
  b->ctl.step_size_max = fmin(fabs(ctl->step_size_max), DBL_MAX);
  b->n = n;
  b->n_out = n_out;
  b->n_parms = n_parms;
  b->n_traj = n_traj;
  b->order = ctl->method == DDE_DOPRI_5 ? 5 : 8;
  cudaDeviceProp prop;
  bool ok = cudaGetDevice(&b->device) == cudaSuccess &&
            cudaGetDeviceProperties(&prop, b->device) == cudaSuccess;
  if (ok) {
    b->device_sms = prop.multiProcessorCount;
  }
  const size_t hl = (size_t) b->order * n + 2;
  ok = ok && cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreate(&b->ev0) == cudaSuccess && cudaEventCreate(&b->ev1) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_y0, sizeof(double) * n * n_traj) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_parms, sizeof(double) * (n_parms ? n_parms : 1) * n_traj) ==
                 cudaSuccess;
  ok = ok && cudaMalloc(&b->d_t_last, sizeof(double) * n_traj) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_step_size, sizeof(double) * n_traj) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_stats, sizeof(int) * 4 * n_traj) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_code, sizeof(int) * n_traj) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_ring_count, sizeof(unsigned) * n_traj) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_init_k1, sizeof(double) * n * n_traj) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_init_h, sizeof(double) * n_traj) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_y_final, sizeof(double) * n * n_traj) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_next, sizeof(unsigned long long)) == cudaSuccess;
  if (ok && ctl->n_history > 0) {
    const size_t bytes = sizeof(double) * hl * (size_t) ctl->n_history * n_traj;
    if (cudaMalloc(&b->d_ring, bytes) != cudaSuccess) {
      fail("cannot allocate %.2f GB of history rings (n_traj %zu x n_history %llu x %zu doubles): "
           "shard the batch or lower n_history",
           (double) bytes / 1e9, n_traj, (unsigned long long) ctl->n_history, hl);
      (void) cudaGetLastError();
      dde_dopri_batch_free(b);
      return nullptr;
    }
  }
  if (!ok) {
    cudaError_t err = cudaGetLastError();
    fail("device allocation failed (%s)", cudaGetErrorString(err));
    dde_dopri_batch_free(b);
    return nullptr;
  }
  return b;
}

} /* extern "C" */

namespace {
int dopri_upload_async(dde_dopri_batch_t *b, const double *y0, const double *parms,
                       size_t parms_stride);
}

extern "C" {

int dde_dopri_batch_upload(dde_dopri_batch_t *b, const double *y0, const double *parms,
                           size_t parms_stride) {
  const int rc = dopri_upload_async(b, y0, parms, parms_stride);
  if (rc != 0) {
    return rc;
  }
  /* the caller's buffers are the caller's again on return (pinned memory makes the
     copies truly asynchronous) */
  DDE_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}

} /* extern "C" */

namespace {
int dopri_upload_async(dde_dopri_batch_t *b, const double *y0, const double *parms,
                       size_t parms_stride) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  DDE_CUDA(cudaSetDevice(b->device));
  if (y0 != nullptr) {
    DDE_CUDA(cudaMemcpyAsync(b->d_y0, y0, sizeof(double) * b->n * b->n_traj,
                             cudaMemcpyHostToDevice, b->stream));
    b->have_y0 = true;
  }
  if (parms != nullptr && b->n_parms > 0) {
    if (parms_stride != 0 && parms_stride != b->n_parms) {
      return fail("parms_stride must be 0 (shared) or n_parms");
    }
    const size_t count = parms_stride == 0 ? b->n_parms : b->n_parms * b->n_traj;
    DDE_CUDA(cudaMemcpyAsync(b->d_parms, parms, sizeof(double) * count, cudaMemcpyHostToDevice,
                             b->stream));
    b->parms_stride = parms_stride;
  }
  return 0;
}
} /* namespace */

extern "C" {

int dde_dopri_batch_set_times(dde_dopri_batch_t *b, const double *times, size_t n_times,
                              const double *tcrit, size_t n_tcrit) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  if (times == nullptr || n_times < 2) {
    return fail("At least two times must be given"); /* src/r_dopri.c:51-53 */
  }
  if (n_times > 0x7fffffff || n_tcrit > 0x7fffffff) {
    return fail("too many times");
  }
  DDE_CUDA(cudaSetDevice(b->device));
  /* dopri_data_reset's checks, src/dopri.c:168-183 */
  b->reset_code = 0;
  if (times[n_times - 1] == times[0]) {
    b->reset_code = DDE_ERR_ZERO_TIME_DIFFERENCE;
    b->sign = 1.0;
  } else {
    b->sign = copysign(1.0, times[n_times - 1] - times[0]);
    for (size_t i = 0; i + 1 < n_times; ++i) {
      const double t0 = times[i], t1 = times[i + 1];
      if ((b->sign > 0 && t1 < t0) || (b->sign < 0 && t1 > t0)) {
        b->reset_code = DDE_ERR_INCONSISTENT_TIME;
        break;
      }
    }
  }
  /* critical times already behind the start are skipped, src/dopri.c:199-206 */
  b->tcrit_idx0 = 0;
  if (n_tcrit > 0) {
    const double t0 = b->sign * times[0];
    while ((size_t) b->tcrit_idx0 < n_tcrit && b->sign * tcrit[b->tcrit_idx0] <= t0) {
      b->tcrit_idx0++;
    }
  }
  if (n_times > b->cap_times) {
    free_dev(b->d_times);
    DDE_CUDA(cudaMalloc(&b->d_times, sizeof(double) * n_times));
    b->cap_times = n_times;
  }
  if (n_tcrit > b->cap_tcrit) {
    free_dev(b->d_tcrit);
    DDE_CUDA(cudaMalloc(&b->d_tcrit, sizeof(double) * n_tcrit));
    b->cap_tcrit = n_tcrit;
  }
  DDE_CUDA(cudaMemcpyAsync(b->d_times, times, sizeof(double) * n_times, cudaMemcpyHostToDevice,
                           b->stream));
  if (n_tcrit > 0) {
    DDE_CUDA(cudaMemcpyAsync(b->d_tcrit, tcrit, sizeof(double) * n_tcrit,
                             cudaMemcpyHostToDevice, b->stream));
  }
  /* the copies read caller memory that may go away */
  DDE_CUDA(cudaStreamSynchronize(b->stream));
  b->n_times = n_times;
  b->n_tcrit = n_tcrit;
  b->have_events = false; /* events belong to one set of critical times */
  b->nt = b->ctl.return_initial ? n_times : n_times - 1;
  const size_t need_y = b->n * b->nt * b->n_traj;
  const size_t need_o = b->n_out * b->nt * b->n_traj;
  if (need_y > b->cap_y_out) {
    free_dev(b->d_y_out);
    if (cudaMalloc(&b->d_y_out, sizeof(double) * need_y) != cudaSuccess) {
      (void) cudaGetLastError();
      b->cap_y_out = 0;
      return fail("cannot allocate %.2f GB for the output array [n x nt x n_traj]",
                  (double) (sizeof(double) * need_y) / 1e9);
    }
    b->cap_y_out = need_y;
  }
  if (need_o > b->cap_out) {
    free_dev(b->d_out);
    DDE_CUDA(cudaMalloc(&b->d_out, sizeof(double) * need_o));
    b->cap_out = need_o;
  }
  b->have_times = true;
  return 0;
}

} /* extern "C" */

namespace {

int dopri_run_impl(dde_dopri_batch_t *b, int restart) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  if (!b->have_times || !b->have_y0) {
    return fail("upload y0 and set times before running");
  }
  if (b->ctl.n_history > 0 && b->d_ring == nullptr) {
    return fail("the handle has no history rings (a failed grow_history): free it");
  }
  DDE_CUDA(cudaSetDevice(b->device));
  DDE_CUDA(cudaEventRecord(b->ev0, b->stream));
  /* the R glue hands the solver zero-filled result arrays, src/r_dopri.c:169 */
  DDE_CUDA(cudaMemsetAsync(b->d_y_out, 0, sizeof(double) * b->n * b->nt * b->n_traj, b->stream));
  if (b->n_out > 0) {
    DDE_CUDA(cudaMemsetAsync(b->d_out, 0, sizeof(double) * b->n_out * b->nt * b->n_traj,
                             b->stream));
  }
  b->last_launches = 0;
  if (b->reset_code != 0) {
    /* dopri_integrate returns right after reset, src/dopri.c:298-300 */
    DDE_CUDA(cudaMemsetAsync(b->d_stats, 0, sizeof(int) * 4 * b->n_traj, b->stream));
    if (!restart) {
      DDE_CUDA(cudaMemsetAsync(b->d_t_last, 0, sizeof(double) * b->n_traj, b->stream));
      DDE_CUDA(cudaMemsetAsync(b->d_step_size, 0, sizeof(double) * b->n_traj, b->stream));
      DDE_CUDA(cudaMemsetAsync(b->d_ring_count, 0, sizeof(unsigned) * b->n_traj, b->stream));
    }
    const int threads = 256;
    const long long blocks = ((long long) b->n_traj + threads - 1) / threads;
    fill_int_kernel<<<(unsigned) blocks, threads, 0, b->stream>>>(b->d_code, (long long) b->n_traj,
                                                                  b->reset_code);
    DDE_CUDA(cudaGetLastError());
    b->last_launches = 1;
  } else {
    DDE_CUDA(cudaMemsetAsync(b->d_next, 0, sizeof(unsigned long long), b->stream));
    dde::DopriArgs a;
    memset(&a, 0, sizeof a);
    a.n_traj = (long long) b->n_traj;
    a.n = (int) b->n;
    a.n_out = (int) b->n_out;
    a.n_parms = (int) b->n_parms;
    a.parms_stride = (long long) b->parms_stride;
    a.y0 = b->d_y0;
    a.parms = b->d_parms;
    a.times = b->d_times;
    a.tcrit = b->d_tcrit;
    a.n_times = (int) b->n_times;
    a.n_tcrit = (int) b->n_tcrit;
    a.tcrit_idx0 = b->tcrit_idx0;
    a.event = b->have_events ? b->d_event : nullptr;
    a.atol = b->ctl.atol;
    a.rtol = b->ctl.rtol;
    a.step_size_min = b->ctl.step_size_min;
    a.step_size_max = b->ctl.step_size_max;
    a.step_size_initial = b->ctl.step_size_initial;
    a.step_max_n = b->ctl.step_max_n;
    a.stiff_check = b->ctl.stiff_check;
    a.n_history = (int) b->ctl.n_history;
    a.step_size_min_allow = b->ctl.step_size_min_allow;
    a.return_initial = b->ctl.return_initial;
    a.sign = b->sign;
    a.y_out = b->d_y_out;
    a.out = b->d_out;
    a.stats = b->d_stats;
    a.code = b->d_code;
    a.t_last = b->d_t_last;
    a.step_size = b->d_step_size;
    a.ring = b->d_ring;
    a.init_k1 = b->d_init_k1;
    a.init_h = b->d_init_h;
    a.ring_count = b->d_ring_count;
    a.restart = restart;
    a.y_final = b->d_y_final;
    a.next_traj = b->d_next;
    a.order = b->have_order ? b->d_order : nullptr;
    if (b->streaming) {
      DDE_CUDA(cudaMemsetAsync(b->d_chunk_done, 0, sizeof(unsigned) * b->n_chunks, b->stream));
      a.chunk_done = b->d_chunk_done;
      a.chunk_flag = b->d_chunk_flag;
      a.chunk_shift = DDE_CHUNK_SHIFT;
    }
    int launches = 0;
    cudaError_t err = b->model->launch_dopri(b->ctl.method, a, b->device_sms, b->stream, &launches);
    if (err != cudaSuccess) {
      return fail("kernel launch failed for model '%s': %s", b->model->name,
                  cudaGetErrorString(err));
    }
    b->last_launches = launches;
  }
  DDE_CUDA(cudaEventRecord(b->ev1, b->stream));
  b->ran = true;
  return 0;
}

} /* namespace */

extern "C" {

/* grow_history (dopri_data_alloc's argument, src/dopri.c:92-95: a ring that
 * grows instead of overwriting).  Nothing is allocated inside a kernel: the run
 * is repeated with a larger ring until no trajectory's ring has wrapped -- the
 * results are then what an unbounded ring gives.  Blocking. */
static int dopri_run_grow(dde_dopri_batch_t *b) {
  for (;;) {
    const int rc = dopri_run_impl(b, 0);
    if (rc != 0) {
      return rc;
    }
    const size_t cap = (size_t) b->ctl.n_history;
    if (!b->ctl.grow_history || cap == 0 || b->reset_code != 0) {
      return 0;
    }
    unsigned *d_max = reinterpret_cast<unsigned *>(b->d_next); /* free once the run is over */
    DDE_CUDA(cudaMemsetAsync(d_max, 0, sizeof(unsigned), b->stream));
    const int threads = 256;
    const long long blocks = ((long long) b->n_traj + threads - 1) / threads;
    max_uint_kernel<<<(unsigned) blocks, threads, 0, b->stream>>>(b->d_ring_count,
                                                                  (long long) b->n_traj, d_max);
    DDE_CUDA(cudaGetLastError());
    unsigned most = 0;
    DDE_CUDA(cudaMemcpyAsync(&most, d_max, sizeof most, cudaMemcpyDeviceToHost, b->stream));
    DDE_CUDA(cudaStreamSynchronize(b->stream));
    if ((size_t) most <= cap) {
      return 0;
    }
    size_t bigger = (size_t) most > 2 * cap ? (size_t) most : 2 * cap;
    bigger = bigger > 0x7fffffffu ? 0x7fffffffu : bigger; /* the kernels index entries with int */
    const size_t hl = (size_t) b->order * b->n + 2;
    /* the old rings go first (the run is repeated from scratch, and both sizes together
       may not fit); a failure leaves the handle without rings, which dopri_run_impl
       refuses to run on until the rings can be allocated again */
    free_dev(b->d_ring);
    b->d_ring = nullptr;
    if (cudaMalloc(&b->d_ring, sizeof(double) * hl * bigger * b->n_traj) != cudaSuccess) {
      (void) cudaGetLastError();
      b->d_ring = nullptr;
      b->ran = false;
      /* back to the capacity asked for, so that a later run starts over from there */
      if (cudaMalloc(&b->d_ring, sizeof(double) * hl * cap * b->n_traj) != cudaSuccess) {
        (void) cudaGetLastError();
        b->d_ring = nullptr;
      }
      return fail("grow_history: cannot allocate %.2f GB of history rings (%zu entries per "
                  "trajectory)", (double) (sizeof(double) * hl * bigger * b->n_traj) / 1e9, bigger);
    }
    b->ctl.n_history = bigger;
  }
}

int dde_dopri_batch_run(dde_dopri_batch_t *b) {
  if (b != nullptr && b->ctl.grow_history) {
    return dopri_run_grow(b);
  }
  return dopri_run_impl(b, 0);
}

uint64_t dde_dopri_batch_n_history(dde_dopri_batch_t *b) {
  return b == nullptr ? 0 : b->ctl.n_history;
}

/* Run and download in one call, overlapping the two: see DopriArgs::chunk_done.
 * The copies of a chunk start as soon as the kernel reports it complete, on a
 * second stream; what is left when the kernel ends is copied then. */
int dde_dopri_batch_run_download(dde_dopri_batch_t *b, double *y_out, double *out,
                                 int32_t *stats, int32_t *code, double *t_last,
                                 double *step_size) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  if (b->ctl.grow_history) { /* the run may have to be repeated: no streaming */
    const int rc = dopri_run_grow(b);
    return rc ? rc : dde_dopri_batch_download(b, y_out, out, stats, code, t_last, step_size);
  }
  DDE_CUDA(cudaSetDevice(b->device));
  const size_t chunk = (size_t) 1 << DDE_CHUNK_SHIFT;
  const size_t n_chunks = (b->n_traj + chunk - 1) / chunk;
  if (b->copy_stream == nullptr) {
    DDE_CUDA(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
  }
  if (b->n_chunks != n_chunks) {
    free_dev(b->d_chunk_done);
    if (b->h_chunk_flag) {
      cudaFreeHost(b->h_chunk_flag);
      b->h_chunk_flag = nullptr;
    }
    DDE_CUDA(cudaMalloc(&b->d_chunk_done, sizeof(unsigned) * n_chunks));
    DDE_CUDA(cudaHostAlloc(&b->h_chunk_flag, sizeof(int) * n_chunks, cudaHostAllocMapped));
    DDE_CUDA(cudaHostGetDevicePointer(&b->d_chunk_flag, b->h_chunk_flag, 0));
    b->n_chunks = n_chunks;
  }
  /* no earlier run of this handle may still be raising flags */
  DDE_CUDA(cudaStreamSynchronize(b->stream));
  volatile int *flag = b->h_chunk_flag;
  for (size_t c = 0; c < n_chunks; ++c) {
    flag[c] = 0;
  }
  b->streaming = true;
  const int rc = dopri_run_impl(b, 0);
  b->streaming = false;
  if (rc != 0) {
    return rc;
  }
  std::vector<char> copied(n_chunks, 0);
  size_t left = n_chunks;
  auto copy_chunk = [&](size_t c) -> cudaError_t {
    const size_t j0 = c * chunk;
    const size_t nj = (j0 + chunk < b->n_traj ? j0 + chunk : b->n_traj) - j0;
    cudaError_t err = cudaSuccess;
    auto d2h = [&](void *dst, const void *src, size_t elem_bytes) {
      if (dst != nullptr && src != nullptr && err == cudaSuccess) {
        err = cudaMemcpyAsync((char *) dst + j0 * elem_bytes, (const char *) src + j0 * elem_bytes,
                              nj * elem_bytes, cudaMemcpyDeviceToHost, b->copy_stream);
      }
    };
    d2h(y_out, b->d_y_out, sizeof(double) * b->n * b->nt);
    d2h(b->n_out > 0 ? out : nullptr, b->d_out, sizeof(double) * b->n_out * b->nt);
    d2h(stats, b->d_stats, sizeof(int) * 4);
    d2h(code, b->d_code, sizeof(int));
    d2h(t_last, b->d_t_last, sizeof(double));
    d2h(step_size, b->d_step_size, sizeof(double));
    copied[c] = 1;
    --left;
    return err;
  };
  while (left > 0) {
    const cudaError_t q = cudaEventQuery(b->ev1);
    if (q != cudaSuccess && q != cudaErrorNotReady) {
      return fail("CUDA error while the batch was running: %s", cudaGetErrorString(q));
    }
    for (size_t c = 0; c < n_chunks; ++c) {
      if (!copied[c] && (q == cudaSuccess || flag[c] != 0)) {
        DDE_CUDA(copy_chunk(c));
      }
    }
    if (left > 0) {
      std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
  }
  DDE_CUDA(cudaStreamSynchronize(b->copy_stream));
  DDE_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}

/* is_event / events of dopri_integrate, src/dopri.h:170-171, src/r_dopri.c:77-103 */
int dde_dopri_batch_set_events(dde_dopri_batch_t *b, const int32_t *event, size_t n_tcrit) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  if (!b->have_times) {
    return fail("set the times (and critical times) before the events");
  }
  if (n_tcrit != b->n_tcrit) {
    return fail("events must name one entry per critical time (%zu), got %zu", b->n_tcrit,
                n_tcrit);
  }
  if (event == nullptr && n_tcrit > 0) {
    return fail("event must not be NULL when there are critical times");
  }
  bool any = false;
  for (size_t i = 0; i < n_tcrit; ++i) {
    if (event[i] >= b->model->n_events) {
      return fail("model '%s' has %d event functions, event %d requested", b->model->name,
                  b->model->n_events, (int) event[i]);
    }
    any = any || event[i] >= 0;
  }
  DDE_CUDA(cudaSetDevice(b->device));
  if (any) {
    if (n_tcrit > b->cap_event) { /* one buffer per handle, grown when needed */
      free_dev(b->d_event);
      b->d_event = nullptr;
      b->cap_event = 0;
      DDE_CUDA(cudaMalloc(&b->d_event, sizeof(int) * n_tcrit));
      b->cap_event = n_tcrit;
    }
    DDE_CUDA(cudaMemcpyAsync(b->d_event, event, sizeof(int) * n_tcrit, cudaMemcpyHostToDevice,
                             b->stream));
    DDE_CUDA(cudaStreamSynchronize(b->stream));
  }
  b->have_events = any;
  return 0;
}

/* The order in which the persistent threads take trajectories (default: index order).
 * Results do not depend on it -- every trajectory is integrated exactly as before -- but
 * the time of a batch does: it ends when its last trajectory does, so handing out the
 * expensive ones first shortens the tail in which most lanes have run out of work. */
int dde_dopri_batch_set_order(dde_dopri_batch_t *b, const uint32_t *order) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  if (order == nullptr) {
    b->have_order = false;
    return 0;
  }
  if (b->n_traj > 0xffffffffULL) {
    return fail("a hand-out order holds 32-bit indices");
  }
  std::vector<bool> seen(b->n_traj, false);
  for (size_t i = 0; i < b->n_traj; ++i) {
    if (order[i] >= b->n_traj || seen[order[i]]) {
      return fail("order must be a permutation of 0 .. n_traj - 1 (entry %zu)", i);
    }
    seen[order[i]] = true;
  }
  DDE_CUDA(cudaSetDevice(b->device));
  if (b->d_order == nullptr) {
    DDE_CUDA(cudaMalloc(&b->d_order, sizeof(unsigned) * (b->n_traj + DDE_ORDER_BINS)));
  }
  DDE_CUDA(cudaMemcpyAsync(b->d_order, order, sizeof(unsigned) * b->n_traj, cudaMemcpyHostToDevice,
                           b->stream));
  DDE_CUDA(cudaStreamSynchronize(b->stream));
  b->have_order = true;
  return 0;
}

/* ... longest first, by the number of step attempts each trajectory made in the LAST run of
 * this handle: for a handle that is run again on the same or on slowly changing inputs (a
 * fitting loop, a scan refined around the last one), where the last run's costs predict the
 * next run's.  Built on the device (a counting sort of the statistics), asynchronous on the
 * handle's stream.  Stays in force until dde_dopri_batch_set_order(b, NULL). */
int dde_dopri_batch_order_by_last_run(dde_dopri_batch_t *b) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  if (!b->ran) {
    return fail("no run of this handle to take the order from");
  }
  if (b->n_traj > 0xffffffffULL) {
    return fail("a hand-out order holds 32-bit indices");
  }
  DDE_CUDA(cudaSetDevice(b->device));
  if (b->d_order == nullptr) {
    DDE_CUDA(cudaMalloc(&b->d_order, sizeof(unsigned) * (b->n_traj + DDE_ORDER_BINS)));
  }
  unsigned *hist = b->d_order + b->n_traj;
  DDE_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned) * DDE_ORDER_BINS, b->stream));
  const int threads = 256;
  const unsigned blocks = (unsigned) (((long long) b->n_traj + threads - 1) / threads);
  order_hist_kernel<<<blocks, threads, 0, b->stream>>>(b->d_stats, (long long) b->n_traj, hist);
  order_scan_kernel<<<1, 256, 0, b->stream>>>(hist);
  order_scatter_kernel<<<blocks, threads, 0, b->stream>>>(b->d_stats, (long long) b->n_traj, hist,
                                                         b->d_order);
  DDE_CUDA(cudaGetLastError());
  b->have_order = true;
  return 0;
}

/* dopri_continue, R/dopri.R:469-511 + src/r_dopri.c:193-262 */
int dde_dopri_batch_continue(dde_dopri_batch_t *b, const double *y, const double *times,
                             size_t n_times, const double *tcrit, size_t n_tcrit) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  if (!b->ran) {
    return fail("nothing has been run on this handle: there is nothing to continue");
  }
  if (times == nullptr || n_times < 2) {
    return fail("At least two times must be given"); /* src/r_dopri.c:211-213 */
  }
  if (b->reset_code == 0 && b->sign != copysign(1.0, times[n_times - 1] - times[0])) {
    return fail("Incorrect sign for the times"); /* src/r_dopri.c:218-220 */
  }
  DDE_CUDA(cudaSetDevice(b->device));
  const size_t nt = b->n_traj, n = b->n;
  const size_t hl = (size_t) b->order * n + 2;
  const bool grow = b->ctl.grow_history && b->ctl.n_history > 0 && b->reset_code == 0;
  /* grow_history: the continuation overwrites what it starts from, so keep a copy
     to start again from if a ring wraps (blocking, like a growing first run) */
  double *kept_ring = nullptr, *kept_t_last = nullptr, *kept_step = nullptr, *kept_y_final = nullptr;
  int *kept_code = nullptr;
  unsigned *kept_count = nullptr;
  const size_t snap_cap = (size_t) b->ctl.n_history;
  auto release = [&]() {
    free_dev(kept_ring);
    free_dev(kept_t_last);
    free_dev(kept_step);
    free_dev(kept_y_final);
    free_dev(kept_code);
    free_dev(kept_count);
  };
  if (grow) {
    DDE_CUDA(cudaStreamSynchronize(b->stream));
    bool ok = cudaMalloc(&kept_ring, sizeof(double) * hl * snap_cap * nt) == cudaSuccess &&
              cudaMalloc(&kept_t_last, sizeof(double) * nt) == cudaSuccess &&
              cudaMalloc(&kept_step, sizeof(double) * nt) == cudaSuccess &&
              cudaMalloc(&kept_y_final, sizeof(double) * n * nt) == cudaSuccess &&
              cudaMalloc(&kept_code, sizeof(int) * nt) == cudaSuccess &&
              cudaMalloc(&kept_count, sizeof(unsigned) * nt) == cudaSuccess;
    ok = ok && cudaMemcpy(kept_ring, b->d_ring, sizeof(double) * hl * snap_cap * nt,
                          cudaMemcpyDeviceToDevice) == cudaSuccess &&
         cudaMemcpy(kept_t_last, b->d_t_last, sizeof(double) * nt, cudaMemcpyDeviceToDevice) ==
             cudaSuccess &&
         cudaMemcpy(kept_step, b->d_step_size, sizeof(double) * nt, cudaMemcpyDeviceToDevice) ==
             cudaSuccess &&
         cudaMemcpy(kept_y_final, b->d_y_final, sizeof(double) * n * nt, cudaMemcpyDeviceToDevice) ==
             cudaSuccess &&
         cudaMemcpy(kept_code, b->d_code, sizeof(int) * nt, cudaMemcpyDeviceToDevice) == cudaSuccess &&
         cudaMemcpy(kept_count, b->d_ring_count, sizeof(unsigned) * nt, cudaMemcpyDeviceToDevice) ==
             cudaSuccess;
    if (!ok) {
      (void) cudaGetLastError();
      release();
      return fail("grow_history: cannot allocate the restart copy of the history rings "
                  "(%.2f GB)", (double) (sizeof(double) * hl * snap_cap * nt) / 1e9);
    }
  }
  for (;;) {
    /* the times first: a refused set of times leaves the state as it was */
    int rc = dde_dopri_batch_set_times(b, times, n_times, tcrit, n_tcrit);
    if (rc == 0) {
      cudaError_t e1;
      if (y != nullptr) { /* a new state: also the pre-history from now on, src/dopri.c:157 */
        e1 = cudaMemcpyAsync(b->d_y0, y, sizeof(double) * n * nt, cudaMemcpyHostToDevice,
                             b->stream);
        /* the caller's buffer is the caller's again on return */
        e1 = e1 != cudaSuccess ? e1 : cudaStreamSynchronize(b->stream);
      } else {
        e1 = cudaMemcpyAsync(b->d_y0, b->d_y_final, sizeof(double) * n * nt,
                             cudaMemcpyDeviceToDevice, b->stream);
      }
      if (e1 != cudaSuccess) {
        release();
        return fail("CUDA error: %s", cudaGetErrorString(e1));
      }
    }
    rc = rc ? rc : dopri_run_impl(b, 1);
    if (rc != 0 || !grow) {
      release();
      return rc;
    }
    const size_t cap = (size_t) b->ctl.n_history;
    unsigned *d_max = reinterpret_cast<unsigned *>(b->d_next);
    unsigned most = 0;
    cudaError_t err = cudaMemsetAsync(d_max, 0, sizeof(unsigned), b->stream);
    const int threads = 256;
    max_uint_kernel<<<(unsigned) (((long long) nt + threads - 1) / threads), threads, 0,
                      b->stream>>>(b->d_ring_count, (long long) nt, d_max);
    err = err != cudaSuccess ? err : cudaGetLastError();
    err = err != cudaSuccess ? err
                             : cudaMemcpyAsync(&most, d_max, sizeof most, cudaMemcpyDeviceToHost,
                                               b->stream);
    err = err != cudaSuccess ? err : cudaStreamSynchronize(b->stream);
    if (err != cudaSuccess) {
      release();
      return fail("CUDA error: %s", cudaGetErrorString(err));
    }
    if ((size_t) most <= cap) {
      release();
      return 0;
    }
    /* a ring wrapped: back to the state before the continuation, in larger rings */
    size_t bigger = (size_t) most > 2 * cap ? (size_t) most : 2 * cap;
    bigger = bigger > 0x7fffffffu ? 0x7fffffffu : bigger;
    double *new_ring = nullptr;
    if (cudaMalloc(&new_ring, sizeof(double) * hl * bigger * nt) != cudaSuccess) {
      (void) cudaGetLastError();
      /* put the state the continuation started from back (rings of the old capacity: the
         kept copy), so that the handle is what it was before the call */
      cudaError_t e2 = cudaMemcpyAsync(b->d_ring, kept_ring, sizeof(double) * hl * snap_cap * nt,
                                       cudaMemcpyDeviceToDevice, b->stream);
      e2 = e2 != cudaSuccess ? e2 : cudaMemcpyAsync(b->d_t_last, kept_t_last, sizeof(double) * nt,
                                                    cudaMemcpyDeviceToDevice, b->stream);
      e2 = e2 != cudaSuccess ? e2 : cudaMemcpyAsync(b->d_step_size, kept_step, sizeof(double) * nt,
                                                    cudaMemcpyDeviceToDevice, b->stream);
      e2 = e2 != cudaSuccess ? e2 : cudaMemcpyAsync(b->d_y_final, kept_y_final,
                                                    sizeof(double) * n * nt,
                                                    cudaMemcpyDeviceToDevice, b->stream);
      e2 = e2 != cudaSuccess ? e2 : cudaMemcpyAsync(b->d_code, kept_code, sizeof(int) * nt,
                                                    cudaMemcpyDeviceToDevice, b->stream);
      e2 = e2 != cudaSuccess ? e2 : cudaMemcpyAsync(b->d_ring_count, kept_count,
                                                    sizeof(unsigned) * nt,
                                                    cudaMemcpyDeviceToDevice, b->stream);
      e2 = e2 != cudaSuccess ? e2 : cudaStreamSynchronize(b->stream);
      if (e2 != cudaSuccess || b->ctl.n_history != snap_cap) {
        (void) cudaGetLastError();
        b->ran = false; /* the state could not be restored: nothing to continue from */
      }
      release();
      return fail("grow_history: cannot allocate %.2f GB of history rings (%zu entries per "
                  "trajectory)", (double) (sizeof(double) * hl * bigger * nt) / 1e9, bigger);
    }
    free_dev(b->d_ring);
    b->d_ring = new_ring;
    const long long cells = (long long) nt * (long long) snap_cap;
    ring_regrow_kernel<<<(unsigned) ((cells + threads - 1) / threads), threads, 0, b->stream>>>(
        kept_ring, b->d_ring, kept_count, (long long) nt, (unsigned) snap_cap, (unsigned) bigger,
        (int) hl);
    err = cudaGetLastError();
    err = err != cudaSuccess ? err : cudaMemcpyAsync(b->d_t_last, kept_t_last, sizeof(double) * nt,
                                                     cudaMemcpyDeviceToDevice, b->stream);
    err = err != cudaSuccess ? err : cudaMemcpyAsync(b->d_step_size, kept_step, sizeof(double) * nt,
                                                     cudaMemcpyDeviceToDevice, b->stream);
    err = err != cudaSuccess ? err : cudaMemcpyAsync(b->d_y_final, kept_y_final,
                                                     sizeof(double) * n * nt,
                                                     cudaMemcpyDeviceToDevice, b->stream);
    err = err != cudaSuccess ? err : cudaMemcpyAsync(b->d_code, kept_code, sizeof(int) * nt,
                                                     cudaMemcpyDeviceToDevice, b->stream);
    err = err != cudaSuccess ? err : cudaMemcpyAsync(b->d_ring_count, kept_count,
                                                     sizeof(unsigned) * nt,
                                                     cudaMemcpyDeviceToDevice, b->stream);
    if (err != cudaSuccess) {
      release();
      return fail("CUDA error: %s", cudaGetErrorString(err));
    }
    b->ctl.n_history = bigger;
  }
}

/* dopri_data_copy, src/dopri.c:92-134: an independent handle with the same
   state (rings included), e.g. to continue one solve in two directions */
dde_dopri_batch_t *dde_dopri_batch_copy(dde_dopri_batch_t *src) {
  if (src == nullptr) {
    fail("NULL handle");
    return nullptr;
  }
  if (cudaSetDevice(src->device) != cudaSuccess || cudaStreamSynchronize(src->stream) != cudaSuccess) {
    fail("CUDA error in dde_dopri_batch_copy");
    (void) cudaGetLastError();
    return nullptr;
  }
  dde_dopri_batch_t *b = dde_dopri_batch_alloc(src->model->name, &src->ctl, src->n, src->n_out,
                                               src->n_parms, src->n_traj);
  if (b == nullptr) {
    return nullptr;
  }
  b->ctl = src->ctl; /* already clamped */
  b->parms_stride = src->parms_stride;
  b->have_y0 = src->have_y0;
  b->sign = src->sign;
  b->reset_code = src->reset_code;
  const size_t nt = src->n_traj, n = src->n;
  const size_t hl = (size_t) src->order * n + 2;
  cudaError_t err = cudaSuccess;
#define DDE_COPY(field, bytes)                                                              \
  if (err == cudaSuccess && src->field != nullptr && (bytes) > 0) {                         \
    err = cudaMemcpy(b->field, src->field, (bytes), cudaMemcpyDeviceToDevice);              \
  }
  DDE_COPY(d_y0, sizeof(double) * n * nt)
  DDE_COPY(d_parms, sizeof(double) * (src->n_parms ? src->n_parms : 1) * nt)
  DDE_COPY(d_y_final, sizeof(double) * n * nt)
  DDE_COPY(d_t_last, sizeof(double) * nt)
  DDE_COPY(d_step_size, sizeof(double) * nt)
  DDE_COPY(d_stats, sizeof(int) * 4 * nt)
  DDE_COPY(d_code, sizeof(int) * nt)
  DDE_COPY(d_ring_count, sizeof(unsigned) * nt)
  DDE_COPY(d_ring, sizeof(double) * hl * (size_t) src->ctl.n_history * nt)
#undef DDE_COPY
  if (err == cudaSuccess && src->have_times) {
    /* results of the last run stay readable through the copy as well */
    std::vector<double> times(src->n_times), tcrit(src->n_tcrit ? src->n_tcrit : 1);
    err = cudaMemcpy(times.data(), src->d_times, sizeof(double) * src->n_times,
                     cudaMemcpyDeviceToHost);
    if (err == cudaSuccess && src->n_tcrit > 0) {
      err = cudaMemcpy(tcrit.data(), src->d_tcrit, sizeof(double) * src->n_tcrit,
                       cudaMemcpyDeviceToHost);
    }
    if (err == cudaSuccess &&
        dde_dopri_batch_set_times(b, times.data(), src->n_times, tcrit.data(), src->n_tcrit) != 0) {
      dde_dopri_batch_free(b);
      return nullptr;
    }
    if (err == cudaSuccess) {
      err = cudaMemcpy(b->d_y_out, src->d_y_out, sizeof(double) * n * src->nt * nt,
                       cudaMemcpyDeviceToDevice);
    }
    if (err == cudaSuccess && src->n_out > 0) {
      err = cudaMemcpy(b->d_out, src->d_out, sizeof(double) * src->n_out * src->nt * nt,
                       cudaMemcpyDeviceToDevice);
    }
  }
  if (err != cudaSuccess) {
    fail("CUDA error in dde_dopri_batch_copy: %s", cudaGetErrorString(err));
    (void) cudaGetLastError();
    dde_dopri_batch_free(b);
    return nullptr;
  }
  b->ran = src->ran;
  return b;
}

int dde_dopri_batch_sync(dde_dopri_batch_t *b) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  DDE_CUDA(cudaSetDevice(b->device));
  DDE_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}

int dde_dopri_batch_download(dde_dopri_batch_t *b, double *y_out, double *out, int32_t *stats,
                             int32_t *code, double *t_last, double *step_size) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  if (!b->ran) {
    return fail("nothing has been run on this handle");
  }
  DDE_CUDA(cudaSetDevice(b->device));
  if (y_out) {
    DDE_CUDA(cudaMemcpyAsync(y_out, b->d_y_out, sizeof(double) * b->n * b->nt * b->n_traj,
                             cudaMemcpyDeviceToHost, b->stream));
  }
  if (out && b->n_out > 0) {
    DDE_CUDA(cudaMemcpyAsync(out, b->d_out, sizeof(double) * b->n_out * b->nt * b->n_traj,
                             cudaMemcpyDeviceToHost, b->stream));
  }
  if (stats) {
    DDE_CUDA(cudaMemcpyAsync(stats, b->d_stats, sizeof(int) * 4 * b->n_traj,
                             cudaMemcpyDeviceToHost, b->stream));
  }
  if (code) {
    DDE_CUDA(cudaMemcpyAsync(code, b->d_code, sizeof(int) * b->n_traj, cudaMemcpyDeviceToHost,
                             b->stream));
  }
  if (t_last) {
    DDE_CUDA(cudaMemcpyAsync(t_last, b->d_t_last, sizeof(double) * b->n_traj,
                             cudaMemcpyDeviceToHost, b->stream));
  }
  if (step_size) {
    DDE_CUDA(cudaMemcpyAsync(step_size, b->d_step_size, sizeof(double) * b->n_traj,
                             cudaMemcpyDeviceToHost, b->stream));
  }
  DDE_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}

/* Results of trajectories first .. first + count - 1 only (host arrays sized for
 * `count` trajectories).  Synchronises. */
int dde_dopri_batch_download_range(dde_dopri_batch_t *b, size_t first, size_t count,
                                   double *y_out, double *out, int32_t *stats, int32_t *code,
                                   double *t_last, double *step_size) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  if (!b->ran) {
    return fail("nothing has been run on this handle");
  }
  if (first > b->n_traj || count > b->n_traj - first) {
    return fail("trajectory range %zu + %zu exceeds the batch (%zu)", first, count, b->n_traj);
  }
  DDE_CUDA(cudaSetDevice(b->device));
  cudaError_t err = cudaSuccess;
  auto d2h = [&](void *dst, const void *src, size_t elem_bytes) {
    if (dst != nullptr && src != nullptr && err == cudaSuccess && count > 0) {
      err = cudaMemcpyAsync(dst, (const char *) src + first * elem_bytes, count * elem_bytes,
                            cudaMemcpyDeviceToHost, b->stream);
    }
  };
  d2h(y_out, b->d_y_out, sizeof(double) * b->n * b->nt);
  d2h(b->n_out > 0 ? out : nullptr, b->d_out, sizeof(double) * b->n_out * b->nt);
  d2h(stats, b->d_stats, sizeof(int) * 4);
  d2h(code, b->d_code, sizeof(int));
  d2h(t_last, b->d_t_last, sizeof(double));
  d2h(step_size, b->d_step_size, sizeof(double));
  DDE_CUDA(err);
  DDE_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}

double dde_dopri_batch_last_ms(dde_dopri_batch_t *b) {
  if (b == nullptr || !b->ran) {
    return -1.0;
  }
  cudaSetDevice(b->device);
  if (cudaEventSynchronize(b->ev1) != cudaSuccess) {
    (void) cudaGetLastError();
    return -1.0;
  }
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, b->ev0, b->ev1) != cudaSuccess) {
    (void) cudaGetLastError();
    return -1.0;
  }
  return (double) ms;
}

int dde_dopri_batch_last_launches(dde_dopri_batch_t *b) {
  return b == nullptr ? 0 : b->last_launches;
}

long dde_dopri_batch_history(dde_dopri_batch_t *b, size_t j, double *history,
                             size_t max_entries) {
  if (b == nullptr || !b->ran || j >= b->n_traj) {
    fail("bad handle or trajectory index");
    return -1;
  }
  if (cudaSetDevice(b->device) != cudaSuccess || cudaStreamSynchronize(b->stream) != cudaSuccess) {
    fail("CUDA error in history export");
    return -1;
  }
  const size_t cap = (size_t) b->ctl.n_history;
  if (cap == 0) {
    return 0;
  }
  unsigned count = 0;
  if (cudaMemcpy(&count, b->d_ring_count + j, sizeof count, cudaMemcpyDeviceToHost) !=
      cudaSuccess) {
    fail("CUDA error in history export");
    return -1;
  }
  const size_t hl = (size_t) b->order * b->n + 2;
  const size_t used = count < cap ? count : cap;
  if (history != nullptr && used > 0) {
    std::vector<double> ring(cap * hl);
    if (cudaMemcpy(ring.data(), b->d_ring + j * cap * hl, sizeof(double) * cap * hl,
                   cudaMemcpyDeviceToHost) != cudaSuccess) {
      fail("CUDA error in history export");
      return -1;
    }
    /* oldest first, like ring_buffer_read from the tail (src/r_dopri.c:401-404) */
    const size_t first = count - used;
    for (size_t e = 0; e < used && e < max_entries; ++e) {
      const size_t slot = (first + e) % cap;
      memcpy(history + e * hl, ring.data() + slot * hl, sizeof(double) * hl);
    }
  }
  return (long) used;
}

void *dde_dopri_batch_device_ptr(dde_dopri_batch_t *b, int which) {
  if (b == nullptr) {
    return nullptr;
  }
  switch (which) {
    case 0: return b->d_y_out;
    case 1: return b->d_out;
    case 2: return b->d_stats;
    case 3: return b->d_code;
    case 4: return b->d_t_last;
    case 5: return b->d_step_size;
    case 6: return b->d_y0;
    case 7: return b->d_parms;
    default: return nullptr;
  }
}

int dde_dopri_batch(const char *model, const dde_dopri_control *ctl, size_t n, size_t n_out,
                    size_t n_traj, const double *y0, const double *parms, size_t n_parms,
                    size_t parms_stride, const double *times, size_t n_times,
                    const double *tcrit, size_t n_tcrit, double *y_out, double *out,
                    int32_t *stats, int32_t *code, double *t_last, double *step_size) {
  if (y0 == nullptr) {
    return fail("y0 must not be NULL");
  }
  if (n_parms > 0 && parms == nullptr) {
    return fail("parms must not be NULL when n_parms > 0");
  }
  dde_dopri_batch_t *b = dde_dopri_batch_alloc(model, ctl, n, n_out, n_parms, n_traj);
  if (b == nullptr) {
    return 1;
  }
  int rc = dde_dopri_batch_upload(b, y0, parms, parms_stride);
  rc = rc ? rc : dde_dopri_batch_set_times(b, times, n_times, tcrit, n_tcrit);
  rc = rc ? rc : dde_dopri_batch_run_download(b, y_out, out, stats, code, t_last, step_size);
  dde_dopri_batch_free(b);
  return rc;
}

/* ---- dopri_interpolate ------------------------------------------------------ */

} /* extern "C" */

namespace {

/* One thread per (query time, trajectory): locate the entry like R's
 * findInterval (R/dopri.R:602) and evaluate all n components. */
__global__ void interpolate_kernel(const double *history, int hl, int n_entries, int n,
                                   long long n_traj, const double *t, int n_t, double *y) {
  const long long gid = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if (gid >= n_traj * n_t) {
    return;
  }
  const long long j = gid / n_t;
  const int q = (int) (gid % n_t);
  const int order = (hl - 2) / n;
  const double *h = history + (size_t) j * (size_t) n_entries * (size_t) hl;
  const double tq = t[q];
  const int it = hl - 2;
  int lo = 0, hi = n_entries; /* last entry with time <= tq */
  while (hi - lo > 1) {
    const int mid = lo + (hi - lo) / 2;
    if (h[(size_t) mid * hl + it] <= tq) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
  const double *e = h + (size_t) lo * hl;
  const double theta = (tq - e[it]) / e[it + 1];
  const double theta1 = 1.0 - theta;
  double *dst = y + (size_t) j * (size_t) n * (size_t) n_t;
  for (int i = 0; i < n; ++i) {
    dst[(size_t) i * n_t + q] = order == 5 ? dde::interp5(n, theta, theta1, e + i)
                                           : dde::interp853(n, theta, theta1, e + i);
  }
}

/* The same for trajectories whose history is still where the solver left it: the rings of
 * a batch handle, read in place (logical entry e of trajectory j in slot e % cap; the
 * oldest surviving one is count - min(count, cap)).  The range check of R/dopri.R:596-600
 * is made here too: a time outside [first entry's start, last entry's end] gives NaN and
 * records the lowest offending trajectory in *bad. */
__global__ void interpolate_rings_kernel(const double *ring, const unsigned *ring_count, int cap,
                                         int hl, int n, long long n_traj, const double *t,
                                         int n_t, double *y, unsigned long long *bad) {
  const long long gid = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if (gid >= n_traj * n_t) {
    return;
  }
  const long long j = gid / n_t;
  const int q = (int) (gid % n_t);
  const int order = (hl - 2) / n;
  const int it = hl - 2;
  const double *r = ring + (size_t) j * (size_t) cap * (size_t) hl;
  const unsigned count = ring_count[j];
  const unsigned used = count < (unsigned) cap ? count : (unsigned) cap;
  const double tq = t[q];
  double *dst = y + (size_t) j * (size_t) n * (size_t) n_t;
#define DDE_RING_ENTRY(e) (r + (size_t) ((e) % (unsigned) cap) * (size_t) hl)
  bool inside = used > 0;
  if (inside) {
    const double *le = DDE_RING_ENTRY(count - 1);
    inside = !(tq < DDE_RING_ENTRY(count - used)[it] || tq > le[it] + le[it + 1]);
  }
  if (!inside) {
    atomicMin(bad, (unsigned long long) j);
    for (int i = 0; i < n; ++i) {
      dst[(size_t) i * n_t + q] = dde::na_real();
    }
    return;
  }
  unsigned lo = count - used, hi = count; /* last entry with time <= tq */
  while (hi - lo > 1) {
    const unsigned mid = lo + (hi - lo) / 2;
    if (DDE_RING_ENTRY(mid)[it] <= tq) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
  const double *e = DDE_RING_ENTRY(lo);
#undef DDE_RING_ENTRY
  const double theta = (tq - e[it]) / e[it + 1];
  const double theta1 = 1.0 - theta;
  for (int i = 0; i < n; ++i) {
    dst[(size_t) i * n_t + q] = order == 5 ? dde::interp5(n, theta, theta1, e + i)
                                           : dde::interp853(n, theta, theta1, e + i);
  }
}

} /* namespace */

extern "C" {

int dde_dopri_interpolate_batch(const double *history, size_t history_len, size_t n_entries,
                                size_t n, size_t n_traj, const double *t, size_t n_t,
                                double *y) {
  if (history == nullptr || t == nullptr || y == nullptr) {
    return fail("NULL argument");
  }
  if (n == 0 || history_len < 2 || (history_len - 2) % n != 0) {
    return fail("Corrupt history object: incorrect number of rows"); /* R/dopri.R:590-594 */
  }
  const size_t order = (history_len - 2) / n;
  if (order != 5 && order != 8) {
    return fail("Corrupt history object: incorrect number of rows");
  }
  if (n_entries == 0 || n_t == 0 || n_traj == 0) {
    return fail("empty history or no times");
  }
  /* range check, R/dopri.R:596-600 */
  double tmin = t[0], tmax = t[0];
  for (size_t q = 1; q < n_t; ++q) {
    tmin = fmin(tmin, t[q]);
    tmax = fmax(tmax, t[q]);
  }
  for (size_t j = 0; j < n_traj; ++j) {
    const double *h = history + j * n_entries * history_len;
    const double first = h[history_len - 2];
    const double *le = h + (n_entries - 1) * history_len;
    const double end = le[history_len - 2] + le[history_len - 1];
    if (tmin < first || tmax > end) {
      return fail("Time falls outside of range of known history [%g, %g]", first, end);
    }
  }
  double *d_h = nullptr, *d_t = nullptr, *d_y = nullptr;
  const size_t hb = sizeof(double) * history_len * n_entries * n_traj;
  const size_t yb = sizeof(double) * n_t * n * n_traj;
  cudaError_t err = cudaMalloc(&d_h, hb);
  err = err ? err : cudaMalloc(&d_t, sizeof(double) * n_t);
  err = err ? err : cudaMalloc(&d_y, yb);
  err = err ? err : cudaMemcpy(d_h, history, hb, cudaMemcpyHostToDevice);
  err = err ? err : cudaMemcpy(d_t, t, sizeof(double) * n_t, cudaMemcpyHostToDevice);
  if (err == cudaSuccess) {
    const long long total = (long long) n_traj * (long long) n_t;
    const int threads = 128;
    interpolate_kernel<<<(unsigned) ((total + threads - 1) / threads), threads>>>(
        d_h, (int) history_len, (int) n_entries, (int) n, (long long) n_traj, d_t, (int) n_t, d_y);
    err = cudaGetLastError();
  }
  err = err ? err : cudaMemcpy(y, d_y, yb, cudaMemcpyDeviceToHost);
  cudaFree(d_h);
  cudaFree(d_t);
  cudaFree(d_y);
  if (err != cudaSuccess) {
    (void) cudaGetLastError();
    return fail("CUDA error in dde_dopri_interpolate_batch: %s", cudaGetErrorString(err));
  }
  return 0;
}

int dde_dopri_batch_interpolate(dde_dopri_batch_t *b, const double *t, size_t n_t, double *y) {
  if (b == nullptr || t == nullptr || y == nullptr) {
    return fail("NULL argument");
  }
  if (!b->ran) {
    return fail("nothing has been run on this handle: there is no history");
  }
  if (b->ctl.n_history == 0) {
    return fail("the handle keeps no history (n_history = 0)");
  }
  if (n_t == 0) {
    return fail("empty history or no times");
  }
  DDE_CUDA(cudaSetDevice(b->device));
  const size_t hl = (size_t) b->order * b->n + 2;
  const size_t yb = sizeof(double) * n_t * b->n * b->n_traj;
  double *d_t = nullptr, *d_y = nullptr;
  unsigned long long *d_bad = nullptr;
  unsigned long long bad = ~0ULL;
  cudaError_t err = cudaMalloc(&d_t, sizeof(double) * n_t);
  err = err ? err : cudaMalloc(&d_y, yb);
  err = err ? err : cudaMalloc(&d_bad, sizeof bad);
  err = err ? err : cudaMemcpyAsync(d_t, t, sizeof(double) * n_t, cudaMemcpyHostToDevice,
                                    b->stream);
  err = err ? err : cudaMemcpyAsync(d_bad, &bad, sizeof bad, cudaMemcpyHostToDevice, b->stream);
  if (err == cudaSuccess) {
    /* on the handle's stream: after the run that writes the rings */
    const long long total = (long long) b->n_traj * (long long) n_t;
    const int threads = 128;
    interpolate_rings_kernel<<<(unsigned) ((total + threads - 1) / threads), threads, 0,
                               b->stream>>>(b->d_ring, b->d_ring_count, (int) b->ctl.n_history,
                                            (int) hl, (int) b->n, (long long) b->n_traj, d_t,
                                            (int) n_t, d_y, d_bad);
    err = cudaGetLastError();
  }
  err = err ? err : cudaMemcpyAsync(y, d_y, yb, cudaMemcpyDeviceToHost, b->stream);
  err = err ? err : cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, b->stream);
  err = err ? err : cudaStreamSynchronize(b->stream);
  cudaFree(d_t);
  cudaFree(d_y);
  cudaFree(d_bad);
  if (err != cudaSuccess) {
    (void) cudaGetLastError();
    return fail("CUDA error in dde_dopri_batch_interpolate: %s", cudaGetErrorString(err));
  }
  if (bad != ~0ULL) { /* R/dopri.R:596-600, for the first trajectory it applies to */
    double first = 0.0, end = 0.0;
    std::vector<double> h((size_t) b->ctl.n_history * hl);
    const long used = dde_dopri_batch_history(b, (size_t) bad, h.data(), (size_t) b->ctl.n_history);
    if (used > 0) {
      first = h[hl - 2];
      end = h[(size_t) (used - 1) * hl + hl - 2] + h[(size_t) (used - 1) * hl + hl - 1];
    }
    return fail("Time falls outside of range of known history [%g, %g] (trajectory %llu)", first,
                end, bad);
  }
  return 0;
}

/* ---- difeq -------------------------------------------------------------------- */

void dde_difeq_batch_free(dde_difeq_batch_t *b) {
  if (b == nullptr) {
    return;
  }
  cudaSetDevice(b->device);
  free_dev(b->d_y0);
  free_dev(b->d_parms);
  free_dev(b->d_y_out);
  free_dev(b->d_out);
  free_dev(b->d_y_final);
  free_dev(b->d_ring);
  free_dev(b->d_steps);
  free_dev(b->d_next);
  free_dev(b->d_code);
  free_dev(b->d_ring_count);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

dde_difeq_batch_t *dde_difeq_batch_alloc(const char *model, size_t n, size_t n_out,
                                         size_t n_parms, size_t n_rep, size_t n_history) {
  if (model == nullptr) {
    fail("model must not be NULL");
    return nullptr;
  }
  dde::ModelDesc *m = find_model(model, 2, n);
  if (m == nullptr) {
    if (model_name_known(model)) {
      fail("model '%s' is not compiled for difeq with n = %zu", model, n);
    } else {
      fail("unknown model '%s'", model);
    }
    return nullptr;
  }
  if (check_model_shape(m, n, n_parms, n_out) != 0) {
    return nullptr;
  }
  if (n_rep == 0) {
    fail("n_rep must be positive");
    return nullptr;
  }
  if (n_history > 0 && n_history < 2) {
    fail("If given, n_history must be at least 2"); /* R/difeq.R:119-121 */
    return nullptr;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    fail("no CUDA device available: dde_b200 has no CPU fallback");
    (void) cudaGetLastError();
    return nullptr;
  }
  dde_difeq_batch_t *b = new dde_difeq_batch_t();
  b->model = m;
  b->n = n;
  b->n_out = n_out;
  b->n_parms = n_parms;
  b->n_rep = n_rep;
  b->n_history = n_history;
  cudaDeviceProp prop;
  bool ok = cudaGetDevice(&b->device) == cudaSuccess &&
            cudaGetDeviceProperties(&prop, b->device) == cudaSuccess;
  if (ok) {
    b->device_sms = prop.multiProcessorCount;
  }
  ok = ok && cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreate(&b->ev0) == cudaSuccess && cudaEventCreate(&b->ev1) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_y0, sizeof(double) * n * n_rep) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_parms, sizeof(double) * (n_parms ? n_parms : 1) * n_rep) ==
                 cudaSuccess;
  ok = ok && cudaMalloc(&b->d_y_final, sizeof(double) * n * n_rep) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_code, sizeof(int) * n_rep) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_next, sizeof(unsigned long long)) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_ring_count, sizeof(unsigned) * n_rep) == cudaSuccess;
  if (ok && n_history > 0) {
    ok = cudaMalloc(&b->d_ring, sizeof(double) * (1 + n + n_out) * n_history * n_rep) ==
         cudaSuccess;
  }
  if (!ok) {
    cudaError_t err = cudaGetLastError();
    fail("device allocation failed (%s)", cudaGetErrorString(err));
    dde_difeq_batch_free(b);
    return nullptr;
  }
  return b;
}

int dde_difeq_batch_upload(dde_difeq_batch_t *b, const double *y0, const double *parms,
                           size_t parms_stride) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  DDE_CUDA(cudaSetDevice(b->device));
  if (y0 != nullptr) {
    DDE_CUDA(cudaMemcpyAsync(b->d_y0, y0, sizeof(double) * b->n * b->n_rep,
                             cudaMemcpyHostToDevice, b->stream));
    b->have_y0 = true;
  }
  if (parms != nullptr && b->n_parms > 0) {
    if (parms_stride != 0 && parms_stride != b->n_parms) {
      return fail("parms_stride must be 0 (shared) or n_parms");
    }
    const size_t count = parms_stride == 0 ? b->n_parms : b->n_parms * b->n_rep;
    DDE_CUDA(cudaMemcpyAsync(b->d_parms, parms, sizeof(double) * count, cudaMemcpyHostToDevice,
                             b->stream));
    b->parms_stride = parms_stride;
  }
  DDE_CUDA(cudaStreamSynchronize(b->stream)); /* the caller's buffers are the caller's again */
  return 0;
}

int dde_difeq_batch_set_steps(dde_difeq_batch_t *b, const uint64_t *steps, size_t n_steps,
                              int32_t return_initial) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  if (steps == nullptr || n_steps < 2) {
    return fail("At least two steps must be given");
  }
  /* difeq_data_reset, src/difeq.c:85-93 */
  if (steps[n_steps - 1] == steps[0]) {
    return fail("Initialisation failure: Beginning and end steps are the same");
  }
  for (size_t i = 0; i + 1 < n_steps; ++i) {
    if (steps[i + 1] <= steps[i]) {
      return fail("Initialisation failure: Steps not strictly increasing");
    }
  }
  if (n_steps > 0x7fffffff || steps[n_steps - 1] > 0x7fffffffULL) {
    return fail("steps must fit an int (yprev takes int steps, src/difeq.c:274)");
  }
  DDE_CUDA(cudaSetDevice(b->device));
  if (n_steps > b->cap_steps) {
    free_dev(b->d_steps);
    DDE_CUDA(cudaMalloc(&b->d_steps, sizeof(unsigned long long) * n_steps));
    b->cap_steps = n_steps;
  }
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "uint64_t layout");
  DDE_CUDA(cudaMemcpyAsync(b->d_steps, steps, sizeof(uint64_t) * n_steps, cudaMemcpyHostToDevice,
                           b->stream));
  DDE_CUDA(cudaStreamSynchronize(b->stream));
  b->n_steps = n_steps;
  b->steps_first = steps[0];
  b->steps_last = steps[n_steps - 1];
  b->h_steps.assign(steps, steps + n_steps);
  b->return_initial = return_initial != 0;
  b->nt = return_initial ? n_steps : n_steps - 1;
  const size_t need_y = b->n * b->nt * b->n_rep;
  const size_t need_o = b->n_out * b->nt * b->n_rep;
  if (need_y > b->cap_y_out) {
    free_dev(b->d_y_out);
    if (cudaMalloc(&b->d_y_out, sizeof(double) * need_y) != cudaSuccess) {
      (void) cudaGetLastError();
      b->cap_y_out = 0;
      return fail("cannot allocate %.2f GB for the output array [n x nt x n_rep]",
                  (double) (sizeof(double) * need_y) / 1e9);
    }
    b->cap_y_out = need_y;
  }
  if (need_o > b->cap_out) {
    free_dev(b->d_out);
    DDE_CUDA(cudaMalloc(&b->d_out, sizeof(double) * need_o));
    b->cap_out = need_o;
  }
  b->have_steps = true;
  return 0;
}

} /* extern "C" */

namespace {

int difeq_run_impl(dde_difeq_batch_t *b, int restart) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  if (!b->have_steps || !b->have_y0) {
    return fail("upload y0 and set steps before running");
  }
  DDE_CUDA(cudaSetDevice(b->device));
  DDE_CUDA(cudaEventRecord(b->ev0, b->stream));
  DDE_CUDA(cudaMemsetAsync(b->d_y_out, 0, sizeof(double) * b->n * b->nt * b->n_rep, b->stream));
  if (b->n_out > 0) {
    DDE_CUDA(cudaMemsetAsync(b->d_out, 0, sizeof(double) * b->n_out * b->nt * b->n_rep,
                             b->stream));
  }
  DDE_CUDA(cudaMemsetAsync(b->d_next, 0, sizeof(unsigned long long), b->stream));
  dde::DifeqArgs a;
  memset(&a, 0, sizeof a);
  a.n_rep = (long long) b->n_rep;
  a.n = (int) b->n;
  a.n_out = (int) b->n_out;
  a.n_parms = (int) b->n_parms;
  a.parms_stride = (long long) b->parms_stride;
  a.y0 = b->d_y0;
  a.parms = b->d_parms;
  a.steps = b->d_steps;
  a.n_steps = (int) b->n_steps;
  a.n_history = (int) b->n_history;
  a.return_initial = b->return_initial;
  a.y_out = b->d_y_out;
  a.out = b->d_out;
  a.code = b->d_code;
  a.y_final = b->d_y_final;
  a.ring = b->d_ring;
  a.next_rep = b->d_next;
  a.restart = restart;
  a.restartable = b->restartable ? 1 : 0;
  a.ring_count = b->d_ring_count;
  if (!restart) {
    b->step_origin = b->steps_first;
  }
  b->continued = restart != 0;
  b->kept_history = b->restartable;
  a.step_origin = b->step_origin;
  int launches = 0;
  cudaError_t err = b->model->launch_difeq(a, b->device_sms, b->stream, &launches);
  if (err != cudaSuccess) {
    return fail("kernel launch failed for model '%s': %s", b->model->name,
                cudaGetErrorString(err));
  }
  DDE_CUDA(cudaEventRecord(b->ev1, b->stream));
  b->last_step = b->steps_last;
  b->ran = true;
  return 0;
}

} /* namespace */

extern "C" {

int dde_difeq_batch_run(dde_difeq_batch_t *b) { return difeq_run_impl(b, 0); }

int dde_difeq_batch_restartable(dde_difeq_batch_t *b, int32_t on) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  b->restartable = on != 0;
  return 0;
}

/* difeq_continue, R/difeq.R + src/r_difeq.c:114-165 */
int dde_difeq_batch_continue(dde_difeq_batch_t *b, const double *y, const uint64_t *steps,
                             size_t n_steps, int32_t return_initial) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  if (!b->ran) {
    return fail("nothing has been run on this handle: there is nothing to continue");
  }
  if (steps == nullptr || n_steps < 2) {
    return fail("At least two steps must be given"); /* src/r_difeq.c:136-138 */
  }
  if (steps[0] != b->last_step) {
    return fail("Incorrect initial step on simulation restart"); /* src/r_difeq.c:139-141 */
  }
  if (b->n_history > 0 && !b->kept_history) {
    return fail("the last run was not restartable: call dde_difeq_batch_restartable(b, 1) "
                "before running (the reference's `restartable = TRUE`, R/difeq.R)");
  }
  DDE_CUDA(cudaSetDevice(b->device));
  if (y != nullptr) {
    DDE_CUDA(cudaMemcpyAsync(b->d_y0, y, sizeof(double) * b->n * b->n_rep,
                             cudaMemcpyHostToDevice, b->stream));
  } else { /* obj->y1, src/r_difeq.c:122-123 */
    DDE_CUDA(cudaMemcpyAsync(b->d_y0, b->d_y_final, sizeof(double) * b->n * b->n_rep,
                             cudaMemcpyDeviceToDevice, b->stream));
  }
  const int rc = dde_difeq_batch_set_steps(b, steps, n_steps, return_initial);
  if (rc != 0) {
    return rc;
  }
  return difeq_run_impl(b, 1);
}

/* difeq_data_copy, src/difeq.c:49-72 */
dde_difeq_batch_t *dde_difeq_batch_copy(dde_difeq_batch_t *src) {
  if (src == nullptr) {
    fail("NULL handle");
    return nullptr;
  }
  if (cudaSetDevice(src->device) != cudaSuccess || cudaStreamSynchronize(src->stream) != cudaSuccess) {
    fail("CUDA error in dde_difeq_batch_copy");
    (void) cudaGetLastError();
    return nullptr;
  }
  dde_difeq_batch_t *b = dde_difeq_batch_alloc(src->model->name, src->n, src->n_out,
                                               src->n_parms, src->n_rep, src->n_history);
  if (b == nullptr) {
    return nullptr;
  }
  b->parms_stride = src->parms_stride;
  b->have_y0 = src->have_y0;
  b->step_origin = src->step_origin;
  b->last_step = src->last_step;
  b->restartable = src->restartable;
  b->kept_history = src->kept_history;
  const size_t nr = src->n_rep, n = src->n;
  cudaError_t err = cudaSuccess;
#define DDE_COPY(field, bytes)                                                              \
  if (err == cudaSuccess && src->field != nullptr && (bytes) > 0) {                         \
    err = cudaMemcpy(b->field, src->field, (bytes), cudaMemcpyDeviceToDevice);              \
  }
  DDE_COPY(d_y0, sizeof(double) * n * nr)
  DDE_COPY(d_parms, sizeof(double) * (src->n_parms ? src->n_parms : 1) * nr)
  DDE_COPY(d_y_final, sizeof(double) * n * nr)
  DDE_COPY(d_code, sizeof(int) * nr)
  DDE_COPY(d_ring_count, sizeof(unsigned) * nr)
  DDE_COPY(d_ring, sizeof(double) * (1 + n + src->n_out) * src->n_history * nr)
#undef DDE_COPY
  if (err == cudaSuccess && src->have_steps) {
    std::vector<uint64_t> steps(src->n_steps);
    err = cudaMemcpy(steps.data(), src->d_steps, sizeof(uint64_t) * src->n_steps,
                     cudaMemcpyDeviceToHost);
    if (err == cudaSuccess &&
        dde_difeq_batch_set_steps(b, steps.data(), src->n_steps, src->return_initial) != 0) {
      dde_difeq_batch_free(b);
      return nullptr;
    }
    if (err == cudaSuccess) {
      err = cudaMemcpy(b->d_y_out, src->d_y_out, sizeof(double) * n * src->nt * nr,
                       cudaMemcpyDeviceToDevice);
    }
    if (err == cudaSuccess && src->n_out > 0) {
      err = cudaMemcpy(b->d_out, src->d_out, sizeof(double) * src->n_out * src->nt * nr,
                       cudaMemcpyDeviceToDevice);
    }
  }
  if (err != cudaSuccess) {
    fail("CUDA error in dde_difeq_batch_copy: %s", cudaGetErrorString(err));
    (void) cudaGetLastError();
    dde_difeq_batch_free(b);
    return nullptr;
  }
  b->ran = src->ran;
  return b;
}

int dde_difeq_batch_sync(dde_difeq_batch_t *b) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  DDE_CUDA(cudaSetDevice(b->device));
  DDE_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}

int dde_difeq_batch_download(dde_difeq_batch_t *b, double *y_out, double *out, int32_t *code,
                             double *y_final) {
  if (b == nullptr) {
    return fail("NULL handle");
  }
  if (!b->ran) {
    return fail("nothing has been run on this handle");
  }
  DDE_CUDA(cudaSetDevice(b->device));
  if (y_out) {
    DDE_CUDA(cudaMemcpyAsync(y_out, b->d_y_out, sizeof(double) * b->n * b->nt * b->n_rep,
                             cudaMemcpyDeviceToHost, b->stream));
  }
  if (out && b->n_out > 0) {
    DDE_CUDA(cudaMemcpyAsync(out, b->d_out, sizeof(double) * b->n_out * b->nt * b->n_rep,
                             cudaMemcpyDeviceToHost, b->stream));
  }
  if (code) {
    DDE_CUDA(cudaMemcpyAsync(code, b->d_code, sizeof(int) * b->n_rep, cudaMemcpyDeviceToHost,
                             b->stream));
  }
  if (y_final) {
    DDE_CUDA(cudaMemcpyAsync(y_final, b->d_y_final, sizeof(double) * b->n * b->n_rep,
                             cudaMemcpyDeviceToHost, b->stream));
  }
  DDE_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}

/* The `history` attribute of a difference equation (src/r_difeq.c:178-186,
 * ring_buffer_read from the tail): entries {step, y[n], out[n_out]}, oldest
 * first.  The run must have been restartable (its history is then in HBM). */
long dde_difeq_batch_history(dde_difeq_batch_t *b, size_t j, double *history,
                             size_t max_entries) {
  if (b == nullptr || !b->ran || j >= b->n_rep) {
    fail("bad handle or replicate index");
    return -1;
  }
  const size_t cap = b->n_history;
  if (cap == 0) {
    return 0;
  }
  if (!b->kept_history) {
    fail("the last run was not restartable: call dde_difeq_batch_restartable(b, 1) before "
         "running to keep its history");
    return -1;
  }
  if (cudaSetDevice(b->device) != cudaSuccess || cudaStreamSynchronize(b->stream) != cudaSuccess) {
    fail("CUDA error in history export");
    return -1;
  }
  unsigned count = 0;
  if (cudaMemcpy(&count, b->d_ring_count + j, sizeof count, cudaMemcpyDeviceToHost) !=
      cudaSuccess) {
    fail("CUDA error in history export");
    return -1;
  }
  const size_t hl = 1 + b->n + b->n_out;
  const size_t used = count < cap ? count : cap;
  if (history != nullptr && used > 0) {
    std::vector<double> ring(cap * hl);
    if (cudaMemcpy(ring.data(), b->d_ring + j * cap * hl, sizeof(double) * cap * hl,
                   cudaMemcpyDeviceToHost) != cudaSuccess) {
      fail("CUDA error in history export");
      return -1;
    }
    /* The output columns.  The reference writes a step's outputs through a pointer
       (out_next, src/difeq.c:154-156,203-213,249-255) that moves to the entry being written
       only when an output ROW is stored, so what an entry's columns end up holding is an
       artefact of which steps were asked for: the row stored while the pointer rested on
       the entry's ring slot (later steps overwrite it until then), zeros where it never
       rested (the ring is calloc'ed), NA in the first entry (fill_na, :144) -- reproduced
       here from the stored rows for a first run; after difeq_continue (the pointer is set
       up again per run) the columns are NA. */
    const unsigned long long na_bits = 0x7FF00000000007A2ULL;
    std::vector<double> slot_out;
    if (b->n_out > 0 && !b->continued && b->h_steps.size() == b->n_steps && b->n_steps >= 2) {
      const size_t n_out = b->n_out, nt = b->nt;
      std::vector<double> rows(nt * n_out);
      if (cudaMemcpy(rows.data(), b->d_out + j * nt * n_out, sizeof(double) * nt * n_out,
                     cudaMemcpyDeviceToHost) != cudaSuccess) {
        fail("CUDA error in history export");
        return -1;
      }
      /* the reference's ring has one slot more than it holds entries (mrc-ide/ring keeps
         the slot under the head free: SURVEY.md appendix C), and the pointer is into that
         memory: entry e lives in slot e % (cap + 1) there */
      const size_t phys = cap + 1;
      slot_out.assign(phys * n_out, 0.0);
      for (size_t i = 0; i < n_out; ++i) {
        memcpy(&slot_out[i], &na_bits, sizeof(double)); /* entry 0, slot 0 */
      }
      size_t ptr_slot = 1 % phys; /* out_next after the first advance: entry 1 */
      bool store_next = b->return_initial != 0;
      size_t row = 0, idx = 1; /* steps[0] is the start; steps[idx] the next output step */
      const unsigned long long s0 = b->h_steps[0], s_last = b->h_steps[b->n_steps - 1];
      for (unsigned long long st = s0; st < s_last; ++st) {
        /* target(y(st)) has written its outputs to the pointer's slot */
        if (store_next) {
          memcpy(&slot_out[ptr_slot * n_out], &rows[row * n_out], sizeof(double) * n_out);
          ++row;
          ptr_slot = (size_t) ((st - s0 + 1) % phys); /* the entry y(st + 1) goes to */
          store_next = false;
        }
        if (idx < b->n_steps && st + 1 == b->h_steps[idx]) {
          store_next = true;
          ++idx;
        }
      }
      if (store_next && row < nt) { /* src/difeq.c:249-255 */
        memcpy(&slot_out[ptr_slot * n_out], &rows[row * n_out], sizeof(double) * n_out);
      }
    }
    const size_t first = count - used;
    for (size_t e = 0; e < used && e < max_entries; ++e) {
      const size_t slot = (first + e) % cap;
      double *dst = history + e * hl;
      /* entry idx holds the state after idx steps from the handle's first step */
      dst[0] = (double) (b->step_origin + first + e);
      memcpy(dst + 1, ring.data() + slot * hl + 1, sizeof(double) * b->n);
      for (size_t i = 0; i < b->n_out; ++i) {
        if (slot_out.empty()) {
          memcpy(dst + 1 + b->n + i, &na_bits, sizeof(double));
        } else {
          dst[1 + b->n + i] = slot_out[((first + e) % (cap + 1)) * b->n_out + i];
        }
      }
    }
  }
  return (long) used;
}

double dde_difeq_batch_last_ms(dde_difeq_batch_t *b) {
  if (b == nullptr || !b->ran) {
    return -1.0;
  }
  cudaSetDevice(b->device);
  if (cudaEventSynchronize(b->ev1) != cudaSuccess) {
    (void) cudaGetLastError();
    return -1.0;
  }
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, b->ev0, b->ev1) != cudaSuccess) {
    (void) cudaGetLastError();
    return -1.0;
  }
  return (double) ms;
}

int dde_difeq_batch(const char *model, size_t n, size_t n_out, size_t n_rep, const double *y0,
                    const double *parms, size_t n_parms, size_t parms_stride,
                    const uint64_t *steps, size_t n_steps, size_t n_history,
                    int32_t return_initial, double *y_out, double *out, int32_t *code,
                    double *y_final) {
  if (y0 == nullptr) {
    return fail("y0 must not be NULL");
  }
  if (n_parms > 0 && parms == nullptr) {
    return fail("parms must not be NULL when n_parms > 0");
  }
  dde_difeq_batch_t *b = dde_difeq_batch_alloc(model, n, n_out, n_parms, n_rep, n_history);
  if (b == nullptr) {
    return 1;
  }
  int rc = dde_difeq_batch_upload(b, y0, parms, parms_stride);
  rc = rc ? rc : dde_difeq_batch_set_steps(b, steps, n_steps, return_initial);
  rc = rc ? rc : dde_difeq_batch_run(b);
  rc = rc ? rc : dde_difeq_batch_download(b, y_out, out, code, y_final);
  dde_difeq_batch_free(b);
  return rc;
}

/* ---- multi-GPU sharding ------------------------------------------------------------ */

int dde_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    (void) cudaGetLastError();
    return 0;
  }
  return count;
}

void dde_shard_range(size_t n_total, size_t world, size_t rank, size_t *first, size_t *count) {
  if (world == 0) {
    world = 1;
  }
  const size_t base = n_total / world, extra = n_total % world;
  const size_t f = rank * base + (rank < extra ? rank : extra);
  if (first) *first = f;
  if (count) *count = base + (rank < extra ? 1 : 0);
}

} /* extern "C" */

namespace {

/* Run `work(shard, first, count)` for every non-empty shard on its own thread
   with that shard's device current; collect the first failure. */
template <class Work>
int run_sharded(size_t n_total, const int32_t *devices, size_t n_devices, Work work) {
  std::vector<int> devs;
  if (devices != nullptr && n_devices > 0) {
    devs.assign(devices, devices + n_devices);
  } else {
    const int count = dde_device_count();
    for (int d = 0; d < count; ++d) {
      devs.push_back(d);
    }
  }
  if (devs.empty()) {
    return fail("no CUDA device available: dde_b200 has no CPU fallback");
  }
  const int visible = dde_device_count();
  for (int d : devs) {
    if (d < 0 || d >= visible) {
      return fail("device %d is not visible (%d devices)", d, visible);
    }
  }
  const size_t world = devs.size();
  std::vector<int> rc(world, 0);
  std::vector<std::string> msg(world);
  std::vector<std::thread> threads;
  for (size_t r = 0; r < world; ++r) {
    size_t first = 0, count = 0;
    dde_shard_range(n_total, world, r, &first, &count);
    if (count == 0) {
      continue;
    }
    threads.emplace_back([&, r, first, count]() {
      if (cudaSetDevice(devs[r]) != cudaSuccess) {
        (void) cudaGetLastError();
        rc[r] = 1;
        msg[r] = "cudaSetDevice failed";
        return;
      }
      rc[r] = work(first, count);
      if (rc[r] != 0) {
        msg[r] = g_last_error; /* thread-local */
      }
    });
  }
  for (std::thread &t : threads) {
    t.join();
  }
  for (size_t r = 0; r < world; ++r) {
    if (rc[r] != 0) {
      return fail("shard %zu (device %d): %s", r, devs[r], msg[r].c_str());
    }
  }
  return 0;
}

template <class T>
T *offset(T *p, size_t k) {
  return p == nullptr ? nullptr : p + k;
}

} /* namespace */

extern "C" {

int dde_dopri_batch_sharded_events(const char *model, const dde_dopri_control *ctl, size_t n,
                                   size_t n_out, size_t n_traj, const double *y0,
                                   const double *parms, size_t n_parms, size_t parms_stride,
                                   const double *times, size_t n_times, const double *tcrit,
                                   size_t n_tcrit, const int32_t *event, double *y_out,
                                   double *out, int32_t *stats, int32_t *code, double *t_last,
                                   double *step_size, const int32_t *devices, size_t n_devices) {
  if (ctl == nullptr || n_times < 2) {
    return fail("ctl must not be NULL and at least two times must be given");
  }
  if (y0 == nullptr) {
    return fail("y0 must not be NULL");
  }
  if (n_parms > 0 && parms == nullptr) {
    return fail("parms must not be NULL when n_parms > 0");
  }
  const size_t nt = ctl->return_initial ? n_times : n_times - 1;
  return run_sharded(n_traj, devices, n_devices, [&](size_t first, size_t count) {
    dde_dopri_batch_t *b = dde_dopri_batch_alloc(model, ctl, n, n_out, n_parms, count);
    if (b == nullptr) {
      return 1;
    }
    int rc = dde_dopri_batch_upload(b, offset(y0, first * n), offset(parms, first * parms_stride),
                                    parms_stride);
    rc = rc ? rc : dde_dopri_batch_set_times(b, times, n_times, tcrit, n_tcrit);
    if (rc == 0 && event != nullptr) {
      rc = dde_dopri_batch_set_events(b, event, n_tcrit);
    }
    rc = rc ? rc
            : dde_dopri_batch_run_download(b, offset(y_out, first * n * nt),
                                           offset(out, first * n_out * nt),
                                           offset(stats, first * 4), offset(code, first),
                                           offset(t_last, first), offset(step_size, first));
    dde_dopri_batch_free(b);
    return rc;
  });
}

int dde_dopri_batch_sharded(const char *model, const dde_dopri_control *ctl, size_t n,
                            size_t n_out, size_t n_traj, const double *y0, const double *parms,
                            size_t n_parms, size_t parms_stride, const double *times,
                            size_t n_times, const double *tcrit, size_t n_tcrit, double *y_out,
                            double *out, int32_t *stats, int32_t *code, double *t_last,
                            double *step_size, const int32_t *devices, size_t n_devices) {
  return dde_dopri_batch_sharded_events(model, ctl, n, n_out, n_traj, y0, parms, n_parms,
                                        parms_stride, times, n_times, tcrit, n_tcrit, nullptr,
                                        y_out, out, stats, code, t_last, step_size, devices,
                                        n_devices);
}

/* ---- a batch resident on several GPUs: one dde_dopri_batch_t per device -------- */

struct dde_dopri_shards_s {
  std::vector<dde_dopri_batch_t *> shard;
  std::vector<int> device;
  std::vector<size_t> first, count;
  size_t n = 0, n_out = 0, n_traj = 0, nt = 0;
};

void dde_dopri_shards_free(dde_dopri_shards_t *h) {
  if (h == nullptr) {
    return;
  }
  int current = 0;
  (void) cudaGetDevice(&current);
  for (size_t r = 0; r < h->shard.size(); ++r) {
    if (h->shard[r] != nullptr) {
      (void) cudaSetDevice(h->device[r]);
      dde_dopri_batch_free(h->shard[r]);
    }
  }
  (void) cudaSetDevice(current);
  delete h;
}

dde_dopri_shards_t *dde_dopri_shards_alloc(const char *model, const dde_dopri_control *ctl,
                                           size_t n, size_t n_out, size_t n_parms, size_t n_traj,
                                           const int32_t *devices, size_t n_devices) {
  std::vector<int> devs;
  if (devices != nullptr && n_devices > 0) {
    devs.assign(devices, devices + n_devices);
  } else {
    for (int d = 0; d < dde_device_count(); ++d) {
      devs.push_back(d);
    }
  }
  if (devs.empty()) {
    fail("no CUDA device available: dde_b200 has no CPU fallback");
    return nullptr;
  }
  const int visible = dde_device_count();
  int current = 0;
  (void) cudaGetDevice(&current);
  dde_dopri_shards_t *h = new dde_dopri_shards_s();
  h->n = n;
  h->n_out = n_out;
  h->n_traj = n_traj;
  for (size_t r = 0; r < devs.size(); ++r) {
    size_t first = 0, count = 0;
    dde_shard_range(n_traj, devs.size(), r, &first, &count);
    if (count == 0) {
      continue;
    }
    if (devs[r] < 0 || devs[r] >= visible) {
      fail("device %d is not visible (%d devices)", devs[r], visible);
      dde_dopri_shards_free(h);
      return nullptr;
    }
    (void) cudaSetDevice(devs[r]);
    dde_dopri_batch_t *b = dde_dopri_batch_alloc(model, ctl, n, n_out, n_parms, count);
    if (b == nullptr) {
      const std::string msg = g_last_error;
      (void) cudaSetDevice(current);
      dde_dopri_shards_free(h);
      fail("shard %zu (device %d): %s", r, devs[r], msg.c_str());
      return nullptr;
    }
    h->shard.push_back(b);
    h->device.push_back(devs[r]);
    h->first.push_back(first);
    h->count.push_back(count);
  }
  (void) cudaSetDevice(current);
  return h;
}

size_t dde_dopri_shards_count(dde_dopri_shards_t *h) { return h ? h->shard.size() : 0; }

} /* extern "C" */

namespace {
/* `work(r)` for every shard, with its device current: in turn (asynchronous calls), or on
   one host thread each (blocking calls); the first failure is reported */
template <class Work>
int for_each_shard(dde_dopri_shards_t *h, bool threads, Work work) {
  if (h == nullptr) {
    return fail("handle must not be NULL");
  }
  const size_t world = h->shard.size();
  std::vector<int> rc(world, 0);
  std::vector<std::string> msg(world);
  int current = 0;
  (void) cudaGetDevice(&current);
  auto one = [&](size_t r) {
    if (cudaSetDevice(h->device[r]) != cudaSuccess) {
      (void) cudaGetLastError();
      rc[r] = 1;
      msg[r] = "cudaSetDevice failed";
      return;
    }
    rc[r] = work(r);
    if (rc[r] != 0) {
      msg[r] = g_last_error; /* thread-local */
    }
  };
  if (threads && world > 1) {
    std::vector<std::thread> pool;
    for (size_t r = 0; r < world; ++r) {
      pool.emplace_back(one, r);
    }
    for (std::thread &t : pool) {
      t.join();
    }
  } else {
    for (size_t r = 0; r < world; ++r) {
      one(r);
    }
    (void) cudaSetDevice(current);
  }
  for (size_t r = 0; r < world; ++r) {
    if (rc[r] != 0) {
      return fail("shard %zu (device %d): %s", r, h->device[r], msg[r].c_str());
    }
  }
  return 0;
}
} /* namespace */

extern "C" {

int dde_dopri_shards_upload(dde_dopri_shards_t *h, const double *y0, const double *parms,
                            size_t parms_stride) {
  /* every shard's copies are queued before the first is waited for */
  const int rc = for_each_shard(h, false, [&](size_t r) {
    return dopri_upload_async(h->shard[r], offset(y0, h->first[r] * h->n),
                              offset(parms, h->first[r] * parms_stride), parms_stride);
  });
  if (rc != 0) {
    return rc;
  }
  return for_each_shard(h, false, [&](size_t r) { return dde_dopri_batch_sync(h->shard[r]); });
}

int dde_dopri_shards_set_times(dde_dopri_shards_t *h, const double *times, size_t n_times,
                               const double *tcrit, size_t n_tcrit) {
  const int rc = for_each_shard(h, false, [&](size_t r) {
    return dde_dopri_batch_set_times(h->shard[r], times, n_times, tcrit, n_tcrit);
  });
  if (rc == 0) {
    h->nt = h->shard.empty() ? 0 : h->shard[0]->nt;
  }
  return rc;
}

int dde_dopri_shards_set_events(dde_dopri_shards_t *h, const int32_t *event, size_t n_tcrit) {
  return for_each_shard(h, false, [&](size_t r) {
    return dde_dopri_batch_set_events(h->shard[r], event, n_tcrit);
  });
}

/* every shard's own order, from its own last run (dde_dopri_batch_order_by_last_run);
   enable == 0 returns to index order */
int dde_dopri_shards_order_by_last_run(dde_dopri_shards_t *h, int32_t enable) {
  return for_each_shard(h, false, [&](size_t r) {
    return enable ? dde_dopri_batch_order_by_last_run(h->shard[r])
                  : dde_dopri_batch_set_order(h->shard[r], nullptr);
  });
}

int dde_dopri_shards_run_download(dde_dopri_shards_t *h, double *y_out, double *out,
                                  int32_t *stats, int32_t *code, double *t_last,
                                  double *step_size) {
  return for_each_shard(h, true, [&](size_t r) {
    const size_t f = h->first[r];
    return dde_dopri_batch_run_download(h->shard[r], offset(y_out, f * h->n * h->nt),
                                        offset(out, f * h->n_out * h->nt), offset(stats, f * 4),
                                        offset(code, f), offset(t_last, f), offset(step_size, f));
  });
}

double dde_dopri_shards_last_ms(dde_dopri_shards_t *h) {
  double ms = 0.0;
  if (h != nullptr) {
    int current = 0;
    (void) cudaGetDevice(&current);
    for (size_t r = 0; r < h->shard.size(); ++r) {
      (void) cudaSetDevice(h->device[r]);
      const double m = dde_dopri_batch_last_ms(h->shard[r]);
      ms = m > ms ? m : ms;
    }
    (void) cudaSetDevice(current);
  }
  return ms;
}

int dde_difeq_batch_sharded(const char *model, size_t n, size_t n_out, size_t n_rep,
                            const double *y0, const double *parms, size_t n_parms,
                            size_t parms_stride, const uint64_t *steps, size_t n_steps,
                            size_t n_history, int32_t return_initial, double *y_out, double *out,
                            int32_t *code, double *y_final, const int32_t *devices,
                            size_t n_devices) {
  if (n_steps < 2) {
    return fail("At least two steps must be given");
  }
  const size_t nt = return_initial ? n_steps : n_steps - 1;
  return run_sharded(n_rep, devices, n_devices, [&](size_t first, size_t count) {
    return dde_difeq_batch(model, n, n_out, count, offset(y0, first * n),
                           offset(parms, first * parms_stride), n_parms, parms_stride, steps,
                           n_steps, n_history, return_initial, offset(y_out, first * n * nt),
                           offset(out, first * n_out * nt), offset(code, first),
                           offset(y_final, first * n));
  });
}

} /* extern "C" */
