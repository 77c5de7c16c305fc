/* dopri_coop_kernel.cuh -- batched Dormand-Prince integrator for LARGE models:
 * one trajectory per thread block (one warp for n <= 128, four warps for
 * n = 512), the state spread over the block's threads.
 *
 * Same reference semantics as dopri_kernels.cuh (it restates the same code:
 * /root/reference/src/dopri.c:138-219,291-572, src/dopri_5.c, src/dopri_853.c),
 * different mapping:
 *   - thread `tid` owns components tid, tid + TPT, tid + 2 TPT, ... (at most
 *     CMAX of them) of y, k1..k10 and of the dense-output coefficients, in
 *     registers; every per-component loop of the reference becomes a loop over
 *     the thread's components, so ring commits, output rows and lag reads are
 *     coalesced along the component index;
 *   - the model's deriv() is called by the whole block at once on a stage
 *     vector staged in shared memory (the cooperative model ABI of
 *     <dde/dde.h>: DDE_FOR / DDE_SCRATCH / DDE_SYNC), one inlined instance
 *     inside a block-uniform stage loop;
 *   - control flow is uniform across the block (one t, one h, one
 *     accept/reject decision per block), so there is no divergence at all;
 *     blocks are persistent and fetch trajectories from an atomic counter;
 *   - the weighted norms (src/dopri_5.c:97-104, src/dopri_853.c:262-278,
 *     src/dopri.c:537-561) are summed in the reference's order i = 0..n-1 by
 *     one thread from terms the block computed in parallel: a tree reduction
 *     would round differently and step counts would drift;
 *   - history ring in HBM, ring[traj][slot][order*n + 2], entry layout as the
 *     reference's; a look-up is a block-wide operation: one thread locates the
 *     entry (cached bracket, else bisection over the stored step times), all
 *     threads interpolate their share of the requested components straight
 *     from the ring (L2 / L1 resident between the stages of a step).
 */
#ifndef DDE_DOPRI_COOP_KERNEL_CUH
#define DDE_DOPRI_COOP_KERNEL_CUH

#include <float.h>
#include <stdint.h>

#include "dde_math.h"
#include "dopri_args.h"
#include "tableau.h"

#ifndef DDE_COOP_SCRATCH_BYTES
#define DDE_COOP_SCRATCH_BYTES 8192 /* what a model may claim with DDE_SCRATCH */
#endif

namespace dde {

/* per-block solver context: what the lag functions need */
struct CoopCtx {
  double sign, t0s;        /* +-1; sign * start time (src/dopri.c:173,186) */
  const double *y0;        /* pre-history state (src/dopri.c:777-778) */
  double *ring;            /* this trajectory's ring */
  int n, hl, order, cap;
  unsigned count;          /* entries committed */
  int head;                /* slot of the next commit */
  int error, code;         /* obj->error, obj->code */
  /* cached bracket: entry c_idx covers [c_ts, c_tn) in sign-adjusted time */
  unsigned c_idx;
  int c_valid;
  double c_ts, c_tn, c_told, c_h;
  /* broadcast slots */
  unsigned b_idx;
  int b_ok;
  double b_told, b_h;
  double b_val[4];
  long long b_traj;
};

__shared__ CoopCtx s_cc;
extern __shared__ double s_coop_dyn[]; /* yin[n] dydt[n] red[2n] scratch[...] */

__device__ __forceinline__ double coop_na_real() {
  return __longlong_as_double(0x7FF00000000007A2LL);
}

__device__ __forceinline__ double coop_inf() {
  return __longlong_as_double(0x7ff0000000000000LL);
}

__device__ __forceinline__ const double *coop_entry(unsigned idx) {
  int slot = s_cc.head - (int) (s_cc.count - idx);
  slot = slot < 0 ? slot + s_cc.cap : slot;
  return s_cc.ring + (size_t) slot * (size_t) s_cc.hl;
}

/* Block-wide: the last committed entry whose time is <= t (>= t backwards),
 * src/dopri.c:742-769.  Thread 0 searches, everybody gets the answer.  On
 * failure ERR_YLAG_FAIL is recorded and false returned. */
__device__ __forceinline__ bool coop_locate(double t, unsigned *idx, double *t_old, double *h) {
  if (threadIdx.x == 0) {
    const double sign = s_cc.sign;
    const double st = sign * t;
    if (s_cc.c_valid && st >= s_cc.c_ts && st < s_cc.c_tn) {
      s_cc.b_ok = 1;
      s_cc.b_idx = s_cc.c_idx;
      s_cc.b_told = s_cc.c_told;
      s_cc.b_h = s_cc.c_h;
    } else {
      const int it = s_cc.hl - 2;
      const unsigned count = s_cc.count;
      const unsigned used = count < (unsigned) s_cc.cap ? count : (unsigned) s_cc.cap;
      unsigned lo = count - used, hi = count;
      bool ok = used > 0 && sign * coop_entry(lo)[it] <= st;
      if (ok) {
        /* the entry wanted is usually the cached one's successor */
        if (s_cc.c_valid && s_cc.c_idx >= lo && s_cc.c_idx < hi) {
          const unsigned g = s_cc.c_idx;
          if (st >= s_cc.c_tn) {
            lo = g + 1 < hi ? g + 1 : g; /* exists when c_tn is finite */
            if (lo + 1 < hi && !(sign * coop_entry(lo + 1)[it] <= st)) {
              hi = lo + 1;
            }
          } else {
            hi = g; /* st < c_ts */
          }
        }
        while (hi - lo > 1) {
          const unsigned mid = lo + (hi - lo) / 2;
          if (sign * coop_entry(mid)[it] <= st) {
            lo = mid;
          } else {
            hi = mid;
          }
        }
        const double *e = coop_entry(lo);
        s_cc.c_valid = 1;
        s_cc.c_idx = lo;
        s_cc.c_told = e[it];
        s_cc.c_h = e[it + 1];
        s_cc.c_ts = sign * e[it];
        s_cc.c_tn = lo + 1 < count ? sign * coop_entry(lo + 1)[it] : coop_inf();
        s_cc.b_ok = 1;
        s_cc.b_idx = lo;
        s_cc.b_told = s_cc.c_told;
        s_cc.b_h = s_cc.c_h;
      } else {
        s_cc.error = 1;
        s_cc.code = -6; /* ERR_YLAG_FAIL */
        s_cc.b_ok = 0;
      }
    }
  }
  __syncthreads();
  const bool ok = s_cc.b_ok != 0;
  *idx = s_cc.b_idx;
  *t_old = s_cc.b_told;
  *h = s_cc.b_h;
  return ok;
}

__device__ __forceinline__ double coop_interp(const double *c, int order, int n, double theta,
                                              double theta1) {
  if (order == 5) {
    return c[0] + theta * (c[n] + theta1 * (c[2 * n] + theta * (c[3 * n] + theta1 * c[4 * n])));
  }
  const double tmp = c[4 * n] + theta * (c[5 * n] + theta1 * (c[6 * n] + theta * c[7 * n]));
  return c[0] + theta * (c[n] + theta1 * (c[2 * n] + theta * (c[3 * n] + theta1 * tmp)));
}

template <class Index>
__device__ __forceinline__ void coop_ylag_indexed(double t, const Index *idx, size_t nidx,
                                                  double *y) {
  if (s_cc.sign * t <= s_cc.t0s) {
    for (size_t i = threadIdx.x; i < nidx; i += blockDim.x) {
      y[i] = s_cc.y0[idx ? (size_t) idx[i] : i];
    }
    __syncthreads();
    return;
  }
  unsigned e_idx;
  double t_old, h;
  if (coop_locate(t, &e_idx, &t_old, &h)) {
    const double *e = coop_entry(e_idx);
    const int n = s_cc.n, order = s_cc.order;
    const double theta = (t - t_old) / h;
    const double theta1 = 1 - theta;
    for (size_t i = threadIdx.x; i < nidx; i += blockDim.x) {
      y[i] = coop_interp(e + (idx ? (size_t) idx[i] : i), order, n, theta, theta1);
    }
  }
  __syncthreads(); /* results visible; the broadcast slots may be reused */
}

} /* namespace dde */

/* ---- the model-facing functions, collective versions ------------------------ */

static __device__ __forceinline__ char *dde_coop_scratch(void) {
  /* after yin, dydt, red[2n] */
  return (char *) (dde::s_coop_dyn + 4 * (size_t) dde::s_cc.n);
}

static __device__ __forceinline__ double ylag_1(double t, size_t i) {
  if (dde::s_cc.sign * t <= dde::s_cc.t0s) {
    return dde::s_cc.y0[i];
  }
  unsigned e_idx;
  double t_old, h;
  double ret = dde::coop_na_real();
  if (dde::coop_locate(t, &e_idx, &t_old, &h)) {
    const double theta = (t - t_old) / h;
    const double theta1 = 1 - theta;
    ret = dde::coop_interp(dde::coop_entry(e_idx) + i, dde::s_cc.order, dde::s_cc.n, theta, theta1);
  }
  __syncthreads();
  return ret;
}

static __device__ __forceinline__ void ylag_all(double t, double *y) {
  dde::coop_ylag_indexed<size_t>(t, nullptr, (size_t) dde::s_cc.n, y);
}

static __device__ __forceinline__ void ylag_vec(double t, const size_t *idx, size_t nidx,
                                                double *y) {
  dde::coop_ylag_indexed<size_t>(t, idx, nidx, y);
}

static __device__ __forceinline__ void ylag_vec_int(double t, const int *idx, size_t nidx,
                                                    double *y) {
  dde::coop_ylag_indexed<int>(t, idx, nidx, y);
}

/* difference equations have no block-per-replicate kernel */
static __device__ __forceinline__ double yprev_1(int, size_t) { return dde::coop_na_real(); }
static __device__ __forceinline__ void yprev_all(int, double *) {}
static __device__ __forceinline__ void yprev_vec(int, const size_t *, size_t, double *) {}
static __device__ __forceinline__ void yprev_vec_int(int, const int *, size_t, double *) {}

namespace dde {

__device__ __forceinline__ double csq(double x) { return x * x; }

/* Sum term[i], i = 0 .. n-1, in that order (the reference's loops), from the
 * per-thread terms of components tid + j TPT.  One or two sums at a time. */
template <int TPT, int CMAX, int NSUM>
__device__ __forceinline__ void coop_seqsum(int n, const double (*term)[CMAX], double *result) {
  double *red = s_coop_dyn + 2 * (size_t) n;
  const int tid = threadIdx.x;
#pragma unroll
  for (int k = 0; k < NSUM; ++k) {
#pragma unroll
    for (int j = 0; j < CMAX; ++j) {
      const int i = tid + j * TPT;
      if (i < n) {
        red[(size_t) k * n + i] = term[k][j];
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    double acc[NSUM];
#pragma unroll
    for (int k = 0; k < NSUM; ++k) {
      acc[k] = 0.0;
    }
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int k = 0; k < NSUM; ++k) {
        acc[k] += red[(size_t) k * n + i];
      }
    }
#pragma unroll
    for (int k = 0; k < NSUM; ++k) {
      s_cc.b_val[k] = acc[k];
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NSUM; ++k) {
    result[k] = s_cc.b_val[k];
  }
  __syncthreads(); /* b_val and red may be reused */
}

template <class M, int METHOD, int TPT, int CMAX>
__global__ void __launch_bounds__(TPT) dopri_coop_kernel(const DopriArgs a) {
  constexpr int ORDER = METHOD == 0 ? 5 : 8;
  constexpr int NST_STEP = METHOD == 0 ? 6 : 11;
  constexpr int NST = METHOD == 0 ? 6 : 16;
  const int n = a.n;
  const int hl = ORDER * n + 2;
  const int tid = threadIdx.x;
  double *yin_s = s_coop_dyn;
  double *dydt_s = s_coop_dyn + n;

  constexpr double step_factor_min = METHOD == 0 ? 0.2 : 0.333;
  constexpr double step_factor_max = METHOD == 0 ? 10.0 : 6.0;
  constexpr double step_beta = METHOD == 0 ? 0.04 : 0.0;
  constexpr double step_factor_safe = 0.9;
  const double step_constant = METHOD == 0 ? 0.2 - step_beta * 0.75 : 0.125 - step_beta * 0.2;
  const double sign = a.sign, atol = a.atol, rtol = a.rtol;
  const int n_times = a.n_times;
  const int nt = a.return_initial ? n_times : n_times - 1;
  const double t_end = a.times[n_times - 1];

#define DDE_CLOOP(j, i)                                        \
  _Pragma("unroll") for (int j = 0; j < CMAX; ++j)             \
    if (const int i = tid + j * TPT; i < n)

  for (;;) {
    /* ---- next trajectory (persistent blocks) ---- */
    __syncthreads();
    if (tid == 0) {
      s_cc.b_traj = (long long) atomicAdd(a.next_traj, 1ULL);
    }
    __syncthreads();
    const long long traj = s_cc.b_traj;
    if (traj >= a.n_traj) {
      break;
    }
    const double *p = a.parms + traj * a.parms_stride;
    const double *y0 = a.y0 + traj * n;
    double *ring = a.ring + (size_t) traj * (size_t) a.n_history * (size_t) hl;
    const bool cannot_restart =
        a.restart && (a.code[traj] != 1 || a.t_last[traj] != a.times[0]);
    const double ssi = a.restart ? a.step_size[traj] : a.step_size_initial;
    __syncthreads(); /* everybody has read the previous run's code */
    if (cannot_restart) { /* src/r_dopri.c:215-217 */
      if (tid == 0) {
        a.code[traj] = -10;
        a.stats[4 * traj + 0] = a.stats[4 * traj + 1] = 0;
        a.stats[4 * traj + 2] = a.stats[4 * traj + 3] = 0;
      }
      continue;
    }
    if (tid == 0) { /* dopri_data_reset, src/dopri.c:138-219 */
      s_cc.sign = sign;
      s_cc.t0s = sign * a.times[0];
      s_cc.y0 = y0;
      s_cc.ring = ring;
      s_cc.n = n;
      s_cc.hl = hl;
      s_cc.order = ORDER;
      s_cc.cap = a.n_history;
      s_cc.count = 0;
      s_cc.head = 0;
      if (a.restart && a.n_history > 0) { /* the ring the last run left, src/dopri.c:185-191 */
        const unsigned cap = (unsigned) a.n_history;
        const unsigned count = a.ring_count[traj];
        s_cc.count = count;
        s_cc.head = (int) (count % cap);
        if (count > 0) {
          const unsigned used = count < cap ? count : cap;
          s_cc.t0s = sign * ring[(size_t) ((count - used) % cap) * (size_t) hl + (hl - 2)];
        }
      }
      s_cc.error = 0;
      s_cc.code = 0;
      s_cc.c_valid = 0;
    }
    __syncthreads();

    double y[CMAX], k1[CMAX];
    DDE_CLOOP(j, i) { y[j] = y0[i]; }
    double t = a.times[0];
    int times_idx = 1, tcrit_idx = a.tcrit_idx0;
    int n_eval = 0, n_step = 0, n_accept = 0, n_reject = 0;
    unsigned long long stiff_n_stiff = 0, stiff_n_nonstiff = 0;
    double fac_old = 1e-4, h_save = 0.0, step_size_exit = ssi;
    bool stop = false, last = false, reject = false;
    double t_stop = t_end;
    if (tcrit_idx < a.n_tcrit && sign * a.tcrit[tcrit_idx] < sign * t_end) {
      t_stop = a.tcrit[tcrit_idx];
    }
    double *yo = a.y_out + (size_t) traj * (size_t) nt * (size_t) n;
    if (a.return_initial) { /* src/dopri.c:320-327 */
      DDE_CLOOP(j, i) { yo[i] = y[j]; }
      yo += n;
    }

    /* one RHS evaluation by the whole block, src/dopri.c:831-835 */
#define DDE_COOP_EVAL(tt, src, dst)                                  \
  do {                                                               \
    DDE_CLOOP(j_, i_) { yin_s[i_] = (src)[j_]; }                     \
    __syncthreads();                                                 \
    M::deriv((size_t) n, (tt), yin_s, dydt_s, p);                    \
    __syncthreads();                                                 \
    DDE_CLOOP(j_, i_) { (dst)[j_] = dydt_s[i_]; }                    \
    ++n_eval;                                                        \
  } while (0)

    bool done = false;
    double h = 0.0;
    /* k1 = f(t0, y0), finiteness check, dopri_h_init: src/dopri.c:329-339,527-572 */
    DDE_COOP_EVAL(t, y, k1);
    {
      double bad[1][CMAX];
#pragma unroll
      for (int j = 0; j < CMAX; ++j) {
        bad[0][j] = 0.0;
      }
      DDE_CLOOP(j, i) { bad[0][j] = isfinite(k1[j]) ? 0.0 : 1.0; }
      double nbad;
      coop_seqsum<TPT, CMAX, 1>(n, bad, &nbad);
      if (nbad > 0.0) {
        if (tid == 0) {
          s_cc.code = -8; /* DDE_ERR_NONFINITE_DERIV */
        }
        done = true;
      }
    }
    if (!done) {
      if (ssi > 0.0) {
        h = ssi;
      } else {
        double term[2][CMAX];
        DDE_CLOOP(j, i) {
          const double sk = atol + rtol * fabs(y[j]);
          term[0][j] = csq(k1[j] / sk);
          term[1][j] = csq(y[j] / sk);
        }
        double norms[2];
        coop_seqsum<TPT, CMAX, 2>(n, term, norms);
        const double norm_f = norms[0], norm_y = norms[1];
        h = (norm_f <= 1e-10 || norm_y <= 1e-10) ? 1e-6 : sqrt(norm_y / norm_f) * 0.01;
        h = copysign(fmin(h, a.step_size_max), sign);
        double ye[CMAX], f1[CMAX];
        DDE_CLOOP(j, i) { ye[j] = y[j] + h * k1[j]; }
        DDE_COOP_EVAL(t + h, ye, f1);
        double term2[1][CMAX];
        DDE_CLOOP(j, i) {
          const double sk = atol + rtol * fabs(y[j]);
          term2[0][j] = csq((f1[j] - k1[j]) / sk);
        }
        double der2;
        coop_seqsum<TPT, CMAX, 1>(n, term2, &der2);
        der2 = sqrt(der2) / h;
        const double der12 = fmax(fabs(der2), sqrt(norm_f));
        const double h1 = (der12 <= 1e-15) ? fmax(1e-6, fabs(h) * 1e-3)
                                           : dde_pow(0.01 / der12, 1.0 / ORDER);
        h = fmin(fmin(100 * fabs(h), h1), a.step_size_max);
        h = copysign(h, sign);
      }
    }

    /* ---- the adaptive loop, src/dopri.c:343-510 ---- */
    while (!done) {
      bool force_this_step = false;
      int fail_code = 0;
      if ((unsigned long long) n_step > a.step_max_n) {
        fail_code = -3;
      } else if (fabs(h) <= a.step_size_min) {
        if (a.step_size_min_allow) {
          h = copysign(a.step_size_min, sign);
          force_this_step = true;
        } else {
          fail_code = -4;
        }
      } else if (fabs(h) <= fabs(t) * DBL_EPSILON) {
        fail_code = -5;
      }
      if (fail_code != 0) {
        if (tid == 0) {
          s_cc.error = 1;
          s_cc.code = fail_code;
        }
        break;
      }
      if ((t + 1.01 * h - t_stop) * sign > 0.0) {
        h = t_stop - t;
        stop = true;
      }
      if ((t + 1.01 * h - t_end) * sign > 0.0) {
        h_save = h;
        h = t_end - t;
        last = true;
      }
      ++n_step;

      double k2[CMAX], k3[CMAX], k4[CMAX], k5[CMAX], k6[CMAX], k7[CMAX], k8[CMAX], k9[CMAX],
          k10[CMAX], ysave[CMAX], hd[ORDER * CMAX], yin[CMAX], kt[CMAX];
      bool accepted = false, step_done = false, lag_failed = false, gave_up = false;
      double h_new = 0.0;

#pragma unroll 1
      for (int st = 0; st < NST; ++st) {
        if (st >= NST_STEP && !accepted) {
          break;
        }
        double tt;
        /* -- the input of evaluation st (block-uniform switch) -- */
        if (METHOD == 0) {
          switch (st) { /* dopri5_step, src/dopri_5.c:54-80 */
            case 0:
              DDE_CLOOP(j, i) { yin[j] = y[j] + h * DP5_A21 * k1[j]; }
              tt = t + DP5_C2 * h;
              break;
            case 1:
              DDE_CLOOP(j, i) { yin[j] = y[j] + h * (DP5_A31 * k1[j] + DP5_A32 * k2[j]); }
              tt = t + DP5_C3 * h;
              break;
            case 2:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP5_A41 * k1[j] + DP5_A42 * k2[j] + DP5_A43 * k3[j]);
              }
              tt = t + DP5_C4 * h;
              break;
            case 3:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP5_A51 * k1[j] + DP5_A52 * k2[j] + DP5_A53 * k3[j] +
                                     DP5_A54 * k4[j]);
              }
              tt = t + DP5_C5 * h;
              break;
            case 4:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP5_A61 * k1[j] + DP5_A62 * k2[j] + DP5_A63 * k3[j] +
                                     DP5_A64 * k4[j] + DP5_A65 * k5[j]);
                ysave[j] = yin[j];
              }
              tt = t + h;
              break;
            default:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP5_A71 * k1[j] + DP5_A73 * k3[j] + DP5_A74 * k4[j] +
                                     DP5_A75 * k5[j] + DP5_A76 * k6[j]);
              }
              tt = t + h;
              break;
          }
        } else {
          switch (st) { /* dopri853_step, src/dopri_853.c:178-240; :304,330-347 */
            case 0:
              DDE_CLOOP(j, i) { yin[j] = y[j] + h * DP8_A21 * k1[j]; }
              tt = t + DP8_C2 * h;
              break;
            case 1:
              DDE_CLOOP(j, i) { yin[j] = y[j] + h * (DP8_A31 * k1[j] + DP8_A32 * k2[j]); }
              tt = t + DP8_C3 * h;
              break;
            case 2:
              DDE_CLOOP(j, i) { yin[j] = y[j] + h * (DP8_A41 * k1[j] + DP8_A43 * k3[j]); }
              tt = t + DP8_C4 * h;
              break;
            case 3:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP8_A51 * k1[j] + DP8_A53 * k3[j] + DP8_A54 * k4[j]);
              }
              tt = t + DP8_C5 * h;
              break;
            case 4:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP8_A61 * k1[j] + DP8_A64 * k4[j] + DP8_A65 * k5[j]);
              }
              tt = t + DP8_C6 * h;
              break;
            case 5:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP8_A71 * k1[j] + DP8_A74 * k4[j] + DP8_A75 * k5[j] +
                                     DP8_A76 * k6[j]);
              }
              tt = t + DP8_C7 * h;
              break;
            case 6:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP8_A81 * k1[j] + DP8_A84 * k4[j] + DP8_A85 * k5[j] +
                                     DP8_A86 * k6[j] + DP8_A87 * k7[j]);
              }
              tt = t + DP8_C8 * h;
              break;
            case 7:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP8_A91 * k1[j] + DP8_A94 * k4[j] + DP8_A95 * k5[j] +
                                     DP8_A96 * k6[j] + DP8_A97 * k7[j] + DP8_A98 * k8[j]);
              }
              tt = t + DP8_C9 * h;
              break;
            case 8:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP8_A101 * k1[j] + DP8_A104 * k4[j] + DP8_A105 * k5[j] +
                                     DP8_A106 * k6[j] + DP8_A107 * k7[j] + DP8_A108 * k8[j] +
                                     DP8_A109 * k9[j]);
              }
              tt = t + DP8_C10 * h;
              break;
            case 9:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP8_A111 * k1[j] + DP8_A114 * k4[j] + DP8_A115 * k5[j] +
                                     DP8_A116 * k6[j] + DP8_A117 * k7[j] + DP8_A118 * k8[j] +
                                     DP8_A119 * k9[j] + DP8_A1110 * k10[j]);
              }
              tt = t + DP8_C11 * h;
              break;
            case 10:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP8_A121 * k1[j] + DP8_A124 * k4[j] + DP8_A125 * k5[j] +
                                     DP8_A126 * k6[j] + DP8_A127 * k7[j] + DP8_A128 * k8[j] +
                                     DP8_A129 * k9[j] + DP8_A1210 * k10[j] + DP8_A1211 * k2[j]);
                ysave[j] = yin[j];
              }
              tt = t + h;
              break;
            case 11: /* the redundant evaluation at the OLD time, src/dopri.c:400-403 */
              DDE_CLOOP(j, i) { yin[j] = k5[j]; }
              tt = t;
              break;
            case 12: /* src/dopri_853.c:304 */
              DDE_CLOOP(j, i) { yin[j] = k5[j]; }
              tt = t + h;
              break;
            case 13:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP8_A141 * k1[j] + DP8_A147 * k7[j] + DP8_A148 * k8[j] +
                                     DP8_A149 * k9[j] + DP8_A1410 * k10[j] + DP8_A1411 * k2[j] +
                                     DP8_A1412 * k3[j] + DP8_A1413 * k4[j]);
              }
              tt = t + DP8_C14 * h;
              break;
            case 14:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP8_A151 * k1[j] + DP8_A156 * k6[j] + DP8_A157 * k7[j] +
                                     DP8_A158 * k8[j] + DP8_A1511 * k2[j] + DP8_A1512 * k3[j] +
                                     DP8_A1513 * k4[j] + DP8_A1514 * k10[j]);
              }
              tt = t + DP8_C15 * h;
              break;
            default:
              DDE_CLOOP(j, i) {
                yin[j] = y[j] + h * (DP8_A161 * k1[j] + DP8_A166 * k6[j] + DP8_A167 * k7[j] +
                                     DP8_A168 * k8[j] + DP8_A169 * k9[j] + DP8_A1613 * k4[j] +
                                     DP8_A1614 * k10[j] + DP8_A1615 * k2[j]);
              }
              tt = t + DP8_C16 * h;
              break;
          }
        }

        /* -- the one hot copy of the model -- */
        DDE_COOP_EVAL(tt, yin, kt);

        /* -- file the result -- */
        bool attempt_done = false;
        if (METHOD == 0) {
          switch (st) {
            case 0: DDE_CLOOP(j, i) { k2[j] = kt[j]; } break;
            case 1: DDE_CLOOP(j, i) { k3[j] = kt[j]; } break;
            case 2: DDE_CLOOP(j, i) { k4[j] = kt[j]; } break;
            case 3: DDE_CLOOP(j, i) { k5[j] = kt[j]; } break;
            case 4: DDE_CLOOP(j, i) { k6[j] = kt[j]; } break;
            default:
              DDE_CLOOP(j, i) { k2[j] = kt[j]; }
              attempt_done = true;
              break;
          }
        } else {
          switch (st) {
            case 0: DDE_CLOOP(j, i) { k2[j] = kt[j]; } break;
            case 1: DDE_CLOOP(j, i) { k3[j] = kt[j]; } break;
            case 2: DDE_CLOOP(j, i) { k4[j] = kt[j]; } break;
            case 3: DDE_CLOOP(j, i) { k5[j] = kt[j]; } break;
            case 4: DDE_CLOOP(j, i) { k6[j] = kt[j]; } break;
            case 5: DDE_CLOOP(j, i) { k7[j] = kt[j]; } break;
            case 6: DDE_CLOOP(j, i) { k8[j] = kt[j]; } break;
            case 7: DDE_CLOOP(j, i) { k9[j] = kt[j]; } break;
            case 8: DDE_CLOOP(j, i) { k10[j] = kt[j]; } break;
            case 9: DDE_CLOOP(j, i) { k2[j] = kt[j]; } break;
            case 10: /* src/dopri_853.c:242-246 */
              DDE_CLOOP(j, i) {
                k3[j] = kt[j];
                k4[j] = DP8_B1 * k1[j] + DP8_B6 * k6[j] + DP8_B7 * k7[j] + DP8_B8 * k8[j] +
                        DP8_B9 * k9[j] + DP8_B10 * k10[j] + DP8_B11 * k2[j] + DP8_B12 * k3[j];
                k5[j] = y[j] + h * k4[j];
              }
              attempt_done = true;
              break;
            case 11: DDE_CLOOP(j, i) { k4[j] = kt[j]; } break;
            case 12: /* src/dopri_853.c:306-327 */
              DDE_CLOOP(j, i) {
                k4[j] = kt[j];
                const double ydiff = k5[j] - y[j];
                const double bspl = h * k1[j] - ydiff;
                hd[0 * CMAX + j] = y[j];
                hd[1 * CMAX + j] = ydiff;
                hd[2 * CMAX + j] = bspl;
                hd[3 * CMAX + j] = ydiff - h * k4[j] - bspl;
                hd[4 * CMAX + j] = DP8_D41 * k1[j] + DP8_D46 * k6[j] + DP8_D47 * k7[j] +
                                   DP8_D48 * k8[j] + DP8_D49 * k9[j] + DP8_D410 * k10[j] +
                                   DP8_D411 * k2[j] + DP8_D412 * k3[j];
                hd[5 * CMAX + j] = DP8_D51 * k1[j] + DP8_D56 * k6[j] + DP8_D57 * k7[j] +
                                   DP8_D58 * k8[j] + DP8_D59 * k9[j] + DP8_D510 * k10[j] +
                                   DP8_D511 * k2[j] + DP8_D512 * k3[j];
                hd[6 * CMAX + j] = DP8_D61 * k1[j] + DP8_D66 * k6[j] + DP8_D67 * k7[j] +
                                   DP8_D68 * k8[j] + DP8_D69 * k9[j] + DP8_D610 * k10[j] +
                                   DP8_D611 * k2[j] + DP8_D612 * k3[j];
                hd[7 * CMAX + j] = DP8_D71 * k1[j] + DP8_D76 * k6[j] + DP8_D77 * k7[j] +
                                   DP8_D78 * k8[j] + DP8_D79 * k9[j] + DP8_D710 * k10[j] +
                                   DP8_D711 * k2[j] + DP8_D712 * k3[j];
              }
              break;
            case 13: DDE_CLOOP(j, i) { k10[j] = kt[j]; } break;
            case 14: DDE_CLOOP(j, i) { k2[j] = kt[j]; } break;
            default: /* 15, src/dopri_853.c:350-366 */
              DDE_CLOOP(j, i) {
                k3[j] = kt[j];
                hd[4 * CMAX + j] = h * (hd[4 * CMAX + j] + DP8_D413 * k4[j] + DP8_D414 * k10[j] +
                                        DP8_D415 * k2[j] + DP8_D416 * k3[j]);
                hd[5 * CMAX + j] = h * (hd[5 * CMAX + j] + DP8_D513 * k4[j] + DP8_D514 * k10[j] +
                                        DP8_D515 * k2[j] + DP8_D516 * k3[j]);
                hd[6 * CMAX + j] = h * (hd[6 * CMAX + j] + DP8_D613 * k4[j] + DP8_D614 * k10[j] +
                                        DP8_D615 * k2[j] + DP8_D616 * k3[j]);
                hd[7 * CMAX + j] = h * (hd[7 * CMAX + j] + DP8_D713 * k4[j] + DP8_D714 * k10[j] +
                                        DP8_D715 * k2[j] + DP8_D716 * k3[j]);
              }
              step_done = true;
              break;
          }
        }

        if (attempt_done) {
          if (s_cc.error) { /* a lag look-up failed during dopri_step, src/dopri.c:387-389 */
            lag_failed = true;
            break;
          }
          double err;
          if (METHOD == 0) { /* dopri5_error, src/dopri_5.c:97-104 (see dopri_kernels.cuh) */
            double term[1][CMAX];
            DDE_CLOOP(j, i) {
              const double ke = h * (DP5_E1 * k1[j] + DP5_E3 * k3[j] + DP5_E4 * k4[j] +
                                     DP5_E5 * k5[j] + DP5_E6 * k6[j] + DP5_E7 * k2[j]);
              const double sk = atol + rtol * fmax(fabs(y[j]), fabs(yin[j]));
              term[0][j] = csq(ke / sk);
            }
            coop_seqsum<TPT, CMAX, 1>(n, term, &err);
            err = sqrt(err / (double) n);
          } else { /* dopri853_error, src/dopri_853.c:249-283 */
            double term[2][CMAX];
            DDE_CLOOP(j, i) {
              const double sk = atol + rtol * fmax(fabs(y[j]), fabs(k5[j]));
              double erri = k4[j] - DP8_BHH1 * k1[j] - DP8_BHH2 * k9[j] - DP8_BHH3 * k3[j];
              term[1][j] = csq(erri / sk);
              erri = DP8_ER1 * k1[j] + DP8_ER6 * k6[j] + DP8_ER7 * k7[j] + DP8_ER8 * k8[j] +
                     DP8_ER9 * k9[j] + DP8_ER10 * k10[j] + DP8_ER11 * k2[j] + DP8_ER12 * k3[j];
              term[0][j] = csq(erri / sk);
            }
            double sums[2];
            coop_seqsum<TPT, CMAX, 2>(n, term, sums);
            double deno = sums[0] + 0.01 * sums[1];
            deno = deno < 0 ? 1.0 : deno;
            err = sign * sums[0] * sqrt(1.0 / ((double) n * deno));
          }
          /* dopri_h_new, src/dopri.c:516-525 */
          const double fac11 = dde_pow(err, step_constant);
          const double facb = METHOD == 0 ? dde_pow(fac_old, step_beta) : 1.0;
          double fac = fac11 / facb;
          fac = fmax(1.0 / step_factor_max, fmin(1.0 / step_factor_min, fac / step_factor_safe));
          h_new = h / fac;
          accepted = (err <= 1) || force_this_step;
          if (accepted) {
            fac_old = fmax(err, 1e-4);
            ++n_accept;
            if (METHOD == 0) {
              step_done = true;
            }
          } else { /* src/dopri.c:495-509 */
            h_new = h / fmin(1 / step_factor_min, fac11 / step_factor_safe);
            reject = true;
            if (n_accept >= 1) {
              ++n_reject;
            }
            last = false;
            stop = false;
            h = h_new;
          }
        }

        /* dopri_test_stiff, src/dopri.c:668-693 */
        if ((METHOD == 0 ? step_done : st == 11) && a.stiff_check != 0 &&
            (stiff_n_stiff > 0 || (unsigned long long) n_accept % a.stiff_check == 0)) {
          double term[2][CMAX];
          DDE_CLOOP(j, i) {
            if (METHOD == 0) {
              term[0][j] = csq(k2[j] - k6[j]);
              term[1][j] = csq(yin[j] - ysave[j]);
            } else {
              term[0][j] = csq(k4[j] - k3[j]);
              term[1][j] = csq(k5[j] - ysave[j]);
            }
          }
          double sums[2];
          coop_seqsum<TPT, CMAX, 2>(n, term, sums);
          const bool stiff = sums[1] > 0 &&
                             fabs(h) * sqrt(sums[0] / sums[1]) > (METHOD == 0 ? 3.25 : 6.1);
          if (stiff) {
            stiff_n_nonstiff = 0;
            if (stiff_n_stiff++ >= 15) {
              gave_up = true;
            }
          } else if (stiff_n_stiff > 0) {
            if (stiff_n_nonstiff++ >= 6) {
              stiff_n_stiff = 0;
            }
          }
          if (gave_up) {
            break;
          }
        }
      } /* stages */

      if (lag_failed) {
        break;
      }
      if (gave_up) {
        if (tid == 0) {
          s_cc.error = 1;
          s_cc.code = -7; /* ERR_STIFF */
        }
        break;
      }
      if (!step_done) {
        continue; /* rejected: next attempt */
      }

      /* ---- the rest of an accepted step, src/dopri.c:410-494 ---- */
      if (METHOD == 0) { /* dopri5_save_history, src/dopri_5.c:106-118,84-89; yin is y1 */
        DDE_CLOOP(j, i) {
          const double ydiff = yin[j] - y[j];
          const double bspl = h * k1[j] - ydiff;
          hd[0 * CMAX + j] = y[j];
          hd[1 * CMAX + j] = ydiff;
          hd[2 * CMAX + j] = bspl;
          hd[3 * CMAX + j] = -h * k2[j] + ydiff - bspl;
          hd[4 * CMAX + j] = h * (DP5_D1 * k1[j] + DP5_D3 * k3[j] + DP5_D4 * k4[j] +
                                  DP5_D5 * k5[j] + DP5_D6 * k6[j] + DP5_D7 * k2[j]);
        }
        DDE_CLOOP(j, i) {
          k1[j] = k2[j];
          y[j] = yin[j];
        }
      } else {
        DDE_CLOOP(j, i) {
          k1[j] = k4[j];
          y[j] = k5[j];
        }
      }
      const double t_old = t;
      t += h;

      /* output rows, src/dopri.c:437-456 */
      while (times_idx < n_times) {
        const double tq = a.times[times_idx];
        if (!(last || sign * tq <= sign * t)) {
          break;
        }
        const double theta = (tq - t_old) / h;
        const double theta1 = 1 - theta;
        DDE_CLOOP(j, i) {
          double c[ORDER];
#pragma unroll
          for (int q = 0; q < ORDER; ++q) {
            c[q] = hd[q * CMAX + j];
          }
          yo[i] = coop_interp(c, ORDER, 1, theta, theta1);
        }
        yo += n;
        ++times_idx;
      }

      /* ring_buffer_head_advance, src/dopri.c:460 */
      if (a.n_history > 0) {
        double *dst = ring + (size_t) s_cc.head * (size_t) hl;
        DDE_CLOOP(j, i) {
#pragma unroll
          for (int q = 0; q < ORDER; ++q) {
            dst[(size_t) q * n + i] = hd[q * CMAX + j];
          }
        }
        __syncthreads(); /* everybody has read the context */
        if (tid == 0) {
          dst[ORDER * n] = t_old;
          dst[ORDER * n + 1] = h;
          const unsigned newest = s_cc.count;
          s_cc.head = s_cc.head + 1 == a.n_history ? 0 : s_cc.head + 1;
          s_cc.count = newest + 1;
          if (s_cc.c_valid) {
            const unsigned tail = s_cc.count > (unsigned) a.n_history
                                      ? s_cc.count - (unsigned) a.n_history : 0u;
            if (s_cc.c_idx < tail) {
              s_cc.c_valid = 0; /* recycled by the overwrite policy */
            } else if (s_cc.c_idx + 1 == newest) {
              s_cc.c_tn = sign * t_old; /* its successor exists now */
            }
          }
        }
        __syncthreads();
      }

      if (last) {
        step_size_exit = h_save; /* src/dopri.c:463 */
        if (tid == 0) {
          s_cc.code = 1; /* OK_COMPLETE */
        }
        break;
      }
      if (fabs(h_new) >= a.step_size_max) {
        h_new = copysign(a.step_size_max, sign);
      }
      if (reject) {
        h_new = copysign(fmin(fabs(h_new), fabs(h)), sign);
        reject = false;
      }
      if (stop) { /* critical times, src/dopri.c:476-491 */
        while (tcrit_idx < a.n_tcrit && sign * a.tcrit[tcrit_idx] <= sign * t) {
          ++tcrit_idx;
        }
        t_stop = (tcrit_idx < a.n_tcrit && sign * a.tcrit[tcrit_idx] < sign * t_end)
                     ? a.tcrit[tcrit_idx] : t_end;
        stop = false;
      } else {
        h = h_new;
      }
    } /* adaptive loop */

    /* ---- retire, src/r_dopri.c:412-428 ---- */
    __syncthreads();
    if (tid == 0) {
      a.stats[4 * traj + 0] = n_eval;
      a.stats[4 * traj + 1] = n_step;
      a.stats[4 * traj + 2] = n_accept;
      a.stats[4 * traj + 3] = n_reject;
      a.code[traj] = s_cc.code;
      a.t_last[traj] = t;
      a.step_size[traj] = step_size_exit;
      a.ring_count[traj] = s_cc.count;
    }
    if (a.y_final != nullptr) {
      DDE_CLOOP(j, i) { a.y_final[traj * n + i] = y[j]; }
    }
#undef DDE_COOP_EVAL
  }
#undef DDE_CLOOP
}

} /* namespace dde */

#endif /* DDE_DOPRI_COOP_KERNEL_CUH */
