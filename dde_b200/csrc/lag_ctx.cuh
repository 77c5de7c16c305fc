/* lag_ctx.cuh -- per-trajectory history rings in HBM and the device side of
 * the model plugin API: ylag_1/all/vec/vec_int and yprev_1/all/vec/vec_int.
 *
 * Reference behaviour restated here:
 *   ylag_*        /root/reference/src/dopri.c:776-828
 *   find_time     src/dopri.c:742-769 (result only; the search path is free,
 *                 SURVEY.md A.2#11)
 *   interpolate   src/dopri.c:582-666, src/dopri_5.c:120-127,
 *                 src/dopri_853.c:369-380
 *   yprev_*       src/difeq.c:261-324
 *   ring          restated semantics of mrc-ide/ring
 *
 * The model ABI carries no solver handle (src/dopri.h:49-50); the reference
 * finds the active solver through a process global (src/dopri.c:245).  On
 * the device the context is implicit too: a structure-of-arrays block in
 * shared memory indexed by the calling thread, filled by the kernel before
 * it first calls the model.  Everything a lag query touches on its fast path
 * is there (ring base, commit counter, head slot, t0, search hint); entry
 * contents are read from the trajectory's ring in HBM through L1/L2.
 *
 * Ring layout in HBM: ring[traj][slot][history_len] doubles, slot-major per
 * trajectory, each entry laid out exactly as the reference's
 * (src/dopri.h:94-126: order*n coefficients, then t, then h), so the
 * `history` attribute is a straight copy.  The in-progress ("head") entry
 * lives in the integrating thread's registers and becomes searchable only
 * when committed, which is what ring_buffer_head_advance does
 * (src/dopri.c:460).
 */
#ifndef DDE_LAG_CTX_CUH
#define DDE_LAG_CTX_CUH

#include <stdint.h>

#include "dde_math.h"

#ifndef DDE_TPB
#define DDE_TPB 128 /* threads per block of the thread-per-trajectory kernels */
#endif
#define DDE_WIN 4 /* entries in a thread's staged lag window (power of two) */
/* shared memory a block may spend on staged coefficients; above this the
   window keeps only step times and sizes and coefficients are read from HBM */
#define DDE_WIN_COEF_BUDGET (72 * 1024)

namespace dde {

/* batch-uniform part */
struct LagUniform {
  double sign;   /* +1 forward, -1 backward (src/dopri.c:173) */
  int cap;       /* ring capacity in entries = n_history */
  int n;         /* number of equations */
  int hl;        /* entry length in doubles: order*n + 2, or 1 + n + n_out */
  int order;     /* 5 or 8 (dopri); 0 for difeq */
  int staged;    /* window coefficients live in shared memory */
};

__shared__ LagUniform s_lag_u;
/* per-thread part, structure of arrays: conflict-free LDS/STS */
__shared__ double *s_ring[DDE_TPB];        /* this trajectory's ring base */
__shared__ const double *s_y0[DDE_TPB];    /* pre-history state (src/dopri.c:777-778) */
__shared__ double s_t0[DDE_TPB];           /* sign * start time (src/dopri.c:186) */
__shared__ unsigned s_count[DDE_TPB];      /* entries committed so far */
__shared__ int s_head[DDE_TPB];            /* slot the next commit goes to */
__shared__ int s_code[DDE_TPB];            /* obj->code */
__shared__ int s_error[DDE_TPB];           /* obj->error */
__shared__ long long s_step[DDE_TPB];      /* difeq: obj->step */
__shared__ long long s_step0[DDE_TPB];     /* difeq: obj->step0 */

/* The staged lag window ("shared-memory staging of the bracketing history
 * steps"): DDE_WIN consecutive committed entries base .. base + wn - 1 of the
 * thread's ring, held in shared memory -- sign-adjusted start time and step
 * size always, the dense-output coefficients too when they fit -- entry
 * `idx` in slot idx % DDE_WIN.  wtop is the (exclusive) upper end of the
 * window: the start time of entry base + wn, or +inf while the window's last
 * entry is the newest committed one (a query beyond it extrapolates that
 * entry's polynomial).  Unused slots hold +inf as
 * start time so that the look-up needs no bounds.  A query inside
 * [wt[base], wtop) is served without touching HBM and without a search.
 *
 * The window follows the lag by itself: the integrator tops it up once per
 * step attempt (lag_window_ensure), while all lanes of a warp are in step, up
 * to the highest time the previous attempts queried relative to the end of
 * the step (dq_hi, learnt from misses); newly committed entries are appended
 * straight from registers when the window sits at the head of the ring.  A
 * query outside the window takes the slow path: bisection over the step
 * times stored in the ring (the search path is free, only its result is
 * defined: src/dopri.c:742-769), then a reload of the
 * window at the entry found. */
__shared__ double s_wt[DDE_WIN][DDE_TPB];
__shared__ double s_wh[DDE_WIN][DDE_TPB];
__shared__ double s_wtop[DDE_TPB];
__shared__ unsigned s_wbase[DDE_TPB];
__shared__ unsigned s_wn[DDE_TPB];
__shared__ double s_att_hi[DDE_TPB]; /* sign * (t + h) of the attempt in progress */
__shared__ double s_dq_hi[DDE_TPB];  /* highest query seen, relative to s_att_hi; -inf: none yet */
extern __shared__ double s_wc[];     /* [DDE_WIN][order * n][DDE_TPB] when staged */

#ifdef DDE_LAG_STATS
/* debug build only: [0] misses above, [1] misses below, [2] misses on an empty
   window, [3] entries appended by lag_window_ensure, [4] failures, [5] bisection probes */
static __device__ unsigned long long dde_lag_stats[8];
#define DDE_LAG_COUNT(i) atomicAdd(&dde_lag_stats[i], 1ULL)
#else
#define DDE_LAG_COUNT(i) ((void) 0)
#endif

__device__ __forceinline__ double na_real() {
  /* R's NA_real_ (quiet NaN, low word 1954); payloads do not survive device
     arithmetic, parity treats every NaN alike */
  return __longlong_as_double(0x7FF00000000007A2LL);
}

__device__ __forceinline__ double pos_inf() {
  return __longlong_as_double(0x7ff0000000000000LL);
}

/* slot of the entry committed `back` commits ago (back = 1: most recent) */
__device__ __forceinline__ int ring_slot_back(int head, int back, int cap) {
  int slot = head - back;
  return slot < 0 ? slot + cap : slot;
}

/* ring slot of logical entry idx (0 = first entry ever committed) */
__device__ __forceinline__ const double *ring_entry(int s, unsigned idx) {
  const int slot = ring_slot_back(s_head[s], (int) (s_count[s] - idx), s_lag_u.cap);
  return s_ring[s] + (size_t) slot * (size_t) s_lag_u.hl;
}

__device__ __forceinline__ void lag_window_reset(int s) {
#pragma unroll
  for (int k = 0; k < DDE_WIN; ++k) {
    s_wt[k][s] = pos_inf();
  }
  s_wtop[s] = pos_inf();
  s_wbase[s] = 0;
  s_wn[s] = 0;
  s_att_hi[s] = 0.0;
  s_dq_hi[s] = -pos_inf();
}

/* Copy committed entry idx from the ring into its window slot.  Returns the
 * sign-adjusted start time of the entry after it, +inf if there is none. */
__device__ __forceinline__ double lag_window_load(int s, unsigned idx) {
  const int hl = s_lag_u.hl;
  const int nc = hl - 2;
  const double sign = s_lag_u.sign;
  const unsigned count = s_count[s];
  const double *e = ring_entry(s, idx);
  const int w = (int) (idx & (DDE_WIN - 1));
  double top = pos_inf();
  if (idx + 1 < count) {
    top = sign * ring_entry(s, idx + 1)[nc];
  }
  s_wt[w][s] = sign * e[nc];
  s_wh[w][s] = e[nc + 1];
  if (s_lag_u.staged) {
    double *dst = s_wc + (size_t) w * nc * DDE_TPB + s;
    for (int c = 0; c < nc; ++c) {
      dst[(size_t) c * DDE_TPB] = e[c];
    }
  }
  return top;
}

/* Once per step attempt, all lanes together: extend the window upwards until
 * it covers sign * (t + h) + dq_hi, dropping entries from its lower end as
 * needed.  Purely a prefetch: any query it fails to anticipate still finds
 * its entry through the slow path. */
__device__ __forceinline__ void lag_window_ensure(int s, double att_hi) {
  s_att_hi[s] = att_hi;
  const double need = att_hi + s_dq_hi[s];
#pragma unroll 1
  for (int it = 0; it < DDE_WIN; ++it) {
    unsigned wn = s_wn[s];
    if (!(wn > 0 && need >= s_wtop[s])) {
      break;
    }
    unsigned base = s_wbase[s];
    if (wn == DDE_WIN) {
      ++base;
      --wn;
    }
    DDE_LAG_COUNT(3);
    const double top = lag_window_load(s, base + wn); /* overwrites the dropped slot */
    s_wbase[s] = base;
    s_wn[s] = wn + 1;
    s_wtop[s] = top;
  }
}

/* Bookkeeping after the integrating thread has written the head entry to
 * its ring slot: advance the ring (ring_buffer_head_advance, src/dopri.c:460)
 * and keep the window consistent -- drop an entry the overwrite policy just
 * recycled, append the new entry if the window ends at the head of the ring.
 * hd: the new entry's coefficients (order * n, in registers), then its start
 * time and step size. */
template <int NCOEF_MAX>
__device__ __forceinline__ void ring_advance(int s, int cap, const double *hd, int nc,
                                             double t_old, double h) {
  int head = s_head[s];
  head = head + 1 == cap ? 0 : head + 1;
  s_head[s] = head;
  const unsigned count = s_count[s] + 1;
  s_count[s] = count;
  const unsigned newest = count - 1;
  unsigned wn = s_wn[s], base = s_wbase[s];
  const unsigned tail = count > (unsigned) cap ? count - (unsigned) cap : 0u;
  if (wn > 0 && base < tail) {
    s_wt[base & (DDE_WIN - 1)][s] = pos_inf();
    ++base;
    --wn;
    s_wbase[s] = base;
    s_wn[s] = wn;
  }
  const double st_old = s_lag_u.sign * t_old;
  bool append = false;
  if (wn == 0) {
    base = newest;
    s_wbase[s] = base;
    append = true;
  } else if (base + wn == newest) {
    if (wn < DDE_WIN) {
      append = true;
    } else {
      s_wtop[s] = st_old;
    }
  }
  if (append) {
    const int w = (int) (newest & (DDE_WIN - 1));
    s_wt[w][s] = st_old;
    s_wh[w][s] = h;
    if (s_lag_u.staged) {
      double *dst = s_wc + (size_t) w * nc * DDE_TPB + s;
#pragma unroll
      for (int c = 0; c < NCOEF_MAX; ++c) {
        if (c < nc) {
          dst[(size_t) c * DDE_TPB] = hd[c];
        }
      }
    }
    s_wn[s] = wn + 1;
    s_wtop[s] = pos_inf();
  }
}

struct LagHit {
  bool ok;         /* false: no entry (ERR_YLAG_FAIL recorded) */
  unsigned idx;    /* logical index of the entry */
  double t_old, h; /* its start time and step */
};

/* Slow path of the look-up: the last committed entry whose time is <= t
 * (>= t when integrating backwards), src/dopri.c:742-769, by bisection over
 * the ring, then the window is reloaded starting at that entry.  Returns
 * false after recording ERR_YLAG_FAIL when there is no such entry. */
static __device__ __noinline__ bool lag_window_miss(int s, double st) {
  const int cap = s_lag_u.cap;
  const int idx_t = s_lag_u.hl - 2;
  const double sign = s_lag_u.sign;
  const unsigned count = s_count[s];
  const unsigned used = count < (unsigned) cap ? count : (unsigned) cap;
#define DDE_PRED(j) (sign * ring_entry(s, (j))[idx_t] <= st)
  unsigned lo = count - used, hi = count; /* want: pred(lo) holds, pred(hi) fails */
  if (used == 0 || !DDE_PRED(lo)) {
    DDE_LAG_COUNT(4);
    s_error[s] = 1;
    s_code[s] = -6; /* ERR_YLAG_FAIL */
    return false;
  }
  /* start from the window: the entry wanted is usually next to it */
  const unsigned wn = s_wn[s];
  if (wn == 0) {
    DDE_LAG_COUNT(2);
  }
  if (wn > 0) {
    const unsigned base = s_wbase[s];
    DDE_LAG_COUNT(st >= s_wtop[s] ? 0 : 1);
    if (st >= s_wtop[s]) {
      const unsigned g = base + wn; /* exists: wtop is finite */
      if (g > lo && g < hi) {
        lo = g; /* its start time is wtop <= st */
        if (g + 1 < hi) {
          if (DDE_PRED(g + 1)) {
            lo = g + 1;
          } else {
            hi = g + 1;
          }
        }
      }
    } else if (base > lo && base < hi) {
      hi = base; /* st < wt[base] */
      if (DDE_PRED(base - 1)) {
        lo = base - 1;
      }
    }
  }
  while (hi - lo > 1) {
    const unsigned mid = lo + (hi - lo) / 2;
    DDE_LAG_COUNT(5);
    if (DDE_PRED(mid)) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
#undef DDE_PRED
  /* remember how far above the end of the step this query reached */
  if (st - s_att_hi[s] > s_dq_hi[s]) {
    s_dq_hi[s] = st - s_att_hi[s];
  }
  /* reload the window: entries lo .. lo + DDE_WIN - 1 as far as committed */
  double top = pos_inf();
  unsigned k = 0;
#pragma unroll 1
  for (int j = 0; j < DDE_WIN; ++j) {
    if (lo + j < count) {
      top = lag_window_load(s, lo + j);
      k = j + 1;
    } else {
      s_wt[(lo + j) & (DDE_WIN - 1)][s] = pos_inf();
    }
  }
  s_wbase[s] = lo;
  s_wn[s] = k;
  s_wtop[s] = top;
  return true;
}

/* The entry a lag query at time t reads, src/dopri.c:742-769.  On failure
 * obj->error / ERR_YLAG_FAIL are set and ok is false. */
__device__ __forceinline__ LagHit find_time(int s, double t) {
  const double sign = s_lag_u.sign;
  const double st = sign * t; /* entries are ordered by sign * time */
  LagHit hit;
  unsigned base = s_wbase[s];
  hit.ok = true;
  if (!(s_wn[s] > 0 && st >= s_wt[base & (DDE_WIN - 1)][s] && st < s_wtop[s])) {
    hit.ok = lag_window_miss(s, st);
    base = s_wbase[s];
  }
  /* unused slots hold +inf, so this counts the entries that start at or before st */
  unsigned j = 0;
#pragma unroll
  for (int k = 1; k < DDE_WIN; ++k) {
    j += st >= s_wt[(base + k) & (DDE_WIN - 1)][s] ? 1u : 0u;
  }
  hit.idx = base + j;
  const int w = (int) (hit.idx & (DDE_WIN - 1));
  hit.t_old = sign * s_wt[w][s];
  hit.h = s_wh[w][s];
  return hit;
}

/* Dense-output polynomial of one component; c points at the component's
 * first coefficient, coefficients are `n` doubles apart
 * (src/dopri_5.c:120-127, src/dopri_853.c:369-380). */
__device__ __forceinline__ double interp5(int n, double theta, double theta1,
                                          const double *c) {
  return c[0] + theta * (c[n] + theta1 * (c[2 * n] + theta * (c[3 * n] + theta1 * c[4 * n])));
}

__device__ __forceinline__ double interp853(int n, double theta, double theta1,
                                            const double *c) {
  const double tmp = c[4 * n] + theta * (c[5 * n] + theta1 * (c[6 * n] + theta * c[7 * n]));
  return c[0] + theta * (c[n] + theta1 * (c[2 * n] + theta * (c[3 * n] + theta1 * tmp)));
}

__device__ __forceinline__ double interp_entry(const double *c, int order, int stride,
                                               double theta, double theta1) {
  return order == 5 ? interp5(stride, theta, theta1, c) : interp853(stride, theta, theta1, c);
}

/* Components idx[0..nidx) (NULL: 0..nidx-1) of the dense output of the entry
 * `hit` at time t, src/dopri.c:582-666.  Two copies so that the staged
 * window is read with shared-memory loads and the ring with global ones. */
template <class Index>
__device__ __forceinline__ void lag_eval(int s, const LagHit &hit, double t, const Index *idx,
                                         size_t nidx, double *y) {
  const int n = s_lag_u.n, order = s_lag_u.order;
  const double theta = (t - hit.t_old) / hit.h;
  const double theta1 = 1 - theta;
  if (s_lag_u.staged) {
    const int w = (int) (hit.idx & (DDE_WIN - 1));
    const double *c = s_wc + (size_t) w * (order * n) * DDE_TPB + s;
    for (size_t i = 0; i < nidx; ++i) {
      const size_t comp = idx ? (size_t) idx[i] : i;
      y[i] = interp_entry(c + comp * DDE_TPB, order, n * DDE_TPB, theta, theta1);
    }
  } else {
    const double *c = ring_entry(s, hit.idx);
    for (size_t i = 0; i < nidx; ++i) {
      const size_t comp = idx ? (size_t) idx[i] : i;
      y[i] = interp_entry(c + comp, order, n, theta, theta1);
    }
  }
}

} /* namespace dde */

/* ---- the model-facing functions (declared in <dde/dde.h>) -------------- */

static __device__ __forceinline__ double ylag_1(double t, size_t i) {
  const int s = threadIdx.x;
  if (dde::s_lag_u.sign * t <= dde::s_t0[s]) {
    return dde::s_y0[s][i];
  }
  const dde::LagHit hit = dde::find_time(s, t);
  if (!hit.ok) {
    return dde::na_real();
  }
  double y;
  dde::lag_eval<size_t>(s, hit, t, &i, 1, &y);
  return y;
}

static __device__ __forceinline__ void ylag_all(double t, double *y) {
  const int s = threadIdx.x;
  const int n = dde::s_lag_u.n;
  if (dde::s_lag_u.sign * t <= dde::s_t0[s]) {
    for (int i = 0; i < n; ++i) {
      y[i] = dde::s_y0[s][i];
    }
    return;
  }
  const dde::LagHit hit = dde::find_time(s, t);
  if (hit.ok) {
    dde::lag_eval<size_t>(s, hit, t, nullptr, (size_t) n, y);
  }
}

template <class Index>
static __device__ __forceinline__ void dde_ylag_indexed(double t, const Index *idx,
                                                        size_t nidx, double *y) {
  const int s = threadIdx.x;
  if (dde::s_lag_u.sign * t <= dde::s_t0[s]) {
    for (size_t i = 0; i < nidx; ++i) {
      y[i] = dde::s_y0[s][idx[i]];
    }
    return;
  }
  const dde::LagHit hit = dde::find_time(s, t);
  if (hit.ok) {
    dde::lag_eval<Index>(s, hit, t, idx, nidx, y);
  }
}

static __device__ __forceinline__ void ylag_vec(double t, const size_t *idx, size_t nidx,
                                                double *y) {
  dde_ylag_indexed<size_t>(t, idx, nidx, y);
}

static __device__ __forceinline__ void ylag_vec_int(double t, const int *idx, size_t nidx,
                                                    double *y) {
  dde_ylag_indexed<int>(t, idx, nidx, y);
}

/* difeq: the y block of the entry for `step`, or NULL + error
 * (src/difeq.c:261-272).  Entry = {step, y[n], out[n_out]}. */
static __device__ __forceinline__ const double *dde_find_step(int s, int step) {
  const int offset = (int) dde::s_step[s] - step;
  const int cap = dde::s_lag_u.cap;
  const unsigned count = dde::s_count[s];
  const unsigned used = count < (unsigned) cap ? count : (unsigned) cap;
  if (cap == 0 || offset < 0 || (unsigned) offset >= used) {
    dde::s_error[s] = 1;
    dde::s_code[s] = -9; /* DDE_ERR_DIFEQ_HISTORY */
    return nullptr;
  }
  const int slot = dde::ring_slot_back(dde::s_head[s], offset + 1, cap);
  return dde::s_ring[s] + (size_t) slot * dde::s_lag_u.hl + 1;
}

static __device__ __forceinline__ double yprev_1(int step, size_t i) {
  const int s = threadIdx.x;
  if (step <= (int) dde::s_step0[s]) {
    return dde::s_y0[s][i];
  }
  const double *e = dde_find_step(s, step);
  return e == nullptr ? dde::na_real() : e[i];
}

static __device__ __forceinline__ void yprev_all(int step, double *y) {
  const int s = threadIdx.x;
  const int n = dde::s_lag_u.n;
  if (step <= (int) dde::s_step0[s]) {
    for (int i = 0; i < n; ++i) {
      y[i] = dde::s_y0[s][i];
    }
    return;
  }
  const double *e = dde_find_step(s, step);
  if (e != nullptr) {
    for (int i = 0; i < n; ++i) {
      y[i] = e[i];
    }
  }
}

template <class Index>
static __device__ __forceinline__ void dde_yprev_indexed(int step, const Index *idx,
                                                         size_t nidx, double *y) {
  const int s = threadIdx.x;
  if (step <= (int) dde::s_step0[s]) {
    for (size_t i = 0; i < nidx; ++i) {
      y[i] = dde::s_y0[s][idx[i]];
    }
    return;
  }
  const double *e = dde_find_step(s, step);
  if (e != nullptr) {
    for (size_t i = 0; i < nidx; ++i) {
      y[i] = e[idx[i]];
    }
  }
}

static __device__ __forceinline__ void yprev_vec(int step, const size_t *idx, size_t nidx,
                                                 double *y) {
  dde_yprev_indexed<size_t>(step, idx, nidx, y);
}

static __device__ __forceinline__ void yprev_vec_int(int step, const int *idx, size_t nidx,
                                                     double *y) {
  dde_yprev_indexed<int>(step, idx, nidx, y);
}

#endif /* DDE_LAG_CTX_CUH */
