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
#define DDE_WIN_EMPTY 0xffffffffu

namespace dde {

/* batch-uniform part */
struct LagUniform {
  double sign;   /* +1 forward, -1 backward (src/dopri.c:173) */
  int cap;       /* ring capacity in entries = n_history */
  int n;         /* number of equations */
  int hl;        /* entry length in doubles: order*n + 2, or 1 + n + n_out */
  int order;     /* 5 or 8 (dopri); 0 for difeq */
};

__shared__ LagUniform s_lag_u;
/* per-thread part, structure of arrays: conflict-free LDS/STS */
__shared__ double *s_ring[DDE_TPB];        /* this trajectory's ring base */
__shared__ const double *s_y0[DDE_TPB];    /* pre-history state (src/dopri.c:777-778) */
__shared__ double s_t0[DDE_TPB];           /* sign * start time (src/dopri.c:186) */
__shared__ unsigned s_count[DDE_TPB];      /* entries committed so far */
__shared__ int s_head[DDE_TPB];            /* slot the next commit goes to */
__shared__ unsigned s_hint[DDE_TPB];       /* logical index the last search found */
/* Staged bracket of the lag window: two adjacent committed entries a = s_win
 * and b = a + 1 with their sign-adjusted start times, the (exclusive) upper
 * bound of b, and their step sizes.  A query that falls in [ta, tc) needs no
 * search and no dependent load: slot and theta come from shared memory while
 * the coefficient loads from the ring are already in flight. */
__shared__ unsigned s_win[DDE_TPB];        /* logical index of a; DDE_WIN_EMPTY = none */
__shared__ double s_win_ta[DDE_TPB], s_win_tb[DDE_TPB], s_win_tc[DDE_TPB];
__shared__ double s_win_ha[DDE_TPB], s_win_hb[DDE_TPB];
__shared__ int s_code[DDE_TPB];            /* obj->code */
__shared__ int s_error[DDE_TPB];           /* obj->error */
__shared__ long long s_step[DDE_TPB];      /* difeq: obj->step */
__shared__ long long s_step0[DDE_TPB];     /* difeq: obj->step0 */

__device__ __forceinline__ double na_real() {
  /* R's NA_real_ (quiet NaN, low word 1954); payloads do not survive device
     arithmetic, parity treats every NaN alike */
  return __longlong_as_double(0x7FF00000000007A2LL);
}

/* slot of the entry committed `back` commits ago (back = 1: most recent) */
__device__ __forceinline__ int ring_slot_back(int head, int back, int cap) {
  int slot = head - back;
  return slot < 0 ? slot + cap : slot;
}

/* Bookkeeping after the integrating thread has written the head entry to
 * its slot: advance the ring (ring_buffer_head_advance, src/dopri.c:460) and
 * drop the staged bracket if the commit changes what it stands for -- its
 * upper bound was open (b or b's successor did not exist yet), or the
 * overwrite policy just recycled one of its slots. */
__device__ __forceinline__ void ring_advance(int s, int cap) {
  int head = s_head[s];
  head = head + 1 == cap ? 0 : head + 1;
  s_head[s] = head;
  const unsigned count = s_count[s] + 1;
  s_count[s] = count;
  const unsigned win = s_win[s];
  if (win != DDE_WIN_EMPTY) {
    const bool open_top = win + 3 >= count;                /* a + 2 was not committed before */
    const bool recycled = count > (unsigned) cap && win < count - (unsigned) cap;
    if (open_top || recycled) {
      s_win[s] = DDE_WIN_EMPTY;
    }
  }
}

struct LagHit {
  const double *entry; /* NULL: no entry (error already recorded) */
  double t_old, h;     /* the entry's start time and step */
};

/* The last committed entry whose time is <= t (>= t when integrating
 * backwards), src/dopri.c:742-769.  On failure obj->error / ERR_YLAG_FAIL
 * are set and entry is NULL. */
__device__ __forceinline__ LagHit find_time(int s, double t) {
  const int cap = s_lag_u.cap;
  const int hl = s_lag_u.hl;
  const double sign = s_lag_u.sign;
  const double st = sign * t; /* entries are ordered by sign * time */
  const unsigned count = s_count[s];
  const int head = s_head[s];
  const double *ring = s_ring[s];
  LagHit hit;
  /* fast path: the staged bracket */
  const unsigned win = s_win[s];
  if (win != DDE_WIN_EMPTY && st >= s_win_ta[s] && st < s_win_tc[s]) {
    const bool second = st >= s_win_tb[s];
    const unsigned idx = win + (second ? 1u : 0u);
    hit.entry = ring + (size_t) ring_slot_back(head, (int) (count - idx), cap) * hl;
    hit.t_old = sign * (second ? s_win_tb[s] : s_win_ta[s]);
    hit.h = second ? s_win_hb[s] : s_win_ha[s];
    return hit;
  }
  /* slow path: bisection over the stored step times, started from the last
     result (the search path is free, only its result is defined) */
  const unsigned used = count < (unsigned) cap ? count : (unsigned) cap;
  const int idx_t = hl - 2;
#define DDE_ENTRY(j) (ring + (size_t) ring_slot_back(head, (int) (count - (j)), cap) * hl)
#define DDE_PRED(j) (sign * DDE_ENTRY(j)[idx_t] <= st)
  unsigned lo = count - used, hi = count; /* want: pred(lo) holds, pred(hi) fails */
  bool ok = used > 0;
  if (ok) {
    const unsigned h = s_hint[s];
    if (h >= lo && h < hi && DDE_PRED(h)) {
      lo = h;
      if (h + 1 < hi) {
        if (DDE_PRED(h + 1)) {
          lo = h + 1;
        } else {
          hi = h + 1;
        }
      }
    } else {
      ok = DDE_PRED(lo);
      if (h > lo && h < hi) {
        hi = h;
      }
    }
  }
  if (!ok) {
    s_error[s] = 1;
    s_code[s] = -6; /* ERR_YLAG_FAIL */
    hit.entry = nullptr;
    hit.t_old = 0.0;
    hit.h = 1.0;
    return hit;
  }
  while (hi - lo > 1) {
    const unsigned mid = lo + (hi - lo) / 2;
    if (DDE_PRED(mid)) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
  s_hint[s] = lo;
  /* stage the bracket [lo, lo + 1] */
  const double *ea = DDE_ENTRY(lo);
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const double ta = ea[idx_t], ha = ea[idx_t + 1];
  double tb = inf, tc = inf, hb = 1.0;
  if (lo + 1 < count) {
    const double *eb = DDE_ENTRY(lo + 1);
    tb = sign * eb[idx_t];
    hb = eb[idx_t + 1];
    if (lo + 2 < count) {
      tc = sign * DDE_ENTRY(lo + 2)[idx_t];
    }
  }
#undef DDE_PRED
#undef DDE_ENTRY
  s_win[s] = lo;
  s_win_ta[s] = sign * ta;
  s_win_tb[s] = tb;
  s_win_tc[s] = tc;
  s_win_ha[s] = ha;
  s_win_hb[s] = hb;
  hit.entry = ea;
  hit.t_old = ta;
  hit.h = ha;
  return hit;
}

/* Dense-output polynomial of one component; c points at the component's
 * first coefficient, coefficients are n apart. */
__device__ __forceinline__ double interp5(int n, double theta, double theta1,
                                          const double *c) {
  return c[0] + theta * (c[n] + theta1 * (c[2 * n] + theta * (c[3 * n] + theta1 * c[4 * n])));
}

__device__ __forceinline__ double interp853(int n, double theta, double theta1,
                                            const double *c) {
  const double tmp = c[4 * n] + theta * (c[5 * n] + theta1 * (c[6 * n] + theta * c[7 * n]));
  return c[0] + theta * (c[n] + theta1 * (c[2 * n] + theta * (c[3 * n] + theta1 * tmp)));
}

__device__ __forceinline__ double interp_entry(const double *e, int order, int n,
                                               double theta, double theta1, size_t i) {
  return order == 5 ? interp5(n, theta, theta1, e + i) : interp853(n, theta, theta1, e + i);
}

} /* namespace dde */

/* ---- the model-facing functions (declared in <dde/dde.h>) -------------- */

static __device__ __forceinline__ double ylag_1(double t, size_t i) {
  const int s = threadIdx.x;
  if (dde::s_lag_u.sign * t <= dde::s_t0[s]) {
    return dde::s_y0[s][i];
  }
  const dde::LagHit hit = dde::find_time(s, t);
  if (hit.entry == nullptr) {
    return dde::na_real();
  }
  const int n = dde::s_lag_u.n, order = dde::s_lag_u.order;
  const double theta = (t - hit.t_old) / hit.h;
  const double theta1 = 1 - theta;
  return dde::interp_entry(hit.entry, order, n, theta, theta1, i);
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
  if (hit.entry != nullptr) {
    const int order = dde::s_lag_u.order;
    const double theta = (t - hit.t_old) / hit.h;
    const double theta1 = 1 - theta;
    for (int i = 0; i < n; ++i) {
      y[i] = dde::interp_entry(hit.entry, order, n, theta, theta1, (size_t) i);
    }
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
  if (hit.entry != nullptr) {
    const int n = dde::s_lag_u.n, order = dde::s_lag_u.order;
    const double theta = (t - hit.t_old) / hit.h;
    const double theta1 = 1 - theta;
    for (size_t i = 0; i < nidx; ++i) {
      y[i] = dde::interp_entry(hit.entry, order, n, theta, theta1, (size_t) idx[i]);
    }
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
