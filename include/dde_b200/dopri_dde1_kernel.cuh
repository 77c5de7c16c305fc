/* dopri_dde1_kernel.cuh -- dop853 for ONE delay equation with one declared lag, one
 * trajectory per thread: the stepping kernel of the Mackey-Glass batch
 * (BASELINE.json configs[1]).
 *
 * Same arithmetic, operation for operation, as dopri_tpt_kernel (dopri_kernels.cuh) and
 * therefore as the reference:
 *   adaptive loop            /root/reference/src/dopri.c:291-514
 *   dop853 step/error/dense  src/dopri_853.c:159-380
 *   controller               src/dopri.c:516-525
 *   history look-up          src/dopri.c:742-788, src/dopri_853.c:369-380
 * What differs is the shape of the program.  The general kernel spends 3/4 of its
 * instructions around the arithmetic (profiles/r1g: 35 % of 3.0e10 warp instructions are
 * FP64, issue slots 42 % busy with 12 warps per SM, the hot path past the instruction
 * cache); this one is built from an instruction budget:
 *
 *   * The solver's context is registers.  The model declares its lag (M::lags,
 *     dde_model_def.h), so the kernel makes the look-up itself, with the window state in
 *     registers, and hands the model the answer (the memo of lag_ctx.cuh: ylag_1 is one
 *     comparison).  Shared memory holds the staged window and that memo, nothing else:
 *     432 bytes per thread with a window of five entries, four 128-thread blocks per SM.
 *   * A window entry is five 16-byte pairs -- (c0,c1) (c2,c3) (c4,c5) (c6,c7) (t,h) --
 *     at [slot][pair][thread]: a look-up is W loads of start times at fixed offsets, a
 *     count, and five 128-bit loads off one computed address.  No branch on its path:
 *     the division is the compiler's fast path without its test (dde_div_fast), the
 *     pre-history is a select, and anything irregular (the entry not in the window, an
 *     irregular quotient) clears a flag that sends the CONSUMER to the out-of-line
 *     search of the ring.
 *   * Evaluation st + 1's look-up is made during evaluation st: the time is known (t and
 *     h are fixed over an attempt) and the chain look-up -> division -> Horner form (~30
 *     dependent FP64 operations) is independent of evaluation st's pow chain, so the two
 *     sit in one basic block and the scheduler interleaves them.
 *   * The sixteen evaluations of a step are one loop around ONE inlined copy of the model
 *     and one warp-uniform switch that files the result and forms the next stage input;
 *     lanes walk an attempt in lock-step, each with its own t and h.
 *   * The reference's redundant evaluation at the old time (src/dopri.c:400-403: only the
 *     stiffness test reads its result, and this kernel is not launched with that test on)
 *     is counted, not computed; its look-up -- the one thing about it that can be observed,
 *     through ERR_YLAG_FAIL -- is the attempt's lowest query, which the window is
 *     positioned on at the start of every attempt, so whether it would have succeeded
 *     is known (and otherwise it is made, out of line).
 *
 * Launched (dde_model.cuh) for dop853, forward time, a fresh start, no critical times,
 * no output function, no stiffness test; everything else runs on dopri_tpt_kernel.
 */
#ifndef DDE_DOPRI_DDE1_KERNEL_CUH
#define DDE_DOPRI_DDE1_KERNEL_CUH

#include "dopri_kernels.cuh"

#ifndef DDE1_WIN
#define DDE1_WIN 5 /* entries in the staged window */
#endif
#ifndef DDE1_WALK
#define DDE1_WALK 4 /* ring entries probed per round trip by a query above the window */
#endif
#ifndef DDE1_MIN_BLOCKS
#define DDE1_MIN_BLOCKS 4 /* resident 128-thread blocks per SM (128 registers) */
#endif
/* one window slot, all threads: 5 pairs x 16 bytes */
#define DDE1_PAIR_BYTES (16u * DDE_TPB)
#define DDE1_SLOT_BYTES (5u * DDE1_PAIR_BYTES)
/* shared memory of a block: the pow / exp tables, then the window */
#define DDE1_SMEM_BYTES (DDE_TAB_SHARED_BYTES + DDE1_WIN * DDE1_SLOT_BYTES)

namespace dde {

/* What the model sees instead of the global history functions and libm when its source
 * is compiled a second time INSIDE a struct derived from this one (dde_model_scoped.inc):
 * unqualified calls to ylag_1 / pow / exp in the model's functions -- now member
 * functions -- find these members first.  The solver's side of the model ABI
 * (src/dopri.h:49-50 carries no solver handle; the reference uses a process global,
 * src/dopri.c:245; the general kernels a per-thread block in shared memory) is then a few
 * registers of the calling thread: the declared query and its answer, where the tables
 * of pow are, whether the model asked for something it did not declare. */
struct Dde1Scope {
  double memo_t, memo_v; /* the query answered ahead of this evaluation, and the answer */
  unsigned tab;          /* shared-window address of the pow / exp tables */
  bool undeclared;       /* a look-up that was not declared (M::lags): DDE_ERR_LAG_UNDECLARED */
  __device__ __forceinline__ double ylag_1(double t, size_t) {
    /* "exactly that time": the same bits.  (Compared as integers, so that the compiler,
       which sees the model compute t - tau twice from the same operands, can tell that a
       value equals itself -- as doubles it would have to allow for NaN -- and folds the
       whole test away for a model that keeps its declaration.) */
    const bool declared = __double_as_longlong(t) == __double_as_longlong(memo_t);
    undeclared |= !declared;
    return declared ? memo_v : na_real();
  }
  __device__ __forceinline__ void ylag_all(double t, double *y) { y[0] = ylag_1(t, 0); }
  __device__ __forceinline__ void ylag_vec(double t, const size_t *, size_t nidx, double *y) {
    for (size_t i = 0; i < nidx; ++i) {
      y[i] = ylag_1(t, 0);
    }
  }
  __device__ __forceinline__ void ylag_vec_int(double t, const int *, size_t nidx, double *y) {
    for (size_t i = 0; i < nidx; ++i) {
      y[i] = ylag_1(t, 0);
    }
  }
  __device__ __forceinline__ double dde_scope_pow(double x, double y) {
    return dde_pow_shared(tab, x, y);
  }
  __device__ __forceinline__ double dde_scope_exp(double x) { return dde_exp(x); }
};

/* Time of right-hand-side evaluation j of a dop853 step (attempt: 0 .. 10; accepted steps
 * go on with 11 .. 15): src/dopri_853.c:178-240, src/dopri.c:400-403,
 * src/dopri_853.c:304,330-347.  The one place these are written, so that a look-up made
 * ahead of an evaluation sees the evaluation's time bit for bit. */
__device__ __forceinline__ double dp8_stage_time(int j, double t, double h) {
  switch (j) {
    case 0: return t + DP8_C2 * h;
    case 1: return t + DP8_C3 * h;
    case 2: return t + DP8_C4 * h;
    case 3: return t + DP8_C5 * h;
    case 4: return t + DP8_C6 * h;
    case 5: return t + DP8_C7 * h;
    case 6: return t + DP8_C8 * h;
    case 7: return t + DP8_C9 * h;
    case 8: return t + DP8_C10 * h;
    case 9: return t + DP8_C11 * h;
    case 11: return t;
    case 13: return t + DP8_C14 * h;
    case 14: return t + DP8_C15 * h;
    case 15: return t + DP8_C16 * h;
    default: return t + h; /* 10, 12 */
  }
}

/* The state of a thread's staged window, in registers.  Entries wbase .. wbase + wn - 1
 * of the trajectory's ring (logical indices: 0 = first ever committed), entry idx in slot
 * idx % DDE1_WIN; unused slots hold +inf as start time, so that the number of slots that
 * start at or before a query locates its entry without a search.  `below` is the
 * shared-window address of the slot BEFORE the base entry's (cyclically), `wtop` the
 * start time of entry wbase + wn, or +inf while the window's last entry is the newest
 * committed one (a query beyond it extrapolates that entry's polynomial). */
struct Dde1Win {
  unsigned first; /* shared-window address of slot 0, this thread's column */
  unsigned below;
  unsigned wbase, wn;
  double wtop;
};

__device__ __forceinline__ unsigned dde1_wrap(const Dde1Win &w, unsigned addr) {
  return addr >= w.first + DDE1_WIN * DDE1_SLOT_BYTES ? addr - DDE1_WIN * DDE1_SLOT_BYTES : addr;
}

/* the slot of the window's k-th entry (k = 0: the base) */
__device__ __forceinline__ unsigned dde1_slot(const Dde1Win &w, unsigned k) {
  return dde1_wrap(w, w.below + (k + 1u) * DDE1_SLOT_BYTES);
}

__device__ __forceinline__ void dde1_window_reset(Dde1Win &w) {
#pragma unroll
  for (int k = 0; k < DDE1_WIN; ++k) {
    dde1_sts(w.first + k * DDE1_SLOT_BYTES + 4u * DDE1_PAIR_BYTES, pos_inf());
  }
  w.below = w.first + (DDE1_WIN - 1) * DDE1_SLOT_BYTES;
  w.wbase = 0;
  w.wn = 0;
  w.wtop = pos_inf();
}

__device__ __forceinline__ void dde1_window_evict(Dde1Win &w) {
  w.below = dde1_wrap(w, w.below + DDE1_SLOT_BYTES); /* now the evicted entry's slot */
  dde1_sts(w.below + 4u * DDE1_PAIR_BYTES, pos_inf());
  ++w.wbase;
  --w.wn;
}

/* The value ylag_1(tq, 0) returns (src/dopri.c:776-788), from the staged window, without
 * a branch and without side effects; *ok tells whether it is that value -- or is not to
 * be had that way (the entry is not in the window, there is none, an irregular quotient):
 * the caller then asks dde1_lookup_ring.  The arithmetic is that of find_time + lag_eval
 * (lag_ctx.cuh), operation for operation. */
__device__ __forceinline__ double dde1_lookup(const Dde1Win &w, double t0, double y0v, double tq,
                                              bool *ok) {
  unsigned cnt = 0;
#pragma unroll
  for (int k = 0; k < DDE1_WIN; ++k) {
    const double ts = dde1_lds(w.first + k * DDE1_SLOT_BYTES + 4u * DDE1_PAIR_BYTES);
    cnt += tq >= ts ? 1u : 0u;
  }
  /* the entry's slot: in range whatever cnt is */
  const unsigned ea = dde1_wrap(w, w.below + cnt * DDE1_SLOT_BYTES);
  DDE_CHECK(ea >= w.first && ea < w.first + DDE1_WIN * DDE1_SLOT_BYTES &&
            (ea - w.first) % DDE1_SLOT_BYTES == 0 && cnt <= DDE1_WIN);
  double c0, c1, c2, c3, c4, c5, c6, c7, ts, hh;
  dde1_lds2(ea, c0, c1);
  dde1_lds2(ea + DDE1_PAIR_BYTES, c2, c3);
  dde1_lds2(ea + 2u * DDE1_PAIR_BYTES, c4, c5);
  dde1_lds2(ea + 3u * DDE1_PAIR_BYTES, c6, c7);
  dde1_lds2(ea + 4u * DDE1_PAIR_BYTES, ts, hh);
  bool dok;
  const double num = tq - ts;
  const double theta = dde_div_fast(num, hh, &dok);
  const double theta1 = 1 - theta;
  const double tmp = c4 + theta * (c5 + theta1 * (c6 + theta * c7));
  const double acc = c0 + theta * (c1 + theta1 * (c2 + theta * (c3 + theta1 * tmp)));
  const bool pre = tq <= t0; /* src/dopri.c:777-778 */
  *ok = pre | ((cnt > 0u) & (tq < w.wtop) & dok);
  return pre ? y0v : acc;
}

/* The same value from the trajectory's ring in HBM, for a query the window could not
 * answer: the last committed entry whose time is <= tq (src/dopri.c:742-769: only the
 * result of the search is defined), or ERR_YLAG_FAIL.  A fresh forward run: logical entry
 * idx sits in slot idx % cap.  `top` is the first entry above the window (wbase + wn) and
 * `wtop` its start time (+inf: there is none): a query at or above it -- a step attempt
 * that spans more old entries than the window holds -- is found by walking up from there,
 * DDE1_WALK probes per round trip; anything else by bisection.  Out of line and by value. */
struct Dde1Found {
  double value;
  bool failed; /* ERR_YLAG_FAIL */
};

static __device__ __noinline__ Dde1Found dde1_lookup_ring(const double *ring, unsigned count,
                                                          int cap, double t0, double y0v,
                                                          double tq, unsigned top, double wtop) {
  Dde1Found f;
  f.failed = false;
  f.value = y0v;
  DDE_LAG_COUNT(6);
  if (tq <= t0) {
    return f;
  }
  const unsigned used = count < (unsigned) cap ? count : (unsigned) cap;
  unsigned lo = count - used, hi = count;
#define DDE1_TIME(j)                                                                       \
  (DDE_LAG_COUNT(5), DDE_CHECK((j) < count && count - (j) <= (unsigned) cap),              \
   ring[(size_t) ((j) % (unsigned) cap) * 10u + 8u])
  if (tq >= wtop) { /* entry `top` exists and starts at wtop */
    DDE_LAG_COUNT(0);
    DDE_CHECK(top < count && count - top <= (unsigned) cap);
    lo = top;
    /* the start times of the next DDE1_WALK entries in ONE round trip to L2 / HBM (a lone
       lane walks while its warp waits: what counts is round trips, not instructions).
       They are sorted, so how many of them are <= tq is how far to go. */
    unsigned slot = (lo + 1) % (unsigned) cap;
    for (;;) {
      double ts[DDE1_WALK];
#pragma unroll
      for (int k = 0; k < DDE1_WALK; ++k) {
        DDE_CHECK(!(lo + 1 + k < count) || count - (lo + 1 + k) <= (unsigned) cap);
        ts[k] = lo + 1 + k < count ? (DDE_LAG_COUNT(5), ring[(size_t) slot * 10u + 8u]) : pos_inf();
        slot = slot + 1 == (unsigned) cap ? 0u : slot + 1;
      }
      unsigned adv = 0;
#pragma unroll
      for (int k = 0; k < DDE1_WALK; ++k) {
        adv += ts[k] <= tq ? 1u : 0u;
      }
      lo += adv;
      if (adv < DDE1_WALK) {
        break;
      }
    }
  } else {
    DDE_LAG_COUNT(1);
    if (used == 0 || !(DDE1_TIME(lo) <= tq)) {
      DDE_LAG_COUNT(4);
      f.failed = true;
      f.value = na_real();
      return f;
    }
    while (hi - lo > 1) {
      const unsigned mid = lo + (hi - lo) / 2;
      if (DDE1_TIME(mid) <= tq) {
        lo = mid;
      } else {
        hi = mid;
      }
    }
  }
#undef DDE1_TIME
  DDE_CHECK(lo < count && count - lo <= (unsigned) cap);
  const double *c = ring + (size_t) (lo % (unsigned) cap) * 10u;
  const double theta = (tq - c[8]) / c[9];
  const double theta1 = 1 - theta;
  f.value = interp853(1, theta, theta1, c);
  return f;
}

/* Once per step attempt, all lanes together: slide the window so that it starts at the
 * entry holding the attempt's lowest query and reaches up to its highest one (or is
 * full).  Returns whether the lowest query can be answered (from the pre-history or the
 * window's base entry). */
__device__ __forceinline__ bool dde1_window_ensure(Dde1Win &w, const double *ring, unsigned count,
                                                   int cap, double t0, double lo, double hi) {
  /* 1. drop the entries no query of this attempt reaches (an entry ends where the next
        one starts) */
#pragma unroll 1
  while (w.wn > 1 && dde1_lds(dde1_slot(w, 1) + 4u * DDE1_PAIR_BYTES) <= lo) {
    dde1_window_evict(w);
  }
  /* 2. extend upwards.  When any lane of the warp has to, every lane fills its free slots
        with the entries that follow (it will want them within a few steps) */
  if (__any_sync(__activemask(), hi >= w.wtop && w.wn < DDE1_WIN)) {
    unsigned idx = w.wbase + w.wn;
    if (w.wn > 0 && idx < count && w.wn < DDE1_WIN) {
      unsigned slot = idx % (unsigned) cap;
      const double *e = ring + (size_t) slot * 10u;
      unsigned dst = dde1_slot(w, w.wn);
      double te = 0.0, he = 0.0;
#pragma unroll 1
      do {
        DDE_LAG_COUNT(3);
        DDE_CHECK(slot < (unsigned) cap && idx < count && count - idx <= (unsigned) cap &&
                  dst >= w.first && dst < w.first + DDE1_WIN * DDE1_SLOT_BYTES);
        const double2 p0 = *(const double2 *) (e), p1 = *(const double2 *) (e + 2),
                      p2 = *(const double2 *) (e + 4), p3 = *(const double2 *) (e + 6),
                      p4 = *(const double2 *) (e + 8);
        dde1_sts2(dst, p0.x, p0.y);
        dde1_sts2(dst + DDE1_PAIR_BYTES, p1.x, p1.y);
        dde1_sts2(dst + 2u * DDE1_PAIR_BYTES, p2.x, p2.y);
        dde1_sts2(dst + 3u * DDE1_PAIR_BYTES, p3.x, p3.y);
        dde1_sts2(dst + 4u * DDE1_PAIR_BYTES, p4.x, p4.y);
        te = p4.x;
        he = p4.y;
        ++w.wn;
        ++idx;
        dst = dde1_wrap(w, dst + DDE1_SLOT_BYTES);
        e += 10;
        if (++slot == (unsigned) cap) {
          slot = 0;
          e = ring;
        }
      } while (idx < count && w.wn < DDE1_WIN);
      /* an entry ends where the next one starts, bit for bit (t += h, src/dopri.c:423) */
      w.wtop = idx < count ? te + he : pos_inf();
    }
  }
  return lo <= t0 || (w.wn > 0 && dde1_lds(dde1_slot(w, 0) + 4u * DDE1_PAIR_BYTES) <= lo);
}

/* After the integrating thread has written the head entry to its ring slot
 * (ring_buffer_head_advance, src/dopri.c:460): drop a window entry the overwrite policy
 * just recycled, append the new entry if the window ends at the head of the ring.
 * count: entries committed including the new one. */
__device__ __forceinline__ void dde1_window_commit(Dde1Win &w, unsigned count, int cap,
                                                   const double *hd, double t_old, double h) {
  const unsigned newest = count - 1;
  const unsigned tail = count > (unsigned) cap ? count - (unsigned) cap : 0u;
  if (w.wn > 0 && w.wbase < tail) {
    dde1_window_evict(w);
  }
  bool append = false;
  if (w.wn == 0) {
    w.wbase = newest;
    append = true;
  } else if (w.wbase + w.wn == newest) {
    if (w.wn < DDE1_WIN) {
      append = true;
    } else {
      w.wtop = t_old;
    }
  }
  if (append) {
    const unsigned dst = dde1_slot(w, w.wn);
    DDE_CHECK(w.wn < DDE1_WIN && dst >= w.first && dst < w.first + DDE1_WIN * DDE1_SLOT_BYTES);
    dde1_sts2(dst, hd[0], hd[1]);
    dde1_sts2(dst + DDE1_PAIR_BYTES, hd[2], hd[3]);
    dde1_sts2(dst + 2u * DDE1_PAIR_BYTES, hd[4], hd[5]);
    dde1_sts2(dst + 3u * DDE1_PAIR_BYTES, hd[6], hd[7]);
    dde1_sts2(dst + 4u * DDE1_PAIR_BYTES, t_old, h);
    ++w.wn;
    w.wtop = pos_inf();
  }
}

template <class M>
__global__ void __launch_bounds__(DDE_TPB, DDE1_MIN_BLOCKS) dopri_dde1_kernel(const DopriArgs a) {
  static_assert(M::N == 1 && M::N_LAGS == 1 && M::NP >= 0, "one equation, one declared lag");
  constexpr int ORDER = 8;
  constexpr int PMAX = M::NP > 0 ? M::NP : 1;
  constexpr int NST_STEP = 11, NST = 16;
  constexpr int hl = ORDER + 2;
  const int s = threadIdx.x;

  /* method constants, src/dopri.c:67-78 and :142-153 */
  constexpr double step_factor_min = 0.333, step_factor_max = 6.0, step_factor_safe = 0.9;
  const double step_constant = 0.125 - 0.0 * 0.2;
  const double atol = a.atol, rtol = a.rtol;
  const int n_times = a.n_times;
  const int nt = a.return_initial ? n_times : n_times - 1;
  const double t0 = a.times[0];
  const double t_end = a.times[n_times - 1];
  const int cap = a.n_history;

  /* shared memory: the block's copy of the pow / exp tables, then the window */
  unsigned smem = (unsigned) __cvta_generic_to_shared(dde_smem);
#ifndef DDE_HOST_EMU
  asm volatile("" : "+r"(smem)); /* opaque: formed once, not re-derived at every use */
#endif
  dde1_tables_fill(smem);
  __syncthreads();
  typename M::Scoped sc; /* the model's side of the solver (Dde1Scope) */
  sc.tab = smem;
  sc.undeclared = false;
  sc.memo_t = sc.memo_v = 0.0;
  Dde1Win w;
  w.first = smem + (unsigned) DDE_TAB_SHARED_BYTES + 16u * (unsigned) s;
#ifndef DDE_HOST_EMU
  asm volatile("" : "+r"(w.first)); /* ... nor this thread's column from the thread index */
#endif
  w.below = w.first;
  w.wbase = w.wn = 0;
  w.wtop = pos_inf();

  /* ---- per-lane solver object (src/dopri.h:55-161) ---- */
  long long traj = -1;
  bool active = false, done = false;
  bool error = false; /* obj->error */
  int code = 0;       /* obj->code */
  double p[PMAX];
  double y = 0.0, k1 = 0.0, y0v = 0.0;
  double t = 0.0, h = 0.0, fac_old = 1e-4, h_save = 0.0;
  bool stop = false, last = false, reject = false;
  int times_idx = 1;
  double t_out = 0.0;
  int n_eval0 = 0, n_step = 0, n_accept = 0, n_reject = 0;
  unsigned count = 0; /* ring entries committed; the head slot is count % cap */
  int head = 0;
  double *ring = nullptr;
  double *yo_next = nullptr;

  for (;;) {
    if (!active) {
      /* persistent threads: take the next trajectory, or leave */
      traj = (long long) atomicAdd(a.next_traj, 1ULL);
      if (traj >= a.n_traj) {
        break;
      }
      if (a.order != nullptr) {
        traj = (long long) a.order[traj];
      }
      active = true;
      done = false;
#pragma unroll
      for (int i = 0; i < PMAX; ++i) {
        if (i < M::NP) {
          p[i] = a.parms[traj * a.parms_stride + i];
        }
      }
      y = a.y0[traj];
      y0v = y; /* the pre-history, src/dopri.c:777-778 */
      /* dopri_data_reset, src/dopri.c:138-219 (time checks: once, on the host) */
      t = t0;
      times_idx = 1;
      yo_next = a.y_out + ((size_t) traj * (size_t) nt + (a.return_initial ? 1u : 0u));
      t_out = n_times > 1 ? a.times[1] : 0.0;
      n_step = n_accept = n_reject = 0;
      ring = a.ring + (size_t) traj * (size_t) cap * (size_t) hl;
      count = 0;
      head = 0;
      dde1_window_reset(w);
      sc.undeclared = false;
      code = 0;
      error = false;
      fac_old = 1e-4;
      stop = last = reject = false;
      h_save = 0.0;
      /* k1 = f(t0, y0) and the first step size come from dopri_init_kernel */
      k1 = a.init_k1[traj];
      h = a.init_h[traj];
      n_eval0 = a.stats[4 * traj + 0];
      const int code0 = a.code[traj];
      if (code0 == -8 || code0 == -10) {
        code = code0;
        done = true;
      } else if (code0 != 0) { /* a lag look-up failed during start-up */
        code = code0;
        error = true;
      }
    }

    /* ---- one trip of the adaptive loop, src/dopri.c:343-510 ---- */
    bool run = !done;
    bool force_this_step = false;
    if (run) {
      if ((unsigned long long) n_step > a.step_max_n) {
        error = true;
        code = -3; /* ERR_TOO_MANY_STEPS */
        done = true;
      } else if (fabs(h) <= a.step_size_min) {
        if (a.step_size_min_allow) {
          h = copysign(a.step_size_min, 1.0);
          force_this_step = true;
        } else {
          error = true;
          code = -4; /* ERR_STEP_SIZE_TOO_SMALL */
          done = true;
        }
      } else if (fabs(h) <= fabs(t) * DBL_EPSILON) {
        error = true;
        code = -5; /* ERR_STEP_SIZE_VANISHED */
        done = true;
      }
      run = !done;
    }
    /* ---- the evaluations of the step, in lock-step across the warp ----
       Position q of the loop is evaluation q for q <= 10 and evaluation q + 1 after that
       (evaluation 11, the redundant one, is counted, not computed: see the head).  `tt` is
       the time of the evaluation in hand, `tn` that of the one after it: its look-up is
       made now, next to this evaluation's pow. */
    double yin = 0.0, tt = t, tn = t;
    double k2 = 0.0, k3 = 0.0, k4 = 0.0, k5 = 0.0, k6 = 0.0, k7 = 0.0, k8 = 0.0, k9 = 0.0,
           k10 = 0.0, k14 = 0.0, k15 = 0.0;
    double h_new = 0.0;
    double zn = 0.0;  /* the answer to the next evaluation's look-up ... */
    bool okn = false; /* ... if it is one */
    bool lo_ok = true; /* the look-up of the redundant evaluation would succeed */
    double lo = 0.0;
    if (run) {
      /* no critical times: the next stop is the end (src/dopri.c:305-311) */
      if (t + 1.01 * h - t_end > 0.0) {
        h = t_end - t;
        stop = true;
      }
      if (t + 1.01 * h - t_end > 0.0) {
        h_save = h;
        h = t_end - t;
        last = true;
      }
      ++n_step;
      /* the window follows the declared lag: [lags(t), lags(t + h)] */
      double la, lb;
      M::lags(t, p, &la);
      M::lags(t + h, p, &lb);
      lo = la;
      lo_ok = dde1_window_ensure(w, ring, count, cap, t0, fmin(la, lb), fmax(la, lb));
      /* input of evaluation 0, src/dopri_853.c:178-181, and its look-up */
      yin = y + h * DP8_A21 * k1;
      tt = dp8_stage_time(0, t, h);
      tn = dp8_stage_time(1, t, h);
      double tq;
      M::lags(tt, p, &tq);
      zn = dde1_lookup(w, t0, y0v, tq, &okn);
    }
    int my_stages = run ? NST_STEP : 0;
#pragma unroll 1
    for (int st = 0; st < NST - 1; ++st) {
      if (st >= my_stages) {
        continue;
      }
      /* this evaluation's look-up was made one evaluation ago */
      double tq;
      M::lags(tt, p, &tq);
      double z = zn;
      if (__builtin_expect(!okn, 0)) {
        const Dde1Found f =
            dde1_lookup_ring(ring, count, cap, t0, y0v, tq, w.wbase + w.wn, w.wtop);
        z = f.value;
        if (f.failed && !error) {
          error = true;
          code = -6; /* ERR_YLAG_FAIL */
        }
      }
      /* -- dopri_eval, src/dopri.c:831-835, next to the look-up of the evaluation after it.
            Two copies of the model: evaluation 10 has nothing to look up ahead (12 asks what
            it asked itself) and the last one nothing at all, and a branch around the look-up
            would take it out of pow's basic block in the other thirteen. -- */
      double kt;
      sc.memo_t = tq;
      sc.memo_v = z;
      if (st == NST_STEP - 1 || st == NST - 2) {
        zn = z;
        okn = true;
        M::deriv_scoped(sc, (size_t) 1, tt, &yin, &kt, p);
      } else {
        double tqn;
        M::lags(tn, p, &tqn);
        zn = dde1_lookup(w, t0, y0v, tqn, &okn);
        M::deriv_scoped(sc, (size_t) 1, tt, &yin, &kt, p);
      }

      /* -- file the result and form the input of the next evaluation (one warp-uniform
            switch per evaluation), src/dopri_853.c:178-246,304,330-347 -- */
      bool attempt_done = false, step_done = false;
      switch (st) {
        case 0:
          k2 = kt;
          yin = y + h * (DP8_A31 * k1 + DP8_A32 * kt);
          tt = tn;
          tn = dp8_stage_time(2, t, h);
          break;
        case 1:
          k3 = kt;
          yin = y + h * (DP8_A41 * k1 + DP8_A43 * kt);
          tt = tn;
          tn = dp8_stage_time(3, t, h);
          break;
        case 2:
          k4 = kt;
          yin = y + h * (DP8_A51 * k1 + DP8_A53 * k3 + DP8_A54 * kt);
          tt = tn;
          tn = dp8_stage_time(4, t, h);
          break;
        case 3:
          k5 = kt;
          yin = y + h * (DP8_A61 * k1 + DP8_A64 * k4 + DP8_A65 * kt);
          tt = tn;
          tn = dp8_stage_time(5, t, h);
          break;
        case 4:
          k6 = kt;
          yin = y + h * (DP8_A71 * k1 + DP8_A74 * k4 + DP8_A75 * k5 + DP8_A76 * kt);
          tt = tn;
          tn = dp8_stage_time(6, t, h);
          break;
        case 5:
          k7 = kt;
          yin = y + h * (DP8_A81 * k1 + DP8_A84 * k4 + DP8_A85 * k5 + DP8_A86 * k6 + DP8_A87 * kt);
          tt = tn;
          tn = dp8_stage_time(7, t, h);
          break;
        case 6:
          k8 = kt;
          yin = y + h * (DP8_A91 * k1 + DP8_A94 * k4 + DP8_A95 * k5 + DP8_A96 * k6 + DP8_A97 * k7 +
                         DP8_A98 * kt);
          tt = tn;
          tn = dp8_stage_time(8, t, h);
          break;
        case 7:
          k9 = kt;
          yin = y + h * (DP8_A101 * k1 + DP8_A104 * k4 + DP8_A105 * k5 + DP8_A106 * k6 +
                         DP8_A107 * k7 + DP8_A108 * k8 + DP8_A109 * kt);
          tt = tn;
          tn = dp8_stage_time(9, t, h);
          break;
        case 8:
          k10 = kt;
          yin = y + h * (DP8_A111 * k1 + DP8_A114 * k4 + DP8_A115 * k5 + DP8_A116 * k6 +
                         DP8_A117 * k7 + DP8_A118 * k8 + DP8_A119 * k9 + DP8_A1110 * kt);
          tt = tn;
          tn = dp8_stage_time(10, t, h);
          break;
        case 9: /* k2 is free again: stage 11's value goes there */
          k2 = kt;
          yin = y + h * (DP8_A121 * k1 + DP8_A124 * k4 + DP8_A125 * k5 + DP8_A126 * k6 +
                         DP8_A127 * k7 + DP8_A128 * k8 + DP8_A129 * k9 + DP8_A1210 * k10 +
                         DP8_A1211 * kt);
          tt = tn;
          tn = dp8_stage_time(12, t, h); /* the same time: evaluation 12 asks what 10 asked */
          break;
        case 10: /* src/dopri_853.c:242-246 */
          k3 = kt;
          k4 = DP8_B1 * k1 + DP8_B6 * k6 + DP8_B7 * k7 + DP8_B8 * k8 + DP8_B9 * k9 + DP8_B10 * k10 +
               DP8_B11 * k2 + DP8_B12 * kt;
          k5 = y + h * k4;
          attempt_done = true;
          break;
        case 11: /* evaluation 12: k4 = f(t + h, k5), src/dopri_853.c:304 */
          k4 = kt;
          yin = y + h * (DP8_A141 * k1 + DP8_A147 * k7 + DP8_A148 * k8 + DP8_A149 * k9 +
                         DP8_A1410 * k10 + DP8_A1411 * k2 + DP8_A1412 * k3 + DP8_A1413 * kt);
          tt = tn;
          tn = dp8_stage_time(14, t, h);
          break;
        case 12: /* evaluation 13 */
          k14 = kt;
          yin = y + h * (DP8_A151 * k1 + DP8_A156 * k6 + DP8_A157 * k7 + DP8_A158 * k8 +
                         DP8_A1511 * k2 + DP8_A1512 * k3 + DP8_A1513 * k4 + DP8_A1514 * kt);
          tt = tn;
          tn = dp8_stage_time(15, t, h);
          break;
        case 13: /* evaluation 14 */
          k15 = kt;
          yin = y + h * (DP8_A161 * k1 + DP8_A166 * k6 + DP8_A167 * k7 + DP8_A168 * k8 +
                         DP8_A169 * k9 + DP8_A1613 * k4 + DP8_A1614 * k14 + DP8_A1615 * kt);
          tt = tn;
          break;
        default: /* 14: evaluation 15, k16 = kt */
          step_done = true;
          break;
      }

      if (attempt_done) {
        if (__builtin_expect(sc.undeclared && !error, 0)) {
          error = true;
          code = -11; /* DDE_ERR_LAG_UNDECLARED */
        }
        if (__builtin_expect(error, 0)) {
          done = true; /* a lag look-up failed during dopri_step, src/dopri.c:387-389 */
          my_stages = 0;
          continue;
        }
        /* dopri853_error, src/dopri_853.c:249-283: scaled by sign (+1 here), not |h|,
           guarded by deno < 0; k3 is evaluation 10's value, kt */
        const double sk = atol + rtol * fmax(fabs(y), fabs(k5));
        double erri = k4 - DP8_BHH1 * k1 - DP8_BHH2 * k9 - DP8_BHH3 * kt;
        const double err2 = 0.0 + sq(erri / sk);
        erri = DP8_ER1 * k1 + DP8_ER6 * k6 + DP8_ER7 * k7 + DP8_ER8 * k8 + DP8_ER9 * k9 +
               DP8_ER10 * k10 + DP8_ER11 * k2 + DP8_ER12 * kt;
        double err = 0.0 + sq(erri / sk);
        double deno = err + 0.01 * err2;
        deno = deno < 0 ? 1.0 : deno;
        err = 1.0 * err * sqrt(1.0 / ((double) 1 * deno));
        /* dopri_h_new, src/dopri.c:516-525 (step_beta = 0: pow(fac_old, 0) == 1) */
        const double fac11 = dde1_pow_ctrl(smem, err, step_constant);
        double fac = fac11 / 1.0;
        fac = fmax(1.0 / step_factor_max, fmin(1.0 / step_factor_min, fac / step_factor_safe));
        const bool accepted = (err <= 1) || force_this_step;
        /* one division for both outcomes: a rejected lane divides by its own factor
           (src/dopri.c:495-499) */
        const double fac_rej = fmin(1 / step_factor_min, fac11 / step_factor_safe);
        h_new = h / (accepted ? fac : fac_rej);
        if (accepted) {
          my_stages = NST - 1;
          fac_old = fmax(err, 1e-4); /* src/dopri.c:398-399 */
          ++n_accept;
          /* the redundant evaluation at the OLD time, src/dopri.c:400-403 (SURVEY.md
             A.1#2): it counts (n_eval below), and its look-up can fail */
          if (__builtin_expect(!lo_ok, 0)) {
            const Dde1Found f =
                dde1_lookup_ring(ring, count, cap, t0, y0v, lo, w.wbase + w.wn, w.wtop);
            if (f.failed && !error) {
              error = true;
              code = -6; /* ERR_YLAG_FAIL */
            }
          }
          /* input of evaluation 12, src/dopri_853.c:304 */
          yin = k5;
          tt = tn;
          tn = dp8_stage_time(13, t, h);
        } else {
          /* rejected, src/dopri.c:495-509 */
          reject = true;
          if (n_accept >= 1) {
            ++n_reject;
          }
          last = false;
          stop = false;
          h = h_new;
        }
      }

      if (step_done) {
        /* ---- the rest of an accepted step, src/dopri.c:410-494 ---- */
        /* dopri853_save_history, src/dopri_853.c:306-327,350-366 */
        double hd[ORDER];
        {
          const double ydiff = k5 - y;
          const double bspl = h * k1 - ydiff;
          hd[0] = y;
          hd[1] = ydiff;
          hd[2] = bspl;
          hd[3] = ydiff - h * k4 - bspl;
          const double k11 = k2, k12 = k3;
          hd[4] = h * (DP8_D41 * k1 + DP8_D46 * k6 + DP8_D47 * k7 + DP8_D48 * k8 + DP8_D49 * k9 +
                       DP8_D410 * k10 + DP8_D411 * k11 + DP8_D412 * k12 + DP8_D413 * k4 +
                       DP8_D414 * k14 + DP8_D415 * k15 + DP8_D416 * kt);
          hd[5] = h * (DP8_D51 * k1 + DP8_D56 * k6 + DP8_D57 * k7 + DP8_D58 * k8 + DP8_D59 * k9 +
                       DP8_D510 * k10 + DP8_D511 * k11 + DP8_D512 * k12 + DP8_D513 * k4 +
                       DP8_D514 * k14 + DP8_D515 * k15 + DP8_D516 * kt);
          hd[6] = h * (DP8_D61 * k1 + DP8_D66 * k6 + DP8_D67 * k7 + DP8_D68 * k8 + DP8_D69 * k9 +
                       DP8_D610 * k10 + DP8_D611 * k11 + DP8_D612 * k12 + DP8_D613 * k4 +
                       DP8_D614 * k14 + DP8_D615 * k15 + DP8_D616 * kt);
          hd[7] = h * (DP8_D71 * k1 + DP8_D76 * k6 + DP8_D77 * k7 + DP8_D78 * k8 + DP8_D79 * k9 +
                       DP8_D710 * k10 + DP8_D711 * k11 + DP8_D712 * k12 + DP8_D713 * k4 +
                       DP8_D714 * k14 + DP8_D715 * k15 + DP8_D716 * kt);
        }
        k1 = k4; /* src/dopri.c:420-421 */
        y = k5;
        const double t_old = t;
        t += h;

        /* output rows, src/dopri.c:437-456: interpolate from the head */
        for (;;) {
          if (times_idx >= n_times) {
            break;
          }
          const double tr = t_out;
          if (!(last || tr <= t)) {
            break;
          }
          const double t_next = times_idx + 1 < n_times ? a.times[times_idx + 1] : 0.0;
          const double theta = (tr - t_old) / h;
          const double theta1 = 1 - theta;
          *yo_next++ = interp853(1, theta, theta1, hd);
          ++times_idx;
          t_out = t_next;
        }

        /* ring_buffer_head_advance, src/dopri.c:460 */
        {
          DDE_CHECK(head >= 0 && head < cap && (unsigned) head == count % (unsigned) cap);
          double *dst = ring + (size_t) head * (size_t) hl;
          *(double2 *) (dst) = make_double2(hd[0], hd[1]);
          *(double2 *) (dst + 2) = make_double2(hd[2], hd[3]);
          *(double2 *) (dst + 4) = make_double2(hd[4], hd[5]);
          *(double2 *) (dst + 6) = make_double2(hd[6], hd[7]);
          *(double2 *) (dst + 8) = make_double2(t_old, h);
          head = head + 1 == cap ? 0 : head + 1;
          ++count;
          /* make room first: the entries no query of the NEXT attempt reaches (it starts at
             the new t, so its lowest query is lags(t)) go now, not at that attempt's start,
             so that the new entry still finds a slot and the window stays at the head of the
             ring -- appended to from registers, never re-read from HBM */
          {
            double lo_next;
            M::lags(t, p, &lo_next);
#pragma unroll 1
            while (w.wn > 1 && dde1_lds(dde1_slot(w, 1) + 4u * DDE1_PAIR_BYTES) <= lo_next) {
              dde1_window_evict(w);
            }
          }
          dde1_window_commit(w, count, cap, hd, t_old, h);
        }

        if (last) {
          code = 1; /* OK_COMPLETE */
          done = true;
          my_stages = 0;
          continue;
        }
        if (fabs(h_new) >= a.step_size_max) {
          h_new = copysign(a.step_size_max, 1.0);
        }
        if (reject) {
          h_new = copysign(fmin(fabs(h_new), fabs(h)), 1.0);
          reject = false;
        }
        if (stop) { /* src/dopri.c:476-491: no critical times to pass */
          stop = false;
        } else {
          h = h_new;
        }
      }
    }

    if (done) {
      /* retire: statistics as r_dopri_cleanup reads them, src/r_dopri.c:412-428; every
         attempt made its 11 evaluations and every accepted step 5 more */
      a.stats[4 * traj + 0] = n_eval0 + NST_STEP * n_step + (NST - NST_STEP) * n_accept;
      a.stats[4 * traj + 1] = n_step;
      a.stats[4 * traj + 2] = n_accept;
      a.stats[4 * traj + 3] = n_reject;
      a.code[traj] = code;
      a.step_size[traj] = code == 1 ? h_save /* src/dopri.c:463 */ : a.step_size_initial;
      if (code != -10) {
        a.t_last[traj] = t;
        if (a.y_final != nullptr) {
          a.y_final[traj] = y;
        }
      }
      a.ring_count[traj] = count;
      chunk_retire(a, traj);
      active = false;
    }
  }
}

} /* namespace dde */

#endif /* DDE_DOPRI_DDE1_KERNEL_CUH */
