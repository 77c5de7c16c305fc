/* difeq_kernels.cuh -- batched difference-equation iteration, one replicate
 * per thread.
 *
 * Device restatement of difeq_run (/root/reference/src/difeq.c:127-258) and
 * of the replicate loop around it (src/r_difeq.c:94-103) for n_rep
 * independent replicates at once.  Each replicate has FRESH history state:
 * the documented contract of difeq_replicate (R/difeq_replicate.R:1-4), not
 * the reference's one-ring-for-all-replicates artefact (SURVEY.md finding 5).
 *
 * What is kept from the reference:
 *   - the initial state is committed to the ring as the entry for steps[0]
 *     before the first target() call (src/difeq.c:140-156), so yprev(step)
 *     at offset 0 is the current state;
 *   - yprev(step <= step0) is y0, any other miss is an error
 *     (src/difeq.c:261-281) that ends the replicate (Rf_error there);
 *   - rows are stored when step == steps[idx] (src/difeq.c:215-225);
 *   - outputs computed by target(step = s, y(s)) belong to row s, hence the
 *     one-iteration delay and the trailing target() call
 *     (src/difeq.c:204-213,249-255).
 * Ring entries are {step as double, y[n], out[n_out]} (src/difeq.c:31-37);
 * the out block is NA-filled as src/difeq.c:146 does -- the reference's
 * later writes into stale out slots are not observable through yprev or
 * through the replicate results and are not reproduced.
 *
 * The per-step arithmetic is whatever the model does in FP64, compiled
 * without contraction, so deterministic models are bit-exact.
 */
#ifndef DDE_DIFEQ_KERNELS_CUH
#define DDE_DIFEQ_KERNELS_CUH

#include "dopri_kernels.cuh"

namespace dde {

struct DifeqArgs {
  long long n_rep;
  int n, n_out, n_parms;
  long long parms_stride;
  const double *y0;                 /* [n x n_rep] */
  const double *parms;
  const unsigned long long *steps;  /* [n_steps], strictly increasing */
  int n_steps;
  int n_history, return_initial;
  double *y_out;   /* [n x nt x n_rep] */
  double *out;     /* [n_out x nt x n_rep] */
  int *code;       /* [n_rep] */
  double *y_final; /* [n x n_rep], obj->y1 (src/difeq.c:244) */
  double *ring;    /* [n_rep][n_history][1 + n + n_out] */
  int staged;      /* history in (dynamic) shared memory instead */
  /* restart (difeq_continue, src/r_difeq.c:114-165): ring, ring_count and code
     of the previous run are inputs; step_origin is the step the handle's very
     first run started at (entry k of a ring holds step_origin + k) */
  int restart;
  int restartable;      /* leave what a later restart needs in HBM */
  unsigned *ring_count; /* [n_rep] entries committed */
  unsigned long long step_origin;
  unsigned long long *next_rep;
};

/* What a difference equation sees instead of the global yprev_* functions when its source
 * is compiled a second time INSIDE a struct derived from this one (dde_model_scoped.inc with
 * DDE_SCOPED_BASE ::dde::DifeqScope): the replicate's history state -- the current step, the
 * first one, where the ring's head is, how many entries it holds -- as members, i.e. in
 * registers of the thread that iterates the replicate.  The model ABI carries no solver
 * handle (src/difeq.h:19-21; the reference keeps a process global, src/difeq.c:12), so the
 * global functions keep that state in a per-thread block of shared memory and every step
 * writes and re-reads it (lag_ctx.cuh: sd_step, sd_head, sd_count, sd_error ...: 58
 * instructions per step of the delayed SIR map for 8 flops of model); with the state in
 * registers the compiler sees through the look-up.  Same semantics, src/difeq.c:261-281. */
struct DifeqScope {
  long long step_now; /* the step target() is being called for */
  long long step0;    /* the first step (on a restart: the oldest surviving entry's) */
  const double *y0;
  const double *ring; /* the replicate's ring in HBM (history not on chip) */
  unsigned hist_col;  /* history on chip: shared-window address of this thread's column */
  unsigned count;     /* entries committed */
  int head;           /* slot of the next commit */
  int cap, n, hl;
  bool staged, error;
  int code;
  __device__ __forceinline__ bool find(int step, int *slot) {
    const int offset = (int) step_now - step;
    const unsigned used = count < (unsigned) cap ? count : (unsigned) cap;
    /* cap == 0 || offset < 0 || offset >= used, as one comparison: no capacity is nothing
       used, a negative offset a huge unsigned one */
    if ((unsigned) offset >= used) {
      error = true;
      code = -9; /* DDE_ERR_DIFEQ_HISTORY */
      return false;
    }
    *slot = ring_slot_back(head, offset + 1, cap);
    return true;
  }
  __device__ __forceinline__ double read(int slot, size_t comp) const {
    if (staged) {
      return lds_f64(hist_col + (unsigned) ((slot * n + (int) comp) * (DDE_TPB * 8)));
    }
    return ring[(size_t) slot * (size_t) hl + 1 + comp];
  }
  __device__ __forceinline__ double yprev_1(int step, size_t i) {
    if (step <= (int) step0) {
      return y0[i];
    }
    int slot;
    return find(step, &slot) ? read(slot, i) : na_real();
  }
  __device__ __forceinline__ void yprev_all(int step, double *y) {
    if (step <= (int) step0) {
      for (int i = 0; i < n; ++i) {
        y[i] = y0[i];
      }
      return;
    }
    int slot;
    if (find(step, &slot)) {
      for (int i = 0; i < n; ++i) {
        y[i] = read(slot, (size_t) i);
      }
    }
  }
  template <class Index>
  __device__ __forceinline__ void yprev_indexed(int step, const Index *idx, size_t nidx,
                                                double *y) {
    if (step <= (int) step0) {
      for (size_t i = 0; i < nidx; ++i) {
        y[i] = y0[idx[i]];
      }
      return;
    }
    int slot;
    if (find(step, &slot)) {
      for (size_t i = 0; i < nidx; ++i) {
        y[i] = read(slot, (size_t) idx[i]);
      }
    }
  }
  __device__ __forceinline__ void yprev_vec(int step, const size_t *idx, size_t nidx, double *y) {
    yprev_indexed<size_t>(step, idx, nidx, y);
  }
  __device__ __forceinline__ void yprev_vec_int(int step, const int *idx, size_t nidx, double *y) {
    yprev_indexed<int>(step, idx, nidx, y);
  }
  __device__ __forceinline__ double dde_scope_pow(double x, double y) { return dde_pow(x, y); }
  __device__ __forceinline__ double dde_scope_exp(double x) { return dde_exp(x); }
};

/* RESTART: compiled with the difeq_continue paths (start from / leave a history
 * in HBM); the plain instantiation is ~10 % faster on C4 without them. */
/* STAGED: the history lives in shared memory (a.staged, decided at launch); a
 * template argument so that yprev_* see it in the per-thread cfg word at compile time. */
template <class M, bool RESTART, bool STAGED>
__global__ void __launch_bounds__(DDE_TPB) difeq_tpt_kernel(const DifeqArgs a) {
  constexpr int NC = M::N;
  constexpr int NMAX = NC > 0 ? NC : DDE_DYN_MAXN;
  constexpr int PMAX = M::NP > 0 ? M::NP : (M::NP == 0 ? 1 : DDE_DYN_MAXP);
  constexpr int OMAX = M::N_OUT > 0 ? M::N_OUT : (M::N_OUT == 0 ? 1 : DDE_DYN_MAXOUT);
  const int n = NC > 0 ? NC : a.n;
  const int n_parms = M::NP >= 0 ? M::NP : a.n_parms;
  const int n_out = M::N_OUT == 0 ? 0 : a.n_out; /* 0 = output function not requested */
  const int hl = 1 + n + n_out;
  const int s = threadIdx.x;
  const int cap = a.n_history;

  if (threadIdx.x == 0) {
    s_lag_u.sign = 1.0;
    s_lag_u.cap = cap;
    s_lag_u.n = n;
    s_lag_u.hl = hl;
    s_lag_u.order = 0;
    s_lag_u.staged = STAGED ? 1 : 0;
  }
  __syncthreads();
  /* what yprev_* read: n, "history on chip" and, then, its capacity (< 128: it fits
     the shared-memory budget) */
  const int cfg = n | ((STAGED ? 1 : 0) << 24) | (STAGED ? (int) ((unsigned) cap << 25) : 0);
  /* this thread's column of the on-chip history as a shared-window address; opaque
     to the optimiser so that it is formed once, not from the special registers in
     every step */
  unsigned hist_col = (unsigned) __cvta_generic_to_shared(sd_hist + s);
  asm volatile("" : "+r"(hist_col));

  const int n_steps = a.n_steps;
  const int nt = a.return_initial ? n_steps : n_steps - 1;
  const unsigned long long step_first = a.steps[0];
  const bool has_output = n_out > 0;

  for (;;) {
    const long long rep = (long long) atomicAdd(a.next_rep, 1ULL);
    if (rep >= a.n_rep) {
      break;
    }
    if (RESTART && a.restart && a.code[rep] != 1) {
      /* a replicate whose last run failed has no state to continue from
         ("Incorrect initial step on simulation restart", src/r_difeq.c:139-141) */
      a.code[rep] = -10;
      continue;
    }
    double p[PMAX];
#pragma unroll
    for (int i = 0; i < PMAX; ++i) {
      if (i < n_parms) {
        p[i] = a.parms[rep * a.parms_stride + i];
      }
    }
    /* rows of one array: see the note on the model's dydt in dopri_init_kernel
       (dopri_kernels.cuh) -- here a local array of the inlined target against y_next / out */
    struct {
      double y[NMAX], y_next[NMAX], o[OMAX];
    } loc;
    double *const y = loc.y, *const y_next = loc.y_next, *const o = loc.o;
    const double *y0 = a.y0 + rep * n;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      y[i] = y0[i];
    }
    /* difeq_data_reset, src/difeq.c:74-105 (step checks run on the host) */
    unsigned long long step = step_first;
    int steps_idx = 1;
    double *ring = a.ring + (size_t) rep * (size_t) cap * (size_t) hl;
    /* the history state the model's yprev_* read: members of `sc` (registers) for a model
       compiled in scope, the per-thread block in shared memory otherwise */
    constexpr bool SCOPED = M::SCOPED;
    typename M::Scoped sc;
    sc.step_now = (long long) step;
    sc.step0 = (long long) step_first;
    sc.y0 = y0;
    sc.ring = ring;
    sc.hist_col = hist_col;
    sc.count = 0;
    sc.head = 0;
    sc.cap = cap;
    sc.n = n;
    sc.hl = hl;
    sc.staged = STAGED;
    sc.error = false;
    sc.code = 0;
    if (!SCOPED) {
      sd_ring[s] = ring;
      sd_y0[s] = y0;
      sd_step0[s] = (long long) step_first;
      sd_step[s] = (long long) step;
      sd_code[s] = 0;
      sd_error[s] = 0;
      sd_cfg[s] = cfg;
    }

    int head = 0;
    unsigned count = 0;
    if (RESTART && a.restart && cap > 0) {
      /* the history the last run left; step0 is its oldest surviving entry
         (difeq_data_reset, src/difeq.c:96-102) */
      count = a.ring_count[rep];
      head = (int) (count % (unsigned) cap);
      const unsigned used = count < (unsigned) cap ? count : (unsigned) cap;
      if (count > 0) {
        sc.step0 = (long long) (a.step_origin + (count - used));
        if (!SCOPED) {
          sd_step0[s] = sc.step0;
        }
      }
      if (STAGED) { /* back on chip */
        for (unsigned k = 0; k < used; ++k) {
          const double *src = ring + (size_t) k * (size_t) hl + 1;
          double *dst = sd_hist + (size_t) k * n * DDE_TPB + s;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            dst[(size_t) i * DDE_TPB] = src[i];
          }
        }
      }
    }
    sc.count = count;
    sc.head = head;
    if (!SCOPED) {
      sd_count[s] = count;
      sd_head[s] = head;
    }
#define DDE_DIFEQ_COMMIT(yy)                                  \
  do {                                                        \
    if (cap > 0) {                                            \
      if (STAGED) {                                           \
        /* on-chip history: only y is ever read back */       \
        const unsigned dst = hist_col + (unsigned) (head * n * DDE_TPB * 8); \
        _Pragma("unroll") for (int i = 0; i < n; ++i) {       \
          asm volatile("st.shared.f64 [%0], %1;" ::"r"(dst + (unsigned) (i * DDE_TPB * 8)), \
                       "d"((yy)[i]) : "memory");              \
        }                                                     \
      } else {                                                \
        double *dst = ring + (size_t) head * (size_t) hl;     \
        dst[0] = (double) step; /* src/difeq.c:327-330 */     \
        _Pragma("unroll") for (int i = 0; i < n; ++i) {       \
          dst[1 + i] = (yy)[i];                               \
        }                                                     \
        for (int i = 0; i < n_out; ++i) {                     \
          dst[1 + n + i] = na_real();                         \
        }                                                     \
      }                                                       \
      head = head + 1 == cap ? 0 : head + 1;                  \
      ++count;                                                \
      sc.head = head;                                         \
      sc.count = count;                                       \
      if (!SCOPED) {                                          \
        sd_head[s] = head;                                    \
        sd_count[s] = count;                                  \
      }                                                       \
    }                                                         \
  } while (0)
#define DDE_DIFEQ_TARGET()                                                           \
  do {                                                                               \
    sc.step_now = (long long) step;                                                  \
    if (SCOPED) {                                                                    \
      M::target_scoped(sc, (size_t) n, (size_t) step, y, y_next, (size_t) n_out, o, p); \
    } else {                                                                         \
      sd_step[s] = (long long) step;                                                 \
      if (NC > 0) { /* what yprev_* will read: lets the optimiser drop the other path */ \
        __builtin_assume(sd_cfg[s] == cfg);                                          \
      }                                                                              \
      M::target((size_t) n, (size_t) step, y, y_next, (size_t) n_out, o, p);          \
    }                                                                                \
  } while (0)
#define DDE_DIFEQ_ERRORED() (SCOPED ? sc.error : sd_error[s] != 0)
#define DDE_DIFEQ_CODE() (SCOPED ? sc.code : sd_code[s])

    /* the initial state is the first entry, src/difeq.c:140-156; on a restart
       the ring already ends with the state the last run stopped at, and the
       reference does not commit again (first_entry is false) */
    if (!(RESTART && a.restart && count > 0)) {
      DDE_DIFEQ_COMMIT(y);
    }

    double *wy = a.y_out + (size_t) rep * (size_t) nt * (size_t) n;
    double *wo = has_output ? a.out + (size_t) rep * (size_t) nt * (size_t) n_out : nullptr;
    bool store_next_output = false;
    if (a.return_initial) {
#pragma unroll
      for (int i = 0; i < n; ++i) {
        wy[i] = y[i];
      }
      wy += n;
      store_next_output = true;
    }

    bool failed = false;
    /* the next step a row is wanted at: read when the row index moves, not per step */
    unsigned long long step_out = steps_idx < n_steps ? a.steps[steps_idx] : ~0ULL;
    for (;;) {
      DDE_DIFEQ_TARGET();
      if (DDE_DIFEQ_ERRORED()) {
        failed = true; /* Rf_error in difeq_find_step, src/difeq.c:267-270 */
        break;
      }
      ++step;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        y[i] = y_next[i];
      }
      if (has_output && store_next_output) {
        /* `o` was computed from the previous y, whose row was just stored */
#pragma unroll
        for (int i = 0; i < OMAX; ++i) {
          if (i < n_out) {
            wo[i] = o[i];
          }
        }
        wo += n_out;
        store_next_output = false;
      }
      bool at_end = false; /* the last requested step is the last row (one test per step) */
      if (step == step_out) {
#pragma unroll
        for (int i = 0; i < n; ++i) {
          wy[i] = y[i];
        }
        wy += n;
        store_next_output = true;
        ++steps_idx;
        at_end = steps_idx >= n_steps;
        step_out = at_end ? ~0ULL : a.steps[steps_idx];
      }
      DDE_DIFEQ_COMMIT(y);
      if (at_end) {
        if (a.y_final != nullptr) {
#pragma unroll
          for (int i = 0; i < n; ++i) {
            a.y_final[rep * n + i] = y[i];
          }
        }
        break;
      }
    }
    if (!failed && has_output && store_next_output) {
      /* trailing call for the outputs of the last stored row,
         src/difeq.c:249-255 */
      DDE_DIFEQ_TARGET();
      if (DDE_DIFEQ_ERRORED()) {
        failed = true;
      } else {
#pragma unroll
        for (int i = 0; i < OMAX; ++i) {
          if (i < n_out) {
            wo[i] = o[i];
          }
        }
      }
    }
    a.code[rep] = failed ? DDE_DIFEQ_CODE() : 1; /* OK_COMPLETE */
#undef DDE_DIFEQ_COMMIT
#undef DDE_DIFEQ_TARGET
#undef DDE_DIFEQ_ERRORED
#undef DDE_DIFEQ_CODE
    if (RESTART && cap > 0 && a.restartable) {
      /* what a later difeq_continue needs: the count, and an on-chip history
         written back to the replicate's ring (slot order is the same) */
      a.ring_count[rep] = count;
      if (STAGED) {
        const unsigned used = count < (unsigned) cap ? count : (unsigned) cap;
        for (unsigned k = 0; k < used; ++k) {
          double *dst = ring + (size_t) k * (size_t) hl + 1;
          const double *src = sd_hist + (size_t) k * n * DDE_TPB + s;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            dst[i] = src[(size_t) i * DDE_TPB];
          }
        }
      }
    }
  }
}

} /* namespace dde */

#endif /* DDE_DIFEQ_KERNELS_CUH */
