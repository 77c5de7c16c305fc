/* dopri_kernels.cuh -- batched Dormand-Prince integrator, one trajectory per
 * thread, state in registers.
 *
 * This is the device restatement of the reference's solver core for n_traj
 * independent solver objects at once:
 *   adaptive loop            /root/reference/src/dopri.c:291-514
 *   reset                    src/dopri.c:138-219
 *   controller / h_init      src/dopri.c:516-572
 *   dopri5 step/error/dense  src/dopri_5.c:42-127
 *   dop853 step/error/dense  src/dopri_853.c:159-380
 *   stiffness test           src/dopri.c:668-693, dopri_5.c:129-140, dopri_853.c:382-394
 * Every floating-point expression keeps the reference's operand order and is
 * compiled without FMA contraction (-fmad=false), and pow() is dde_pow, so a
 * trajectory takes bit-for-bit the same steps as the reference C solver
 * (finding 1).  The reference's documented departures
 * from Hairer's Fortran are reproduced on purpose.
 *
 * Design (B200): the whole solver object of src/dopri.h:55-161 lives in the
 * thread's registers -- y, k1..k7 (dopri5) or k1..k10 (dop853) and the dense
 * coefficients of the in-progress step -- so for small n the only HBM
 * traffic is: y0/parms in, output rows out, ring commits, lag look-ups.
 * The model's deriv() is inlined into the kernel (one translation unit,
 * template on the model), which is what keeps y/dydt in registers despite
 * the pointer-based model ABI.
 *
 * Control flow.  Each lane carries its own trajectory, t and h, but a warp
 * walks the stages of a step attempt in lock-step: the attempt is a loop over
 * RHS evaluations ("stages") with a warp-uniform stage counter, a switch that
 * forms the stage input, ONE inlined instance of the model's deriv(), and a
 * switch that files the result.  That keeps the code small enough for the
 * instruction caches (18 inlined copies of deriv + pow + lag look-up were
 * 160 KB and stalled on instruction fetch, profiles/r1a) and makes the
 * expensive part -- deriv, pow, the ring look-up -- execute convergently.
 * Lanes that rejected the step idle through the accept-only stages (dop853's
 * 5 extra evaluations); lanes whose trajectory has ended are refilled with
 * the next one from an atomic counter (persistent threads), so a warp's cost
 * follows the mean, not the max, of its lanes' step counts.
 */
#ifndef DDE_DOPRI_KERNELS_CUH
#define DDE_DOPRI_KERNELS_CUH

#include <float.h>

#include "dopri_args.h"
#include "lag_ctx.cuh"
#include "tableau.h"

namespace dde {

#ifndef DDE_MIN_BLOCKS
/* Resident blocks per SM the register allocation of a scalar model (n = 1)
   must allow: 3 x 128 threads at <= 168 registers; the staged window
   (DDE_WIN 6) is sized to the shared memory that leaves per block.  Larger
   compile-time n keep their whole state in registers instead (1 block). */
#define DDE_MIN_BLOCKS 3
#endif
#ifndef DDE_MIN_BLOCKS_SMALL
#define DDE_MIN_BLOCKS_SMALL 2 /* the same for compile-time n = 2 .. 4 */
#endif
#ifndef DDE_MIN_BLOCKS_UNROLLED
/* ... and for the unrolled dopri5 kernel of a model with n = 2 .. 4 (no history: its shared
   memory is the context and the staged output rows); 3 blocks: 168 registers, a
   48-byte spill (measured on the Lorenz sweep: 1000 rows 54.3 -> 49.7 ms per 1 M
   trajectories, 10 rows 37.2 -> 25.1 ms) */
#define DDE_MIN_BLOCKS_UNROLLED 3
#endif
#ifndef DDE_ROW_GROUP
#define DDE_ROW_GROUP 2 /* output rows the unrolled kernel forms per trip of its row loop */
#endif
#ifndef DDE_STAGE_ROW_DOUBLES
/* doubles of output rows a thread of the unrolled kernel holds back until they fill an
   aligned 32-byte sector of the result array (lcm(n, 4) for n <= 4; pitch 13) */
#define DDE_STAGE_ROW_DOUBLES 12
#endif
__device__ __forceinline__ double sq(double x) { return x * x; }

/* shared memory by shared-window address */
#ifdef DDE_HOST_EMU /* tests/emu: an "address" is a byte offset into dde_smem */
static inline size_t __cvta_generic_to_shared(const void *p) {
  return (size_t) ((const char *) p - (const char *) dde_smem);
}
__device__ __forceinline__ double lds_f64(unsigned addr) { return dde_smem[addr / 8]; }
__device__ __forceinline__ void sts_f64(unsigned addr, double v) { dde_smem[addr / 8] = v; }
#else
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
#endif

/* four doubles to one aligned 32-byte sector of global memory: one 256-bit store */
__device__ __forceinline__ void stg_sector(double *dst, double a, double b, double c, double d) {
#ifdef DDE_HOST_EMU
  dst[0] = a;
  dst[1] = b;
  dst[2] = c;
  dst[3] = d;
#else
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(a), "d"(b), "d"(c), "d"(d)
               : "memory");
#endif
}

/* the step-size controller's pow (once per attempt) is out of line, so that the hot
   loop holds one copy of pow, the model's (measured: +1.2 %) */
#ifndef DDE_CTRL_POW_INLINE
static __device__ __noinline__ double dde_pow_ctrl(double x, double y) { return dde_pow(x, y); }
static __device__ __noinline__ DdePow2 dde_pow2_ctrl(double x, double y1, double y2) {
  return dde_pow2(x, y1, y2);
}
#else
#define dde_pow_ctrl dde_pow
#define dde_pow2_ctrl dde_pow2
#endif

/* shared memory by shared-window address: 64- and 128-bit accesses.  volatile, so the
   compiler keeps them in program order (the window is written and read by the same
   thread in different phases of a step); ptxas schedules them like any other access */
#ifdef DDE_HOST_EMU
__device__ __forceinline__ double dde1_lds(unsigned addr) {
  return *(const double *) ((const char *) dde_smem + addr);
}
__device__ __forceinline__ void dde1_sts(unsigned addr, double v) {
  *(double *) ((char *) dde_smem + addr) = v;
}
__device__ __forceinline__ void dde1_lds2(unsigned addr, double &a, double &b) {
  a = dde1_lds(addr);
  b = dde1_lds(addr + 8);
}
__device__ __forceinline__ void dde1_sts2(unsigned addr, double a, double b) {
  dde1_sts(addr, a);
  dde1_sts(addr + 8, b);
}
#else
__device__ __forceinline__ double dde1_lds(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dde1_sts(unsigned addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void dde1_lds2(unsigned addr, double &a, double &b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void dde1_sts2(unsigned addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
#endif

/* the step-size controller's pow (once per attempt), out of line like dde_pow_ctrl, from the
   block's copy of the tables */
static __device__ __noinline__ double dde1_pow_ctrl(unsigned tab, double x, double y) {
  return dde_pow_shared(tab, x, y);
}
static __device__ __noinline__ DdePow2 dde1_pow2_ctrl(unsigned tab, double x, double y1, double y2) {
  return dde_pow2_shared(tab, x, y1, y2);
}

/* the block's copy of the tables (dde_math.h: dde_pow_shared); a barrier follows */
__device__ __forceinline__ void dde1_tables_fill(unsigned base) {
  for (unsigned i = threadIdx.x; i < 128u; i += blockDim.x) {
    double invc, logc, logctail;
    dde_pow_log_entry((int) i, &invc, &logc, &logctail);
    dde1_sts2(base + 16u * i, invc, logc);
    dde1_sts(base + 2048u + 8u * i, logctail);
#if defined(__CUDA_ARCH__)
    const ulonglong2 e = dde_exp_tab_d[i];
    dde1_sts2(base + 3072u + 16u * i, dde_asdouble((uint64_t) e.x), dde_asdouble((uint64_t) e.y));
#else
    dde1_sts2(base + 3072u + 16u * i, dde_asdouble(DDE_EXP_TAB(2 * i)),
              dde_asdouble(DDE_EXP_TAB(2 * i + 1)));
#endif
  }
}

/* Ring state a trajectory starts from: empty, or -- on a restart -- what the
 * previous run left, in which case t0 is the oldest surviving entry's time
 * (dopri_data_reset, src/dopri.c:185-191). */
struct RingStart {
  unsigned count;
  int head;
  double t0s; /* sign * t0 */
};

__device__ __forceinline__ RingStart ring_start(const DopriArgs &a, long long traj,
                                                const double *ring, int hl) {
  RingStart r;
  r.count = 0;
  r.head = 0;
  r.t0s = a.sign * a.times[0];
  if (a.restart && a.n_history > 0) {
    const unsigned cap = (unsigned) a.n_history;
    r.count = a.ring_count[traj];
    r.head = (int) (r.count % cap);
    if (r.count > 0) {
      const unsigned used = r.count < cap ? r.count : cap;
      const unsigned tail_slot = (r.count - used) % cap;
      r.t0s = a.sign * ring[(size_t) tail_slot * (size_t) hl + (hl - 2)];
    }
  }
  return r;
}

/* MODE, chosen at launch (dde_model.cuh):
 *   0  the stage loop as a loop: a warp-uniform stage counter, one inlined copy of
 *      the model, one switch that files its result and forms the next input;
 *   1  the stage loop fully unrolled -- one inlined copy of the model per
 *      evaluation, no switch, output rows staged in shared memory.  For a small
 *      right-hand side without history (Lorenz: 9 flops) the loop + switch cost
 *      several times the model itself (fixed n <= 4, no history ring);
 *   2  unrolled too, but around ONE out-of-line copy of the model (arguments and
 *      result by value, in registers): for a right-hand side with pow and a lag
 *      look-up, whose sixteen inlined copies would overflow the instruction
 *      cache (profiles/r1a) while the loop + switch still cost ~10 % (n = 1). */
template <int N>
struct DdeVec {
  double v[N];
};

/* the model, out of line (MODE 2): arguments and result by value, in registers */
template <class M, int NMAX, int PMAX, int CFG>
static __device__ __noinline__ DdeVec<NMAX> dde_deriv_call(double tt, DdeVec<NMAX> yin,
                                                           DdeVec<PMAX> p) {
  const int s = threadIdx.x;
  s_tt[s] = tt;
  __builtin_assume(s_cfg[s] == CFG);
  DdeVec<NMAX> kt;
  M::deriv((size_t) NMAX, tt, yin.v, kt.v, p.v);
  return kt;
}

template <class M, int METHOD, int MODE = 0>
__global__ void
__launch_bounds__(DDE_TPB, (M::N == 1 ? DDE_MIN_BLOCKS
                            : M::N > 4 || M::N < 1
                                ? 1
                                : (MODE == 1 && METHOD == 0 ? DDE_MIN_BLOCKS_UNROLLED
                                                           : DDE_MIN_BLOCKS_SMALL)))
dopri_tpt_kernel(const DopriArgs a) {
  constexpr bool UNROLLED = MODE != 0;
  constexpr int NC = M::N;
  constexpr int NMAX = NC > 0 ? NC : DDE_DYN_MAXN;
  constexpr int PMAX = M::NP > 0 ? M::NP : (M::NP == 0 ? 1 : DDE_DYN_MAXP);
  constexpr int OMAX = M::N_OUT > 0 ? M::N_OUT : (M::N_OUT == 0 ? 1 : DDE_DYN_MAXOUT);
  constexpr int ORDER = METHOD == 0 ? 5 : 8;
  /* RHS evaluations per attempt, and extra ones per accepted step
     (src/dopri_5.c:42-95; src/dopri_853.c:159-247, src/dopri.c:400-403,
     src/dopri_853.c:304,330-347) */
  constexpr int NST_STEP = METHOD == 0 ? 6 : 11;
  constexpr int NST = METHOD == 0 ? 6 : 16;
  /* window coefficients staged in shared memory (decided by the shape alone) */
  constexpr bool STAGED =
      NC > 0 && sizeof(double) * DDE_WIN * ORDER * NC * DDE_TPB <= DDE_WIN_COEF_BUDGET;
  const int n = NC > 0 ? NC : a.n;
  const int n_parms = M::NP >= 0 ? M::NP : a.n_parms;
  const int n_out = M::N_OUT == 0 ? 0 : a.n_out; /* 0 = output function not requested */
  const int hl = ORDER * n + 2;
  const int s = threadIdx.x;

  if (threadIdx.x == 0) {
    s_lag_u.sign = a.sign;
    s_lag_u.cap = a.n_history;
    s_lag_u.n = n;
    s_lag_u.hl = hl;
    s_lag_u.order = ORDER;
    s_lag_u.staged = NC > 0 ? (STAGED ? 1 : 0) : a.win_staged;
  }
  /* the unrolled kernel has shared memory to spare (no lag window): the step-size
     controller's pow reads a copy of its tables there, behind the threads' row buffers --
     a fixed ~30 cycles per gather instead of an L1 miss now and then (6 % of the stall
     samples of the Lorenz sweep) */
  constexpr bool TAB_SHARED = MODE == 1 && NC > 0 && NC <= 4;
  unsigned tab = 0;
  if (TAB_SHARED) {
    tab = (unsigned) __cvta_generic_to_shared(s_wc) +
          8u * (unsigned) ((DDE_STAGE_ROW_DOUBLES + 1) * DDE_TPB);
    tab = (tab + 15u) & ~15u;
#ifndef DDE_HOST_EMU
    asm volatile("" : "+r"(tab));
#endif
    dde1_tables_fill(tab);
  }
  __syncthreads();

  /* method constants, src/dopri.c:67-78 and :142-153 */
  constexpr double step_factor_min = METHOD == 0 ? 0.2 : 0.333;
  constexpr double step_factor_max = METHOD == 0 ? 10.0 : 6.0;
  constexpr double step_beta = METHOD == 0 ? 0.04 : 0.0;
  constexpr double step_factor_safe = 0.9;
  const double step_constant = METHOD == 0 ? 0.2 - step_beta * 0.75 : 0.125 - step_beta * 0.2;
  const double sign = a.sign;
  const double atol = a.atol, rtol = a.rtol;
  const int n_times = a.n_times;
  const int nt = a.return_initial ? n_times : n_times - 1;
  const double t_end = a.times[n_times - 1];

  /* ---- per-lane solver object (src/dopri.h:55-161), in registers ---- */
  long long traj = -1;
  bool active = false; /* this lane holds a trajectory */
  bool done = false;   /* ... whose integration has ended */
  double p[PMAX];
  /* y, k1 and the vectors every evaluation is handed (its input, its result) are rows of
     ONE array that is live from the first line to the last, so that the compiler cannot
     give a local array of the inlined model the stack slot of the model's own dydt (the
     note in dopri_init_kernel below) */
  double loc[4][NMAX];
  double *const y = loc[0], *const k1 = loc[1], *const yin = loc[2], *const kt = loc[3];
  double t = 0.0, h = 0.0, h_save = 0.0;
  /* pow(fac_old, step_beta) of the controller (src/dopri.c:521), and its value for the
     fac_old = 1e-4 every integration starts from and never goes below */
  const double facb_floor = METHOD == 0 ? dde_pow(1e-4, step_beta) : 1.0;
  double facb = facb_floor;
  bool stop = false, last = false, reject = false;
  int times_idx = 1, tcrit_idx = 0;
  /* times[times_idx]: read once per output row, not once per step, and requested
     before the previous row's interpolation so that a dense output grid does not
     wait for it */
  double t_out = 0.0;
  int n_eval = 0, n_step = 0, n_accept = 0, n_reject = 0;
  unsigned long long stiff_n_stiff = 0, stiff_n_nonstiff = 0;

  /* Output rows of the unrolled kernel leave as whole 32-byte sectors.  A dense output grid
     is bound by the transactions of its row stores (three 8-byte stores per row and lane,
     each a partial write of a sector another store completes); so a thread keeps the
     doubles of its rows in a small buffer of its own in shared memory (where a model with
     history keeps its lag window), addressed by their position in the result array modulo
     lcm(n, 4), and whenever the four doubles of an aligned sector are there it writes them
     with one 256-bit store.  Nothing is shared between lanes: no votes, no shuffles, no
     second pass over the rows.  The doubles before the first and after the last aligned
     sector of a trajectory's block (there are none when n * n_rows is a multiple of four)
     are written one by one. */
  constexpr bool STAGE_OUT = MODE == 1 && NC > 0 && NC <= 4;
  constexpr int SBUF = NC == 3 ? 12 : 4; /* lcm(n, 4), n = 1 .. 4 */
  static_assert(!STAGE_OUT || SBUF <= DDE_STAGE_ROW_DOUBLES, "sector buffer");
  /* the thread's buffer as a shared-window address (odd pitch), opaque to the
     optimiser: it would otherwise re-derive it from the special registers per row */
  unsigned stage =
      (unsigned) __cvta_generic_to_shared(s_wc + (size_t) s * (DDE_STAGE_ROW_DOUBLES + 1));
#ifndef DDE_HOST_EMU
  asm volatile("" : "+r"(stage));
#endif
  double *yo_next = nullptr; /* where the trajectory's next output row goes */
  double *sec_dst = nullptr; /* the aligned sector the oldest double held back belongs to */
  int sec_have = 0;          /* doubles of the result array from sec_dst up to yo_next */
  int sec_skip = 0;          /* ... of which the first are not this trajectory's */
  int sec_rd = 0, sec_wr = 0; /* buffer positions of sec_dst and of yo_next */

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
        if (i < n_parms) {
          p[i] = a.parms[traj * a.parms_stride + i];
        }
      }
      const double *y0 = a.y0 + traj * n;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        y[i] = y0[i];
      }
      /* dopri_data_reset, src/dopri.c:138-219 (the time checks run once on
         the host: times are shared by all trajectories) */
      t = a.times[0];
      times_idx = 1;
      yo_next = a.y_out + ((size_t) traj * (size_t) nt + (a.return_initial ? 1u : 0u)) * (size_t) n;
      if (STAGE_OUT) {
        /* position of the trajectory's first row in the result array, in doubles */
        const unsigned long long g = (unsigned long long) (yo_next - a.y_out);
        sec_skip = (int) (g & 3ull);
        sec_have = sec_skip;
        sec_dst = yo_next - sec_skip;
        sec_wr = (int) (g % (unsigned) SBUF);
        sec_rd = sec_wr - sec_skip; /* >= 0: SBUF is a multiple of four */
      }
      t_out = n_times > 1 ? a.times[1] : 0.0;
      tcrit_idx = a.tcrit_idx0;
      n_eval = n_step = n_accept = n_reject = 0;
      stiff_n_stiff = stiff_n_nonstiff = 0;
      s_ring[s] = a.ring + (size_t) traj * (size_t) a.n_history * (size_t) hl;
      if (n == 1) {
        s_y0_value[s] = y[0]; /* the value itself, see lag_ctx.cuh */
      } else {
        s_y0[s] = y0;
      }
      const RingStart rs = ring_start(a, traj, s_ring[s], hl);
      s_t0[s] = rs.t0s;
      s_count[s] = rs.count;
      s_head[s] = rs.head;
      lag_window_reset(s);
      s_code[s] = 0; /* NOT_SET */
      s_error[s] = 0;

      /* dopri_integrate prologue, src/dopri.c:300-339 */
      facb = facb_floor; /* fac_old = 1e-4 */
      stop = last = reject = false;
      h_save = 0.0;
      /* k1 = f(t0, y0) and the first step size come from dopri_init_kernel */
      const double *k0 = a.init_k1 + traj * n;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        k1[i] = k0[i];
      }
      h = a.init_h[traj];
      n_eval = a.stats[4 * traj + 0];
      const int code0 = a.code[traj];
      if (code0 == -8 || code0 == -10) { /* non-finite initial derivative, src/dopri.c:331-336;
                                            or not restartable from times[0] */
        s_code[s] = code0;
        done = true;
      } else if (code0 != 0) { /* a lag look-up failed during start-up */
        s_code[s] = code0;
        s_error[s] = 1;
      }
    }

    /* ---- one trip of the adaptive loop, src/dopri.c:343-510 ---- */
    bool run = !done;
    bool force_this_step = false;
    if (run) {
      if ((unsigned long long) n_step > a.step_max_n) {
        s_error[s] = 1;
        s_code[s] = -3; /* ERR_TOO_MANY_STEPS */
        done = true;
      } else if (fabs(h) <= a.step_size_min) {
        if (a.step_size_min_allow) {
          h = copysign(a.step_size_min, sign);
          force_this_step = true;
        } else {
          s_error[s] = 1;
          s_code[s] = -4; /* ERR_STEP_SIZE_TOO_SMALL */
          done = true;
        }
      } else if (fabs(h) <= fabs(t) * DBL_EPSILON) {
        s_error[s] = 1;
        s_code[s] = -5; /* ERR_STEP_SIZE_VANISHED */
        done = true;
      }
      run = !done;
    }
    if (run) {
      /* the next stop: a critical time before the end, else the end
         (src/dopri.c:305-311,484-490; recomputed, not carried) */
      double t_stop = t_end;
      if (tcrit_idx < a.n_tcrit && sign * a.tcrit[tcrit_idx] < sign * t_end) {
        t_stop = a.tcrit[tcrit_idx];
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
      if (MODE != 1 && a.n_history > 0) { /* MODE 1 is launched without a ring */
        /* top up the staged lag window while the warp is in step */
        lag_window_ensure<ORDER * NMAX, (NC > 0)>(s, sign * t, sign * (t + h));
      }
    }

    /* stage storage: k1 is carried (FSAL); dopri5 uses k2..k6, dop853
       k2..k10.  ysave: the stage input the stiffness test needs (dopri5's
       `ysti`, dop853's last y1).  hd: dense-output coefficients of this
       step, i.e. the ring's head entry, in registers until committed. */
    double k2[NMAX], k3[NMAX], k4[NMAX], k5[NMAX], k6[NMAX], k7[NMAX], k8[NMAX], k9[NMAX],
        k10[NMAX], k14[NMAX], k15[NMAX], ysave[NMAX], hd[ORDER * NMAX];
    bool accepted = false;
    double h_new = 0.0;
    /* the stage input and its time, prepared one evaluation ahead */
    double tt = t;
    if (run) { /* input of evaluation 0: src/dopri_5.c:54-57, src/dopri_853.c:178-181 */
#pragma unroll
      for (int i = 0; i < n; ++i) {
        yin[i] = y[i] + h * (METHOD == 0 ? DP5_A21 : DP8_A21) * k1[i];
      }
      tt = t + (METHOD == 0 ? DP5_C2 : DP8_C2) * h;
    }

    /* how many evaluations this lane takes part in: none, the attempt's, or --
       once accepted -- the accept-only ones too */
    int my_stages = run ? NST_STEP : 0;
    constexpr int STAGE_UNROLL = UNROLLED ? NST : 1;
#pragma unroll STAGE_UNROLL
    for (int st = 0; st < NST; ++st) {
      if (st >= my_stages) {
        continue;
      }
      /* -- 1. dopri_eval, src/dopri.c:831-835: the one hot copy of the model, on
            the stage input `yin` at time `tt` that the previous trip through the
            switch below (or the attempt's prologue) prepared -- */
      if (MODE == 2) {
        DdeVec<NMAX> vin;
        DdeVec<PMAX> vp;
#pragma unroll
        for (int i = 0; i < NMAX; ++i) vin.v[i] = yin[i];
#pragma unroll
        for (int i = 0; i < PMAX; ++i) vp.v[i] = p[i];
        const DdeVec<NMAX> vk =
            dde_deriv_call<M, NMAX, PMAX, (NC | (ORDER << 16) | ((STAGED ? 1 : 0) << 24))>(tt, vin,
                                                                                         vp);
#pragma unroll
        for (int i = 0; i < NMAX; ++i) kt[i] = vk.v[i];
      } else {
        s_tt[s] = tt; /* lag look-ups learn their offset from it */
        if (NC > 0) {
          /* the lag functions read the method / width / staging flags from the
             thread's context: tell the optimiser what they will find */
          __builtin_assume(s_cfg[s] == (NC | (ORDER << 16) | ((STAGED ? 1 : 0) << 24)));
        }
        M::deriv((size_t) n, tt, yin, kt, p);
      }
      ++n_eval;

      /* -- 2. file the result and form the input of the next evaluation (one
            warp-uniform switch per evaluation); after the last stage of the
            attempt: error estimate, controller, accept / reject -- */
      bool attempt_done = false; /* all NST_STEP evaluations of the attempt are in */
      bool step_done = false;    /* ... and so are the accept-only ones */
      if (METHOD == 0) {
        switch (st) {
          case 0:
#pragma unroll
            for (int i = 0; i < n; ++i) k2[i] = kt[i];
            /* input of evaluation 1 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP5_A31 * k1[i] + DP5_A32 * k2[i]);
            }
            tt = t + DP5_C3 * h;
            break;
          case 1:
#pragma unroll
            for (int i = 0; i < n; ++i) k3[i] = kt[i];
            /* input of evaluation 2 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP5_A41 * k1[i] + DP5_A42 * k2[i] + DP5_A43 * k3[i]);
            }
            tt = t + DP5_C4 * h;
            break;
          case 2:
#pragma unroll
            for (int i = 0; i < n; ++i) k4[i] = kt[i];
            /* input of evaluation 3 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP5_A51 * k1[i] + DP5_A52 * k2[i] + DP5_A53 * k3[i] +
                                   DP5_A54 * k4[i]);
            }
            tt = t + DP5_C5 * h;
            break;
          case 3:
#pragma unroll
            for (int i = 0; i < n; ++i) k5[i] = kt[i];
            /* input of evaluation 4 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP5_A61 * k1[i] + DP5_A62 * k2[i] + DP5_A63 * k3[i] +
                                   DP5_A64 * k4[i] + DP5_A65 * k5[i]);
              ysave[i] = yin[i];
            }
            tt = t + h;
            break;
          case 4:
#pragma unroll
            for (int i = 0; i < n; ++i) k6[i] = kt[i];
            /* input of evaluation 5 */ /* 5 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP5_A71 * k1[i] + DP5_A73 * k3[i] + DP5_A74 * k4[i] +
                                   DP5_A75 * k5[i] + DP5_A76 * k6[i]);
            }
            tt = t + h;
            break;
          default: /* 5: k2 = f(t + h, y1); yin is y1 */
#pragma unroll
            for (int i = 0; i < n; ++i) k2[i] = kt[i];
            attempt_done = true;
            break;
        }
      } else {
        switch (st) {
          case 0:
#pragma unroll
            for (int i = 0; i < n; ++i) k2[i] = kt[i];
            /* input of evaluation 1 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP8_A31 * k1[i] + DP8_A32 * k2[i]);
            }
            tt = t + DP8_C3 * h;
            break;
          case 1:
#pragma unroll
            for (int i = 0; i < n; ++i) k3[i] = kt[i];
            /* input of evaluation 2 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP8_A41 * k1[i] + DP8_A43 * k3[i]);
            }
            tt = t + DP8_C4 * h;
            break;
          case 2:
#pragma unroll
            for (int i = 0; i < n; ++i) k4[i] = kt[i];
            /* input of evaluation 3 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP8_A51 * k1[i] + DP8_A53 * k3[i] + DP8_A54 * k4[i]);
            }
            tt = t + DP8_C5 * h;
            break;
          case 3:
#pragma unroll
            for (int i = 0; i < n; ++i) k5[i] = kt[i];
            /* input of evaluation 4 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP8_A61 * k1[i] + DP8_A64 * k4[i] + DP8_A65 * k5[i]);
            }
            tt = t + DP8_C6 * h;
            break;
          case 4:
#pragma unroll
            for (int i = 0; i < n; ++i) k6[i] = kt[i];
            /* input of evaluation 5 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP8_A71 * k1[i] + DP8_A74 * k4[i] + DP8_A75 * k5[i] +
                                   DP8_A76 * k6[i]);
            }
            tt = t + DP8_C7 * h;
            break;
          case 5:
#pragma unroll
            for (int i = 0; i < n; ++i) k7[i] = kt[i];
            /* input of evaluation 6 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP8_A81 * k1[i] + DP8_A84 * k4[i] + DP8_A85 * k5[i] +
                                   DP8_A86 * k6[i] + DP8_A87 * k7[i]);
            }
            tt = t + DP8_C8 * h;
            break;
          case 6:
#pragma unroll
            for (int i = 0; i < n; ++i) k8[i] = kt[i];
            /* input of evaluation 7 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP8_A91 * k1[i] + DP8_A94 * k4[i] + DP8_A95 * k5[i] +
                                   DP8_A96 * k6[i] + DP8_A97 * k7[i] + DP8_A98 * k8[i]);
            }
            tt = t + DP8_C9 * h;
            break;
          case 7:
#pragma unroll
            for (int i = 0; i < n; ++i) k9[i] = kt[i];
            /* input of evaluation 8 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP8_A101 * k1[i] + DP8_A104 * k4[i] + DP8_A105 * k5[i] +
                                   DP8_A106 * k6[i] + DP8_A107 * k7[i] + DP8_A108 * k8[i] +
                                   DP8_A109 * k9[i]);
            }
            tt = t + DP8_C10 * h;
            break;
          case 8:
#pragma unroll
            for (int i = 0; i < n; ++i) k10[i] = kt[i];
            /* input of evaluation 9 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP8_A111 * k1[i] + DP8_A114 * k4[i] + DP8_A115 * k5[i] +
                                   DP8_A116 * k6[i] + DP8_A117 * k7[i] + DP8_A118 * k8[i] +
                                   DP8_A119 * k9[i] + DP8_A1110 * k10[i]);
            }
            tt = t + DP8_C11 * h;
            break;
          case 9:
#pragma unroll
            for (int i = 0; i < n; ++i) k2[i] = kt[i];
            /* input of evaluation 10 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP8_A121 * k1[i] + DP8_A124 * k4[i] + DP8_A125 * k5[i] +
                                   DP8_A126 * k6[i] + DP8_A127 * k7[i] + DP8_A128 * k8[i] +
                                   DP8_A129 * k9[i] + DP8_A1210 * k10[i] + DP8_A1211 * k2[i]);
              ysave[i] = yin[i];
            }
            tt = t + h;
            break;
          case 10: /* src/dopri_853.c:242-246 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              k3[i] = kt[i];
              k4[i] = DP8_B1 * k1[i] + DP8_B6 * k6[i] + DP8_B7 * k7[i] + DP8_B8 * k8[i] +
                      DP8_B9 * k9[i] + DP8_B10 * k10[i] + DP8_B11 * k2[i] + DP8_B12 * k3[i];
              k5[i] = y[i] + h * k4[i];
            }
            attempt_done = true;
            break;
          case 11: /* k4 = f(t_old, k5): only the stiffness test reads it */
#pragma unroll
            for (int i = 0; i < n; ++i) k4[i] = kt[i];
            /* input of evaluation 12 */ /* dopri853_save_history, src/dopri_853.c:304 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = k5[i];
            }
            tt = t + h;
            break;
          case 12: /* k4 = f(t + h, k5), src/dopri_853.c:304 */
#pragma unroll
            for (int i = 0; i < n; ++i) k4[i] = kt[i];
            /* input of evaluation 13 */ /* src/dopri_853.c:330-335 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP8_A141 * k1[i] + DP8_A147 * k7[i] + DP8_A148 * k8[i] +
                                   DP8_A149 * k9[i] + DP8_A1410 * k10[i] + DP8_A1411 * k2[i] +
                                   DP8_A1412 * k3[i] + DP8_A1413 * k4[i]);
            }
            tt = t + DP8_C14 * h;
            break;
          case 13:
#pragma unroll
            for (int i = 0; i < n; ++i) k14[i] = kt[i];
            /* input of evaluation 14 */ /* src/dopri_853.c:336-341 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP8_A151 * k1[i] + DP8_A156 * k6[i] + DP8_A157 * k7[i] +
                                   DP8_A158 * k8[i] + DP8_A1511 * k2[i] + DP8_A1512 * k3[i] +
                                   DP8_A1513 * k4[i] + DP8_A1514 * k14[i]);
            }
            tt = t + DP8_C15 * h;
            break;
          case 14:
#pragma unroll
            for (int i = 0; i < n; ++i) k15[i] = kt[i];
            /* input of evaluation 15 */ /* 15, src/dopri_853.c:342-347 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = y[i] + h * (DP8_A161 * k1[i] + DP8_A166 * k6[i] + DP8_A167 * k7[i] +
                                   DP8_A168 * k8[i] + DP8_A169 * k9[i] + DP8_A1613 * k4[i] +
                                   DP8_A1614 * k14[i] + DP8_A1615 * k15[i]);
            }
            tt = t + DP8_C16 * h;
            break;
          default: /* 15, src/dopri_853.c:306-327,350-366 */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              const double k16 = kt[i];
              const double ydiff = k5[i] - y[i];
              const double bspl = h * k1[i] - ydiff;
              hd[i] = y[i];
              hd[n + i] = ydiff;
              hd[2 * n + i] = bspl;
              hd[3 * n + i] = ydiff - h * k4[i] - bspl;
              hd[4 * n + i] = h * (DP8_D41 * k1[i] + DP8_D46 * k6[i] + DP8_D47 * k7[i] +
                                   DP8_D48 * k8[i] + DP8_D49 * k9[i] + DP8_D410 * k10[i] +
                                   DP8_D411 * k2[i] + DP8_D412 * k3[i] + DP8_D413 * k4[i] +
                                   DP8_D414 * k14[i] + DP8_D415 * k15[i] + DP8_D416 * k16);
              hd[5 * n + i] = h * (DP8_D51 * k1[i] + DP8_D56 * k6[i] + DP8_D57 * k7[i] +
                                   DP8_D58 * k8[i] + DP8_D59 * k9[i] + DP8_D510 * k10[i] +
                                   DP8_D511 * k2[i] + DP8_D512 * k3[i] + DP8_D513 * k4[i] +
                                   DP8_D514 * k14[i] + DP8_D515 * k15[i] + DP8_D516 * k16);
              hd[6 * n + i] = h * (DP8_D61 * k1[i] + DP8_D66 * k6[i] + DP8_D67 * k7[i] +
                                   DP8_D68 * k8[i] + DP8_D69 * k9[i] + DP8_D610 * k10[i] +
                                   DP8_D611 * k2[i] + DP8_D612 * k3[i] + DP8_D613 * k4[i] +
                                   DP8_D614 * k14[i] + DP8_D615 * k15[i] + DP8_D616 * k16);
              hd[7 * n + i] = h * (DP8_D71 * k1[i] + DP8_D76 * k6[i] + DP8_D77 * k7[i] +
                                   DP8_D78 * k8[i] + DP8_D79 * k9[i] + DP8_D710 * k10[i] +
                                   DP8_D711 * k2[i] + DP8_D712 * k3[i] + DP8_D713 * k4[i] +
                                   DP8_D714 * k14[i] + DP8_D715 * k15[i] + DP8_D716 * k16);
            }
            step_done = true;
            break;
        }
      }

      if (attempt_done) {
        if (s_error[s]) {
          done = true; /* a lag look-up failed during dopri_step, src/dopri.c:387-389 */
          my_stages = 0;
          continue;
        }
        double err;
        if (METHOD == 0) {
          /* dopri5_error, src/dopri_5.c:97-104.  (The reference first
             overwrites k4 with this combination, :91-94, and writes the
             fifth dense coefficient into the ring head on every attempt,
             :84-89; only accepted steps read the head, so k4 is kept and that
             coefficient is formed on acceptance from the same operands.) */
          err = 0.0;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            const double ke = h * (DP5_E1 * k1[i] + DP5_E3 * k3[i] + DP5_E4 * k4[i] +
                                   DP5_E5 * k5[i] + DP5_E6 * k6[i] + DP5_E7 * k2[i]);
            const double sk = atol + rtol * fmax(fabs(y[i]), fabs(yin[i]));
            err += sq(ke / sk);
          }
          err = sqrt(err / (double) n);
        } else {
          /* dopri853_error, src/dopri_853.c:249-283: scaled by sign, not
             |h|, guarded by deno < 0 */
          err = 0.0;
          double err2 = 0.0;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            const double sk = atol + rtol * fmax(fabs(y[i]), fabs(k5[i]));
            double erri = k4[i] - DP8_BHH1 * k1[i] - DP8_BHH2 * k9[i] - DP8_BHH3 * k3[i];
            err2 += sq(erri / sk);
            erri = DP8_ER1 * k1[i] + DP8_ER6 * k6[i] + DP8_ER7 * k7[i] + DP8_ER8 * k8[i] +
                   DP8_ER9 * k9[i] + DP8_ER10 * k10[i] + DP8_ER11 * k2[i] + DP8_ER12 * k3[i];
            err += sq(erri / sk);
          }
          double deno = err + 0.01 * err2;
          deno = deno < 0 ? 1.0 : deno;
          err = sign * err * sqrt(1.0 / ((double) n * deno));
        }
        /* dopri_h_new, src/dopri.c:516-525 */
        /* dop853 has step_beta = 0 and pow(x, 0) == 1 for every x.  dopri5 wants
           pow(err, 0.17) now and, if this step is accepted, pow(fac_old = max(err, 1e-4),
           0.04) at the attempts that follow it: two powers of one base, whose logarithm
           (half of a pow) is computed once (dde_pow2); facb is the value kept from the step
           that set fac_old */
        double fac11, facb_next = 1.0;
        if (METHOD == 0) {
          const DdePow2 pw = TAB_SHARED ? dde1_pow2_ctrl(tab, err, step_constant, step_beta)
                                        : dde_pow2_ctrl(err, step_constant, step_beta);
          fac11 = pw.a;
          facb_next = pw.b;
        } else {
          fac11 = TAB_SHARED ? dde1_pow_ctrl(tab, err, step_constant)
                             : dde_pow_ctrl(err, step_constant);
        }
        double fac = fac11 / facb;
        fac = fmax(1.0 / step_factor_max, fmin(1.0 / step_factor_min, fac / step_factor_safe));
        accepted = (err <= 1) || force_this_step;
        /* one division for both outcomes -- a rejected lane divides by its own factor,
           src/dopri.c:495-499 -- instead of a second one that only the rejected lanes of the
           warp execute */
        const double fac_rej = fmin(1 / step_factor_min, fac11 / step_factor_safe);
        h_new = h / (accepted ? fac : fac_rej);
        if (accepted) {
          my_stages = NST;
          /* fac_old = fmax(err, 1e-4), src/dopri.c:398-399: kept as pow(fac_old, step_beta) */
          facb = err >= 1e-4 ? facb_next : facb_floor;
          ++n_accept;
          if (METHOD == 0) {
            step_done = true;
          } else {
            /* input of evaluation 11: the redundant one at the OLD time,
               src/dopri.c:400-403: it counts, and its lag
               look-ups can fail */
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yin[i] = k5[i];
            }
            tt = t;
          }
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

      /* dopri_test_stiff, src/dopri.c:668-693: after the redundant
         evaluation for dop853, right after acceptance for dopri5 */
      bool stiff_due = false;
      if (METHOD == 0 ? step_done : st == 11) { /* the warp-uniform test first */
        stiff_due = a.stiff_check != 0 &&
                    (stiff_n_stiff > 0 || (unsigned long long) n_accept % a.stiff_check == 0);
      }
      if (stiff_due) {
        double stnum = 0, stden = 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
          if (METHOD == 0) { /* src/dopri_5.c:129-140; yin is y1 */
            stnum += sq(k2[i] - k6[i]);
            stden += sq(yin[i] - ysave[i]);
          } else { /* src/dopri_853.c:382-394 */
            stnum += sq(k4[i] - k3[i]);
            stden += sq(k5[i] - ysave[i]);
          }
        }
        const bool stiff =
            stden > 0 && fabs(h) * sqrt(stnum / stden) > (METHOD == 0 ? 3.25 : 6.1);
        bool give_up = false;
        if (stiff) {
          stiff_n_nonstiff = 0;
          if (stiff_n_stiff++ >= 15) {
            give_up = true;
          }
        } else if (stiff_n_stiff > 0) {
          if (stiff_n_nonstiff++ >= 6) {
            stiff_n_stiff = 0;
          }
        }
        if (give_up) {
          s_error[s] = 1;
          s_code[s] = -7; /* ERR_STIFF */
          done = true;
          my_stages = 0;
          continue;
        }
      }

      if (step_done) {
        /* ---- the rest of an accepted step, src/dopri.c:410-494 ---- */
        if (METHOD == 0) {
          /* dopri5_save_history, src/dopri_5.c:106-118 and :84-89; yin is y1 */
#pragma unroll
          for (int i = 0; i < n; ++i) {
            const double ydiff = yin[i] - y[i];
            const double bspl = h * k1[i] - ydiff;
            hd[i] = y[i];
            hd[n + i] = ydiff;
            hd[2 * n + i] = bspl;
            hd[3 * n + i] = -h * k2[i] + ydiff - bspl;
            hd[4 * n + i] = h * (DP5_D1 * k1[i] + DP5_D3 * k3[i] + DP5_D4 * k4[i] +
                                 DP5_D5 * k5[i] + DP5_D6 * k6[i] + DP5_D7 * k2[i]);
          }
#pragma unroll
          for (int i = 0; i < n; ++i) {
            k1[i] = k2[i]; /* src/dopri.c:416-417 */
            y[i] = yin[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < n; ++i) {
            k1[i] = k4[i]; /* src/dopri.c:420-421 */
            y[i] = k5[i];
          }
        }
        const double t_old = t;
        t += h;

        /* output rows, src/dopri.c:437-456: interpolate from the head.  The rows of a step
           all divide by its h: the divisor's half of the division is done once */
        /* a row into the thread's sector buffer; the sector it completes, if any, out */
        auto put_row = [&](const double *row) {
#pragma unroll
          for (int i = 0; i < n; ++i) {
            sts_f64(stage + 8u * (unsigned) (sec_wr + i), row[i]);
          }
          sec_wr = sec_wr + n == SBUF ? 0 : sec_wr + n;
          sec_have += n;
          if (sec_have >= 4) { /* n <= 4: at most one sector per row */
            const unsigned src = stage + 8u * (unsigned) sec_rd;
            const double d0 = lds_f64(src), d1 = lds_f64(src + 8u), d2 = lds_f64(src + 16u),
                         d3 = lds_f64(src + 24u);
            if (__builtin_expect(sec_skip == 0, 1)) {
              stg_sector(sec_dst, d0, d1, d2, d3);
            } else { /* the trajectory's block starts inside this sector */
              if (sec_skip <= 1) {
                sec_dst[1] = d1;
              }
              if (sec_skip <= 2) {
                sec_dst[2] = d2;
              }
              sec_dst[3] = d3;
              sec_skip = 0;
            }
            sec_dst += 4;
            sec_rd = sec_rd + 4 == SBUF ? 0 : sec_rd + 4;
            sec_have -= 4;
          }
        };
        double h_rcp = 0.0;
        if (STAGE_OUT && times_idx < n_times && (last || sign * t_out <= sign * t)) {
          h_rcp = dde_div_prepare(h);
        }
        if (STAGE_OUT && n_out == 0) {
          /* DDE_ROW_GROUP rows per trip: a dense grid asks for a few rows per step, a different
             number in every lane, and a row is one dependent chain (its time, the quotient,
             the Horner form); in groups the warp makes fewer trips and each trip overlaps
             several chains.  Every row of the group is computed whether it is due or not
             (the due ones are a prefix: the times ascend), and kept if it is. */
          constexpr int G = DDE_ROW_GROUP;
          while (times_idx < n_times && (last || sign * t_out <= sign * t)) {
            double tq[G + 1];
            tq[0] = t_out;
#pragma unroll
            for (int k = 1; k <= G; ++k) {
              tq[k] = times_idx + k < n_times ? a.times[times_idx + k] : 0.0;
            }
            int due = 1;
#pragma unroll
            for (int k = 1; k < G; ++k) {
              due += (due == k && times_idx + k < n_times && (last || sign * tq[k] <= sign * t))
                         ? 1 : 0;
            }
            double rows[G][NMAX];
#pragma unroll
            for (int k = 0; k < G; ++k) {
              const double theta = dde_div_prepared(tq[k] - t_old, h, h_rcp);
              const double theta1 = 1 - theta;
#pragma unroll
              for (int i = 0; i < n; ++i) {
                rows[k][i] = METHOD == 0 ? interp5(n, theta, theta1, hd + i)
                                         : interp853(n, theta, theta1, hd + i);
              }
            }
            put_row(rows[0]);
#pragma unroll
            for (int k = 1; k < G; ++k) {
              if (k < due) {
                put_row(rows[k]);
              }
            }
            yo_next += due * n;
            times_idx += due;
            double tn = tq[G];
#pragma unroll
            for (int k = G - 1; k >= 1; --k) {
              tn = due == k ? tq[k] : tn;
            }
            t_out = tn;
          }
        }
        for (;;) {
          if (times_idx >= n_times) {
            break;
          }
          const double tq = t_out;
          if (!(last || sign * tq <= sign * t)) {
            break;
          }
          const double t_next = times_idx + 1 < n_times ? a.times[times_idx + 1] : 0.0;
          const double theta =
              STAGE_OUT ? dde_div_prepared(tq - t_old, h, h_rcp) : (tq - t_old) / h;
          const double theta1 = 1 - theta;
          double row[NMAX];
          /* row of this time: times_idx, shifted by one without the initial row; rows
             are emitted in order, so where they go is a running pointer */
          double *yo = yo_next;
          yo_next += n;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            row[i] = METHOD == 0 ? interp5(n, theta, theta1, hd + i)
                                 : interp853(n, theta, theta1, hd + i);
          }
          if (STAGE_OUT) {
            put_row(row);
          } else {
#pragma unroll
            for (int i = 0; i < n; ++i) {
              yo[i] = row[i];
            }
          }
          if (n_out > 0) {
            double o[OMAX];
            const size_t r = (size_t) traj * (size_t) nt +
                             (size_t) (times_idx - (a.return_initial ? 0 : 1));
            double *oo = a.out + r * (size_t) n_out;
            M::output((size_t) n, tq, row, (size_t) n_out, o, p);
#pragma unroll
            for (int i = 0; i < OMAX; ++i) {
              if (i < n_out) {
                oo[i] = o[i];
              }
            }
          }
          ++times_idx;
          t_out = t_next;
        }

        /* ring_buffer_head_advance, src/dopri.c:460: the head entry goes to
           its slot in HBM and becomes visible to lag queries */
        if (MODE != 1 && a.n_history > 0) {
          double *dst = s_ring[s] + (size_t) s_head[s] * (size_t) hl;
#pragma unroll
          for (int i = 0; i < ORDER * n; ++i) {
            dst[i] = hd[i];
          }
          dst[ORDER * n] = t_old;
          dst[ORDER * n + 1] = h;
          ring_advance<ORDER * NMAX>(s, a.n_history, hd, ORDER * n, t_old, h);
        }

        if (last) {
          s_code[s] = 1;           /* OK_COMPLETE */
          done = true;
          my_stages = 0;
          continue;
        }
        if (fabs(h_new) >= a.step_size_max) {
          h_new = copysign(a.step_size_max, sign);
        }
        if (reject) {
          h_new = copysign(fmin(fabs(h_new), fabs(h)), sign);
          reject = false;
        }
        if (stop) {
          /* critical times, src/dopri.c:476-491 */
          while (tcrit_idx < a.n_tcrit && sign * a.tcrit[tcrit_idx] <= sign * t) {
            if (M::N_EVENTS > 0 && a.event != nullptr) {
              /* an event changes the state (and may change the parameters) in
                 place; k1 is NOT re-evaluated, as in the reference */
              const int ev = a.event[tcrit_idx];
              if (ev >= 0) {
                M::event(ev, (size_t) n, t, y, p);
              }
            }
            ++tcrit_idx;
          }
          stop = false;
        } else {
          h = h_new;
        }
      }
    }

    if (done) {
      if (STAGE_OUT) { /* the doubles past the last whole sector */
        for (int k = sec_skip; k < sec_have; ++k) {
          sec_dst[k] = lds_f64(stage + 8u * (unsigned) (sec_rd + k));
        }
        sec_have = sec_skip = 0;
      }
      /* retire: statistics as r_dopri_cleanup reads them, src/r_dopri.c:412-428 */
      a.stats[4 * traj + 0] = n_eval;
      a.stats[4 * traj + 1] = n_step;
      a.stats[4 * traj + 2] = n_accept;
      a.stats[4 * traj + 3] = n_reject;
      a.code[traj] = s_code[s];
      if (s_code[s] == 1) {
        a.step_size[traj] = h_save; /* obj->step_size_initial on return, src/dopri.c:463 */
      } else if (!a.restart) {
        a.step_size[traj] = a.step_size_initial;
      }
      if (s_code[s] != -10) { /* a trajectory that could not restart keeps its old state */
        a.t_last[traj] = t;
        if (a.y_final != nullptr) {
#pragma unroll
          for (int i = 0; i < n; ++i) {
            a.y_final[traj * n + i] = y[i];
          }
        }
      }
      a.ring_count[traj] = s_count[s];
      chunk_retire(a, traj);
      active = false;
    }
  }
}

/* Start-up of every trajectory, one thread each: row 0 of the results, k1 =
 * f(t0, y0), the finiteness check and dopri_h_init -- the part of
 * dopri_integrate before its loop, src/dopri.c:320-339,527-572.  Kept out of
 * the stepping kernel so that kernel's instruction footprint stays small and
 * its lanes never run start-up code one at a time.  Leaves code = 0 (go),
 * -8 (non-finite derivative: trajectory over) or -6 (a lag look-up failed:
 * the stepping kernel ends the trajectory after its first attempt,
 * src/dopri.c:387-389) and n_eval in stats[0]. */
template <class M, int METHOD>
__global__ void __launch_bounds__(DDE_TPB) dopri_init_kernel(const DopriArgs a) {
  constexpr int NC = M::N;
  constexpr int NMAX = NC > 0 ? NC : DDE_DYN_MAXN;
  constexpr int PMAX = M::NP > 0 ? M::NP : (M::NP == 0 ? 1 : DDE_DYN_MAXP);
  constexpr int OMAX = M::N_OUT > 0 ? M::N_OUT : (M::N_OUT == 0 ? 1 : DDE_DYN_MAXOUT);
  constexpr int ORDER = METHOD == 0 ? 5 : 8;
  const int n = NC > 0 ? NC : a.n;
  const int n_parms = M::NP >= 0 ? M::NP : a.n_parms;
  const int n_out = M::N_OUT == 0 ? 0 : a.n_out;
  const int s = threadIdx.x;
  if (threadIdx.x == 0) {
    s_lag_u.sign = a.sign;
    s_lag_u.cap = a.n_history;
    s_lag_u.n = n;
    s_lag_u.hl = ORDER * n + 2;
    s_lag_u.order = ORDER;
    s_lag_u.staged = 0;
  }
  __syncthreads();
  const long long traj = blockIdx.x * (long long) DDE_TPB + threadIdx.x;
  if (traj >= a.n_traj) {
    return;
  }
  if (a.restart && (a.code[traj] != 1 || a.t_last[traj] != a.times[0])) {
    /* r_dopri_continue: "Incorrect initial time on integration restart"
       (src/r_dopri.c:215-217); a failed solve has nothing to continue from */
    a.code[traj] = -10;
    a.stats[4 * traj + 0] = 0;
    a.init_h[traj] = 0.0;
    return;
  }
  /* obj->step_size_initial: on a restart the step the last run would have
     taken next (src/dopri.c:463) */
  const double ssi = a.restart ? a.step_size[traj] : a.step_size_initial;
  const double sign = a.sign, atol = a.atol, rtol = a.rtol;
  const int nt = a.return_initial ? a.n_times : a.n_times - 1;
  /* y, k1 and the vectors of dopri_h_init's second evaluation are rows of ONE array.  The
     model's dydt must be: nvcc 12.9 gives a local array of the inlined model (e.g.
     `double lagged[16]` filled by ylag_all) the stack slot of a dydt array of its own when
     it can place the two in different live ranges elsewhere, although the model reads the
     one while it writes the other (seen with a model that reads lagged[n - 1 - i] for
     dydt[i]: tests/models/growth.h, delayed_growth_all).  Rows of an array that is live
     from the first line to the last cannot be shared out. */
  double p[PMAX], loc[4][NMAX];
  double *const y = loc[0], *const k1 = loc[1], *const f1 = loc[2], *const ye = loc[3];
#pragma unroll
  for (int i = 0; i < PMAX; ++i) {
    if (i < n_parms) {
      p[i] = a.parms[traj * a.parms_stride + i];
    }
  }
  const double *y0 = a.y0 + traj * n;
#pragma unroll
  for (int i = 0; i < n; ++i) {
    y[i] = y0[i];
  }
  const double t = a.times[0];
  if (n == 1) {
    s_y0_value[s] = y[0];
  } else {
    s_y0[s] = y0;
  }
  s_ring[s] = a.ring + (size_t) traj * (size_t) a.n_history * (size_t) (ORDER * n + 2);
  const RingStart rs = ring_start(a, traj, s_ring[s], ORDER * n + 2);
  s_t0[s] = rs.t0s;
  s_count[s] = rs.count;
  s_head[s] = rs.head;
  s_code[s] = 0;
  s_error[s] = 0;
  lag_window_reset(s);
  if (a.return_initial) { /* src/dopri.c:320-327 */
    double *yo = a.y_out + (size_t) traj * (size_t) nt * (size_t) n;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      yo[i] = y[i];
    }
    if (n_out > 0) {
      double *oo = a.out + (size_t) traj * (size_t) nt * (size_t) n_out;
      double o[OMAX];
      M::output((size_t) n, t, y, (size_t) n_out, o, p);
#pragma unroll
      for (int i = 0; i < OMAX; ++i) {
        if (i < n_out) {
          oo[i] = o[i];
        }
      }
    }
  }
  int n_eval = 0;
  double h = 0.0;
  int code = 0;
  M::deriv((size_t) n, t, y, k1, p); /* src/dopri.c:329 */
  ++n_eval;
  bool finite0 = true;
#pragma unroll
  for (int i = 0; i < n; ++i) {
    finite0 = finite0 && isfinite(k1[i]);
  }
  if (!finite0) {
    code = -8; /* DDE_ERR_NONFINITE_DERIV, src/dopri.c:331-336 */
  } else if (ssi > 0.0) {
    h = ssi; /* src/dopri.c:528-530 */
  } else {
    /* dopri_h_init, src/dopri.c:527-572 */
    double norm_f = 0.0, norm_y = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const double sk = atol + rtol * fabs(y[i]);
      norm_f += sq(k1[i] / sk);
      norm_y += sq(y[i] / sk);
    }
    h = (norm_f <= 1e-10 || norm_y <= 1e-10) ? 1e-6 : sqrt(norm_y / norm_f) * 0.01;
    h = copysign(fmin(h, a.step_size_max), sign);
#pragma unroll
    for (int i = 0; i < n; ++i) {
      ye[i] = y[i] + h * k1[i];
    }
    M::deriv((size_t) n, t + h, ye, f1, p);
    ++n_eval;
    double der2 = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const double sk = atol + rtol * fabs(y[i]);
      der2 += sq((f1[i] - k1[i]) / sk);
    }
    der2 = sqrt(der2) / h;
    const double der12 = fmax(fabs(der2), sqrt(norm_f));
    const double h1 = (der12 <= 1e-15) ? fmax(1e-6, fabs(h) * 1e-3)
                                       : dde_pow(0.01 / der12, 1.0 / ORDER);
    h = fmin(fmin(100 * fabs(h), h1), a.step_size_max);
    h = copysign(h, sign);
  }
  if (code == 0 && s_error[s]) {
    code = s_code[s];
  }
  double *k0 = a.init_k1 + traj * n;
#pragma unroll
  for (int i = 0; i < n; ++i) {
    k0[i] = k1[i];
  }
  a.init_h[traj] = h;
  a.stats[4 * traj + 0] = n_eval;
  a.code[traj] = code;
}

} /* namespace dde */

#endif /* DDE_DOPRI_KERNELS_CUH */
