/* dopri_coop_kernel.cuh -- batched Dormand-Prince integrator for LARGE models:
 * one trajectory per thread block (one warp for n <= 128, four warps for
 * n = 512), the state spread over the block's threads.
 *
 * Same reference semantics as dopri_kernels.cuh (it restates the same code:
 * /root/reference/src/dopri.c:138-219,291-572, src/dopri_5.c, src/dopri_853.c),
 * different mapping:
 *   - the solver object (y, k1..k10, the stage input) is a set of shared-memory
 *     vectors; thread `tid` works on components tid, tid + TPT, ... of every
 *     per-component loop of the reference, so ring commits, output rows and lag
 *     reads are coalesced along the component index, registers stay few and
 *     several blocks are resident per SM;
 *   - the stage loops are driven by the tableau as data (lists of
 *     (slot, weight) terms in constant memory, summed left to right exactly as
 *     the reference's expressions), so the code is a few small loops;
 *   - the model's deriv() is called by the whole block at once and reads the
 *     stage vector / writes the derivative slot in place (the cooperative model
 *     ABI of <dde/dde.h>: DDE_FOR / DDE_SCRATCH / DDE_SYNC);
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
#ifndef DDE_COOP_MINB256
#define DDE_COOP_MINB256 3 /* resident 256-thread blocks the register allocation must allow */
#endif
#ifndef DDE_COOP_MAX_TABLE
#define DDE_COOP_MAX_TABLE 1024 /* longest ring whose step times are kept in shared memory */
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
  /* start time (sign-adjusted) and step of every ring slot, in shared memory
     when the ring is short enough: a look-up is then a block-wide scan */
  int tab_off; /* offset of the table in s_coop_dyn (doubles), -1: no table.  An offset, not a
                  pointer, so that the accesses compile to shared-memory loads */
  /* the entries the last few look-ups ended at (logical indices): a step's
     evaluations query each lag within an entry or two, so most look-ups are
     answered by testing these, identically in every thread, without a scan */
  unsigned hint[4];
  int hint_slot[4]; /* their ring slots (no division on the fast path) */
  unsigned hint_next;
  /* [hint_ts, hint_tn): the sign-adjusted times a hinted entry answers for (+inf, .): none.
     Kept up to date where the ring changes (a commit gives the newest entry a successor
     and may recycle the oldest), so that testing a hint is two loads and two comparisons */
  double hint_ts[4], hint_tn[4];
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
extern __shared__ double s_coop_dyn[]; /* 16 solver vectors of n doubles, then model scratch */
#define DDE_COOP_TAB_T (s_coop_dyn + s_cc.tab_off)
#define DDE_COOP_TAB_H (s_coop_dyn + s_cc.tab_off + s_cc.cap)

__device__ __forceinline__ double coop_na_real() {
  return __longlong_as_double(0x7FF00000000007A2LL);
}

__device__ __forceinline__ double coop_inf() {
  return __longlong_as_double(0x7ff0000000000000LL);
}

__device__ __forceinline__ const double *coop_entry(unsigned idx) {
  int slot = s_cc.head - (int) (s_cc.count - idx);
  slot = slot < 0 ? slot + s_cc.cap : slot;
  const double *ring = s_cc.ring;
  __builtin_assume(__isGlobal(ring)); /* global, not generic, loads */
  return ring + (size_t) slot * (size_t) s_cc.hl;
}

/* Block-wide: the last committed entry whose time is <= t (>= t backwards),
 * src/dopri.c:742-769.  Thread 0 searches, everybody gets the answer.  On
 * failure ERR_YLAG_FAIL is recorded and false returned. */
__device__ __forceinline__ bool coop_locate(double t, unsigned *idx, double *t_old, double *h) {
  if (s_cc.tab_off >= 0) {
    /* every thread tests the entries it owns: the wanted one starts at or
       before t while its successor (if any) starts after t */
    const double st = s_cc.sign * t;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double tj = s_cc.hint_ts[q];
      if (tj <= st && st < s_cc.hint_tn[q]) {
        *idx = s_cc.hint[q];
        *t_old = s_cc.sign * tj;
        *h = DDE_COOP_TAB_H[s_cc.hint_slot[q]];
        return true; /* the same answer in every thread: no barrier needed */
      }
    }
    const unsigned count = s_cc.count;
    const unsigned cap = (unsigned) s_cc.cap;
    const unsigned used = count < cap ? count : cap;
    const unsigned tail = count - used;
    if (threadIdx.x == 0) {
      s_cc.b_ok = 0;
    }
    __syncthreads();
    for (unsigned k = threadIdx.x; k < used; k += blockDim.x) {
      const unsigned j = tail + k; /* logical index; slot j % cap */
      const double tj = DDE_COOP_TAB_T[j % cap];
      if (tj <= st && (j + 1 == count || !(DDE_COOP_TAB_T[(j + 1) % cap] <= st))) {
        s_cc.b_ok = 1;
        s_cc.b_idx = j;
        s_cc.b_told = s_cc.sign * tj;
        s_cc.b_h = DDE_COOP_TAB_H[j % cap];
      }
    }
    __syncthreads();
    const bool found = s_cc.b_ok != 0;
    *idx = s_cc.b_idx;
    *t_old = s_cc.b_told;
    *h = s_cc.b_h;
    if (threadIdx.x == 0) { /* (everybody read the hints before the first barrier) */
      if (found) {
        const unsigned q = s_cc.hint_next & 3u, j = s_cc.b_idx;
        s_cc.hint[q] = j;
        s_cc.hint_slot[q] = (int) (j % cap);
        s_cc.hint_ts[q] = DDE_COOP_TAB_T[j % cap];
        s_cc.hint_tn[q] = j + 1 < count ? DDE_COOP_TAB_T[(j + 1) % cap] : coop_inf();
        s_cc.hint_next++;
      } else {
        s_cc.error = 1;
        s_cc.code = -6; /* ERR_YLAG_FAIL */
      }
    }
    __syncthreads();
    return found;
  }
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
  /* after the DDE_COOP_NVEC solver vectors */
  return (char *) (dde::s_coop_dyn + 16 * (size_t) dde::s_cc.n);
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

/* ---- tableaux as data: the stage loops below walk lists of (slot, weight)
 * terms, each sum formed left to right over the same terms in the same order
 * as the reference's hand-written expressions (src/dopri_5.c:54-94,
 * src/dopri_853.c:178-246,315-363).  Slot j is the reference's k[j-1]. */
struct CoopTerm {
  int k;
  double a;
};
struct CoopStage {
  double c; /* node; 1.0 means t + h exactly as the reference computes it */
  int dst;  /* slot that receives f(t + c h, stage input) */
  int n_terms;
  CoopTerm t[9];
};

static __constant__ CoopStage c_dp5_stages[6] = {
    {DP5_C2_LIT, 2, 1, {{1, DP5_A21_LIT}}},
    {DP5_C3_LIT, 3, 2, {{1, DP5_A31_LIT}, {2, DP5_A32_LIT}}},
    {DP5_C4_LIT, 4, 3, {{1, DP5_A41_LIT}, {2, DP5_A42_LIT}, {3, DP5_A43_LIT}}},
    {DP5_C5_LIT, 5, 4, {{1, DP5_A51_LIT}, {2, DP5_A52_LIT}, {3, DP5_A53_LIT}, {4, DP5_A54_LIT}}},
    {1.0, 6, 5,
     {{1, DP5_A61_LIT}, {2, DP5_A62_LIT}, {3, DP5_A63_LIT}, {4, DP5_A64_LIT}, {5, DP5_A65_LIT}}},
    {1.0, 2, 5,
     {{1, DP5_A71_LIT}, {3, DP5_A73_LIT}, {4, DP5_A74_LIT}, {5, DP5_A75_LIT}, {6, DP5_A76_LIT}}},
};
static __constant__ CoopTerm c_dp5_err[6] = {{1, DP5_E1_LIT}, {3, DP5_E3_LIT}, {4, DP5_E4_LIT},
                                             {5, DP5_E5_LIT}, {6, DP5_E6_LIT}, {2, DP5_E7_LIT}};
static __constant__ CoopTerm c_dp5_dense[6] = {{1, DP5_D1_LIT}, {3, DP5_D3_LIT}, {4, DP5_D4_LIT},
                                               {5, DP5_D5_LIT}, {6, DP5_D6_LIT}, {2, DP5_D7_LIT}};
static __constant__ CoopStage c_dp8_stages[11] = {
    {DP8_C2_LIT, 2, 1, {{1, DP8_A21_LIT}}},
    {DP8_C3_LIT, 3, 2, {{1, DP8_A31_LIT}, {2, DP8_A32_LIT}}},
    {DP8_C4_LIT, 4, 2, {{1, DP8_A41_LIT}, {3, DP8_A43_LIT}}},
    {DP8_C5_LIT, 5, 3, {{1, DP8_A51_LIT}, {3, DP8_A53_LIT}, {4, DP8_A54_LIT}}},
    {DP8_C6_LIT, 6, 3, {{1, DP8_A61_LIT}, {4, DP8_A64_LIT}, {5, DP8_A65_LIT}}},
    {DP8_C7_LIT, 7, 4, {{1, DP8_A71_LIT}, {4, DP8_A74_LIT}, {5, DP8_A75_LIT}, {6, DP8_A76_LIT}}},
    {DP8_C8_LIT, 8, 5,
     {{1, DP8_A81_LIT}, {4, DP8_A84_LIT}, {5, DP8_A85_LIT}, {6, DP8_A86_LIT}, {7, DP8_A87_LIT}}},
    {DP8_C9_LIT, 9, 6,
     {{1, DP8_A91_LIT}, {4, DP8_A94_LIT}, {5, DP8_A95_LIT}, {6, DP8_A96_LIT}, {7, DP8_A97_LIT},
      {8, DP8_A98_LIT}}},
    {DP8_C10_LIT, 10, 7,
     {{1, DP8_A101_LIT}, {4, DP8_A104_LIT}, {5, DP8_A105_LIT}, {6, DP8_A106_LIT},
      {7, DP8_A107_LIT}, {8, DP8_A108_LIT}, {9, DP8_A109_LIT}}},
    {DP8_C11_LIT, 2, 8,
     {{1, DP8_A111_LIT}, {4, DP8_A114_LIT}, {5, DP8_A115_LIT}, {6, DP8_A116_LIT},
      {7, DP8_A117_LIT}, {8, DP8_A118_LIT}, {9, DP8_A119_LIT}, {10, DP8_A1110_LIT}}},
    {1.0, 3, 9,
     {{1, DP8_A121_LIT}, {4, DP8_A124_LIT}, {5, DP8_A125_LIT}, {6, DP8_A126_LIT},
      {7, DP8_A127_LIT}, {8, DP8_A128_LIT}, {9, DP8_A129_LIT}, {10, DP8_A1210_LIT},
      {2, DP8_A1211_LIT}}},
};
static __constant__ CoopTerm c_dp8_b[8] = {{1, DP8_B1_LIT}, {6, DP8_B6_LIT},   {7, DP8_B7_LIT},
                                           {8, DP8_B8_LIT}, {9, DP8_B9_LIT},   {10, DP8_B10_LIT},
                                           {2, DP8_B11_LIT}, {3, DP8_B12_LIT}};
static __constant__ CoopTerm c_dp8_er[8] = {{1, DP8_ER1_LIT}, {6, DP8_ER6_LIT},   {7, DP8_ER7_LIT},
                                            {8, DP8_ER8_LIT}, {9, DP8_ER9_LIT},   {10, DP8_ER10_LIT},
                                            {2, DP8_ER11_LIT}, {3, DP8_ER12_LIT}};
static __constant__ CoopStage c_dp8_extra[3] = {
    {DP8_C14_LIT, 10, 8,
     {{1, DP8_A141_LIT}, {7, DP8_A147_LIT}, {8, DP8_A148_LIT}, {9, DP8_A149_LIT},
      {10, DP8_A1410_LIT}, {2, DP8_A1411_LIT}, {3, DP8_A1412_LIT}, {4, DP8_A1413_LIT}}},
    {DP8_C15_LIT, 2, 8,
     {{1, DP8_A151_LIT}, {6, DP8_A156_LIT}, {7, DP8_A157_LIT}, {8, DP8_A158_LIT},
      {2, DP8_A1511_LIT}, {3, DP8_A1512_LIT}, {4, DP8_A1513_LIT}, {10, DP8_A1514_LIT}}},
    {DP8_C16_LIT, 3, 8,
     {{1, DP8_A161_LIT}, {6, DP8_A166_LIT}, {7, DP8_A167_LIT}, {8, DP8_A168_LIT},
      {9, DP8_A169_LIT}, {4, DP8_A1613_LIT}, {10, DP8_A1614_LIT}, {2, DP8_A1615_LIT}}},
};
static __constant__ CoopTerm c_dp8_d1[4][8] = {
    {{1, DP8_D41_LIT}, {6, DP8_D46_LIT}, {7, DP8_D47_LIT}, {8, DP8_D48_LIT}, {9, DP8_D49_LIT},
     {10, DP8_D410_LIT}, {2, DP8_D411_LIT}, {3, DP8_D412_LIT}},
    {{1, DP8_D51_LIT}, {6, DP8_D56_LIT}, {7, DP8_D57_LIT}, {8, DP8_D58_LIT}, {9, DP8_D59_LIT},
     {10, DP8_D510_LIT}, {2, DP8_D511_LIT}, {3, DP8_D512_LIT}},
    {{1, DP8_D61_LIT}, {6, DP8_D66_LIT}, {7, DP8_D67_LIT}, {8, DP8_D68_LIT}, {9, DP8_D69_LIT},
     {10, DP8_D610_LIT}, {2, DP8_D611_LIT}, {3, DP8_D612_LIT}},
    {{1, DP8_D71_LIT}, {6, DP8_D76_LIT}, {7, DP8_D77_LIT}, {8, DP8_D78_LIT}, {9, DP8_D79_LIT},
     {10, DP8_D710_LIT}, {2, DP8_D711_LIT}, {3, DP8_D712_LIT}},
};
static __constant__ CoopTerm c_dp8_d2[4][4] = {
    {{4, DP8_D413_LIT}, {10, DP8_D414_LIT}, {2, DP8_D415_LIT}, {3, DP8_D416_LIT}},
    {{4, DP8_D513_LIT}, {10, DP8_D514_LIT}, {2, DP8_D515_LIT}, {3, DP8_D516_LIT}},
    {{4, DP8_D613_LIT}, {10, DP8_D614_LIT}, {2, DP8_D615_LIT}, {3, DP8_D616_LIT}},
    {{4, DP8_D713_LIT}, {10, DP8_D714_LIT}, {2, DP8_D715_LIT}, {3, DP8_D716_LIT}},
};

#include "coop_tableau_dp8.inc"

/* shared-memory vectors of one trajectory, n doubles each */
#define DDE_COOP_NVEC 16 /* y, k1..k10, stage input, 4 x scratch */

/* Sum red[k][i], i = 0 .. n-1, in that order (the reference's loops), from the
 * terms all threads have written.  NSUM <= 2 sums at a time. */
template <int NSUM>
__device__ __forceinline__ void coop_seqsum(int n, const double *red, double *result) {
  __syncthreads();
  if (threadIdx.x == 0) {
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

/* NFIX: n at compile time (the registered capacity TPT * CMAX, when the batch has exactly that
   many equations: every vector of the solver object then sits at an immediate offset from
   one base, and the per-component loops have a fixed trip count), or 0: n from the arguments */
template <class M, int METHOD, int TPT, int CMAX, int NFIX>
__global__ void
__launch_bounds__(TPT, (TPT >= 256 ? DDE_COOP_MINB256 : (TPT >= 128 ? 4 : 8)))
dopri_coop_kernel(const DopriArgs a) {
  constexpr int ORDER = METHOD == 0 ? 5 : 8;
  const int n = NFIX > 0 ? NFIX : a.n;
  const int hl = ORDER * n + 2;
  const int tid = threadIdx.x;
  /* the solver object of src/dopri.h:55-161 for one trajectory, in shared memory */
  double *Y = s_coop_dyn;               /* obj->y */
  double *K = s_coop_dyn + n;           /* obj->k[0..9]: slot j at K + (j - 1) n */
  double *YIN = s_coop_dyn + 11 * (size_t) n; /* obj->y1: the stage input */
  double *RED = s_coop_dyn + 12 * (size_t) n; /* 4 n: norm terms / dense coefficients in the making */
#define KV(j) (K + (size_t) ((j) -1) * n)
#define DDE_EACH(i) for (int i = tid; i < n; i += TPT)

  constexpr double step_factor_min = METHOD == 0 ? 0.2 : 0.333;
  constexpr double step_factor_max = METHOD == 0 ? 10.0 : 6.0;
  constexpr double step_beta = METHOD == 0 ? 0.04 : 0.0;
  constexpr double step_factor_safe = 0.9;
  const double step_constant = METHOD == 0 ? 0.2 - step_beta * 0.75 : 0.125 - step_beta * 0.2;
  const double sign = a.sign, atol = a.atol, rtol = a.rtol;
  const int n_times = a.n_times;
  const int nt = a.return_initial ? n_times : n_times - 1;
  const double t_end = a.times[n_times - 1];

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
      for (int q = 0; q < 4; ++q) {
        s_cc.hint[q] = 0xffffffffu;
        s_cc.hint_ts[q] = coop_inf();
        s_cc.hint_tn[q] = coop_inf();
      }
      s_cc.hint_next = 0;
      /* the table of step times follows the model's scratch area */
      double *tab = (double *) ((char *) (s_coop_dyn + DDE_COOP_NVEC * (size_t) n) +
                                DDE_COOP_SCRATCH_BYTES);
      const bool have_tab = a.n_history > 0 && a.n_history <= DDE_COOP_MAX_TABLE;
      s_cc.tab_off = have_tab ? (int) (tab - s_coop_dyn) : -1;
    }
    DDE_EACH(i) { Y[i] = y0[i]; }
    __syncthreads();
    if (s_cc.tab_off >= 0 && s_cc.count > 0) { /* restart: the times of the surviving entries */
      const unsigned cap = (unsigned) a.n_history;
      const unsigned used = s_cc.count < cap ? s_cc.count : cap;
      for (unsigned k = tid; k < used; k += TPT) {
        const unsigned slot = (s_cc.count - used + k) % cap;
        const double *e = ring + (size_t) slot * (size_t) hl;
        DDE_COOP_TAB_T[slot] = sign * e[hl - 2];
        DDE_COOP_TAB_H[slot] = e[hl - 1];
      }
      __syncthreads();
    }

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
      DDE_EACH(i) { yo[i] = Y[i]; }
      yo += n;
    }

    /* one RHS evaluation by the whole block, src/dopri.c:831-835: the stage
       vector and the derivative slot are shared-memory vectors, so the model
       reads and writes them in place */
#define DDE_COOP_EVAL(tt, src, dst)                \
  do {                                             \
    __syncthreads();                               \
    M::deriv((size_t) n, (tt), (src), (dst), p);   \
    __syncthreads();                               \
    ++n_eval;                                      \
  } while (0)
    /* stage input y + h * (sum of the stage's terms) into `into`, then its evaluation */
#define DDE_COOP_STAGE(st, first, into)                                              \
  do {                                                                               \
    const CoopStage &st_ = (st);                                                     \
    DDE_EACH(i) {                                                                    \
      if (first) { /* written h * A21 * k1, i.e. (h A21) k1 */                       \
        (into)[i] = Y[i] + h * st_.t[0].a * KV(1)[i];                                \
      } else {                                                                       \
        double acc = st_.t[0].a * KV(st_.t[0].k)[i];                                 \
        for (int q = 1; q < st_.n_terms; ++q) {                                      \
          acc = acc + st_.t[q].a * KV(st_.t[q].k)[i];                                \
        }                                                                            \
        (into)[i] = Y[i] + h * acc;                                                  \
      }                                                                              \
    }                                                                                \
    DDE_COOP_EVAL(st_.c == 1.0 ? t + h : t + st_.c * h, (into), KV(st_.dst));        \
  } while (0)

    bool done = false;
    double h = 0.0;
    /* k1 = f(t0, y0), finiteness check, dopri_h_init: src/dopri.c:329-339,527-572 */
    DDE_COOP_EVAL(t, Y, KV(1));
    {
      DDE_EACH(i) { RED[i] = isfinite(KV(1)[i]) ? 0.0 : 1.0; }
      double nbad;
      coop_seqsum<1>(n, RED, &nbad);
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
        DDE_EACH(i) {
          const double sk = atol + rtol * fabs(Y[i]);
          RED[i] = csq(KV(1)[i] / sk);
          RED[n + i] = csq(Y[i] / sk);
        }
        double norms[2];
        coop_seqsum<2>(n, RED, norms);
        const double norm_f = norms[0], norm_y = norms[1];
        h = (norm_f <= 1e-10 || norm_y <= 1e-10) ? 1e-6 : sqrt(norm_y / norm_f) * 0.01;
        h = copysign(fmin(h, a.step_size_max), sign);
        /* the reference uses k[1] for f1 and k[2] for the Euler point */
        DDE_EACH(i) { KV(3)[i] = Y[i] + h * KV(1)[i]; }
        DDE_COOP_EVAL(t + h, KV(3), KV(2));
        DDE_EACH(i) {
          const double sk = atol + rtol * fabs(Y[i]);
          RED[i] = csq((KV(2)[i] - KV(1)[i]) / sk);
        }
        double der2;
        coop_seqsum<1>(n, RED, &der2);
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

      /* ---- dopri_step + dopri_error ---- */
      double err;
      if (METHOD == 0) {
        /* dopri5_step, src/dopri_5.c:42-95: stage 5 is formed in `ysti` (slot 7) */
#pragma unroll 1
        for (int j = 0; j < 6; ++j) {
          DDE_COOP_STAGE(c_dp5_stages[j], j == 0, j == 4 ? KV(7) : YIN);
        }
        DDE_EACH(i) {
          /* the fifth dense coefficient, formed on every attempt before k4 is
             overwritten (src/dopri_5.c:84-94) */
          double d = c_dp5_dense[0].a * KV(c_dp5_dense[0].k)[i];
          double e = c_dp5_err[0].a * KV(c_dp5_err[0].k)[i];
#pragma unroll
          for (int q = 1; q < 6; ++q) {
            d = d + c_dp5_dense[q].a * KV(c_dp5_dense[q].k)[i];
            e = e + c_dp5_err[q].a * KV(c_dp5_err[q].k)[i];
          }
          RED[3 * (size_t) n + i] = h * d;
          KV(4)[i] = h * e;
          const double sk = atol + rtol * fmax(fabs(Y[i]), fabs(YIN[i]));
          RED[i] = csq(KV(4)[i] / sk);
        }
        if (s_cc.error) { /* a lag look-up failed during dopri_step, src/dopri.c:387-389 */
          break;
        }
        coop_seqsum<1>(n, RED, &err);
        err = sqrt(err / (double) n);
      } else {
        /* dopri853_step, src/dopri_853.c:159-247 */
        /* the stage sums as straight-line code (coop_tableau_dp8.inc, generated from the
           tables above): a term is an LDS off a loop-invariant base and two FP64
           instructions, where walking the tables cost an LDC pair, a multiplication by n and
           an address per term (27 % of the kernel's instructions in round 1) */
#pragma unroll 1
        for (int j = 0; j < 11; ++j) {
          if (j == 0) { /* written h * A21 * k1, i.e. (h A21) k1 */
            DDE_EACH(i) { YIN[i] = Y[i] + h * DP8_A21 * KV(1)[i]; }
          } else {
            DDE_COOP_DP8_STAGE_INPUT(j)
          }
          DDE_COOP_EVAL(DDE_COOP_DP8_STAGE_TIME(j), YIN, KV(DDE_COOP_DP8_STAGE_DST(j)));
        }
        DDE_EACH(i) {
          const double b = DDE_COOP_DP8_B;
          KV(4)[i] = b;
          KV(5)[i] = Y[i] + h * b;
          /* dopri853_error, src/dopri_853.c:249-283 */
          const double sk = atol + rtol * fmax(fabs(Y[i]), fabs(KV(5)[i]));
          double erri = KV(4)[i] - DP8_BHH1 * KV(1)[i] - DP8_BHH2 * KV(9)[i] - DP8_BHH3 * KV(3)[i];
          RED[n + i] = csq(erri / sk);
          erri = DDE_COOP_DP8_ER;
          RED[i] = csq(erri / sk);
        }
        if (s_cc.error) {
          break;
        }
        double sums[2];
        coop_seqsum<2>(n, RED, sums);
        double deno = sums[0] + 0.01 * sums[1];
        deno = deno < 0 ? 1.0 : deno;
        err = sign * sums[0] * sqrt(1.0 / ((double) n * deno));
      }

      /* dopri_h_new, src/dopri.c:516-525 */
      const double fac11 = dde_pow(err, step_constant);
      const double facb = METHOD == 0 ? dde_pow(fac_old, step_beta) : 1.0;
      double fac = fac11 / facb;
      fac = fmax(1.0 / step_factor_max, fmin(1.0 / step_factor_min, fac / step_factor_safe));
      double h_new = h / fac;
      if (!((err <= 1) || force_this_step)) {
        /* rejected, src/dopri.c:495-509 */
        h_new = h / fmin(1 / step_factor_min, fac11 / step_factor_safe);
        reject = true;
        if (n_accept >= 1) {
          ++n_reject;
        }
        last = false;
        stop = false;
        h = h_new;
        continue;
      }

      /* ---- accepted, src/dopri.c:396-494 ---- */
      fac_old = fmax(err, 1e-4);
      ++n_accept;
      if (METHOD == 1) { /* the redundant evaluation at the OLD time, src/dopri.c:400-403 */
        DDE_COOP_EVAL(t, KV(5), KV(4));
      }
      /* dopri_test_stiff, src/dopri.c:668-693 */
      if (a.stiff_check != 0 &&
          (stiff_n_stiff > 0 || (unsigned long long) n_accept % a.stiff_check == 0)) {
        DDE_EACH(i) {
          if (METHOD == 0) { /* src/dopri_5.c:129-140 */
            RED[i] = csq(KV(2)[i] - KV(6)[i]);
            RED[n + i] = csq(YIN[i] - KV(7)[i]);
          } else { /* src/dopri_853.c:382-394 */
            RED[i] = csq(KV(4)[i] - KV(3)[i]);
            RED[n + i] = csq(KV(5)[i] - YIN[i]);
          }
        }
        double sums[2];
        coop_seqsum<2>(n, RED, sums);
        const bool stiff =
            sums[1] > 0 && fabs(h) * sqrt(sums[0] / sums[1]) > (METHOD == 0 ? 3.25 : 6.1);
        bool gave_up = false;
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
          if (tid == 0) {
            s_cc.error = 1;
            s_cc.code = -7; /* ERR_STIFF */
          }
          break;
        }
      }
      if (METHOD == 1) {
        /* dopri853_save_history, src/dopri_853.c:285-367: one more evaluation,
           the first half of coefficients 4..7 (before slots 10, 2, 3 are
           reused), three more stages */
        DDE_COOP_EVAL(t + h, KV(5), KV(4));
        DDE_EACH(i) {
          RED[i] = DDE_COOP_DP8_D1_0;
          RED[(size_t) n + i] = DDE_COOP_DP8_D1_1;
          RED[2 * (size_t) n + i] = DDE_COOP_DP8_D1_2;
          RED[3 * (size_t) n + i] = DDE_COOP_DP8_D1_3;
        }
#pragma unroll 1
        for (int j = 0; j < 3; ++j) {
          DDE_COOP_DP8_EXTRA_INPUT(j)
          DDE_COOP_EVAL(DDE_COOP_DP8_EXTRA_TIME(j), YIN, KV(DDE_COOP_DP8_EXTRA_DST(j)));
        }
      }

      /* dense coefficients of the step, per component: into the ring's head
         slot (ring_buffer_head_advance, src/dopri.c:460) and, evaluated at the
         requested times the step has passed, into the output rows
         (src/dopri.c:437-456) */
      const double t_old = t;
      t += h;
      int rows = 0; /* how many requested times this step emits */
      while (times_idx + rows < n_times &&
             (last || sign * a.times[times_idx + rows] <= sign * t)) {
        ++rows;
      }
      double *dst = a.n_history > 0 ? ring + (size_t) s_cc.head * (size_t) hl : nullptr;
      DDE_EACH(i) {
        double c[ORDER];
        if (METHOD == 0) { /* dopri5_save_history, src/dopri_5.c:106-118; YIN is y1 */
          const double ydiff = YIN[i] - Y[i];
          const double bspl = h * KV(1)[i] - ydiff;
          c[0] = Y[i];
          c[1] = ydiff;
          c[2] = bspl;
          c[3] = -h * KV(2)[i] + ydiff - bspl;
          c[4] = RED[3 * (size_t) n + i];
        } else { /* src/dopri_853.c:306-314,350-363 */
          const double ydiff = KV(5)[i] - Y[i];
          const double bspl = h * KV(1)[i] - ydiff;
          c[0] = Y[i];
          c[1] = ydiff;
          c[2] = bspl;
          c[3] = ydiff - h * KV(4)[i] - bspl;
          c[4] = h * DDE_COOP_DP8_D2_0(RED[i]);
          c[5] = h * DDE_COOP_DP8_D2_1(RED[(size_t) n + i]);
          c[6] = h * DDE_COOP_DP8_D2_2(RED[2 * (size_t) n + i]);
          c[7] = h * DDE_COOP_DP8_D2_3(RED[3 * (size_t) n + i]);
        }
        if (dst != nullptr) {
#pragma unroll
          for (int q = 0; q < ORDER; ++q) {
            dst[(size_t) q * n + i] = c[q];
          }
        }
        for (int r = 0; r < rows; ++r) {
          const double theta = (a.times[times_idx + r] - t_old) / h;
          const double theta1 = 1 - theta;
          yo[(size_t) r * n + i] = coop_interp(c, ORDER, 1, theta, theta1);
        }
      }
      yo += (size_t) rows * n;
      times_idx += rows;
      __syncthreads(); /* every thread is done with y, k1 and the context */
      /* y <- y1 / k5, k1 <- k2 / k4: src/dopri.c:414-424 */
      DDE_EACH(i) {
        Y[i] = METHOD == 0 ? YIN[i] : KV(5)[i];
        KV(1)[i] = METHOD == 0 ? KV(2)[i] : KV(4)[i];
      }
      if (a.n_history > 0 && tid == 0) {
        dst[ORDER * n] = t_old;
        dst[ORDER * n + 1] = h;
        if (s_cc.tab_off >= 0) {
          DDE_COOP_TAB_T[s_cc.head] = sign * t_old;
          DDE_COOP_TAB_H[s_cc.head] = h;
        }
        const unsigned newest = s_cc.count;
        s_cc.head = s_cc.head + 1 == a.n_history ? 0 : s_cc.head + 1;
        s_cc.count = newest + 1;
        if (s_cc.tab_off >= 0) {
          /* the hinted entries: the newest one has a successor now, the oldest may be gone */
          const unsigned gone = newest + 1 > (unsigned) a.n_history
                                    ? newest + 1 - (unsigned) a.n_history : 0u;
          for (int q = 0; q < 4; ++q) {
            const unsigned j = s_cc.hint[q];
            if (j != 0xffffffffu) {
              if (j + 1 == newest) {
                s_cc.hint_tn[q] = sign * t_old;
              }
              if (j < gone) {
                s_cc.hint[q] = 0xffffffffu;
                s_cc.hint_ts[q] = coop_inf();
              }
            }
          }
        }
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
      DDE_EACH(i) { a.y_final[traj * n + i] = Y[i]; }
    }
    if (a.chunk_done != nullptr) { /* every thread of the block wrote result rows */
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        chunk_retire(a, traj);
      }
    }
#undef DDE_COOP_EVAL
#undef DDE_COOP_STAGE
  }
#undef KV
#undef DDE_EACH
}

} /* namespace dde */

#endif /* DDE_DOPRI_COOP_KERNEL_CUH */
