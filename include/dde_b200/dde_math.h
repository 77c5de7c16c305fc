/* dde_math.h -- pow() and exp() that are bit-identical on host and device.
 *
 * The reference's only non-IEEE-exact arithmetic is libm `pow` in the step
 * controller (/root/reference/src/dopri.c:500,517,521,569) and whatever
 * libm calls a model makes.  One ulp of difference there changes accepted /
 * rejected step counts, so the device cannot use CUDA's pow.  This file
 * restates the algorithm the reference links to on this platform: glibc
 * 2.39 x86-64 `pow`/`exp`, FMA ifunc variant (third-party, not under
 * /root/reference; upstream: ARM optimized-routines math/pow.c, math/exp.c
 * by Szabolcs Nagy, also glibc sysdeps/ieee754/dbl-64/e_pow.c, e_exp.c).
 *
 * Published algorithm:
 *   log(x)  = k ln2 + log(c) + log1p(z/c - 1), x = 2^k z, c from a 128-entry
 *             table, result as a double-double (hi, lo);
 *   x^y     = exp(y * (hi + lo)) with the product kept as (ehi, elo);
 *   exp(s)  = 2^(k/128) * (1 + tail + p(r)),  s = k ln2/128 + r.
 * The operation ORDER and the placement of fused multiply-adds below follow
 * the machine code of this image's libm (__pow_fma / __exp_fma), because
 * rounding depends on both; every fma() here is a single IEEE fused
 * operation on both host (-mfma) and device (DFMA), every other operation is
 * an individually rounded IEEE add/mul, so the result is the same bits
 * everywhere.  Compile device code with -fmad=false and host code with
 * -ffp-contract=off so the compiler adds no contractions of its own.
 *
 * Licence of the upstream algorithm: MIT OR Apache-2.0 WITH LLVM-exception (ARM
 * optimized-routines); the notice is carried in THIRD_PARTY_NOTICES.md.
 *
 * Pinned by tests/test_math_cpu.py against libm on >= 10^7 inputs per domain
 * plus the special-value lattice; errno / FP exception flags are not kept.
 */
#ifndef DDE_MATH_H
#define DDE_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "dde_math_tables.h"

#if defined(__CUDACC__)
#define DDE_HD __host__ __device__ __forceinline__
#define DDE_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define DDE_HD static inline
#define DDE_HD_NOINLINE static
#endif

/* The tables live in global memory on the device (6 KB, L1/L2 resident:
 * lanes index them divergently, which a __constant__ bank would serialise)
 * and in .rodata on the host. */
#if defined(__CUDACC__)
/* device copies: one entry = one 32-byte (log) / 16-byte (exp) aligned record,
   fetched with one vector load (plus one scalar load for the log tail): a
   divergent gather pays per load instruction and distinct line */
struct __align__(32) DdePowLogEntry {
  double2 invc_logc;
  double logctail;
  double pad;
};
static __device__ const DdePowLogEntry dde_pow_log_tab_d[128] = {DDE_POW_LOG_TABLE4};
static __device__ const ulonglong2 dde_exp_tab_d[128] = {DDE_EXP_TABLE};
#endif
static const double dde_pow_log_tab_h[3 * 128] = {DDE_POW_LOG_TABLE};
static const unsigned long long dde_exp_tab_h[2 * 128] = {DDE_EXP_TABLE};

#if defined(__CUDA_ARCH__)
#define DDE_FMA(a, b, c) fma((a), (b), (c))
#else
#define DDE_POW_LOG_TAB(i) dde_pow_log_tab_h[i]
#define DDE_EXP_TAB(i) ((uint64_t) dde_exp_tab_h[i])
#define DDE_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

DDE_HD uint64_t dde_asuint64(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t) __double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, sizeof u);
  return u;
#endif
}

DDE_HD double dde_asdouble(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long) u);
#else
  double x;
  memcpy(&x, &u, sizeof x);
  return x;
#endif
}

DDE_HD uint32_t dde_top12(double x) { return (uint32_t) (dde_asuint64(x) >> 52); }

/* Result scaling when the exponent of `scale` may have left the normal
 * range (|s| large): handles overflow and the subnormal range with a single
 * final rounding. */
DDE_HD_NOINLINE double dde_exp_specialcase(double tmp, uint64_t sbits, uint64_t ki) {
  double scale, y;
  if ((ki & 0x80000000ULL) == 0) {
    /* k > 0: the exponent of scale might have overflowed by <= 460 */
    sbits -= 1009ULL << 52;
    scale = dde_asdouble(sbits);
    y = DDE_FMA(scale, tmp, scale);
    return 0x1p1009 * y;
  }
  /* k < 0: care in the subnormal range; sbits is a signed scale */
  sbits += 1022ULL << 52;
  scale = dde_asdouble(sbits);
  const double st = scale * tmp;
  y = scale + st;
  if (fabs(y) < 1.0) {
    double one = 1.0;
    if (y < 0.0) {
      one = -1.0;
    }
    double lo = (scale - y) + st;
    const double hi = one + y;
    lo = ((one - hi) + y) + lo;
    y = (lo + hi) - one;
    if (y == 0.0) {
      y = dde_asdouble(sbits & 0x8000000000000000ULL);
    }
  }
  return 0x1p-1022 * y;
}

/* exp(x + xtail) * (-1)^(sign_bias != 0); |xtail| tiny relative to x. */
DDE_HD double dde_exp_inline(double x, double xtail, uint32_t sign_bias, int has_tail) {
  uint32_t abstop = dde_top12(x) & 0x7ff;
  if (abstop - 0x3c9u >= 0x3fu) { /* |x| < 2^-54 or |x| >= 512 */
    if (abstop - 0x3c9u >= 0x80000000u) {
      /* tiny: exp(x) rounds like 1 + x */
      const double one = 1.0 + x;
      return sign_bias ? -one : one;
    }
    if (abstop >= 0x409u) { /* |x| >= 1024; inf and nan handled by callers */
      if (dde_asuint64(x) >> 63) {
        return sign_bias ? -0.0 : 0.0;
      }
      return sign_bias ? -HUGE_VAL : HUGE_VAL;
    }
    abstop = 0; /* large |x|: finish in dde_exp_specialcase */
  }
  /* x = k ln2/128 + r, |r| <= ln2/256 */
  double kd = DDE_FMA(x, DDE_EXP_INVLN2N, DDE_EXP_SHIFT);
  const uint64_t ki = dde_asuint64(kd);
  kd = kd - DDE_EXP_SHIFT;
  double r = DDE_FMA(kd, DDE_EXP_NEGLN2HIN, x);
  r = DDE_FMA(kd, DDE_EXP_NEGLN2LON, r);
  if (has_tail) {
    r = xtail + r;
  }
  /* 2^(k/128) ~= scale * (1 + tail) */
  const uint64_t idx = 2 * (ki % 128);
  const uint64_t top = (ki + sign_bias) << 45;
#if defined(__CUDA_ARCH__)
  const ulonglong2 entry = __ldg(&dde_exp_tab_d[ki % 128]);
  const double tail = dde_asdouble((uint64_t) entry.x);
  const uint64_t sbits = (uint64_t) entry.y + top;
  (void) idx;
#else
  const double tail = dde_asdouble(DDE_EXP_TAB(idx));
  const uint64_t sbits = DDE_EXP_TAB(idx + 1) + top;
#endif
  const double r2 = r * r;
  const double p23 = DDE_FMA(r, DDE_EXP_C3, DDE_EXP_C2);
  const double p45 = DDE_FMA(r, DDE_EXP_C5, DDE_EXP_C4);
  double tmp = DDE_FMA(p23, r2, r + tail);
  tmp = DDE_FMA(r2 * r2, p45, tmp);
  if (abstop == 0) {
    return dde_exp_specialcase(tmp, sbits, ki);
  }
  const double scale = dde_asdouble(sbits);
  return DDE_FMA(scale, tmp, scale);
}

DDE_HD double dde_exp(double x) {
  const uint32_t abstop = dde_top12(x) & 0x7ff;
  if (abstop >= 0x409u) {
    if (dde_asuint64(x) == 0xfff0000000000000ULL) {
      return 0.0; /* exp(-inf) */
    }
    if (abstop == 0x7ff) {
      return 1.0 + x; /* +inf, nan */
    }
  }
  return dde_exp_inline(x, 0.0, 0, 0);
}

/* log(x) for the normalised bit pattern ix as hi + *tail, in two parts around the table
 * read so that the caller chooses where the table is (global memory, .rodata, or a
 * kernel's copy in shared memory): x = 2^k z, z in [OFF, 2 OFF); the range is cut into
 * 128 subintervals, and this is the subinterval's index. */
DDE_HD int dde_pow_log_index(uint64_t ix) {
  const uint64_t tmp = ix - 0x3fe6955500000000ULL;
  return (int) ((tmp >> 45) % 128);
}

/* ... and this the rest, given entry i = {invc, logc, logctail}. */
DDE_HD double dde_pow_log_finish(uint64_t ix, double invc, double logc, double logctail,
                                 double *tail) {
  const uint64_t tmp = ix - 0x3fe6955500000000ULL;
  const int k = (int) ((int64_t) tmp >> 52);
  const uint64_t iz = ix - (tmp & 0xfffULL << 52);
  const double z = dde_asdouble(iz);
  const double kd = (double) k;

  /* r = z/c - 1 is exactly representable */
  const double r = DDE_FMA(z, invc, -1.0);
  /* k ln2 + log(c) + r */
  const double t1 = DDE_FMA(kd, DDE_POW_LN2HI, logc);
  const double t2 = t1 + r;
  const double lo1 = DDE_FMA(kd, DDE_POW_LN2LO, logctail);
  const double lo2 = (t1 - t2) + r;
  /* + A0 r^2, A0 = -1/2, kept exact with an fma residual */
  const double ar = DDE_POW_A0 * r;
  const double ar2 = r * ar;
  const double ar3 = r * ar2;
  const double hi = t2 + ar2;
  const double lo3 = DDE_FMA(ar, r, -ar2);
  const double lo4 = (t2 - hi) + ar2;
  /* p = log1p(r) - r - A0 r^2 = ar3 * q(r) */
  const double q12 = DDE_FMA(r, DDE_POW_A2, DDE_POW_A1);
  const double q34 = DDE_FMA(r, DDE_POW_A4, DDE_POW_A3);
  const double q56 = DDE_FMA(r, DDE_POW_A6, DDE_POW_A5);
  double q = DDE_FMA(q56, ar2, q34);
  q = DDE_FMA(ar2, q, q12);
  double lo = lo1 + lo2;
  lo = lo + lo3;
  lo = lo + lo4;
  lo = DDE_FMA(q, ar3, lo);
  const double y = hi + lo;
  *tail = (hi - y) + lo;
  return y;
}

/* the default table */
DDE_HD void dde_pow_log_entry(int i, double *invc, double *logc, double *logctail) {
#if defined(__CUDA_ARCH__)
  const double2 invc_logc = __ldg(&dde_pow_log_tab_d[i].invc_logc);
  *invc = invc_logc.x;
  *logc = invc_logc.y;
  *logctail = __ldg(&dde_pow_log_tab_d[i].logctail);
#else
  *invc = DDE_POW_LOG_TAB(3 * i);
  *logc = DDE_POW_LOG_TAB(3 * i + 1);
  *logctail = DDE_POW_LOG_TAB(3 * i + 2);
#endif
}

DDE_HD double dde_pow_log_inline(uint64_t ix, double *tail) {
  double invc, logc, logctail;
  dde_pow_log_entry(dde_pow_log_index(ix), &invc, &logc, &logctail);
  return dde_pow_log_finish(ix, invc, logc, logctail, tail);
}

/* 0 = not an integer, 1 = odd integer, 2 = even integer */
DDE_HD int dde_pow_checkint(uint64_t iy) {
  const int e = (int) ((iy >> 52) & 0x7ff);
  if (e < 0x3ff) {
    return 0;
  }
  if (e > 0x3ff + 52) {
    return 2;
  }
  if (iy & ((1ULL << (0x3ff + 52 - e)) - 1)) {
    return 0;
  }
  if (iy & (1ULL << (0x3ff + 52 - e))) {
    return 1;
  }
  return 2;
}

DDE_HD int dde_pow_zeroinfnan(uint64_t i) {
  return 2 * i - 1 >= 2 * 0x7ff0000000000000ULL - 1;
}

/* x^y for the normalised bit pattern ix of |x| (the common path): log in
 * double-double, product, exp. */
DDE_HD double dde_pow_main(uint64_t ix, double y, uint32_t sign_bias) {
  double lo;
  const double hi = dde_pow_log_inline(ix, &lo);
  const double ehi = y * hi;
  const double elo = DDE_FMA(y, lo, DDE_FMA(hi, y, -ehi));
  return dde_exp_inline(ehi, elo, sign_bias, 1);
}

/* Everything outside "x positive normal, 2^-65 <= |y| < 2^63": zeros,
 * infinities, NaNs, negative and subnormal bases, tiny and huge exponents.
 * Out of line (and by value, so the caller keeps nothing on the stack). */
DDE_HD_NOINLINE double dde_pow_slow(double x, double y) {
  uint32_t sign_bias = 0;
  uint64_t ix = dde_asuint64(x);
  const uint64_t iy = dde_asuint64(y);
  uint32_t topx = dde_top12(x);
  const uint32_t topy = dde_top12(y);
  const uint64_t one_bits = 0x3ff0000000000000ULL;
  const uint64_t inf_bits = 0x7ff0000000000000ULL;
  if (dde_pow_zeroinfnan(iy)) {
    if (2 * iy == 0) {
      return 1.0;
    }
    if (ix == one_bits) {
      return 1.0;
    }
    if (2 * ix > 2 * inf_bits || 2 * iy > 2 * inf_bits) {
      return x + y;
    }
    if (2 * ix == 2 * one_bits) {
      return 1.0;
    }
    if ((2 * ix < 2 * one_bits) == !(iy >> 63)) {
      return 0.0; /* |x| < 1 && y == inf, or |x| > 1 && y == -inf */
    }
    return y * y;
  }
  if (dde_pow_zeroinfnan(ix)) {
    double x2 = x * x;
    if ((ix >> 63) && dde_pow_checkint(iy) == 1) {
      x2 = -x2;
    }
    return (iy >> 63) ? 1 / x2 : x2;
  }
  /* x and y are non-zero finite */
  if (ix >> 63) {
    /* finite x < 0 */
    const int yint = dde_pow_checkint(iy);
    if (yint == 0) {
      return (x - x) / (x - x); /* invalid: NaN */
    }
    if (yint == 1) {
      sign_bias = 0x800u << 7;
    }
    ix &= 0x7fffffffffffffffULL;
    topx &= 0x7ff;
  }
  if ((topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
    /* sign_bias == 0 here because y is not odd */
    if (ix == one_bits) {
      return 1.0;
    }
    if ((topy & 0x7ff) < 0x3be) {
      /* |y| < 2^-65: x^y ~= 1 + y log(x) */
      return ix > one_bits ? 1.0 + y : 1.0 - y;
    }
    return ((ix > one_bits) == (topy < 0x800)) ? HUGE_VAL : 0.0;
  }
  if (topx == 0) {
    /* normalise subnormal x so the exponent becomes negative */
    ix = dde_asuint64(x * 0x1p52);
    ix &= 0x7fffffffffffffffULL;
    ix -= 52ULL << 52;
  }
  return dde_pow_main(ix, y, sign_bias);
}

/* The common path of dde_pow_main without a branch: the value when x is
 * positive normal and the exponent of the final exp() is ordinary, something
 * harmless (never a trap, never an out-of-range table index) otherwise;
 * *exp_rare tells which.  In three parts around the two table reads
 * (dde_pow_log_index/_finish above; the reduction of the exponent and the polynomial
 * below), which dde_pow_common strings together for the default tables and
 * dde_pow_shared for a kernel's copy in shared memory. */
struct DdeExpReduced {
  uint64_t ki; /* entry ki % 128; scale exponent ki << 45 */
  double r;
};

DDE_HD struct DdeExpReduced dde_pow_exp_reduce(double y, double hi, double lo, int *exp_rare) {
  const double ehi = y * hi;
  const double elo = DDE_FMA(y, lo, DDE_FMA(hi, y, -ehi));
  const uint32_t abstop = dde_top12(ehi) & 0x7ff;
  *exp_rare = abstop - 0x3c9u >= 0x3fu;
  /* dde_exp_inline(ehi, elo, 0, 1), its common path */
  double kd = DDE_FMA(ehi, DDE_EXP_INVLN2N, DDE_EXP_SHIFT);
  struct DdeExpReduced red;
  red.ki = dde_asuint64(kd);
  kd = kd - DDE_EXP_SHIFT;
  double r = DDE_FMA(kd, DDE_EXP_NEGLN2HIN, ehi);
  r = DDE_FMA(kd, DDE_EXP_NEGLN2LON, r);
  red.r = elo + r;
  return red;
}

/* given entry ki % 128 = {tail, scale bits} */
DDE_HD double dde_pow_exp_finish(struct DdeExpReduced red, double tail, uint64_t entry_sbits) {
  const uint64_t top = red.ki << 45;
  const uint64_t sbits = entry_sbits + top;
  const double r = red.r;
  const double r2 = r * r;
  const double p23 = DDE_FMA(r, DDE_EXP_C3, DDE_EXP_C2);
  const double p45 = DDE_FMA(r, DDE_EXP_C5, DDE_EXP_C4);
  double tmp = DDE_FMA(p23, r2, r + tail);
  tmp = DDE_FMA(r2 * r2, p45, tmp);
  const double scale = dde_asdouble(sbits);
  return DDE_FMA(scale, tmp, scale);
}

DDE_HD double dde_pow_common(uint64_t ix, double y, int *exp_rare) {
  double lo;
  const double hi = dde_pow_log_inline(ix, &lo);
  const struct DdeExpReduced red = dde_pow_exp_reduce(y, hi, lo, exp_rare);
#if defined(__CUDA_ARCH__)
  const ulonglong2 entry = __ldg(&dde_exp_tab_d[red.ki % 128]);
  return dde_pow_exp_finish(red, dde_asdouble((uint64_t) entry.x), (uint64_t) entry.y);
#else
  const uint64_t idx = 2 * (red.ki % 128);
  return dde_pow_exp_finish(red, dde_asdouble(DDE_EXP_TAB(idx)), DDE_EXP_TAB(idx + 1));
#endif
}

/* Everything dde_pow_common does not cover.  `common` is not used: taking it
 * keeps the common path's arithmetic ahead of the (one) branch in dde_pow, so
 * that path is a single basic block a scheduler can interleave with whatever
 * else the caller has in flight. */
DDE_HD_NOINLINE double dde_pow_rare(double x, double y, double common) {
  (void) common;
  const uint32_t topx = dde_top12(x);
  const uint32_t topy = dde_top12(y);
  if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
    return dde_pow_slow(x, y);
  }
  return dde_pow_main(dde_asuint64(x), y, 0);
}

DDE_HD double dde_pow(double x, double y) {
  const uint32_t topx = dde_top12(x);
  const uint32_t topy = dde_top12(y);
  const int rare =
      (topx - 0x001u >= 0x7ffu - 0x001u) | ((topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu);
  int exp_rare;
  const double common = dde_pow_common(dde_asuint64(x), y, &exp_rare);
  if (rare | exp_rare) {
    return dde_pow_rare(x, y, common);
  }
  return common;
}


/* x^y1 and x^y2 at once.  pow(x, y) is exp(y log x) with log x a double-double that does
 * not depend on y, so two powers of one base share it -- the table read, the polynomial,
 * half of the work -- and each is, bit for bit, what dde_pow(x, y) returns alone.  The
 * dopri5 step-size controller wants err^0.17 now and max(err, 1e-4)^0.04 at the next step
 * (/root/reference/src/dopri.c:517,521: fac_old is the accepted step's error). */
struct DdePow2 {
  double a, b;
};

DDE_HD struct DdePow2 dde_pow2(double x, double y1, double y2) {
  const uint32_t topx = dde_top12(x);
  const int rare_x = topx - 0x001u >= 0x7ffu - 0x001u;
  const int rare1 = rare_x | ((dde_top12(y1) & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu);
  const int rare2 = rare_x | ((dde_top12(y2) & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu);
  double lo;
  const double hi = dde_pow_log_inline(dde_asuint64(x), &lo);
  int er1, er2;
  const struct DdeExpReduced red1 = dde_pow_exp_reduce(y1, hi, lo, &er1);
  const struct DdeExpReduced red2 = dde_pow_exp_reduce(y2, hi, lo, &er2);
#if defined(__CUDA_ARCH__)
  const ulonglong2 e1 = __ldg(&dde_exp_tab_d[red1.ki % 128]);
  const ulonglong2 e2 = __ldg(&dde_exp_tab_d[red2.ki % 128]);
  const double c1 = dde_pow_exp_finish(red1, dde_asdouble((uint64_t) e1.x), (uint64_t) e1.y);
  const double c2 = dde_pow_exp_finish(red2, dde_asdouble((uint64_t) e2.x), (uint64_t) e2.y);
#else
  const uint64_t i1 = 2 * (red1.ki % 128), i2 = 2 * (red2.ki % 128);
  const double c1 = dde_pow_exp_finish(red1, dde_asdouble(DDE_EXP_TAB(i1)), DDE_EXP_TAB(i1 + 1));
  const double c2 = dde_pow_exp_finish(red2, dde_asdouble(DDE_EXP_TAB(i2)), DDE_EXP_TAB(i2 + 1));
#endif
  struct DdePow2 out;
  out.a = (rare1 | er1) ? dde_pow_rare(x, y1, c1) : c1;
  out.b = (rare2 | er2) ? dde_pow_rare(x, y2, c2) : c2;
  return out;
}

#if defined(__CUDACC__) || defined(DDE_HOST_EMU)
/* ---- the same pow with the two tables in a kernel's shared memory --------------------
 * at shared-window address `base`: {invc, logc} x 128, then logctail x 128, then
 * {tail, scale bits} x 128 (DDE_TAB_SHARED_BYTES; filled by dde_tab_shared_fill).  The
 * gathers are divergent (every lane its own entry): from shared memory they cost a
 * fixed ~30 cycles and no cache capacity, where the global tables compete with the
 * history rings for what little L1 a kernel with 220 KB of shared memory leaves. */
#define DDE_TAB_SHARED_BYTES (128 * 16 + 128 * 8 + 128 * 16)
#if defined(DDE_HOST_EMU)
/* tests/emu: a shared-window address is a byte offset into the emulated shared memory */
static inline char *dde_emu_shared(unsigned addr);
#endif

__device__ __forceinline__ double dde_pow_shared(unsigned base, double x, double y) {
  const uint32_t topx = dde_top12(x);
  const uint32_t topy = dde_top12(y);
  const int rare =
      (topx - 0x001u >= 0x7ffu - 0x001u) | ((topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu);
  const uint64_t ix = dde_asuint64(x);
  const unsigned i = (unsigned) dde_pow_log_index(ix);
  double invc, logc, logctail;
#if defined(DDE_HOST_EMU)
  invc = ((const double *) dde_emu_shared(base + 16u * i))[0];
  logc = ((const double *) dde_emu_shared(base + 16u * i))[1];
  logctail = *(const double *) dde_emu_shared(base + 2048u + 8u * i);
#else
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(invc), "=d"(logc) : "r"(base + 16u * i));
  asm("ld.shared.f64 %0, [%1];" : "=d"(logctail) : "r"(base + 2048u + 8u * i));
#endif
  double lo;
  const double hi = dde_pow_log_finish(ix, invc, logc, logctail, &lo);
  int exp_rare;
  const struct DdeExpReduced red = dde_pow_exp_reduce(y, hi, lo, &exp_rare);
  const unsigned k = (unsigned) red.ki % 128u;
  double tail, sb;
#if defined(DDE_HOST_EMU)
  tail = ((const double *) dde_emu_shared(base + 3072u + 16u * k))[0];
  sb = ((const double *) dde_emu_shared(base + 3072u + 16u * k))[1];
#else
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tail), "=d"(sb) : "r"(base + 3072u + 16u * k));
#endif
  const double common = dde_pow_exp_finish(red, tail, dde_asuint64(sb));
  if (rare | exp_rare) {
    return dde_pow_rare(x, y, common);
  }
  return common;
}

/* dde_pow2 from the kernel's copy of the tables */
__device__ __forceinline__ struct DdePow2 dde_pow2_shared(unsigned base, double x, double y1,
                                                          double y2) {
  const uint32_t topx = dde_top12(x);
  const int rare_x = topx - 0x001u >= 0x7ffu - 0x001u;
  const int rare1 = rare_x | ((dde_top12(y1) & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu);
  const int rare2 = rare_x | ((dde_top12(y2) & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu);
  const uint64_t ix = dde_asuint64(x);
  const unsigned i = (unsigned) dde_pow_log_index(ix);
  double invc, logc, logctail;
#if defined(DDE_HOST_EMU)
  invc = ((const double *) dde_emu_shared(base + 16u * i))[0];
  logc = ((const double *) dde_emu_shared(base + 16u * i))[1];
  logctail = *(const double *) dde_emu_shared(base + 2048u + 8u * i);
#else
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(invc), "=d"(logc) : "r"(base + 16u * i));
  asm("ld.shared.f64 %0, [%1];" : "=d"(logctail) : "r"(base + 2048u + 8u * i));
#endif
  double lo;
  const double hi = dde_pow_log_finish(ix, invc, logc, logctail, &lo);
  int er1, er2;
  const struct DdeExpReduced red1 = dde_pow_exp_reduce(y1, hi, lo, &er1);
  const struct DdeExpReduced red2 = dde_pow_exp_reduce(y2, hi, lo, &er2);
  const unsigned k1 = (unsigned) red1.ki % 128u, k2 = (unsigned) red2.ki % 128u;
  double t1, s1, t2, s2;
#if defined(DDE_HOST_EMU)
  t1 = ((const double *) dde_emu_shared(base + 3072u + 16u * k1))[0];
  s1 = ((const double *) dde_emu_shared(base + 3072u + 16u * k1))[1];
  t2 = ((const double *) dde_emu_shared(base + 3072u + 16u * k2))[0];
  s2 = ((const double *) dde_emu_shared(base + 3072u + 16u * k2))[1];
#else
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t1), "=d"(s1) : "r"(base + 3072u + 16u * k1));
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t2), "=d"(s2) : "r"(base + 3072u + 16u * k2));
#endif
  const double c1 = dde_pow_exp_finish(red1, t1, dde_asuint64(s1));
  const double c2 = dde_pow_exp_finish(red2, t2, dde_asuint64(s2));
  struct DdePow2 out;
  out.a = (rare1 | er1) ? dde_pow_rare(x, y1, c1) : c1;
  out.b = (rare2 | er2) ? dde_pow_rare(x, y2, c2) : c2;
  return out;
}
#endif

#endif /* DDE_MATH_H */
