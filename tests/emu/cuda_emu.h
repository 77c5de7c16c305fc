/* cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Just enough of the CUDA device vocabulary for g++ to compile the stepping
 * kernels of include/dde_b200 as ordinary host functions, ONE lane at a time
 * (a warp of one thread, a block whose threads run one after the other), so
 * that the control flow and the arithmetic of the device code can be checked
 * against the oracle in this container, which has no GPU.  It exercises
 * nothing warp-wide (votes and shuffles see a single lane) and is never part
 * of the product: libdde_b200.so has no host fallback.
 */
#ifndef DDE_CUDA_EMU_H
#define DDE_CUDA_EMU_H

#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* every standard header the driver uses comes before the CUDA keywords below
   (libstdc++ spells its own attributes __noinline__ etc.) */
#include <string>
#include <vector>

#define DDE_HOST_EMU 1

struct emu_dim3 {
  unsigned x, y, z;
};
static emu_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1},
                gridDim = {1, 1, 1};

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __constant__
#define __launch_bounds__(...)
#define __align__(x) alignas(x)
#define __restrict__
#define __builtin_assume(x) ((void) 0)

static inline void __syncthreads() {}
static inline int __syncthreads_and(int p) { return p; }
static inline int __syncthreads_or(int p) { return p; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline unsigned __activemask() { return 1u; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline int __any_sync(unsigned, int p) { return p != 0; }
static inline int __all_sync(unsigned, int p) { return p != 0; }
template <class T>
static inline T __shfl_sync(unsigned, T v, int) {
  return v;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int) x); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
  const unsigned long long old = *p;
  *p = old + v;
  return old;
}
static inline unsigned atomicAdd(unsigned *p, unsigned v) {
  const unsigned old = *p;
  *p = old + v;
  return old;
}
template <class T>
static inline T __ldg(const T *p) {
  return *p;
}
static inline double __longlong_as_double(long long v) {
  double d;
  memcpy(&d, &v, sizeof d);
  return d;
}
static inline long long __double_as_longlong(double d) {
  long long v;
  memcpy(&v, &d, sizeof v);
  return v;
}
static inline int __double2hiint(double d) { return (int) (__double_as_longlong(d) >> 32); }
static inline double __hiloint2double(int hi, int lo) {
  return __longlong_as_double((long long) (((unsigned long long) (unsigned) hi << 32) | (unsigned) lo));
}
static inline float __int_as_float(int v) {
  float f;
  memcpy(&f, &v, sizeof f);
  return f;
}
struct double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) {
  double2 v = {x, y};
  return v;
}
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return __builtin_fma(a, b, c); }

#endif /* DDE_CUDA_EMU_H */
