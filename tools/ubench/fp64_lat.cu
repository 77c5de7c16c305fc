/* fp64_lat.cu -- B200 FP64 latency / issue microbenchmarks used to size the
 * stepping kernel's schedule: dependent-chain latency of
 * DFMA / DADD / DMUL, the same with ILP 2 / 4, shared / L1 / L2 load-to-use latency,
 * and MUFU.RCP64H.  One warp, clock64 around 4096 dependent operations.
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o build/fp64_lat tools/ubench/fp64_lat.cu */
#include <cstdio>
#include <cuda_runtime.h>

#define N_IT 1024

template <int OP, int ILP>
__global__ void chain(double *out, long long *cyc, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (OP == 0) x[i] = fma(x[i], b, a);
        if (OP == 1) x[i] = x[i] + b;
        if (OP == 2) x[i] = x[i] * b;
        if (OP == 3) { x[i] = x[i] * b; x[i] = x[i] + a; }
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void lds_chain(int *out, long long *cyc) {
  __shared__ int buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = (i + 32) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) p = buf[p];
  }
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void ldg_chain(const int *tab, int *out, long long *cyc, int warm) {
  int p = threadIdx.x;
  for (int w = 0; w < warm; ++w) { /* warm L1 */
    int q = threadIdx.x;
    for (int it = 0; it < 64; ++it) q = __ldg(&tab[q]);
    p ^= q & 0;
  }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) p = __ldg(&tab[p]);
  }
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void rcp_chain(double *out, long long *cyc, double a) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x));
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void div_chain(double *out, long long *cyc, double a, double b) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) x = x / b + a;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP, int ILP>
static void run(const char *name, int warps_per_block, int blocks) {
  double *out;
  long long *cyc;
  cudaMalloc(&out, sizeof(double) * 1024 * 1024);
  cudaMallocManaged(&cyc, sizeof(long long) * 4096);
  chain<OP, ILP><<<blocks, 32 * warps_per_block>>>(out, cyc, 1.000001, 0.999999);
  cudaDeviceSynchronize();
  chain<OP, ILP><<<blocks, 32 * warps_per_block>>>(out, cyc, 1.000001, 0.999999);
  cudaDeviceSynchronize();
  const double ops = (double) N_IT * 4 * ILP * (OP == 3 ? 2 : 1);
  printf("%-10s ilp %d warps/block %2d blocks %4d: %.2f cycles per op per warp, %.3f warp-ops/cycle/SM\n",
         name, ILP, warps_per_block, blocks, cyc[0] / ops * 1.0,
         ops * warps_per_block * (blocks >= 148 ? blocks / 148.0 : 1.0) / cyc[0]);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0, 1>("dfma", 1, 1);
  run<1, 1>("dadd", 1, 1);
  run<2, 1>("dmul", 1, 1);
  run<3, 1>("dmul+dadd", 1, 1);
  run<0, 2>("dfma", 1, 1);
  run<0, 4>("dfma", 1, 1);
  run<0, 8>("dfma", 1, 1);
  run<3, 4>("dmul+dadd", 1, 1);
  /* 4 warps = one per scheduler; 12 = three per scheduler (the stepping kernel's geometry) */
  run<0, 1>("dfma", 4, 1);
  run<0, 1>("dfma", 12, 1);
  run<0, 2>("dfma", 12, 1);
  run<0, 4>("dfma", 12, 1);
  run<3, 1>("dmul+dadd", 12, 1);
  run<3, 2>("dmul+dadd", 12, 1);
  run<0, 1>("dfma", 16, 1);
  run<0, 1>("dfma", 32, 1);
  run<0, 4>("dfma", 12, 148);
  int *o;
  long long *cyc;
  cudaMalloc(&o, 4096);
  cudaMallocManaged(&cyc, 64);
  lds_chain<<<1, 32>>>(o, cyc);
  cudaDeviceSynchronize();
  printf("lds dependent chain: %.1f cycles per load\n", cyc[0] / (N_IT * 4.0));
  int *tab, h[1 << 20];
  static int hh[1 << 22];
  /* 4 KB table: L1-resident after warm-up; 16 MB stride-chase: L2 */
  for (int i = 0; i < 1024; ++i) h[i] = (i + 32) & 1023;
  cudaMalloc(&tab, sizeof hh);
  cudaMemcpy(tab, h, 4096, cudaMemcpyHostToDevice);
  ldg_chain<<<1, 32>>>(tab, o, cyc, 1);
  cudaDeviceSynchronize();
  printf("ldg (L1 hit) dependent chain: %.1f cycles per load\n", cyc[0] / (N_IT * 4.0));
  for (int i = 0; i < (1 << 22); ++i) hh[i] = (int) (((long long) i + 32 * 1031) % (1 << 22));
  cudaMemcpy(tab, hh, sizeof hh, cudaMemcpyHostToDevice);
  ldg_chain<<<1, 32>>>(tab, o, cyc, 0);
  cudaDeviceSynchronize();
  printf("ldg (L2, 16 MB chase) dependent chain: %.1f cycles per load\n", cyc[0] / (N_IT * 4.0));
  double *od;
  cudaMalloc(&od, 4096);
  rcp_chain<<<1, 32>>>(od, cyc, 1.5);
  cudaDeviceSynchronize();
  printf("mufu.rcp64h dependent chain: %.1f cycles per op\n", cyc[0] / (N_IT * 4.0));
  div_chain<<<1, 32>>>(od, cyc, 1.5, 1.0000001);
  cudaDeviceSynchronize();
  printf("x / b + a dependent chain: %.1f cycles per (div + add)\n", cyc[0] / (N_IT * 4.0));
  return 0;
}
