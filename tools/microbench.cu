// Latency / issue-rate probes for the instruction classes the v2 solve kernel is bound by
// (B200, sm_100a).  One warp, dependent chains timed with clock64().  Build:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N 256

__device__ __forceinline__ double mufu_rcp(double x) {
  double r;
  asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void k(double *out, long long *cyc, double seed) {
  __shared__ double sh[64];
  const int lane = threadIdx.x;
  sh[lane] = seed + lane;
  sh[lane + 32] = seed;
  __syncwarp();
  double x = seed + lane * 1e-3, y = 1.0 + seed, acc = 0;
  long long t0, t1;
  // 0: dependent DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) x = fma(x, y, seed);
  t1 = clock64();
  if (lane == 0) cyc[0] = t1 - t0;
  acc += x;
  // 1: dependent MUFU.RCP64H chain (seed only)
  x = 1.5 + seed;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) x = mufu_rcp(x);
  t1 = clock64();
  if (lane == 0) cyc[1] = t1 - t0;
  acc += x;
  // 2: dependent rcp1 chain (MUFU + 1 Newton = 2 DFMA)
  x = 1.5 + seed;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double r = mufu_rcp(x);
    x = fma(r, fma(-x, r, 1.0), r) + 1.0;
  }
  t1 = clock64();
  if (lane == 0) cyc[2] = t1 - t0;
  acc += x;
  // 3: dependent SHFL (64-bit) chain
  x = seed + lane;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1);
  t1 = clock64();
  if (lane == 0) cyc[3] = t1 - t0;
  acc += x;
  // 4: dependent LDS chain (pointer chase through shared memory)
  int idx = lane;
  volatile int *shi = (volatile int *)sh;
  __syncwarp();
  shi[lane] = (lane + 1) & 31;
  __syncwarp();
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) idx = shi[idx];
  t1 = clock64();
  if (lane == 0) cyc[4] = t1 - t0;
  acc += idx;
  // 5: dependent DMMA chain (accumulator dependency)
  double c0 = 0, c1 = 0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) dmma(c0, c1, y, seed);
  t1 = clock64();
  if (lane == 0) cyc[5] = t1 - t0;
  acc += c0 + c1;
  // 6: 8 independent DMMA accumulators (issue rate)
  double d[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) d[j][0] = d[j][1] = j;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N / 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma(d[j][0], d[j][1], y, seed);
  t1 = clock64();
  if (lane == 0) cyc[6] = t1 - t0;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc += d[j][0] + d[j][1];
  // 7: 8 independent DFMA chains (issue rate)
  double e[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) e[j] = seed + j;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N / 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] = fma(e[j], y, seed);
  t1 = clock64();
  if (lane == 0) cyc[7] = t1 - t0;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc += e[j];
  // 8: DADD -> DSETP -> select chain
  x = seed + 2.0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) x = (x > seed) ? x + y : x - y;
  t1 = clock64();
  if (lane == 0) cyc[8] = t1 - t0;
  acc += x;
  // 9: STS -> LDS same-thread round trip
  x = seed;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    ((volatile double *)sh)[lane + 32] = x;
    x = ((volatile double *)sh)[lane + 32] + 1.0;
  }
  t1 = clock64();
  if (lane == 0) cyc[9] = t1 - t0;
  acc += x;
  out[lane] = acc;
}

int main() {
  double *out;
  long long *cyc, h[16];
  cudaMalloc(&out, 32 * sizeof(double));
  cudaMalloc(&cyc, 16 * sizeof(long long));
  for (int rep = 0; rep < 2; ++rep) {
    k<<<1, 32>>>(out, cyc, 1e-9);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const char *names[] = {"DFMA dependent",       "MUFU.RCP64H dependent", "rcp1+DADD dependent (MUFU+2DFMA+DADD)",
                         "SHFL.64 dependent",    "LDS.32 dependent",      "DMMA dependent (accumulator)",
                         "DMMA x8 independent",  "DFMA x8 independent",   "DADD+DSETP+select dependent",
                         "STS.64->LDS.64+DADD"};
  for (int i = 0; i < 10; ++i) printf("%-40s %7.2f cycles/op\n", names[i], (double)h[i] / N);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
