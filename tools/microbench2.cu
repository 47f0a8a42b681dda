// FP64 issue-rate probes (B200): ILP x warps-per-scheduler.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench2 tools/microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__device__ __forceinline__ double vfma(double a, double b, double c) {
  double d;
  asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c));
  return d;
}
template <int ILP>
__global__ void k(double *out, long long *cyc, double seed) {
  double e[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) e[j] = seed + j + threadIdx.x;
  const double y = 1.0 + seed;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N / ILP; ++i)
#pragma unroll
    for (int j = 0; j < ILP; ++j) e[j] = vfma(e[j], y, seed);
  const long long t1 = clock64();
  double acc = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) acc += e[j];
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
void run(int warps, double *out, long long *cyc) {
  long long h;
  for (int rep = 0; rep < 2; ++rep) {
    k<ILP><<<1, 32 * warps>>>(out, cyc, 1e-9);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("ILP %d  warps/SM %2d (%d per scheduler): %6.2f cycles per DFMA per warp, %5.2f DFMA warp-instr per cycle per SM\n", ILP, warps,
         warps / 4 ? warps / 4 : 1, (double)h / N, (double)N * warps / h);
}
int main() {
  double *out;
  long long *cyc;
  cudaMalloc(&out, 1024 * sizeof(double));
  cudaMalloc(&cyc, 16 * sizeof(long long));
  for (int w : {1, 4, 8, 16, 32}) {
    run<1>(w, out, cyc);
    run<2>(w, out, cyc);
    run<4>(w, out, cyc);
    run<8>(w, out, cyc);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
