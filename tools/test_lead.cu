// Standalone device test of v2::lead_solve against a host GTH (debugging aid).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include -I radex_emcee_b200/csrc -o tools/test_lead tools/test_lead.cu radex_emcee_b200/csrc/moldata.cpp
#include "../radex_emcee_b200/csrc/radex_b200.cu"
#include <random>

__global__ void k_test_lead(const double *Q, int Kp, double *xout, double *totout) {
  extern __shared__ double smem[];
  double *sm = smem;
  const int lane = threadIdx.x, n = 4 * Kp;
  double *B = sm + v2::O_B;
  for (int e = lane; e < v2::NB; e += 32) B[e] = 0.0;
  __syncwarp();
  for (int e = lane; e < n * n; e += 32) {
    const int i = e / n, j = e % n;
    B[i * (n + 2) + j] = Q[i * n + j];
  }
  // M = 0 -> frozen populations 0
  for (int e = lane; e < n * (v2::LDB - n); e += 32) B[v2::o_m(n) + e] = 0.0;
  __syncwarp();
  const double tot = v2::lead_solve(sm, Kp, lane);
  __syncwarp();
  for (int i = lane; i < v2::NL; i += 32) xout[i] = sm[v2::O_XNEW + i];
  if (lane == 0) *totout = tot;
}

int main() {
  std::mt19937_64 rng(3);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  for (int Kp = 3; Kp <= 7; ++Kp) {
    const int n = 4 * Kp;
    std::vector<double> Q(n * n), W(n * n);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) Q[i * n + j] = (i == j) ? 0.0 : U(rng) * pow(10.0, -8.0 * U(rng));
    W = Q;
    std::vector<double> x(n, 0.0);
    for (int k = n - 1; k >= 1; --k) {
      double s = 0;
      for (int j = 0; j < k; ++j) s += W[k * n + j];
      for (int i = 0; i < k; ++i) W[i * n + k] /= s;
      for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) W[i * n + j] += W[i * n + k] * W[k * n + j];
    }
    x[0] = 1;
    for (int k = 1; k < n; ++k) {
      double s = 0;
      for (int i = 0; i < k; ++i) s += x[i] * W[i * n + k];
      x[k] = s;
    }
    double *dQ, *dx, *dt;
    cudaMalloc(&dQ, n * n * 8); cudaMalloc(&dx, 41 * 8); cudaMalloc(&dt, 8);
    cudaMemcpy(dQ, Q.data(), n * n * 8, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_test_lead, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SLAB * 8);
    k_test_lead<<<1, 32, v2::SLAB * 8>>>(dQ, Kp, dx, dt);
    double hx[41], ht;
    cudaMemcpy(hx, dx, 41 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&ht, dt, 8, cudaMemcpyDeviceToHost);
    double worst = 0;
    int wi = -1;
    for (int i = 0; i < n; ++i) {
      const double d = fabs(hx[i] - x[i]) / fabs(x[i]);
      if (d > worst) { worst = d; wi = i; }
    }
    printf("Kp %d: %s worst rel err %.3e at level %d (dev %.6e host %.6e)\n", Kp, cudaGetErrorString(cudaGetLastError()), worst, wi,
           wi >= 0 ? hx[wi] : 0, wi >= 0 ? x[wi] : 0);
  }
  return 0;
}
