// Cycle count of the 64 x 64 diagonal-block factorisation + inverse of gp_chol.cu in isolation (one block of 256 threads).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o chol_diag chol_diag.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../../para-ortho-pca-bo_b200/csrc/gp_chol.cu"
const char *pcabo_set_cuda_error(cudaError_t e) { return cudaGetErrorString(e); }
namespace {
__global__ void __launch_bounds__(NTH, 2) kdiag(const double *A, double *Lout, double *Xout, long long *cyc) {
  extern __shared__ __align__(16) double sm[];
  double *D = sm, *X = sm + T * TLD;
  __shared__ double s_rdiag[T];
  __shared__ int s_info;
  for (int i = threadIdx.x; i < T * T; i += NTH) { const int r = i >> 6, c = i & 63; D[r * TLD + c] = (c <= r) ? A[r * T + c] : 0.0; }
  if (threadIdx.x == 0) s_info = 0;
  __syncthreads();
  long long t0 = clock64();
  diag_factor_invert(D, X, s_rdiag, &s_info, 0);
  long long t1 = clock64();
  for (int i = threadIdx.x; i < T * T; i += NTH) { const int r = i >> 6, c = i & 63; Lout[i] = D[r * TLD + c]; Xout[i] = X[r * TLD + c]; }
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = s_info; }
}
}
int main() {
  const int n = 64;
  static double h[4096], M[4096], L[4096], X[4096];
  srand(3);
  for (int i = 0; i < n * n; ++i) M[i] = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = (i == j) ? 4.0 : 0.0; for (int t = 0; t < n; ++t) s += M[i * n + t] * M[j * n + t]; h[i * n + j] = s; }
  double *dA, *dL, *dX; long long *dc;
  cudaMalloc(&dA, 32768); cudaMalloc(&dL, 32768); cudaMalloc(&dX, 32768); cudaMalloc(&dc, 16);
  cudaMemcpy(dA, h, 32768, cudaMemcpyHostToDevice);
  const size_t smem = 2 * T * TLD * sizeof(double);
  cudaFuncSetAttribute(kdiag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long c[2];
  for (int rep = 0; rep < 3; ++rep) { kdiag<<<1, NTH, smem>>>(dA, dL, dX, dc); cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost); printf("run %d: %lld cycles, info %lld\n", rep, c[0], c[1]); }
  cudaMemcpy(L, dL, 32768, cudaMemcpyDeviceToHost); cudaMemcpy(X, dX, 32768, cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0;
  for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int t = 0; t <= j; ++t) s += L[i * n + t] * L[j * n + t]; e1 = fmax(e1, fabs(s - h[i * n + j])); }
  for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int t = j; t <= i; ++t) s += L[i * n + t] * X[t * n + j]; e2 = fmax(e2, fabs(s - (i == j ? 1.0 : 0.0))); }
  printf("|L L^T - A| = %.2e, |L X - I| = %.2e (%s)\n", e1, e2, cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
