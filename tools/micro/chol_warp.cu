// Cycle counts of the optimiser warp's small dense kernels in isolation (one warp): dev::chol_upper_warp (shared
// memory, flat-mapped entries) on 20 x 20 / 10 x 10 matrices, plain Cholesky and the LEL^T form, and
// dev::tri_solve_warp. (Variants are tried here by adding a second template value V.)
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o chol_warp chol_warp.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include "../../para-ortho-pca-bo_b200/csrc/lbfgsb_core.cuh"

template <int V>
__global__ void k(const double *Ain, int lda, int n, int nsplit, double *Aout, double *rd_out, double *b_io, long long *cyc) {
  __shared__ double A[20 * 20], rd[20], b[20];
  for (int i = threadIdx.x; i < lda * n; i += 32) A[i] = Ain[i];
  if (threadIdx.x < n) b[threadIdx.x] = b_io[threadIdx.x];
  __syncwarp();
  long long t0 = clock64();
  int bad = lbfgsb::dev::chol_upper_warp(A, lda, n, nsplit, rd);
  long long t1 = clock64();
  lbfgsb::dev::tri_solve_warp(A, lda, n, b, rd, true);
  long long t2 = clock64();
  lbfgsb::dev::tri_solve_warp(A, lda, n, b, rd, false);
  long long t3 = clock64();
  for (int i = threadIdx.x; i < lda * n; i += 32) Aout[i] = A[i];
  if (threadIdx.x < n) { rd_out[threadIdx.x] = rd[threadIdx.x]; b_io[threadIdx.x] = b[threadIdx.x]; }
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = bad; }
}

int main() {
  for (int n : {20, 10, 14}) for (int split : {0, 1}) {
    const int lda = 20, nsplit = split ? n / 2 : n;
    double h[400] = {0}, M[400];
    srand(1 + n);
    for (int i = 0; i < n * n; ++i) M[i] = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = (i == j) ? n : 0.0; for (int t = 0; t < n; ++t) s += M[i * n + t] * M[j * n + t]; h[i * lda + j] = s; }
    double hb[20]; for (int i = 0; i < n; ++i) hb[i] = 1.0 + i;
    double *dA, *dO, *dr, *db; long long *dc;
    cudaMalloc(&dA, 3200); cudaMalloc(&dO, 3200); cudaMalloc(&dr, 160); cudaMalloc(&db, 160); cudaMalloc(&dc, 32);
    cudaMemcpy(dA, h, lda * n * 8, cudaMemcpyHostToDevice);
    long long c[2][4]; double U[2][400], rd[2][20], x[2][20];
    for (int v = 0; v < 2; ++v) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemcpy(db, hb, n * 8, cudaMemcpyHostToDevice);
        if (v == 0) k<0><<<1, 32>>>(dA, lda, n, nsplit, dO, dr, db, dc); else k<1><<<1, 32>>>(dA, lda, n, nsplit, dO, dr, db, dc);
        cudaMemcpy(c[v], dc, 32, cudaMemcpyDeviceToHost);
      }
      cudaMemcpy(U[v], dO, lda * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(rd[v], dr, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(x[v], db, n * 8, cudaMemcpyDeviceToHost);
    }
    double du = 0, dd = 0, dx = 0;
    for (int i = 0; i < n; ++i) { dd = fmax(dd, fabs(rd[0][i] - rd[1][i]) / fabs(rd[0][i])); dx = fmax(dx, fabs(x[0][i] - x[1][i]) / (fabs(x[0][i]) + 1e-300)); for (int j = i + 1; j < n; ++j) du = fmax(du, fabs(U[0][i * lda + j] - U[1][i * lda + j])); }
    printf("n = %2d nsplit = %2d: chol %6lld cycles (second run %6lld); solves %lld / %lld; bad %lld %lld; max |dU| %.1e |drd| %.1e |dx| %.1e (%s)\n", n, nsplit,
           c[0][0], c[1][0], c[0][1], c[0][2], c[0][3], c[1][3], du, dd, dx, cudaGetErrorString(cudaDeviceSynchronize()));
  }
  return 0;
}
