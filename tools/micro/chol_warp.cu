// Cycle count of the optimiser warp's small dense kernels in isolation (one warp): dev::chol_upper_warp on a 20 x 20 /
// 10 x 10 matrix and dev::tri_solve_warp. nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o chol_warp chol_warp.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include "../../para-ortho-pca-bo_b200/csrc/lbfgsb_core.cuh"

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
  for (int n : {20, 10}) {
    const int lda = n, nsplit = n;   // plain Cholesky (nsplit = n: no sign flips)
    double h[400], M[400];
    srand(1);
    for (int i = 0; i < n * n; ++i) M[i] = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = (i == j) ? n : 0.0; for (int t = 0; t < n; ++t) s += M[i * n + t] * M[j * n + t]; h[i * n + j] = s; }
    double hb[20]; for (int i = 0; i < n; ++i) hb[i] = 1.0 + i;
    double *dA, *dO, *dr, *db; long long *dc;
    cudaMalloc(&dA, 3200); cudaMalloc(&dO, 3200); cudaMalloc(&dr, 160); cudaMalloc(&db, 160); cudaMalloc(&dc, 32);
    cudaMemcpy(dA, h, n * n * 8, cudaMemcpyHostToDevice);
    long long c[4];
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemcpy(db, hb, n * 8, cudaMemcpyHostToDevice);
      k<<<1, 32>>>(dA, lda, n, nsplit, dO, dr, db, dc);
      cudaMemcpy(c, dc, 32, cudaMemcpyDeviceToHost);
    }
    double x[20]; cudaMemcpy(x, db, n * 8, cudaMemcpyDeviceToHost);
    // check: A x = b
    double err = 0; for (int i = 0; i < n; ++i) { double s = 0; for (int j = 0; j < n; ++j) s += h[i * n + j] * x[j]; err = fmax(err, fabs(s - hb[i])); }
    printf("n = %2d: chol %lld cycles (%.0f / pivot), solve R^T %lld, solve R %lld, bad %lld, |A x - b| = %.2e  (%s)\n", n, c[0], (double)c[0] / n, c[1], c[2], c[3], err,
           cudaGetErrorString(cudaDeviceSynchronize()));
  }
  return 0;
}
