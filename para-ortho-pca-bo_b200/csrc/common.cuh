// Shared device/host helpers for libpcabo_b200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/pcabo_b200.h"

extern const char *pcabo_set_cuda_error(cudaError_t e);

#define PCABO_CUDA_CHECK(expr)                      \
  do {                                              \
    cudaError_t e__ = (expr);                       \
    if (e__ != cudaSuccess) {                       \
      pcabo_set_cuda_error(e__);                    \
      return PCABO_E_CUDA;                          \
    }                                               \
  } while (0)

#define PCABO_LAUNCH_CHECK() PCABO_CUDA_CHECK(cudaGetLastError())

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// gp_chol.cu: Cholesky / inverse / alpha / transpose pipeline on the kernel matrix stored in L (lower triangle)
int pcabo_chol_pipeline(int B, int n, int ld_n, double *L, double *W, double *WT, double *alpha, int *info, cudaStream_t st);

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; `scratch` must hold >= 32 doubles. All threads get the result.
__device__ __forceinline__ double block_sum(double v, double *scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double r = (lane < nw) ? scratch[lane] : 0.0;
  r = warp_sum(r);
  return r;
}

// numpy's float64 pairwise summation order (DOUBLE_pairwise_sum): bit-identical to ndarray.sum().
__device__ inline double numpy_pairwise_sum(const double *a, int n) {
  if (n < 8) {
    if (n == 0) return -0.0;
    double res = a[0];
    for (int i = 1; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  }
  if (n <= 128) {
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i = 8;
    const int lim = n - (n % 8);
    for (; i < lim; i += 8) {
      r0 = __dadd_rn(r0, a[i + 0]); r1 = __dadd_rn(r1, a[i + 1]);
      r2 = __dadd_rn(r2, a[i + 2]); r3 = __dadd_rn(r3, a[i + 3]);
      r4 = __dadd_rn(r4, a[i + 4]); r5 = __dadd_rn(r5, a[i + 5]);
      r6 = __dadd_rn(r6, a[i + 6]); r7 = __dadd_rn(r7, a[i + 7]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                           __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __dadd_rn(numpy_pairwise_sum(a, n2), numpy_pairwise_sum(a + n2, n - n2));
}
