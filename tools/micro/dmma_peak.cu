// Micro-benchmark: sustained DMMA.8x8x4 (mma.sync.m8n8k4.f64) and DFMA rate per SM for a given number of warps and
// independent accumulator chains per warp.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_peak dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void k_dmma(double *out, int iters, double a, double b) {
  double acc[CH][2];
#pragma unroll
  for (int c = 0; c < CH; ++c) { acc[c][0] = threadIdx.x; acc[c][1] = c; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) dmma(acc[c][0], acc[c][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += acc[c][0] + acc[c][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = threadIdx.x + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double *out;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    float ms = timeit([&] { k_dmma<16><<<148, warps * 32>>>(out, iters, 1.0000001, 0.999999); });
    double n = 148.0 * warps * 16.0 * iters;   // DMMA instructions
    printf("DMMA  warps/SM %2d ch 16: %.3f ms  %.2f TFLOP/s  cycles per DMMA per SMSP @%d MHz: %.1f\n", warps, ms,
           n * 512 / ms / 1e9, clk / 1000, (ms * 1e-3 * clk * 1e3) / (n / 148 / 4));
    ms = timeit([&] { k_dmma<4><<<148, warps * 32>>>(out, iters, 1.0000001, 0.999999); });
    n = 148.0 * warps * 4.0 * iters;
    printf("DMMA  warps/SM %2d ch  4: %.3f ms  %.2f TFLOP/s  cycles per DMMA per SMSP: %.1f\n", warps, ms, n * 512 / ms / 1e9,
           (ms * 1e-3 * clk * 1e3) / (n / 148 / 4));
    ms = timeit([&] { k_dfma<16><<<148, warps * 32>>>(out, iters, 1.0000001, 0.999999); });
    n = 148.0 * warps * 32 * 16.0 * iters;     // thread-level FMAs
    printf("DFMA  warps/SM %2d ch 16: %.3f ms  %.2f TFLOP/s\n", warps, ms, n * 2 / ms / 1e9);
  }
  return 0;
}
