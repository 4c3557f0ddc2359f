// Dependent-chain latencies of the FP64 operations the single optimiser warp of lbfgsb_logei_kernel lives on (B200).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int OP>
__global__ void k(double *out, long long *cyc, double seed, int nwarps_active) {
  __shared__ double sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = seed + threadIdx.x;
  __syncthreads();
  if ((threadIdx.x >> 5) >= nwarps_active) return;
  double x = seed + threadIdx.x * 1e-3, y = 1.0000001, z = 0.0, u = seed * 0.5, v = seed * 0.25, w2 = seed * 0.125;
  int idx = threadIdx.x & 31;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, y, 1e-9);
    else if (OP == 1) x = x + y;
    else if (OP == 2) x = x * y;
    else if (OP == 3) x = 1.0 / x + 1.0;
    else if (OP == 4) x = sqrt(x) + 1.0;
    else if (OP == 5) x = rsqrt(x) + 1.0;
    else if (OP == 6) x = exp(-x) + 1.0;
    else if (OP == 7) dmma(x, z, y, y);
    else if (OP == 8) x += __shfl_xor_sync(0xffffffffu, x, 1);
    else if (OP == 9) { idx = (int)sm[idx] & 31; x += idx; }   // LDS + cvt chain
    else if (OP == 10) { x = fma(x, y, 1e-9); u = fma(u, y, 1e-9); v = fma(v, y, 1e-9); w2 = fma(w2, y, 1e-9); }
    else if (OP == 11) x = log(x) + 2.0;
    else if (OP == 12) x = erfc(x * 1e-3) + 1.0;
    else if (OP == 13) x = sm[(i + threadIdx.x) & 63] * x;      // independent LDS feeding a dependent DMUL
    else if (OP == 14) x = __fdividef((float)x, 3.0f) + 1.0;     // float path for reference
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x + z + u + v + w2 + idx;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  const char *names[] = {"dfma dep", "dadd dep", "dmul dep", "1/x + 1", "sqrt + 1", "rsqrt + 1", "exp + 1", "dmma dep", "shfl64 + add",
                         "lds chase + cvt", "4 indep dfma", "log + 2", "erfc + 1", "lds + dmul", "float div"};
  for (int nw = 1; nw <= 8; nw *= 8) {
    printf("active warps in the block: %d\n", nw);
#define RUN(OP) { k<OP><<<1, 256>>>(out, cyc, 1.5, nw); k<OP><<<1, 256>>>(out, cyc, 1.5, nw); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("  %-18s %7.1f cycles/iter\n", names[OP], (double)c / N); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14)
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
