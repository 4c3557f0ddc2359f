// Throughput of cp.async.bulk global->shared row copies into a ring of slots (one producer thread, 8 consumer warps that
// only wait and release), as lbfgsb_logei_kernel stages the panels of W. B200.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../para-ortho-pca-bo_b200/csrc/async_copy.cuh"
__global__ void k(const double *src, long ld, int rows, int rowbytes, int nslots, int npanels, int touch, long long *cyc, double *sink) {
  extern __shared__ __align__(16) unsigned char sm[];
  unsigned long long *full = reinterpret_cast<unsigned long long *>(sm);
  unsigned long long *empty = full + 16;
  double *ring = reinterpret_cast<double *>(sm + 512);
  const int slot_doubles = rows * (rowbytes / 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { for (int i = 0; i < 16; ++i) { acp::mbar_init(&full[i], 1); acp::mbar_init(&empty[i], 8); } acp::fence_barrier_init(); }
  __syncthreads();
  const double *base = src + (size_t)blockIdx.x * ld * 512;
  long long t0 = clock64();
  double acc = 0.0;
  if (warp == 8) {
    for (int p = 0; p < npanels; ++p) {
      const int s = p % nslots;
      if (p >= nslots) acp::mbar_wait(&empty[s], ((p / nslots) - 1) & 1);
      if (lane == 0) { acp::fence_proxy_async(); acp::mbar_arrive_expect_tx(&full[s], (unsigned)rows * rowbytes); }
      __syncwarp();
      const double *g = base + (size_t)((p * rows) % 448) * ld;
      for (int r = lane; r < rows; r += 32) acp::bulk_g2s(ring + (size_t)s * slot_doubles + (size_t)r * (rowbytes / 8), g + (size_t)r * ld, rowbytes, &full[s]);
    }
  } else {
    for (int p = 0; p < npanels; ++p) {
      const int s = p % nslots;
      acp::mbar_wait(&full[s], (p / nslots) & 1);
      if (touch) { const double *q = ring + (size_t)s * slot_doubles; for (int i = tid; i < slot_doubles; i += 256 * touch) acc += q[i]; }
      __syncwarp();
      if (lane == 0) acp::mbar_arrive(&empty[s]);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 12345.678) sink[0] = acc;
}
int main() {
  const long ld = 512;  // doubles
  double *src, *sink; long long *cyc;
  cudaMalloc(&src, 148 * 512 * ld * 8); cudaMemset(src, 0, 148 * 512 * ld * 8);
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int rows, rowbytes, nslots; } cfgs[] = {{16, 3456, 2}, {16, 3456, 3}, {8, 3456, 4}, {8, 3456, 6}, {16, 1728, 4}, {32, 1024, 4}, {1, 55296, 2}, {1, 55296, 3}, {4, 13824, 3}, {64, 512, 4}};
  for (int grid : {1, 148}) for (auto c : cfgs) for (int touch : {0, 1}) {
    const int npanels = 200;
    size_t smem = 512 + (size_t)c.nslots * c.rows * c.rowbytes;
    if (smem > 200 * 1024) continue;
    // a "row" longer than ld doubles just reads on into the following rows (contiguous memory): fine for a bandwidth test
    k<<<grid, 288, smem>>>(src, ld, c.rows, c.rowbytes, c.nslots, npanels, touch, cyc, sink);
    k<<<grid, 288, smem>>>(src, ld, c.rows, c.rowbytes, c.nslots, npanels, touch, cyc, sink);
    long long h[148]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("grid %3d rows %2d x %5d B slots %d touch %d: %7.0f cycles/panel  %6.1f B/clk/SM  (%s)\n", grid, c.rows, c.rowbytes, c.nslots, touch,
           (double)mx / npanels, (double)c.rows * c.rowbytes * npanels / mx, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
