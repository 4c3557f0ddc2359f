// K6/K7: stand-alone log-EI kernel (value, optional gradient / posterior mean / sigma) and row-wise argmax.
// The evaluation itself lives in logei_core.cuh (shared with the device-resident L-BFGS-B kernel).
#include "logei_core.cuh"

namespace {

using namespace logei;

struct LogEiArgs {
  pcabo_gp_model m;
  int npts;
  const double *X;
  double *val, *grad, *mu, *sigma;
};

template <int P, bool GRAD>
__global__ void __launch_bounds__(LE_THREADS) logei_kernel(LogEiArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.y, p0 = blockIdx.x * P;
  const int pcount = min(P, a.npts - p0);
  Ctx c;
  load_ctx(c, a.m, b);
  carve_ctx(c, P, GRAD, sm);
  const size_t o = (size_t)b * a.npts + p0;
  eval_points<P, GRAD>(c, pcount, a.X + o * a.m.kmax, a.m.kmax, a.val + o, GRAD ? a.grad + o * a.m.kmax : nullptr,
                       a.m.kmax, a.m.kmax, 1.0, a.mu ? a.mu + o : nullptr, a.sigma ? a.sigma + o : nullptr);
}

template <int P, bool GRAD>
int launch_logei(const LogEiArgs &a, cudaStream_t st) {
  const size_t smem = scratch_doubles(P, a.m.kmax, a.m.n, GRAD) * sizeof(double);
  if (smem > 227 * 1024) return PCABO_E_UNSUPPORTED;
  PCABO_CUDA_CHECK(cudaFuncSetAttribute(logei_kernel<P, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((a.npts + P - 1) / P, a.m.B);
  logei_kernel<P, GRAD><<<grid, LE_THREADS, smem, st>>>(a);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}

__global__ void argmax_kernel(int mcols, const double *v, double *best_val, int *best_idx) {
  __shared__ double sv[32];
  __shared__ int si[32];
  const int b = blockIdx.x;
  const double *row = v + (size_t)b * mcols;
  double bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < mcols; i += blockDim.x) {
    const double x = row[i];
    if (x > bv || (x == bv && i < bi)) { bv = x; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    bv = lane < nw ? sv[lane] : -INFINITY;
    bi = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { best_val[b] = bv; best_idx[b] = (bi == 0x7fffffff) ? 0 : bi; }
  }
}

}  // namespace

extern "C" int pcabo_logei(const pcabo_gp_model *m, int npts, const double *X, double *val, double *grad, double *mu,
                           double *sigma, void *stream) {
  if (!m || npts <= 0 || !X || !val || m->B <= 0 || m->n < 1 || m->ld_n < m->n || !m->norm_off || !m->norm_coef ||
      !m->ZnT || !m->alpha || !m->W || !m->WT || !m->ymean || !m->ystd || !m->best_f || (!m->lengthscales && !(m->lengthscale > 0.0)))
    return PCABO_E_INVALID;
  if (grad && !m->Zn) return PCABO_E_INVALID;   // the gradient pass reads the row-major training inputs
  if (m->kmax > 104 || m->n > 4096) return PCABO_E_UNSUPPORTED;
  LogEiArgs a;
  a.m = *m; a.npts = npts; a.X = X; a.val = val; a.grad = grad; a.mu = mu; a.sigma = sigma;
  cudaStream_t st = as_stream(stream);
  auto fits = [&](int P, bool g) { return scratch_doubles(P, m->kmax, m->n, g) * sizeof(double) <= 200 * 1024; };
  if (grad) {
    // pick the largest P that fits shared memory and does not over-pad small batches
    if (npts > 8 && fits(16, true)) return launch_logei<16, true>(a, st);
    if (npts > 4 && fits(8, true)) return launch_logei<8, true>(a, st);
    return launch_logei<4, true>(a, st);
  }
  if (npts > 8 && fits(16, false)) return launch_logei<16, false>(a, st);
  if (npts > 4 && fits(8, false)) return launch_logei<8, false>(a, st);
  return launch_logei<4, false>(a, st);
}

extern "C" int pcabo_argmax(int B, int mcols, const double *v, double *best_val, int *best_idx, void *stream) {
  if (B <= 0 || mcols <= 0 || !v || !best_val || !best_idx) return PCABO_E_INVALID;
  argmax_kernel<<<B, 256, 0, as_stream(stream)>>>(mcols, v, best_val, best_idx);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}
