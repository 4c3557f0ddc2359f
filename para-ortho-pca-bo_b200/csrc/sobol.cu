// K8: scrambled Sobol raw samples generated on the device from host-drawn scramble bits.
//
// Replaces botorch draw_sobol_samples -> torch.quasirandom.SobolEngine(k, scramble=True).draw(n, float64) scaled to the
// search box, i.e. the raw_samples=512 points of the reference's optimize_acqf call (PCA_BO.py:607-614, SURVEY App.
// A.5 step 2). The random bits of the scramble (torch.randint from the run's CPU generator) stay on the host for
// RNG-stream parity; everything downstream is integer arithmetic restated from ATen's SobolEngineOps and is
// bit-identical to torch:
//   scrambled direction number  s'[d][j] : bit (29 - p) = parity(ltm_dots[d][p] & s[d][j]),  p = 0..29
//   point 0                     = float32(shift[d]) * 2^-30   (torch keeps the first point in float32)
//   point i >= 1                = (shift[d] xor XOR_{l in bits(gray(i))} s'[d][l]) * 2^-30,  gray(i) = i ^ (i >> 1)
//   sample                      = lo + (hi - lo) * point        (no FMA contraction, as torch evaluates it)
#include "common.cuh"

namespace {

constexpr int MAXBIT = 30;

__global__ void __launch_bounds__(256) sobol_kernel(int npts, int kmax, const int *k, const int *shift,
                                                    const int *ltm_dots, const int *state0, const double *box_lo,
                                                    const double *box_hi, double *X) {
  extern __shared__ int sst[];  // [kk][MAXBIT] scrambled direction numbers
  const int b = blockIdx.x, tid = threadIdx.x;
  const int kk = k ? k[b] : kmax;
  const int *ld = ltm_dots + (size_t)b * kmax * MAXBIT;
  for (int idx = tid; idx < kk * MAXBIT; idx += blockDim.x) {
    const int d = idx / MAXBIT;
    const int v = state0[idx];
    int t2 = 0;
#pragma unroll
    for (int p = 0; p < MAXBIT; ++p) t2 |= (__popc(ld[d * MAXBIT + p] & v) & 1) << (MAXBIT - 1 - p);
    sst[idx] = t2;
  }
  __syncthreads();
  const double recip = 1.0 / 1073741824.0;
  double *Xb = X + (size_t)b * npts * kmax;
  for (int idx = tid; idx < npts * kmax; idx += blockDim.x) {
    const int i = idx / kmax, d = idx % kmax;
    double out = 0.0;
    if (d < kk) {
      const int sh = shift[(size_t)b * kmax + d];
      double raw;
      if (i == 0) {
        raw = (double)((float)sh) * recip;
      } else {
        int q = sh;
        unsigned g = (unsigned)i ^ ((unsigned)i >> 1);
        while (g) {
          const int l = __ffs(g) - 1;
          q ^= sst[d * MAXBIT + l];
          g &= g - 1;
        }
        raw = (double)q * recip;
      }
      const double lo = box_lo[(size_t)b * kmax + d], hi = box_hi[(size_t)b * kmax + d];
      out = __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), raw));
    }
    Xb[idx] = out;
  }
}

__global__ void gather_rows_kernel(int npts, int nsel, int kmax, const double *X, const int *idx, double *out) {
  const int b = blockIdx.x;
  for (int e = threadIdx.x; e < nsel * kmax; e += blockDim.x) {
    const int r = e / kmax, c = e % kmax;
    out[((size_t)b * nsel + r) * kmax + c] = X[((size_t)b * npts + idx[(size_t)b * nsel + r]) * kmax + c];
  }
}

}  // namespace

extern "C" int pcabo_sobol_points(int B, int npts, int kmax, const int *k, const int *shift, const int *ltm_dots,
                                  const int *state0, const double *box_lo, const double *box_hi, double *X,
                                  void *stream) {
  if (B <= 0 || npts <= 0 || kmax <= 0 || !shift || !ltm_dots || !state0 || !box_lo || !box_hi || !X)
    return PCABO_E_INVALID;
  if (kmax > 1024) return PCABO_E_UNSUPPORTED;
  const size_t smem = (size_t)kmax * MAXBIT * sizeof(int);
  if (smem > 48 * 1024)
    PCABO_CUDA_CHECK(cudaFuncSetAttribute(sobol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sobol_kernel<<<B, 256, smem, as_stream(stream)>>>(npts, kmax, k, shift, ltm_dots, state0, box_lo, box_hi, X);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}

extern "C" int pcabo_gather_rows(int B, int npts, int nsel, int kmax, const double *X, const int *idx, double *out,
                                 void *stream) {
  if (B <= 0 || npts <= 0 || nsel <= 0 || kmax <= 0 || !X || !idx || !out) return PCABO_E_INVALID;
  gather_rows_kernel<<<B, 128, 0, as_stream(stream)>>>(npts, nsel, kmax, X, idx, out);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}
