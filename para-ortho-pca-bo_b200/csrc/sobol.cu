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

// ---------------------------------------------------------------------------------------------- MT19937 replay
// torch.randint(2, size, generator=g) on the CPU consumes one 32-bit MT19937 output per element and keeps its lowest
// bit (ATen uniform_int_from_to: random() % 2). The scramble of SobolEngine draws (k, 30) then (k, 30, 30) such bits.
// This kernel replays the generator on the device: input = the generator's 624-word state + position, output = the
// packed shift / lower-triangular rows the Sobol kernel needs and the advanced state, so the host generator can be
// put exactly where torch would have left it.
constexpr int MT_N = 624, MT_M = 397;

__device__ __forceinline__ unsigned mt_twist(unsigned u, unsigned v) {
  return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}

__device__ void mt_regenerate(unsigned *st) {
  const int tid = threadIdx.x, nt = blockDim.x;
  // phase 1: i in [0, 227) uses old st[i], st[i+1], st[i+397]
  for (int base = 0; base < MT_N - MT_M; base += nt) {
    const int i = base + tid;
    unsigned v = 0;
    const bool on = i < MT_N - MT_M;
    if (on) v = st[i + MT_M] ^ mt_twist(st[i], st[i + 1]);
    __syncthreads();
    if (on) st[i] = v;
    __syncthreads();
  }
  // phases 2, 3: i in [227, 454) and [454, 623) use new st[i - 227], old st[i], st[i+1]
  for (int lo = MT_N - MT_M; lo < MT_N - 1; lo += MT_N - MT_M) {
    const int hi = min(lo + MT_N - MT_M, MT_N - 1);
    for (int base = lo; base < hi; base += nt) {
      const int i = base + tid;
      unsigned v = 0;
      const bool on = i < hi;
      if (on) v = st[i - (MT_N - MT_M)] ^ mt_twist(st[i], st[i + 1]);
      __syncthreads();
      if (on) st[i] = v;
      __syncthreads();
    }
  }
  if (tid == 0) st[MT_N - 1] = st[MT_M - 1] ^ mt_twist(st[MT_N - 1], st[0]);
  __syncthreads();
}

__global__ void __launch_bounds__(256) mt_scramble_bits_kernel(int kmax, const int *k, unsigned *mt_state, int *mt_pos,
                                                               int *shift, int *ltm_dots) {
  __shared__ unsigned st[MT_N];
  extern __shared__ int acc[];  // [kmax] shift + [kmax][30] dots
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int kk = k ? k[b] : kmax;
  for (int i = tid; i < MT_N; i += nt) st[i] = mt_state[(size_t)b * MT_N + i];
  for (int i = tid; i < kmax * (MAXBIT + 1); i += nt) acc[i] = 0;
  int pos = mt_pos[b];
  __syncthreads();
  int *sh = acc, *dots = acc + kmax;
  const int total = kk * MAXBIT + kk * MAXBIT * MAXBIT;
  int done = 0;
  while (done < total) {
    if (pos >= MT_N) { mt_regenerate(st); pos = 0; }
    const int take = min(MT_N - pos, total - done);
    for (int t = tid; t < take; t += nt) {
      unsigned y = st[pos + t];
      y ^= y >> 11;
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= y >> 18;
      if (y & 1u) {
        const int e = done + t;
        if (e < kk * MAXBIT) {
          atomicOr(&sh[e / MAXBIT], 1 << (e % MAXBIT));                      // shift = sum bit_j 2^j
        } else {
          const int r = e - kk * MAXBIT;
          const int d = r / (MAXBIT * MAXBIT), p = (r / MAXBIT) % MAXBIT, c = r % MAXBIT;
          if (c < p) atomicOr(&dots[d * MAXBIT + p], 1 << (MAXBIT - 1 - c));   // tril, packed MSB-first
        }
      }
    }
    pos += take;
    done += take;
    __syncthreads();
  }
  for (int i = tid; i < kmax; i += nt) shift[(size_t)b * kmax + i] = (i < kk) ? sh[i] : 0;
  for (int i = tid; i < kmax * MAXBIT; i += nt) {
    const int d = i / MAXBIT, p = i % MAXBIT;
    ltm_dots[(size_t)b * kmax * MAXBIT + i] = (d < kk) ? (dots[i] | (1 << (MAXBIT - 1 - p))) : 0;  // unit diagonal
  }
  for (int i = tid; i < MT_N; i += nt) mt_state[(size_t)b * MT_N + i] = st[i];
  if (tid == 0) mt_pos[b] = pos;
}

// ---------------------------------------------------------------------------------------------- initial conditions
// botorch initialize_q_batch (optim/initializers.py) for one run per block, entirely on the device:
//   Boltzmann weights w = exp(eta (v - mean) / std) of the raw-sample values (eta halved while a weight overflows),
//   torch.multinomial(w, nsel) without replacement = top-nsel of w / q with q ~ Exp(1), where torch's CPU exponential_
//   draws q_i = -log1p(-u_i), u_i = 53 low bits of random64() * 2^-53, random64() = two consecutive MT19937 outputs
//   (high word first) of the run's generator -- replayed here from the device-resident generator state --, the arg-max of
//   the values forced in as the last pick when it was not drawn, and the gather of the picked rows.
// Runs whose values are degenerate (std zero or not finite: botorch picks a random permutation and warns) are only
// flagged, their generator is left untouched, and the host path handles them.
__global__ void __launch_bounds__(256) ic_select_kernel(int nraw, int nsel, int kmax, double eta, const double *vals,
                                                        unsigned *mt_state, int *mt_pos, const double *X, double *ics,
                                                        int *idx_out, int *flag) {
  __shared__ unsigned st[MT_N];
  __shared__ double red[34];
  __shared__ int ired[34];
  __shared__ int picks[32];
  extern __shared__ double dyn[];            // score [nraw] | raw words [2 nraw] (as unsigned)
  double *score = dyn;
  unsigned *words = reinterpret_cast<unsigned *>(dyn + nraw);
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const double *v = vals + (size_t)b * nraw;
  // mean and unbiased standard deviation (two passes)
  double s = 0.0;
  for (int i = tid; i < nraw; i += nt) s += v[i];
  const double mean = block_sum(s, red) / nraw;
  double q2 = 0.0;
  for (int i = tid; i < nraw; i += nt) { const double d = v[i] - mean; q2 = fma(d, d, q2); }
  const double sd = sqrt(block_sum(q2, red) / (nraw - 1));
  if (!(isfinite(sd) && sd != 0.0)) {
    if (tid == 0) flag[b] = 1;
    return;
  }
  if (tid == 0) flag[b] = 0;
  // Boltzmann weights; botorch: while isinf(weights).any(): etaZ *= 0.5
  double scale = eta / sd;
  for (int iter = 0; iter < 64; ++iter) {
    int inf = 0;
    for (int i = tid; i < nraw; i += nt) {
      const double w = exp(scale * (v[i] - mean));
      score[i] = w;
      inf |= isinf(w) ? 1 : 0;
    }
    if (!__syncthreads_or(inf)) break;
    scale *= 0.5;
  }
  // 2 nraw tempered MT19937 outputs from the run's generator
  for (int i = tid; i < MT_N; i += nt) st[i] = mt_state[(size_t)b * MT_N + i];
  int pos = mt_pos[b];
  __syncthreads();
  const int total = 2 * nraw;
  int done = 0;
  while (done < total) {
    if (pos >= MT_N) { mt_regenerate(st); pos = 0; }
    const int take = min(MT_N - pos, total - done);
    for (int t = tid; t < take; t += nt) {
      unsigned y = st[pos + t];
      y ^= y >> 11;
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= y >> 18;
      words[done + t] = y;
    }
    pos += take;
    done += take;
    __syncthreads();
  }
  for (int i = tid; i < MT_N; i += nt) mt_state[(size_t)b * MT_N + i] = st[i];
  if (tid == 0) mt_pos[b] = pos;
  // score = w / q
  for (int i = tid; i < nraw; i += nt) {
    const unsigned long long x = ((unsigned long long)words[2 * i] << 32) | words[2 * i + 1];
    const double u = (double)(x & ((1ULL << 53) - 1)) * (1.0 / 9007199254740992.0);
    const double q = -log1p(-u);
    score[i] = score[i] / q;
  }
  __syncthreads();
  // arg-max of the raw values (first occurrence)
  const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  int amax;
  {
    double bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < nraw; i += nt)
      if (v[i] > bv || (v[i] == bv && i < bi)) { bv = v[i]; bi = i; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red[wid] = bv; ired[wid] = bi; }
    __syncthreads();
    bv = red[0]; bi = ired[0];
    for (int w8 = 1; w8 < nw; ++w8)
      if (red[w8] > bv || (red[w8] == bv && ired[w8] < bi)) { bv = red[w8]; bi = ired[w8]; }
    amax = bi;
    __syncthreads();
  }
  // top-nsel of the scores (descending; a taken entry is retired)
  for (int r = 0; r < nsel; ++r) {
    double bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < nraw; i += nt) {
      const double sc = score[i];
      if (sc > bv || (sc == bv && i < bi)) { bv = sc; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red[wid] = bv; ired[wid] = bi; }
    __syncthreads();
    if (tid == 0) {
      bv = red[0]; bi = ired[0];
      for (int w8 = 1; w8 < nw; ++w8)
        if (red[w8] > bv || (red[w8] == bv && ired[w8] < bi)) { bv = red[w8]; bi = ired[w8]; }
      if (bi == 0x7fffffff) bi = r;            // only NaN scores left: any index
      picks[r] = bi;
      score[bi] = -INFINITY;
    }
    __syncthreads();
  }
  if (tid == 0) {
    bool have = false;
    for (int r = 0; r < nsel; ++r) have = have || picks[r] == amax;
    if (!have) picks[nsel - 1] = amax;
  }
  __syncthreads();
  for (int r = tid; r < nsel; r += nt) idx_out[(size_t)b * nsel + r] = picks[r];
  for (int e = tid; e < nsel * kmax; e += nt) {
    const int r = e / kmax, c = e % kmax;
    ics[((size_t)b * nsel + r) * kmax + c] = X[((size_t)b * nraw + picks[r]) * kmax + c];
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

extern "C" int pcabo_sobol_scramble_bits(int B, int kmax, const int *k, unsigned *mt_state, int *mt_pos, int *shift,
                                        int *ltm_dots, void *stream) {
  if (B <= 0 || kmax <= 0 || !mt_state || !mt_pos || !shift || !ltm_dots) return PCABO_E_INVALID;
  if (kmax > 256) return PCABO_E_UNSUPPORTED;
  const size_t smem = (size_t)kmax * (MAXBIT + 1) * sizeof(int);
  mt_scramble_bits_kernel<<<B, 256, smem, as_stream(stream)>>>(kmax, k, mt_state, mt_pos, shift, ltm_dots);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}

extern "C" int pcabo_select_initial_conditions(int B, int nraw, int nsel, int kmax, double eta, const double *vals,
                                               unsigned *mt_state, int *mt_pos, const double *X, double *ics,
                                               int *idx, int *flag, void *stream) {
  if (B <= 0 || nraw < 2 || nsel <= 0 || nsel > 32 || nsel >= nraw || kmax <= 0 || !vals || !mt_state || !mt_pos || !X ||
      !ics || !idx || !flag)
    return PCABO_E_INVALID;
  const size_t smem = (size_t)nraw * sizeof(double) + 2 * (size_t)nraw * sizeof(unsigned);
  if (smem > 40 * 1024) return PCABO_E_UNSUPPORTED;
  ic_select_kernel<<<B, 256, smem, as_stream(stream)>>>(nraw, nsel, kmax, eta, vals, mt_state, mt_pos, X, ics, idx, flag);
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
