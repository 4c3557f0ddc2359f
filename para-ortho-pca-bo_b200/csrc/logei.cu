// K6/K7: log-EI value (+ gradient, posterior mean / sigma) at batches of raw-space points, and row-wise argmax.
//
// Replaces botorch LogExpectedImprovement(model, best_f, maximize)(X) as the reference builds it at
// PCA_BO.py:199-203 (class map PCA_BO.py:679-686) and evaluates it inside optimize_acqf (PCA_BO.py:607-614),
// including the gradient that botorch obtains by autograd in gen_candidates_scipy. Semantics: SURVEY App. A.3/A.4.
//   xn = (x - off) / coef ; d_i = sqrt(max(|zn_i - xn|^2 / ell^2, 1e-30)) ; k_i = (1 + s5 d + 5/3 d^2) exp(-s5 d)
//   mu = ymean + ystd * k.alpha ; var = max(ystd^2 (k** - |W k|^2), 1e-10), then max(., 1e-12) ; sigma = sqrt(var)
//   u = +-(mu - best_f) / sigma ; logEI = log_h(u) + log(sigma)   (three-branch log_h of botorch _log_ei_helper)
// One CTA handles P points of one run: k-vector build (thread per training point, coalesced ZnT), t = W k
// (warp per row of the lower-triangular W, coalesced), optional s = W^T t (thread per column) for the gradient.
#include "common.cuh"

namespace {

constexpr double SQRT5 = 2.23606797749978969640917366873128;
constexpr double INV_SQRT_2PI = 0.39894228040143267793994605993438;
constexpr double INV_SQRT2 = 0.70710678118654752440084436210485;
constexpr double LOG_2PI = 1.8378770664093454835606594728112;
constexpr double LOG_SQRT_PI_DIV_2 = 0.22579135264472743236309761494744;  // log(pi/2)/2
constexpr double SQRT_PI_DIV_2 = 1.2533141373155002512078826424055;
constexpr double LOG2 = 0.69314718055994530941723212145818;
constexpr int LE_THREADS = 256;

// log_h(u) and d log_h / du, botorch.acquisition.analytic._log_ei_helper in float64
__device__ __forceinline__ void log_h_and_grad(double u, double &lh, double &dlh) {
  if (u > -1.0) {
    const double phi = INV_SQRT_2PI * exp(-0.5 * u * u);
    const double Phi = 0.5 * erfc(-INV_SQRT2 * u);
    const double h = phi + u * Phi;
    lh = log(h);
    dlh = Phi / h;
  } else if (u > -1e6) {
    const double E = erfcx(-INV_SQRT2 * u);
    const double w = log(E * fabs(u)) + LOG_SQRT_PI_DIV_2;
    const double em1 = expm1(w);  // = -(h / phi)
    const double l1me = (w > -LOG2) ? log(-em1) : log1p(-exp(w));
    lh = -0.5 * (u * u + LOG_2PI) + l1me;
    dlh = (E * SQRT_PI_DIV_2) / (-em1);  // Phi / h = (Phi / phi) / (h / phi)
  } else {
    lh = -0.5 * (u * u + LOG_2PI) - 2.0 * log(fabs(u));
    dlh = -u - 2.0 / u;
  }
}

struct LogEiArgs {
  pcabo_gp_model m;
  int npts;
  const double *X;
  double *val, *grad, *mu, *sigma;
};

template <int P, bool GRAD>
__global__ void __launch_bounds__(LE_THREADS) logei_kernel(LogEiArgs a) {
  extern __shared__ __align__(16) double sm[];
  const pcabo_gp_model &m = a.m;
  const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  const int n = m.n, kmax = m.kmax, ld = m.ld_n;
  const int kk = m.k ? m.k[b] : kmax;
  const int np = (n + 1) & ~1;
  const int p0 = chunk * P;
  const int pcount = min(P, a.npts - p0);
  double *xn = sm;                         // [P][kmax]
  double *kst = xn + P * kmax;             // [P][np]   k vector, later c_i
  double *tv = kst + P * np;               // [P][np]   t = W k            (GRAD only)
  double *gf = tv + (GRAD ? P * np : 0);   // [P][np]   d k / d(dist^2) factor (GRAD only)
  double *red = gf + (GRAD ? P * np : 0);  // [8 warps][2P] + [4P]
  const double *ZnT = m.ZnT + (size_t)b * kmax * ld;
  const double *W = m.W + (size_t)b * ld * ld;
  const double *alpha = m.alpha + (size_t)b * ld;
  const double *off = m.norm_off + (size_t)b * kmax;
  const double *coef = m.norm_coef + (size_t)b * kmax;
  const double *Xb = a.X + ((size_t)b * a.npts + p0) * kmax;
  const double inv_ell2 = 1.0 / (m.lengthscale * m.lengthscale);
  const int warp = tid >> 5, lane = tid & 31, nw = LE_THREADS >> 5;

  for (int idx = tid; idx < P * kmax; idx += LE_THREADS) {
    const int p = idx / kmax, c = idx % kmax;
    xn[idx] = (p < pcount && c < kk) ? (Xb[(size_t)p * kmax + c] - off[c]) / coef[c] : 0.0;
  }
  __syncthreads();
  // ---- k vector (and gradient factor); partial mean
  double mpart[P];
#pragma unroll
  for (int p = 0; p < P; ++p) mpart[p] = 0.0;
  for (int i = tid; i < n; i += LE_THREADS) {
    double sq[P];
#pragma unroll
    for (int p = 0; p < P; ++p) sq[p] = 0.0;
    for (int c = 0; c < kk; ++c) {
      const double z = ZnT[(size_t)c * ld + i];
#pragma unroll
      for (int p = 0; p < P; ++p) { const double d = z - xn[p * kmax + c]; sq[p] = fma(d, d, sq[p]); }
    }
    const double al = alpha[i];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const double s2 = sq[p] * inv_ell2;
      const double d = sqrt(fmax(s2, 1e-30));
      const double e = exp(-SQRT5 * d);
      const double kv = (SQRT5 * d + 1.0 + (5.0 / 3.0) * (d * d)) * e;
      kst[p * np + i] = kv;
      mpart[p] = fma(kv, al, mpart[p]);
      if (GRAD) gf[p * np + i] = (s2 > 1e-30) ? -(5.0 / 3.0) * (1.0 + SQRT5 * d) * e * inv_ell2 : 0.0;
    }
  }
  __syncthreads();
  // ---- t = W k (lower triangular), quadratic form v = |t|^2
  double vpart[P];
#pragma unroll
  for (int p = 0; p < P; ++p) vpart[p] = 0.0;
  for (int i = warp; i < n; i += nw) {
    double acc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = 0.0;
    const double *wrow = W + (size_t)i * ld;
    for (int j = lane; j <= i; j += 32) {
      const double wv = wrow[j];
#pragma unroll
      for (int p = 0; p < P; ++p) acc[p] = fma(wv, kst[p * np + j], acc[p]);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const double t = warp_sum(acc[p]);
      vpart[p] = fma(t, t, vpart[p]);
      if (GRAD && lane == 0) tv[p * np + i] = t;
    }
  }
  // ---- block reductions of the mean partials (per thread) and the variance partials (per warp, lanes equal)
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const double ms = warp_sum(mpart[p]);
    if (lane == 0) { red[warp * 2 * P + p] = ms; red[warp * 2 * P + P + p] = vpart[p]; }
  }
  __syncthreads();
  double *res = red + nw * 2 * P;  // [4][P]: A coefficient, B coefficient, (unused), (unused)
  if (tid < P) {
    const int p = tid;
    double ms = 0.0, v = 0.0;
    for (int w8 = 0; w8 < nw; ++w8) { ms += red[w8 * 2 * P + p]; v += red[w8 * 2 * P + P + p]; }
    // k(x*, x*) as gpytorch evaluates it: distance clamped to sqrt(1e-30)
    const double d0 = 1e-15;
    const double kss = (SQRT5 * d0 + 1.0 + (5.0 / 3.0) * (d0 * d0)) * exp(-SQRT5 * d0);
    const double ystd = m.ystd[b];
    const double mu = m.ymean[b] + ystd * ms;
    const double var_raw = ystd * ystd * (kss - v);
    const bool clamped = !(var_raw >= 1e-10);
    const double var = clamped ? 1e-10 : var_raw;
    const double sg = sqrt(var);
    const double sgn = m.maximize ? 1.0 : -1.0;
    const double u = sgn * (mu - m.best_f[b]) / sg;
    double lh, dlh;
    log_h_and_grad(u, lh, dlh);
    if (p < pcount) {
      const size_t o = (size_t)b * a.npts + p0 + p;
      a.val[o] = lh + log(sg);
      if (a.mu) a.mu[o] = mu;
      if (a.sigma) a.sigma[o] = sg;
    }
    if (GRAD) {
      res[p] = dlh * sgn * ystd / sg;                                            // d logEI / d mu_s
      res[P + p] = clamped ? 0.0 : (1.0 - u * dlh) * ystd * ystd / (2.0 * var);  // d logEI / d var_s
    }
  }
  if (!GRAD) return;
  __syncthreads();
  // ---- s = W^T t, then c_i = gf_i (A alpha_i - 2 B s_i)  (stored over the k vector)
  for (int j = tid; j < n; j += LE_THREADS) {
    double acc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = 0.0;
    for (int i = j; i < n; ++i) {
      const double wv = W[(size_t)i * ld + j];
#pragma unroll
      for (int p = 0; p < P; ++p) acc[p] = fma(wv, tv[p * np + i], acc[p]);
    }
    const double al = alpha[j];
#pragma unroll
    for (int p = 0; p < P; ++p) kst[p * np + j] = gf[p * np + j] * (res[p] * al - 2.0 * res[P + p] * acc[p]);
  }
  __syncthreads();
  // ---- grad[p][c] = (1 / coef_c) sum_i c_i (xn_c - zn_ic), warp per (p, c)
  for (int pc = warp; pc < pcount * kmax; pc += nw) {
    const int p = pc / kmax, c = pc % kmax;
    double acc = 0.0;
    if (c < kk) {
      const double xc = xn[p * kmax + c];
      for (int i = lane; i < n; i += 32) acc = fma(kst[p * np + i], xc - ZnT[(size_t)c * ld + i], acc);
      acc = warp_sum(acc);
    }
    if (lane == 0) a.grad[((size_t)b * a.npts + p0 + p) * kmax + c] = (c < kk) ? acc / coef[c] : 0.0;
  }
}

template <int P, bool GRAD>
int launch_logei(const LogEiArgs &a, cudaStream_t st) {
  const int np = (a.m.n + 1) & ~1;
  const size_t smem = ((size_t)P * a.m.kmax + (size_t)(GRAD ? 3 : 1) * P * np + 8 * 2 * P + 4 * P + 8) * sizeof(double);
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
      !m->ZnT || !m->alpha || !m->W || !m->ymean || !m->ystd || !m->best_f || !(m->lengthscale > 0.0))
    return PCABO_E_INVALID;
  if (m->kmax > 104 || m->n > 4096) return PCABO_E_UNSUPPORTED;
  LogEiArgs a;
  a.m = *m; a.npts = npts; a.X = X; a.val = val; a.grad = grad; a.mu = mu; a.sigma = sigma;
  cudaStream_t st = as_stream(stream);
  const int np = (m->n + 1) & ~1;
  if (grad) {
    // 3 P np doubles of shared memory: pick the largest P that fits and does not over-pad small batches
    if (npts > 8 && (size_t)3 * 16 * np * 8 + 16 * m->kmax * 8 < 200 * 1024) return launch_logei<16, true>(a, st);
    if (npts > 4 && (size_t)3 * 8 * np * 8 + 8 * m->kmax * 8 < 200 * 1024) return launch_logei<8, true>(a, st);
    return launch_logei<4, true>(a, st);
  }
  if (npts > 8 && (size_t)16 * np * 8 + 16 * m->kmax * 8 < 200 * 1024) return launch_logei<16, false>(a, st);
  if (npts > 4) return launch_logei<8, false>(a, st);
  return launch_logei<4, false>(a, st);
}

extern "C" int pcabo_argmax(int B, int mcols, const double *v, double *best_val, int *best_idx, void *stream) {
  if (B <= 0 || mcols <= 0 || !v || !best_val || !best_idx) return PCABO_E_INVALID;
  argmax_kernel<<<B, 256, 0, as_stream(stream)>>>(mcols, v, best_val, best_idx);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}
