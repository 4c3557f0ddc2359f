// Device-side log-EI evaluation shared by the stand-alone kernel (logei.cu) and the L-BFGS-B kernel (lbfgsb.cu).
//
// Replaces botorch LogExpectedImprovement(model, best_f, maximize)(X) as the reference builds it at
// PCA_BO.py:199-203 (class map PCA_BO.py:679-686) and evaluates it inside optimize_acqf (PCA_BO.py:607-614),
// including the gradient that botorch obtains by autograd in gen_candidates_scipy. Semantics: SURVEY App. A.3/A.4.
//   xn = (x - off) / coef ; d_i = sqrt(max(|zn_i - xn|^2 / ell^2, 1e-30)) ; k_i = (1 + s5 d + 5/3 d^2) exp(-s5 d)
//   mu = ymean + ystd * k.alpha ; var = max(ystd^2 (k** - |W k|^2), 1e-10), then max(., 1e-12) ; sigma = sqrt(var)
//   u = +-(mu - best_f) / sigma ; logEI = log_h(u) + log(sigma)   (three-branch log_h of botorch _log_ei_helper)
// One block handles P points of one run: k-vector build (thread per training point, coalesced ZnT), t = W k
// (warp per row of the lower-triangular W, coalesced), optional s = W^T t (thread per column) for the gradient.
#pragma once
#include "common.cuh"

namespace logei {

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

// Per-block view of one run's GP caches plus the shared-memory scratch of the evaluation.
struct Ctx {
  int n, kmax, ld, kk, np;
  const double *ZnT, *W, *alpha, *off, *coef;
  double inv_ell2, ymean, ystd, best_f, sgn;
  double *xn;   // [P][kmax]
  double *kst;  // [P][np]
  double *tv;   // [P][np]  (GRAD)
  double *gf;   // [P][np]  (GRAD)
  double *red;  // [8][2P] + [4P]
};

__host__ __device__ inline size_t scratch_doubles(int P, int kmax, int n, bool grad) {
  const int np = (n + 1) & ~1;
  return (size_t)P * kmax + (size_t)(grad ? 3 : 1) * P * np + 8 * 2 * P + 4 * P + 8;
}

__device__ inline void carve_ctx(Ctx &c, int P, bool grad, double *sm) {
  c.xn = sm;
  c.kst = c.xn + P * c.kmax;
  c.tv = c.kst + P * c.np;
  c.gf = c.tv + (grad ? P * c.np : 0);
  c.red = c.gf + (grad ? P * c.np : 0);
}

__device__ inline void load_ctx(Ctx &c, const pcabo_gp_model &m, int b) {
  c.n = m.n; c.kmax = m.kmax; c.ld = m.ld_n;
  c.kk = m.k ? m.k[b] : m.kmax;
  c.np = (m.n + 1) & ~1;
  c.ZnT = m.ZnT + (size_t)b * m.kmax * m.ld_n;
  c.W = m.W + (size_t)b * m.ld_n * m.ld_n;
  c.alpha = m.alpha + (size_t)b * m.ld_n;
  c.off = m.norm_off + (size_t)b * m.kmax;
  c.coef = m.norm_coef + (size_t)b * m.kmax;
  c.inv_ell2 = 1.0 / (m.lengthscale * m.lengthscale);
  c.ymean = m.ymean[b]; c.ystd = m.ystd[b]; c.best_f = m.best_f[b];
  c.sgn = m.maximize ? 1.0 : -1.0;
}

// Evaluate pcount <= P points. X: raw-space coordinates, point p at X + p * xstride (first kk entries used).
// val[p] (p < pcount) <- log-EI; if GRAD: grad[p * gstride + j] <- gscale * d logEI / d x_j for j < kk and 0 for
// kk <= j < gfill. mu / sigma optional. All LE_THREADS threads of the block must call; ends with __syncthreads().
template <int P, bool GRAD>
__device__ void eval_points(const Ctx &c, int pcount, const double *X, int xstride, double *val, double *grad,
                            int gstride, int gfill, double gscale, double *mu_out, double *sigma_out) {
  const int tid = threadIdx.x;
  const int n = c.n, kmax = c.kmax, ld = c.ld, kk = c.kk, np = c.np;
  const int warp = tid >> 5, lane = tid & 31, nw = LE_THREADS >> 5;
  double *xn = c.xn, *kst = c.kst, *tv = c.tv, *gf = c.gf, *red = c.red;
  const double *ZnT = c.ZnT, *W = c.W, *alpha = c.alpha;
  const double inv_ell2 = c.inv_ell2;

  for (int idx = tid; idx < P * kmax; idx += LE_THREADS) {
    const int p = idx / kmax, j = idx % kmax;
    xn[idx] = (p < pcount && j < kk) ? (X[(size_t)p * xstride + j] - c.off[j]) / c.coef[j] : 0.0;
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
    for (int j = 0; j < kk; ++j) {
      const double z = ZnT[(size_t)j * ld + i];
#pragma unroll
      for (int p = 0; p < P; ++p) { const double d = z - xn[p * kmax + j]; sq[p] = fma(d, d, sq[p]); }
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
  double *res = red + nw * 2 * P;  // [2][P]: d logEI / d mu_s, d logEI / d var_s
  if (tid < P) {
    const int p = tid;
    double ms = 0.0, v = 0.0;
    for (int w8 = 0; w8 < nw; ++w8) { ms += red[w8 * 2 * P + p]; v += red[w8 * 2 * P + P + p]; }
    // k(x*, x*) as gpytorch evaluates it: distance clamped to sqrt(1e-30)
    const double d0 = 1e-15;
    const double kss = (SQRT5 * d0 + 1.0 + (5.0 / 3.0) * (d0 * d0)) * exp(-SQRT5 * d0);
    const double ystd = c.ystd;
    const double mu = c.ymean + ystd * ms;
    const double var_raw = ystd * ystd * (kss - v);
    const bool clamped = !(var_raw >= 1e-10);
    const double var = clamped ? 1e-10 : var_raw;
    const double sg = sqrt(var);
    const double u = c.sgn * (mu - c.best_f) / sg;
    double lh, dlh;
    log_h_and_grad(u, lh, dlh);
    if (p < pcount) {
      val[p] = lh + log(sg);
      if (mu_out) mu_out[p] = mu;
      if (sigma_out) sigma_out[p] = sg;
    }
    if (GRAD) {
      res[p] = dlh * c.sgn * ystd / sg;                                          // d logEI / d mu_s
      res[P + p] = clamped ? 0.0 : (1.0 - u * dlh) * ystd * ystd / (2.0 * var);  // d logEI / d var_s
    }
  }
  __syncthreads();
  if (!GRAD) return;
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
  // ---- grad[p][j] = gscale (1 / coef_j) sum_i c_i (xn_j - zn_ij), warp per (p, j)
  for (int pc = warp; pc < pcount * gfill; pc += nw) {
    const int p = pc / gfill, j = pc % gfill;
    double acc = 0.0;
    if (j < kk) {
      const double xc = xn[p * kmax + j];
      for (int i = lane; i < n; i += 32) acc = fma(kst[p * np + i], xc - ZnT[(size_t)j * ld + i], acc);
      acc = warp_sum(acc);
    }
    if (lane == 0) grad[(size_t)p * gstride + j] = (j < kk) ? gscale * acc / c.coef[j] : 0.0;
  }
  __syncthreads();
}

}  // namespace logei
