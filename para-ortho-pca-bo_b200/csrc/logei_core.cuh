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
  int n, kmax, ld, kk, np, acq;
  const double *ZnT, *Zn, *W, *WT, *alpha, *off, *coef;
  double inv_ell2, ymean, ystd, best_f, sgn;
  double *xn;    // [P][kmax]
  double *kst;   // [P][np]  k vector, later the gradient coefficients c_i
  double *tv;    // [P][np]  t = W k
  double *gf;    // [P][np]  d k / d(dist^2) factor (GRAD)
  double *part;  // [256 P]  split-K partial sums of the gradient pass (aliases tv)
  double *red;   // [8][2P] + [4P]
  long long *prof = nullptr;  // optional cycle counters of the evaluation phases (slots 16..22; slot 31 = last stamp)
};

// attribute the cycles since the previous stamp to evaluation phase `id` (thread 0 only; no-op without counters)
__device__ __forceinline__ void ev_tick(const Ctx &c, int id) {
  if (c.prof != nullptr && threadIdx.x == 0) {
    const long long t = clock64();
    if (id >= 0) c.prof[id] += t - c.prof[31];
    c.prof[31] = t;
  }
}

constexpr int RB = 64;               // row / column block of the triangular passes
constexpr int NG = LE_THREADS / RB;  // split-K groups

// tv (t = W k) is dead once s = W^T t is formed, so the split-K partials of the gradient pass share its storage.
__host__ __device__ inline size_t tv_part_doubles(int P, int np, bool grad) {
  const size_t tvd = (size_t)P * np, pd = grad ? (size_t)LE_THREADS * P : 0;
  return tvd > pd ? tvd : pd;
}

__host__ __device__ inline size_t scratch_doubles(int P, int kmax, int n, bool grad) {
  const int np = (n + 1) & ~1;
  return (size_t)P * kmax + (size_t)(grad ? 2 : 1) * P * np + tv_part_doubles(P, np, grad) + 8 * 2 * P + 4 * P + 8;
}

__device__ inline void carve_ctx(Ctx &c, int P, bool grad, double *sm) {
  c.xn = sm;
  c.kst = c.xn + P * c.kmax;
  c.gf = c.kst + P * c.np;
  c.tv = c.gf + (grad ? P * c.np : 0);
  c.part = c.tv;
  c.red = c.tv + tv_part_doubles(P, c.np, grad);
}

__device__ inline void load_ctx(Ctx &c, const pcabo_gp_model &m, int b) {
  c.n = m.n; c.kmax = m.kmax; c.ld = m.ld_n; c.acq = m.acq;
  c.kk = m.k ? m.k[b] : m.kmax;
  c.np = (m.n + 1) & ~1;
  c.ZnT = m.ZnT + (size_t)b * m.kmax * m.ld_n;
  c.Zn = m.Zn + (size_t)b * m.ld_n * m.kmax;
  c.W = m.W + (size_t)b * m.ld_n * m.ld_n;
  c.WT = m.WT + (size_t)b * m.ld_n * m.ld_n;
  c.alpha = m.alpha + (size_t)b * m.ld_n;
  c.off = m.norm_off + (size_t)b * m.kmax;
  c.coef = m.norm_coef + (size_t)b * m.kmax;
  { const double ell = m.lengthscales ? m.lengthscales[b] : m.lengthscale; c.inv_ell2 = 1.0 / (ell * ell); }
  c.ymean = m.ymean[b]; c.ystd = m.ystd[b]; c.best_f = m.best_f[b];
  c.sgn = m.maximize ? 1.0 : -1.0;
}

// acc[p] += sum_{j in [j0, j1)} M[j * ld + col] * v[p * np + j], loads issued eight at a time (the chain is latency bound)
template <int P>
__device__ __forceinline__ void strided_dot(double (&acc)[P], const double *__restrict__ M, long ld, int col, int j0,
                                            int j1, const double *__restrict__ v, int np) {
  int j = j0;
  for (; j + 16 <= j1; j += 16) {
    double w16[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) w16[u] = M[(long)(j + u) * ld + col];
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int p = 0; p < P; ++p) acc[p] = fma(w16[u], v[p * np + j + u], acc[p]);
  }
  for (; j + 4 <= j1; j += 4) {
    double w4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) w4[u] = M[(long)(j + u) * ld + col];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int p = 0; p < P; ++p) acc[p] = fma(w4[u], v[p * np + j + u], acc[p]);
  }
  for (; j < j1; ++j) {
    const double wv = M[(long)j * ld + col];
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = fma(wv, v[p * np + j], acc[p]);
  }
}

// Evaluate pcount <= P points. X: raw-space coordinates, point p at X + p * xstride (first kk entries used).
// val[p] (p < pcount) <- log-EI; if GRAD: grad[p * gstride + j] <- gscale * d logEI / d x_j for j < kk and 0 for
// kk <= j < gfill. mu / sigma optional. All LE_THREADS threads of the block must call; ends with __syncthreads().
//
// Every heavy phase is a thread-per-output loop over coalesced global loads with no cross-lane reduction:
//   t = W k     : thread (g, r) owns row i of a 64-row block and the g-th quarter of j <= i, reading WT[j][i]
//   s = W^T t   : thread (g, c) owns column j of a 64-column block and a quarter of i >= j, reading W[i][j]
//   gradient    : thread (ig, j) owns dimension j and every IG-th training point, reading Zn[i][j]
// the quarters are combined in a fixed order through shared memory (deterministic).
template <int P, bool GRAD>
__device__ void eval_points(const Ctx &c, int pcount, const double *X, int xstride, double *val, double *grad,
                            int gstride, int gfill, double gscale, double *mu_out, double *sigma_out) {
  const int tid = threadIdx.x;
  const int n = c.n, kmax = c.kmax, kk = c.kk, np = c.np;
  const long ld = c.ld;
  const int warp = tid >> 5, lane = tid & 31, nw = LE_THREADS >> 5;
  double *xn = c.xn, *kst = c.kst, *tv = c.tv, *gf = c.gf, *part = c.part, *red = c.red;
  const double *ZnT = c.ZnT, *alpha = c.alpha;
  const double inv_ell2 = c.inv_ell2;

  ev_tick(c, -1);
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
    int j = 0;
    for (; j + 4 <= kk; j += 4) {
      double z4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) z4[u] = ZnT[(size_t)(j + u) * ld + i];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int p = 0; p < P; ++p) { const double d = z4[u] - xn[p * kmax + j + u]; sq[p] = fma(d, d, sq[p]); }
    }
    for (; j < kk; ++j) {
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
  ev_tick(c, 16);
  // ---- t = W k (lower triangular): every thread owns whole rows (no cross-thread combine, no barriers inside the
  //      pass). Row blocks of RB rows are dealt to the NG thread groups longest-with-shortest so the triangular work
  //      balances; within a group the RB lanes read WT[j][i0 .. i0 + RB) for one j: coalesced.
  const int g = tid / RB, r = tid % RB;
  const int nblk = (n + RB - 1) / RB;
  double vpart[P];
#pragma unroll
  for (int p = 0; p < P; ++p) vpart[p] = 0.0;
  for (int base = 0, round = 0; base < nblk; base += NG, ++round) {
    // this group's block of the round: alternate between the long end and the short end of the triangle
    const int blk = base + ((round & 1) ? (NG - 1 - g) : g);
    if (blk >= nblk) continue;
    const int i = blk * RB + r;
    if (i < n) {
      double acc[P];
#pragma unroll
      for (int p = 0; p < P; ++p) acc[p] = 0.0;
      // every lane runs to the end of its row block (WT is zero below its diagonal): uniform trip count, no tail
      // divergence inside the warp
      strided_dot<P>(acc, c.WT, ld, i, 0, min(blk * RB + RB, n), kst, np);
#pragma unroll
      for (int p = 0; p < P; ++p) tv[p * np + i] = acc[p];
    }
  }
  __syncthreads();
  ev_tick(c, 17);
  for (int i = tid; i < n; i += LE_THREADS) {
#pragma unroll
    for (int p = 0; p < P; ++p) { const double t = tv[p * np + i]; vpart[p] = fma(t, t, vpart[p]); }
  }
  // ---- block reductions of the mean and variance partials
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const double ms = warp_sum(mpart[p]);
    const double vs = warp_sum(vpart[p]);
    if (lane == 0) { red[warp * 2 * P + p] = ms; red[warp * 2 * P + P + p] = vs; }
  }
  __syncthreads();
  ev_tick(c, 21);
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
    double value, dmu, dvar;   // acquisition value, d value / d mu_s, d value / d var_s
    if (c.acq == PCABO_ACQ_PI) {
      // botorch ProbabilityOfImprovement: Phi(u); d/du = phi(u); du/dmu = sgn/sigma, du/dvar = -u / (2 var)
      value = 0.5 * erfc(-INV_SQRT2 * u);
      const double phi = INV_SQRT_2PI * exp(-0.5 * u * u);
      dmu = phi * c.sgn * ystd / sg;
      dvar = clamped ? 0.0 : -phi * u * ystd * ystd / (2.0 * var);
    } else {
      double lh, dlh;
      log_h_and_grad(u, lh, dlh);
      value = lh + log(sg);
      dmu = dlh * c.sgn * ystd / sg;
      dvar = clamped ? 0.0 : (1.0 - u * dlh) * ystd * ystd / (2.0 * var);
    }
    if (p < pcount) {
      val[p] = value;
      if (mu_out) mu_out[p] = mu;
      if (sigma_out) sigma_out[p] = sg;
    }
    if (GRAD) {
      res[p] = dmu;
      res[P + p] = dvar;
    }
  }
  ev_tick(c, 22);
  __syncthreads();
  ev_tick(c, 18);
  if constexpr (GRAD) {
  // ---- s = W^T t: every thread owns whole columns (W[i][j], i >= j: consecutive j coalesce), blocks dealt like above
  //      (column blocks get shorter towards the end); then c_j = gf_j (A alpha_j - 2 B s_j) stored over the k vector
  for (int base = 0, round = 0; base < nblk; base += NG, ++round) {
    const int blk = base + ((round & 1) ? (NG - 1 - g) : g);
    if (blk >= nblk) continue;
    const int j = blk * RB + r;
    if (j < n) {
      double acc[P];
#pragma unroll
      for (int p = 0; p < P; ++p) acc[p] = 0.0;
      // rows i >= j of column j; every lane starts at the first row of the column block (W is zero above its
      // diagonal) so that the lanes of a group read the same row of W in every step: coalesced 512-byte segments
      strided_dot<P>(acc, c.W + j, ld, 0, blk * RB, n, tv, np);
      const double al = alpha[j];
#pragma unroll
      for (int p = 0; p < P; ++p) kst[p * np + j] = gf[p * np + j] * (res[p] * al - 2.0 * res[P + p] * acc[p]);
    }
  }
  __syncthreads();
  ev_tick(c, 19);
  // ---- grad[p][j] = gscale (1 / coef_j) sum_i c_i (xn_j - zn_ij): thread (ig, j), every IG-th training point
  {
    const int IG = LE_THREADS / kk;  // >= 2 for kk <= 104
    const int ig = tid / kk, j = tid % kk;
    double acc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = 0.0;
    if (ig < IG) {
      double xj[P];
#pragma unroll
      for (int p = 0; p < P; ++p) xj[p] = xn[p * kmax + j];
      int i = ig;
      for (; i + 3 * IG < n; i += 4 * IG) {
        double z4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) z4[u] = c.Zn[(size_t)(i + u * IG) * kmax + j];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int p = 0; p < P; ++p) acc[p] = fma(kst[p * np + i + u * IG], xj[p] - z4[u], acc[p]);
      }
      for (; i < n; i += IG) {
        const double z = c.Zn[(size_t)i * kmax + j];
#pragma unroll
        for (int p = 0; p < P; ++p) acc[p] = fma(kst[p * np + i], xj[p] - z, acc[p]);
      }
#pragma unroll
      for (int p = 0; p < P; ++p) part[(ig * P + p) * kk + j] = acc[p];
    }
    __syncthreads();
    for (int idx = tid; idx < pcount * gfill; idx += LE_THREADS) {
      const int p = idx / gfill, jj = idx % gfill;
      double sv = 0.0;
      if (jj < kk) {
        for (int q = 0; q < IG; ++q) sv += part[(q * P + p) * kk + jj];
        sv = gscale * sv / c.coef[jj];
      }
      grad[(size_t)p * gstride + jj] = sv;
    }
  }
  __syncthreads();
  ev_tick(c, 20);
  }  // GRAD
}

}  // namespace logei
