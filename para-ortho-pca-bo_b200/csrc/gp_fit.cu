// K3-K5: GP construction + prediction caches, batched over runs.
//
// Replaces PCA_BO._initialize_model (PCA_BO.py:502-545) / Vanilla_BO._initialise_model (Vanilla_BO.py:166-195),
// i.e. botorch SingleTaskGP(train_z, train_obj, MaternKernel(2.5), Standardize(m=1), Normalize(d, bounds)) plus the
// caches gpytorch builds on the first posterior call (K + sigma^2 I -> Cholesky -> alpha = K^-1 y_s, R = L^-T), and the
// search box of PCA_BO.optimize_acqf_and_get_observation (PCA_BO.py:558-573). The GP is never fitted in the
// reference: lengthscale and noise are the fixed prior values passed by the caller.
//
//   gp_prep_kernel   : per-dim min/max -> Normalize (0.1 pad) + search box (0.5 pad, min width 0.1); Standardize;
//                      normalised inputs transposed (ZnT[c][i]) so later kernels read them coalesced.
//   gp_kmat_kernel   : lower triangle of K = matern52(|zn_i - zn_j| / ell) + sigma^2 I, 64x64 tiles.
//   (gp_chol.cu)     : Cholesky, W = L^-1, alpha = W^T (W y_s), W^T as a pipeline of tile kernels (pcabo_chol_pipeline)
//   gp_pack_w_kernel : W repacked in 8-row groups for the bulk copies of the optimiser kernel (w_packed.cuh)
#include "common.cuh"
#include "w_packed.cuh"

namespace {

constexpr double SQRT5 = 2.23606797749978969640917366873128;

// --------------------------------------------------------------------------------------------- prep
struct PrepArgs {
  int B, n, kmax, ld_n, mode;
  const double *Z;
  const int *k;
  const double *f, *fix_lo, *fix_hi;
  double *norm_off, *norm_coef, *box_lo, *box_hi, *ymean, *ystd, *ZnT, *Zn, *ys;
};

__global__ void __launch_bounds__(256) gp_prep_kernel(PrepArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int n = a.n, kmax = a.kmax;
  const int kk = a.k ? a.k[b] : kmax;
  double *pmin = sm;                 // [G][kmax]
  const int G = max(1, nt / kmax);
  double *pmax = pmin + G * kmax;    // [G][kmax]
  double *off = pmax + G * kmax;     // [kmax]
  double *coef = off + kmax;         // [kmax]
  double *red = coef + kmax;         // [34]
  const double *Zb = a.Z + (size_t)b * a.ld_n * kmax;
  const double *fb = a.f + (size_t)b * a.ld_n;
  // min / max per dimension
  {
    const int g = tid / kmax, c = tid % kmax;
    if (g < G) {
      double mn = INFINITY, mx = -INFINITY;
      if (c < kk)
        for (int i = g; i < n; i += G) {
          const double v = Zb[(size_t)i * kmax + c];
          mn = fmin(mn, v);
          mx = fmax(mx, v);
        }
      pmin[g * kmax + c] = mn;
      pmax[g * kmax + c] = mx;
    }
  }
  __syncthreads();
  if (tid < kmax) {
    const int c = tid;
    double lo_n = 0.0, cf = 1.0, blo = 0.0, bhi = 0.0;
    if (c < kk) {
      double mn = INFINITY, mx = -INFINITY;
      for (int g = 0; g < G; ++g) { mn = fmin(mn, pmin[g * kmax + c]); mx = fmax(mx, pmax[g * kmax + c]); }
      const double rng = mx - mn;
      if (a.mode == 0) {
        // PCA_BO.py:512-518 Normalize bounds, botorch Normalize: coefficient = hi - lo, |.| < 1e-12 -> 1
        lo_n = mn - 0.1 * rng;
        const double hi_n = mx + 0.1 * rng;
        cf = hi_n - lo_n;
        if (fabs(cf) < 1e-12) cf = 1.0;
        // PCA_BO.py:558-573 search box
        blo = mn - 0.5 * rng;
        bhi = mx + 0.5 * rng;
        if (bhi - blo < 0.1) {
          const double mid = (bhi + blo) / 2;
          blo = mid - 0.1 / 2;
          bhi = mid + 0.1 / 2;
        }
      } else {
        blo = a.fix_lo[c];
        bhi = a.fix_hi[c];
      }
    }
    off[c] = lo_n;
    coef[c] = cf;
    a.norm_off[(size_t)b * kmax + c] = lo_n;
    a.norm_coef[(size_t)b * kmax + c] = cf;
    a.box_lo[(size_t)b * kmax + c] = blo;
    a.box_hi[(size_t)b * kmax + c] = bhi;
  }
  // Standardize(m=1): mean, unbiased std (two pass), std < 1e-8 -> 1
  double s = 0.0;
  for (int i = tid; i < n; i += nt) s += fb[i];
  s = block_sum(s, red);
  const double mean = s / (double)n;
  double q = 0.0;
  for (int i = tid; i < n; i += nt) { const double d = fb[i] - mean; q = fma(d, d, q); }
  q = block_sum(q, red);
  double sd = (n > 1) ? sqrt(q / (double)(n - 1)) : 1.0;
  if (!(sd >= 1e-8)) sd = 1.0;
  if (tid == 0) { a.ymean[b] = mean; a.ystd[b] = sd; }
  for (int i = tid; i < n; i += nt) a.ys[(size_t)b * a.ld_n + i] = (fb[i] - mean) / sd;
  __syncthreads();
  // normalised, transposed training inputs
  double *ZnTb = a.ZnT + (size_t)b * kmax * a.ld_n;
  for (int idx = tid; idx < kk * n; idx += nt) {
    const int c = idx / n, i = idx % n;
    ZnTb[(size_t)c * a.ld_n + i] = (Zb[(size_t)i * kmax + c] - off[c]) / coef[c];
  }
  double *Znb = a.Zn + (size_t)b * a.ld_n * kmax;
  for (int idx = tid; idx < n * kmax; idx += nt) {
    const int c = idx % kmax;
    Znb[idx] = (c < kk) ? (Zb[idx] - off[c]) / coef[c] : 0.0;
  }
}

// --------------------------------------------------------------------------------------------- kernel matrix
__device__ __forceinline__ double matern52_from_sq(double sq, double inv_ell2) {
  // gpytorch: dist = sqrt(clamp_min(sq / ell^2, 1e-30)); (1 + sqrt5 d + 5/3 d^2) exp(-sqrt5 d)
  const double d = sqrt(fmax(sq * inv_ell2, 1e-30));
  return (SQRT5 * d + 1.0 + (5.0 / 3.0) * (d * d)) * exp(-SQRT5 * d);
}

constexpr int KT = 64;  // kernel-matrix tile

__global__ void __launch_bounds__(256) gp_kmat_kernel(int n, int kmax, int ld_n, const int *k, const double *ZnT,
                                                     double inv_ell2, double noise, const double *jitter,
                                                     const double *ls_run, const double *nz_run, double *Kout) {
  // grid: (tiles_lower, B); tile (ti >= tj)
  __shared__ double zi[8][KT];
  __shared__ double zj[8][KT];
  const int b = blockIdx.y;
  int t = blockIdx.x, ti = 0;
  while (t >= ti + 1) { t -= ti + 1; ++ti; }
  const int tj = t;
  const int kk = k ? k[b] : kmax;
  const double *Zb = ZnT + (size_t)b * kmax * ld_n;
  double *Kb = Kout + (size_t)b * ld_n * ld_n;
  // per-run hyper-parameters (opt-in fit) override the reference's fixed prior values
  if (ls_run) inv_ell2 = 1.0 / (ls_run[b] * ls_run[b]);
  const double diag_add = (nz_run ? nz_run[b] : noise) + (jitter ? jitter[b] : 0.0);   // psd_safe_cholesky adds its jitter to the diagonal
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;  // thread computes rows ty*4.., cols tx*4..
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int c0 = 0; c0 < kk; c0 += 8) {
    for (int idx = threadIdx.x; idx < 8 * KT; idx += blockDim.x) {
      const int c = idx / KT, r = idx % KT;
      const int gi = ti * KT + r, gj = tj * KT + r;
      const bool vc = (c0 + c) < kk;
      zi[c][r] = (vc && gi < n) ? Zb[(size_t)(c0 + c) * ld_n + gi] : 0.0;
      zj[c][r] = (vc && gj < n) ? Zb[(size_t)(c0 + c) * ld_n + gj] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      double av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = zi[c][ty * 4 + i]; bv[i] = zj[c][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double d = av[i] - bv[j]; acc[i][j] = fma(d, d, acc[i][j]); }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gi = ti * KT + ty * 4 + i, gj = tj * KT + tx * 4 + j;
      if (gi < n && gj <= gi) {
        double v = matern52_from_sq(acc[i][j], inv_ell2);
        if (gi == gj) v += diag_add;
        Kb[(size_t)gi * ld_n + gj] = v;
      }
    }
}

// W (row-major lower triangle) -> the packed 8-row groups the optimiser kernel streams (w_packed.cuh). grid (groups, B).
__global__ void __launch_bounds__(256) gp_pack_w_kernel(int n, int ld_n, const double *W, double *Wp, long long wp_stride) {
  const int g = blockIdx.x, b = blockIdx.y;
  const int RS = lbeval::grp_rs(g);
  const double *Wb = W + (size_t)b * ld_n * ld_n;
  double *out = Wp + (size_t)b * (size_t)wp_stride + lbeval::grp_off(g);
  for (int idx = threadIdx.x; idx < 8 * RS; idx += blockDim.x) {
    const int r = idx / RS, c = idx % RS, i = 8 * g + r;
    out[idx] = (i < n && c <= i) ? Wb[(size_t)i * ld_n + c] : 0.0;
  }
}

__global__ void gp_mll_kernel(int n, int ld_n, const double *f, const double *ymean, const double *ystd,
                              const double *L, const double *alpha, double *mll) {
  __shared__ double red[34];
  const int b = blockIdx.x;
  const double mean = ymean[b], sd = ystd[b];
  double quad = 0.0, logdet = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double ys = (f[(size_t)b * ld_n + i] - mean) / sd;
    quad = fma(ys, alpha[(size_t)b * ld_n + i], quad);
    logdet += log(L[(size_t)b * ld_n * ld_n + (size_t)i * ld_n + i]);
  }
  quad = block_sum(quad, red);
  logdet = block_sum(logdet, red);
  if (threadIdx.x == 0) mll[b] = -0.5 * quad - logdet - 0.5 * (double)n * 1.8378770664093453;  // log(2 pi)
}

// --------------------------------------------------------------------------------------------- MLL gradient
// d MLL / d theta = 1/2 alpha^T (dK/dtheta) alpha - 1/2 tr(K^-1 dK/dtheta) for theta = lengthscale and noise, with
// K^-1 = W^T W never formed in memory: one block per 32 x 32 tile (ti >= tj) of the lower triangle builds its tile of
// K^-1 as the product of two column panels of W (rows k >= ti * 32, where W is non-zero) staged through shared memory,
// multiplies it by the tile of dK/dl computed on the fly from the normalised inputs, and adds its share of both traces
// to the per-run outputs. Matern-5/2: k(a) = (1 + a + a^2/3) e^-a, a = sqrt5 r / l  =>  dk/dl = a^2 (1 + a) e^-a / (3 l).
constexpr int GT = 32;   // tile
constexpr int GK = 32;   // rows of W per staging step

__global__ void __launch_bounds__(256) gp_mll_grad_kernel(int n, int kmax, int ld_n, const int *k, const double *Zn,
                                                         const double *W, const double *alpha, double lengthscale,
                                                         const double *ls_run,
                                                         double *grad /* [B][2]: d/dl, d/dnoise */) {
  __shared__ double wi[GK][GT + 1], wj[GK][GT + 1];
  __shared__ double red[34];
  const int b = blockIdx.y;
  int t = blockIdx.x, ti = 0;
  while (t >= ti + 1) { t -= ti + 1; ++ti; }
  const int tj = t;
  const int kk = k ? k[b] : kmax;
  const double *Wb = W + (size_t)b * ld_n * ld_n;
  const double *Zb = Zn + (size_t)b * ld_n * kmax;
  const double *al = alpha + (size_t)b * ld_n;
  const int tid = threadIdx.x;
  // thread owns the 2 x 2 block (r0, r0 + 1) x (c0, c0 + 1) of the tile
  const int r0 = (tid / 16) * 2, c0 = (tid % 16) * 2;
  double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  const int i0 = ti * GT, j0 = tj * GT;
  for (int k0 = i0; k0 < n; k0 += GK) {     // W[k][i] = 0 for k < i: start at the tile's first row index
    for (int idx = tid; idx < GK * GT; idx += 256) {
      const int kr = idx / GT, c = idx % GT;
      const int kg = k0 + kr;
      double a = 0.0, bb = 0.0;
      if (kg < n) {
        const int gi = i0 + c, gj = j0 + c;
        if (gi < n && gi <= kg) a = Wb[(size_t)kg * ld_n + gi];
        if (gj < n && gj <= kg) bb = Wb[(size_t)kg * ld_n + gj];
      }
      wi[kr][c] = a;
      wj[kr][c] = bb;
    }
    __syncthreads();
#pragma unroll 8
    for (int kr = 0; kr < GK; ++kr) {
      const double a0 = wi[kr][r0], a1 = wi[kr][r0 + 1], b0 = wj[kr][c0], b1 = wj[kr][c0 + 1];
      acc[0][0] = fma(a0, b0, acc[0][0]);
      acc[0][1] = fma(a0, b1, acc[0][1]);
      acc[1][0] = fma(a1, b0, acc[1][0]);
      acc[1][1] = fma(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
  // contributions of this tile: off-diagonal tiles stand for both (i, j) and (j, i)
  const double mult = (ti == tj) ? 1.0 : 2.0;
  const double inv_l = 1.0 / (ls_run ? ls_run[b] : lengthscale);
  double g_l = 0.0, g_n = 0.0;
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int gi = i0 + r0 + u, gj = j0 + c0 + v;
      if (gi < n && gj < n) {
        double sq = 0.0;
        for (int d = 0; d < kk; ++d) {
          const double df = Zb[(size_t)gi * kmax + d] - Zb[(size_t)gj * kmax + d];
          sq = fma(df, df, sq);
        }
        const double a = SQRT5 * sqrt(fmax(sq * inv_l * inv_l, 1e-30));
        const double dk = (gi == gj) ? 0.0 : a * a * (1.0 + a) * exp(-a) * inv_l / 3.0;
        g_l += mult * 0.5 * (al[gi] * al[gj] - acc[u][v]) * dk;
        if (gi == gj) g_n += 0.5 * (al[gi] * al[gi] - acc[u][v]);     // dK/dnoise = I
      }
    }
  g_l = block_sum(g_l, red);
  g_n = block_sum(g_n, red);
  if (tid == 0) {
    atomicAdd(&grad[2 * b + 0], g_l);
    atomicAdd(&grad[2 * b + 1], g_n);
  }
}

}  // namespace

extern "C" int pcabo_gp_mll_grad_hyper(int B, int n, int kmax, int ld_n, const int *k, const double *Zn, const double *W,
                                       const double *alpha, double lengthscale, const double *lengthscales, double *grad,
                                       void *stream) {
  if (B <= 0 || n < 1 || kmax < 1 || ld_n < n || !Zn || !W || !alpha || !grad || (!lengthscales && !(lengthscale > 0.0)))
    return PCABO_E_INVALID;
  cudaStream_t st = as_stream(stream);
  PCABO_CUDA_CHECK(cudaMemsetAsync(grad, 0, sizeof(double) * 2 * B, st));
  const int nt = (n + GT - 1) / GT;
  dim3 grid(nt * (nt + 1) / 2, B);
  gp_mll_grad_kernel<<<grid, 256, 0, st>>>(n, kmax, ld_n, k, Zn, W, alpha, lengthscale, lengthscales, grad);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}

extern "C" int pcabo_gp_mll_grad(int B, int n, int kmax, int ld_n, const int *k, const double *Zn, const double *W,
                                 const double *alpha, double lengthscale, double *grad, void *stream) {
  return pcabo_gp_mll_grad_hyper(B, n, kmax, ld_n, k, Zn, W, alpha, lengthscale, nullptr, grad, stream);
}

extern "C" int pcabo_gp_fit_hyper(int B, int n, int kmax, int ld_n, int mode, const double *Z, const int *k,
                                  const double *f, const double *fix_lo, const double *fix_hi, double lengthscale,
                                  double noise, const double *lengthscales, const double *noises, const double *jitter,
                                  double *norm_off, double *norm_coef, double *box_lo, double *box_hi, double *ymean,
                                  double *ystd, double *ZnT, double *Zn, double *L, double *W, double *WT, double *alpha,
                                  int *info, void *stream) {
  if (B <= 0 || n < 1 || kmax < 1 || ld_n < n || !Z || !f || !norm_off || !norm_coef || !box_lo || !box_hi || !ymean ||
      !ystd || !ZnT || !Zn || !L || !W || !WT || !alpha || !info || (!lengthscales && !(lengthscale > 0.0)))
    return PCABO_E_INVALID;
  if (mode != 0 && (!fix_lo || !fix_hi)) return PCABO_E_INVALID;
  if (kmax > 104 || n > 4096) return PCABO_E_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  {
    PrepArgs a;
    a.B = B; a.n = n; a.kmax = kmax; a.ld_n = ld_n; a.mode = mode;
    a.Z = Z; a.k = k; a.f = f; a.fix_lo = fix_lo; a.fix_hi = fix_hi;
    a.norm_off = norm_off; a.norm_coef = norm_coef; a.box_lo = box_lo; a.box_hi = box_hi;
    a.ymean = ymean; a.ystd = ystd; a.ZnT = ZnT; a.Zn = Zn; a.ys = alpha;
    const int G = 256 / kmax > 0 ? 256 / kmax : 1;
    const size_t smem = (size_t)(2 * G * kmax + 2 * kmax + 34) * sizeof(double);
    gp_prep_kernel<<<B, 256, smem, st>>>(a);
    PCABO_LAUNCH_CHECK();
  }
  {
    const int nt = (n + KT - 1) / KT;
    dim3 grid(nt * (nt + 1) / 2, B);
    gp_kmat_kernel<<<grid, 256, 0, st>>>(n, kmax, ld_n, k, ZnT, lengthscales ? 1.0 : 1.0 / (lengthscale * lengthscale), noise,
                                         jitter, lengthscales, noises, L);
    PCABO_LAUNCH_CHECK();
  }
  return pcabo_chol_pipeline(B, n, ld_n, L, W, WT, alpha, info, st);
}

extern "C" int pcabo_gp_fit_jitter(int B, int n, int kmax, int ld_n, int mode, const double *Z, const int *k,
                                   const double *f, const double *fix_lo, const double *fix_hi, double lengthscale,
                                   double noise, const double *jitter, double *norm_off, double *norm_coef,
                                   double *box_lo, double *box_hi, double *ymean, double *ystd, double *ZnT, double *Zn,
                                   double *L, double *W, double *WT, double *alpha, int *info, void *stream) {
  return pcabo_gp_fit_hyper(B, n, kmax, ld_n, mode, Z, k, f, fix_lo, fix_hi, lengthscale, noise, nullptr, nullptr, jitter,
                            norm_off, norm_coef, box_lo, box_hi, ymean, ystd, ZnT, Zn, L, W, WT, alpha, info, stream);
}

extern "C" int pcabo_gp_fit(int B, int n, int kmax, int ld_n, int mode, const double *Z, const int *k, const double *f,
                            const double *fix_lo, const double *fix_hi, double lengthscale, double noise,
                            double *norm_off, double *norm_coef, double *box_lo, double *box_hi, double *ymean,
                            double *ystd, double *ZnT, double *Zn, double *L, double *W, double *WT, double *alpha,
                            int *info, void *stream) {
  return pcabo_gp_fit_jitter(B, n, kmax, ld_n, mode, Z, k, f, fix_lo, fix_hi, lengthscale, noise, nullptr, norm_off,
                             norm_coef, box_lo, box_hi, ymean, ystd, ZnT, Zn, L, W, WT, alpha, info, stream);
}

extern "C" int pcabo_gp_mll(int B, int n, int ld_n, const double *f, const double *ymean, const double *ystd,
                            const double *L, const double *alpha, double *mll, void *stream) {
  if (B <= 0 || n < 1 || ld_n < n || !f || !ymean || !ystd || !L || !alpha || !mll) return PCABO_E_INVALID;
  gp_mll_kernel<<<B, 256, 0, as_stream(stream)>>>(n, ld_n, f, ymean, ystd, L, alpha, mll);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}

extern "C" long long pcabo_gp_packed_doubles(int n) { return n < 1 ? -1 : lbeval::packed_doubles(n); }

extern "C" int pcabo_gp_pack_w(int B, int n, int ld_n, const double *W, double *Wp, long long wp_stride, void *stream) {
  if (B <= 0 || n < 1 || ld_n < n || !W || !Wp || wp_stride < lbeval::packed_doubles(n) || (wp_stride & 1)) return PCABO_E_INVALID;
  dim3 grid(lbeval::ngroups(n), B);
  gp_pack_w_kernel<<<grid, 256, 0, as_stream(stream)>>>(n, ld_n, W, Wp, wp_stride);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}
