// K1: rank weights + weighted PCA (covariance + shared-memory Jacobi eigensolver) + projection, one CTA per run.
// K2: inverse map z -> x with the bounds test.
//
// Replaces PCA_BO._calculate_weights (PCA_BO.py:316-341), PCA_BO._transform_points_to_reduced_space
// (PCA_BO.py:343-408, i.e. sklearn PCA().fit + .transform) and PCA_BO._transform_point_to_original_space
// (PCA_BO.py:410-434) of the reference. Semantics:
//   ranks = argsort(argsort(+-f)) + 1 (ties by index) ; w~ = ln n - ln rank ; w = w~ / sum(w~) [numpy pairwise order]
//   xbar = mean(X) ; Xw = (X - xbar) * sqrt(w) + noise ; m2 = mean(Xw) ; C = (Xw - m2)^T (Xw - m2) / (n - 1)
//   eigen-decomposition of C (two-sided parallel-ordered Jacobi in shared memory), descending, negatives clipped,
//   svd_flip (largest |entry| of every component positive), ratio = lam / sum(lam),
//   k = #(cumsum(ratio) <= thr) + 1 clamped to [1, r]  (or n_components if > 0), Z = (X - xbar) V_k^T - m2 V_k^T.
#include "common.cuh"

namespace {

constexpr int PCA_THREADS = 512;
constexpr int MAX_SWEEPS = 40;

struct PcaArgs {
  int B, n, D, ld_n, R;
  const double *X, *f, *noise, *logtab;
  double thr;
  int ncomp, maxim;
  double *w;
  int *ranks;
  double *xbar, *m2, *V, *lam, *evr;
  int *k;
  double *Z;
};

__host__ __device__ inline size_t pca_smem_doubles(int n, int D, int R) {
  const int n_pad = (n + 1) & ~1;
  const int ldc = D + 1;
  const int D4 = ((D + 3) / 4) * 4;
  const int H = (D + 1) / 2;
  size_t d = 0;
  d += 2 * (size_t)n_pad;          // key, pre
  d += 2 * (size_t)D * ldc;        // C, Vm
  d += ((size_t)R * D4 > (size_t)PCA_THREADS ? (size_t)R * D4 : (size_t)PCA_THREADS);  // row tile / partial buffers
  d += 4 * (size_t)D4;             // xb, mm2, lamv, mproj
  d += 3 * (size_t)(H + 1);        // rc, rs, rt
  d += 40;                         // reduction scratch + scalars
  d += (size_t)(2 * (H + 1) + D4 + 8 + 1) / 2 + 2;  // ints: ip, iq, order, flags
  return d;
}

__global__ void __launch_bounds__(PCA_THREADS, 1) pca_fit_kernel(PcaArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int n = a.n, D = a.D, R = a.R;
  const int n_pad = (n + 1) & ~1;
  const int ldc = D + 1;
  const int D4 = ((D + 3) / 4) * 4;
  const int Dp = (D + 1) & ~1;   // even
  const int H = Dp / 2;
  double *key = sm;
  double *pre = key + n_pad;
  double *C = pre + n_pad;
  double *Vm = C + (size_t)D * ldc;
  double *tile = Vm + (size_t)D * ldc;
  double *xb = tile + ((size_t)R * D4 > (size_t)PCA_THREADS ? (size_t)R * D4 : (size_t)PCA_THREADS);
  double *mm2 = xb + D4;
  double *lamv = mm2 + D4;
  double *mproj = lamv + D4;
  double *rc = mproj + D4;
  double *rs = rc + (H + 1);
  double *rt = rs + (H + 1);
  double *red = rt + (H + 1);  // 40 doubles: [0..31] scratch, [32] S, [33] total
  int *ip = reinterpret_cast<int *>(red + 40);
  int *iq = ip + (H + 1);
  int *order = iq + (H + 1);
  int *flags = order + D4;     // [0] rotated flag, [1] k

  const double *Xb = a.X + (size_t)b * a.ld_n * D;
  const double *fb = a.f + (size_t)b * a.ld_n;
  const double *Nb = a.noise + (size_t)b * n * D;

  // ---- A: keys
  for (int i = tid; i < n; i += nt) key[i] = a.maxim ? -fb[i] : fb[i];
  __syncthreads();
  // ---- B: ranks (stable by index) and pre-weights from the host log table
  const double logn = a.logtab[n];
  for (int i = tid; i < n; i += nt) {
    const double ki = key[i];
    int r = 1;
    for (int j = 0; j < n; ++j) {
      const double kj = key[j];
      r += (kj < ki) || (kj == ki && j < i);
    }
    a.ranks[(size_t)b * a.ld_n + i] = r;
    pre[i] = __dsub_rn(logn, a.logtab[r]);
  }
  __syncthreads();
  // ---- C: normaliser (numpy pairwise order, warp 0) || column-mean partials of X (other warps)
  if (tid < 32) {
    if (tid == 0) red[32] = numpy_pairwise_sum(pre, n);
  } else {
    const int tt = tid - 32, nth = nt - 32;
    const int G = nth / D;  // >= 4 for D <= 104
    const int g = tt / D, d = tt % D;
    if (g < G) {
      double acc = 0.0;
      for (int i = g; i < n; i += G) acc += Xb[(size_t)i * D + d];
      tile[g * D + d] = acc;
    }
  }
  __syncthreads();
  {
    const int G = (nt - 32) / D;
    if (tid < D) {
      double acc = 0.0;
      for (int g = 0; g < G; ++g) acc += tile[g * D + tid];
      xb[tid] = acc / (double)n;
      a.xbar[(size_t)b * D + tid] = xb[tid];
    }
  }
  // ---- D: weights, sqrt(weights)
  {
    const double S = red[32];
    for (int i = tid; i < n; i += nt) {
      const double wi = __ddiv_rn(pre[i], S);
      a.w[(size_t)b * a.ld_n + i] = wi;
      key[i] = sqrt(wi);  // key now holds sqrt(w)
    }
  }
  __syncthreads();
  // ---- E: m2 = column mean of Xw = (X - xbar) * sqrt(w) + noise
  {
    const int G = nt / D;
    const int g = tid / D, d = tid % D;
    if (g < G) {
      double acc = 0.0;
      const double xbd = xb[d];
      for (int i = g; i < n; i += G)
        acc += (Xb[(size_t)i * D + d] - xbd) * key[i] + Nb[(size_t)i * D + d];
      tile[g * D + d] = acc;
    }
    __syncthreads();
    if (tid < D) {
      double acc = 0.0;
      for (int gg = 0; gg < G; ++gg) acc += tile[gg * D + tid];
      mm2[tid] = acc / (double)n;
      a.m2[(size_t)b * D + tid] = mm2[tid];
    }
    __syncthreads();
  }
  // ---- F: covariance C = A^T A / (n - 1), A = Xw - m2, 4x4 register tiles over the lower block triangle
  {
    const int nt4 = D4 / 4;
    const int Tl = nt4 * (nt4 + 1) / 2;
    int G = nt / Tl;
    if (G < 1) G = 1;
    if (G > 16) G = 16;
    const bool active = tid < Tl * G;
    int ta = 0, tb = 0, g = 0;
    if (active) {
      const int t = tid % Tl;
      g = tid / Tl;
      int acc_t = 0;
      for (ta = 0; ta < nt4; ++ta) {
        if (t < acc_t + ta + 1) { tb = t - acc_t; break; }
        acc_t += ta + 1;
      }
    }
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int i0 = 0; i0 < n; i0 += R) {
      const int rows = min(R, n - i0);
      for (int idx = tid; idx < rows * D4; idx += nt) {
        const int r = idx / D4, d = idx % D4;
        double v = 0.0;
        if (d < D) {
          const int i = i0 + r;
          v = (Xb[(size_t)i * D + d] - xb[d]) * key[i] + Nb[(size_t)i * D + d] - mm2[d];
        }
        tile[r * D4 + d] = v;
      }
      __syncthreads();
      if (active) {
        for (int r = g; r < rows; r += G) {
          const double *row = tile + r * D4;
          double av[4], bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { av[i] = row[ta * 4 + i]; bv[i] = row[tb * 4 + i]; }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
      }
      __syncthreads();
    }
    // deterministic reduction of the G row-groups into C (both triangles), then scale
    for (int idx = tid; idx < D * ldc; idx += nt) C[idx] = 0.0;
    __syncthreads();
    for (int gg = 0; gg < G; ++gg) {
      if (active && g == gg) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int p = ta * 4 + i, q = tb * 4 + j;
            if (p < D && q < D && p >= q) C[p * ldc + q] += acc[i][j];
          }
      }
      __syncthreads();
    }
    const double inv = 1.0 / (double)(n - 1);
    for (int idx = tid; idx < D * D; idx += nt) {
      const int p = idx / D, q = idx % D;
      if (p >= q) {
        const double v = C[p * ldc + q] * inv;
        C[p * ldc + q] = v;
        C[q * ldc + p] = v;
      }
    }
    for (int idx = tid; idx < D * D; idx += nt) {
      const int p = idx / D, q = idx % D;
      Vm[p * ldc + q] = (p == q) ? 1.0 : 0.0;
    }
    __syncthreads();
  }
  // ---- G: two-sided Jacobi, round-robin parallel ordering (H disjoint pairs per step, Dp-1 steps per sweep)
  {
    const double EPS = 2.220446049250313e-16;
    const int m = Dp - 1;
    for (int sweep = 0; sweep < MAX_SWEEPS; ++sweep) {
      if (tid == 0) flags[0] = 0;
      __syncthreads();
      for (int step = 0; step < m; ++step) {
        if (tid < H) {
          int p, q;
          if (tid == 0) { p = step % m; q = m; }
          else { p = (step + tid) % m; q = (step - tid + m) % m; }
          if (p > q) { const int t_ = p; p = q; q = t_; }
          double c = 1.0, s = 0.0, t = 0.0;
          if (q < D) {
            const double al = C[p * ldc + p], be = C[q * ldc + q], ga = C[p * ldc + q];
            const double aga = fabs(ga);
            if (aga > EPS * sqrt(fabs(al * be)) && aga > 1e-300) {
              const double tau = (be - al) / (2.0 * ga);
              t = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
              if (t != 0.0) {
                c = 1.0 / sqrt(1.0 + t * t);
                s = t * c;
                flags[0] = 1;
              }
            }
          }
          rc[tid] = c; rs[tid] = s; rt[tid] = t; ip[tid] = p; iq[tid] = q;
        }
        __syncthreads();
        // C <- J^T C J on 2x2 blocks (pa,qa) x (pb,qb), upper block triangle computed and mirrored
        for (int idx = tid; idx < H * H; idx += nt) {
          const int pa_i = idx / H, pb_i = idx % H;
          if (pa_i > pb_i) continue;
          const double ca = rc[pa_i], sa = rs[pa_i], cb = rc[pb_i], sb = rs[pb_i];
          if (sa == 0.0 && sb == 0.0) continue;
          const int pa = ip[pa_i], qa = iq[pa_i], pb = ip[pb_i], qb = iq[pb_i];
          if (pa_i == pb_i) {
            if (qa >= D) continue;
            const double ga = C[pa * ldc + qa], t = rt[pa_i];
            C[pa * ldc + pa] -= t * ga;
            C[qa * ldc + qa] += t * ga;
            C[pa * ldc + qa] = 0.0;
            C[qa * ldc + pa] = 0.0;
            continue;
          }
          const bool va = qa < D, vb = qb < D;  // the dummy index (odd D) carries no data
          double m00 = C[pa * ldc + pb];
          double m01 = vb ? C[pa * ldc + qb] : 0.0;
          double m10 = va ? C[qa * ldc + pb] : 0.0;
          double m11 = (va && vb) ? C[qa * ldc + qb] : 0.0;
          // rows: J_a^T
          const double r00 = ca * m00 - sa * m10, r01 = ca * m01 - sa * m11;
          const double r10 = sa * m00 + ca * m10, r11 = sa * m01 + ca * m11;
          // cols: J_b
          m00 = cb * r00 - sb * r01; m01 = sb * r00 + cb * r01;
          m10 = cb * r10 - sb * r11; m11 = sb * r10 + cb * r11;
          C[pa * ldc + pb] = m00; C[pb * ldc + pa] = m00;
          if (vb) { C[pa * ldc + qb] = m01; C[qb * ldc + pa] = m01; }
          if (va) { C[qa * ldc + pb] = m10; C[pb * ldc + qa] = m10; }
          if (va && vb) { C[qa * ldc + qb] = m11; C[qb * ldc + qa] = m11; }
        }
        // V <- V J
        for (int idx = tid; idx < D * H; idx += nt) {
          const int r = idx / H, pi = idx % H;
          const double s = rs[pi];
          if (s == 0.0) continue;
          const double c = rc[pi];
          const int p = ip[pi], q = iq[pi];
          const double vp = Vm[r * ldc + p], vq = Vm[r * ldc + q];
          Vm[r * ldc + p] = c * vp - s * vq;
          Vm[r * ldc + q] = s * vp + c * vq;
        }
        __syncthreads();
      }
      const int rotated = flags[0];
      __syncthreads();
      if (!rotated) break;
    }
  }
  // ---- H: sort descending (ties by index), clip negatives, svd_flip, ratios, k
  const int rnk = min(n, D);  // sklearn keeps min(n, D) components
  if (tid < D) {
    double v = C[tid * ldc + tid];
    lamv[tid] = v;
  }
  __syncthreads();
  if (tid < D) {
    const double v = lamv[tid];
    int pos = 0;
    for (int j = 0; j < D; ++j) {
      const double u = lamv[j];
      pos += (u > v) || (u == v && j < tid);
    }
    order[pos] = tid;
  }
  __syncthreads();
  // svd_flip: sign of the largest-|entry| (first occurrence) of each component -> mproj (temporary)
  {
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int c = warp; c < D; c += nw) {
      const int col = order[c];
      double best = -1.0;
      int bidx = 0;
      for (int d = lane; d < D; d += 32) {
        const double av = fabs(Vm[d * ldc + col]);
        if (av > best) { best = av; bidx = d; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
      }
      if (lane == 0) mproj[c] = (Vm[bidx * ldc + col] < 0.0) ? -1.0 : 1.0;
    }
  }
  __syncthreads();
  // components_ (rows) into the C storage (only C's diagonal was still needed, already in lamv) and to global
  for (int idx = tid; idx < D * D; idx += nt) {
    const int c = idx / D, d = idx % D;
    const double v = mproj[c] * Vm[d * ldc + order[c]];
    C[c * ldc + d] = v;
    a.V[((size_t)b * D + c) * D + d] = v;
  }
  __syncthreads();
  if (tid == 0) {
    // explained variance (descending, clipped), ratio with numpy's pairwise total, cumulative threshold
    double *sorted = tile;  // reuse
    for (int c = 0; c < D; ++c) {
      double v = lamv[order[c]];
      if (v < 0.0) v = 0.0;
      sorted[c] = (c < rnk) ? v : 0.0;
    }
    const double total = numpy_pairwise_sum(sorted, rnk);
    double cum = 0.0;
    int cnt = 0;
    for (int c = 0; c < D; ++c) {
      const double ratio = (c < rnk) ? sorted[c] / total : 0.0;
      a.lam[(size_t)b * D + c] = sorted[c];
      a.evr[(size_t)b * D + c] = ratio;
      if (c < rnk) {
        cum = __dadd_rn(cum, ratio);
        if (cum <= a.thr) ++cnt;
      }
    }
    int kk = a.ncomp > 0 ? a.ncomp : cnt + 1;
    if (a.ncomp <= 0) kk = max(1, min(kk, rnk));
    kk = max(1, min(kk, D));
    flags[1] = kk;
    a.k[b] = kk;
  }
  __syncthreads();
  const int kk = flags[1];
  // ---- I: projection Z = (X - xbar) Vk^T - m2 Vk^T
  if (tid < D) {
    double acc = 0.0;
    for (int d = 0; d < D; ++d) acc = fma(mm2[d], C[tid * ldc + d], acc);
    mproj[tid] = acc;
  }
  __syncthreads();
  double *Zb = a.Z + (size_t)b * a.ld_n * D;
  for (int i0 = 0; i0 < n; i0 += R) {
    const int rows = min(R, n - i0);
    for (int idx = tid; idx < rows * D; idx += nt) {
      const int r = idx / D, d = idx % D;
      tile[r * D4 + d] = Xb[(size_t)(i0 + r) * D + d] - xb[d];
    }
    __syncthreads();
    for (int idx = tid; idx < rows * D; idx += nt) {
      const int r = idx / D, c = idx % D;
      double z = 0.0;
      if (c < kk) {
        const double *row = tile + r * D4;
        const double *vc = C + c * ldc;
        double acc = 0.0;
        for (int d = 0; d < D; ++d) acc = fma(row[d], vc[d], acc);
        z = acc - mproj[c];
      }
      Zb[(size_t)(i0 + r) * D + c] = z;
    }
    __syncthreads();
  }
}

__global__ void inverse_map_kernel(int B, int D, const double *z, const int *k, const double *V, const double *m2,
                                   const double *xbar, const double *lo, const double *hi, double *x, int *oob) {
  const int b = blockIdx.x;
  __shared__ int s_oob;
  if (threadIdx.x == 0) s_oob = 0;
  __syncthreads();
  const int kk = k ? k[b] : D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double acc = 0.0;
    for (int c = 0; c < kk; ++c) acc = fma(z[(size_t)b * D + c], V[((size_t)b * D + c) * D + d], acc);
    const double xv = (acc + m2[(size_t)b * D + d]) + xbar[(size_t)b * D + d];
    x[(size_t)b * D + d] = xv;
    // PCA_BO.py:248: outside unless all(x >= lo) and all(x <= hi)  (NaN counts as outside)
    if (!(xv >= lo[d]) || !(xv <= hi[d])) s_oob = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) oob[b] = s_oob;
}

}  // namespace

extern "C" int pcabo_weights_pca_fit(int B, int n, int D, int ld_n, const double *X, const double *f,
                                     const double *noise, const double *logtab, double var_threshold,
                                     int n_components, int maximization, double *w, int *ranks, double *xbar,
                                     double *m2, double *V, double *lam, double *evr, int *k, double *Z,
                                     void *stream) {
  if (B <= 0 || n < 2 || D < 1 || ld_n < n || !X || !f || !noise || !logtab || !w || !ranks || !xbar || !m2 || !V ||
      !lam || !evr || !k || !Z)
    return PCABO_E_INVALID;
  if (D > 104 || n > 4096) return PCABO_E_UNSUPPORTED;
  int R = 32;
  const size_t limit = 227 * 1024;
  while (R > 4 && pca_smem_doubles(n, D, R) * sizeof(double) > limit) R /= 2;
  const size_t smem = pca_smem_doubles(n, D, R) * sizeof(double);
  if (smem > limit) return PCABO_E_UNSUPPORTED;
  PCABO_CUDA_CHECK(cudaFuncSetAttribute(pca_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PcaArgs a;
  a.B = B; a.n = n; a.D = D; a.ld_n = ld_n; a.R = R;
  a.X = X; a.f = f; a.noise = noise; a.logtab = logtab;
  a.thr = var_threshold; a.ncomp = n_components; a.maxim = maximization;
  a.w = w; a.ranks = ranks; a.xbar = xbar; a.m2 = m2; a.V = V; a.lam = lam; a.evr = evr; a.k = k; a.Z = Z;
  pca_fit_kernel<<<B, PCA_THREADS, smem, as_stream(stream)>>>(a);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}

extern "C" int pcabo_inverse_map(int B, int D, const double *z, const int *k, const double *V, const double *m2,
                                 const double *xbar, const double *lo, const double *hi, double *x, int *oob,
                                 void *stream) {
  if (B <= 0 || D < 1 || !z || !V || !m2 || !xbar || !lo || !hi || !x || !oob) return PCABO_E_INVALID;
  inverse_map_kernel<<<B, 128, 0, as_stream(stream)>>>(B, D, z, k, V, m2, xbar, lo, hi, x, oob);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}
