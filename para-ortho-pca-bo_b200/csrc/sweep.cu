// K6b/K7: log-EI sweep over a large candidate set against one fitted GP with a fused arg-max (BASELINE config C4:
// n = 450 training points, r <= 40, 2^20 .. 2^26 scrambled-Sobol candidates in the 0.5-padded box).
//
// This is the raw-sample scoring pass of the reference's optimize_acqf (PCA_BO.py:607-614; botorch evaluates
// LogExpectedImprovement on raw_samples = 512 Sobol points in chunks of batch_limit = 5) scaled to millions of
// candidates, so unlike the 5-point evaluations of the BO loop it is a dense FP64 contraction and FP64-pipe bound:
//     T = W K*      W = L^-1 (n x n lower triangular),  K*[j][c] = matern52(|zn_j - xn_c| / ell)   (n x N_cand)
//     var_c = k** - sum_i T[i][c]^2 ,  mu_c = alpha . K*[:, c]  ->  log-EI_c  ->  arg-max
// One CTA owns a tile of 32 candidates and keeps the whole n x 32 tile of T in registers as FP64 tensor-core
// accumulators (mma.sync.aligned.m8n8k4.f64, the only FP64 MMA shape on sm_100: SASS DMMA.8x8x4). The k dimension is
// streamed in panels of 16 training points: the W panel arrives by cp.async (double buffered, 16-byte copies), the K*
// panel is computed on the fly from the staged training rows, so neither W nor K* is ever re-read per candidate and
// candidates are generated in-kernel from the scrambled Sobol state (no HBM traffic for them at all).
#include "logei_core.cuh"

namespace {

using logei::SQRT5;

constexpr int SW_THREADS = 512;
constexpr int NWARP = SW_THREADS / 32;   // 16
constexpr int TC = 32;        // candidates per tile
constexpr int CT = TC / 8;    // 8-wide column tiles
constexpr int KB = 16;        // training points per k panel
constexpr int WS_LD = 20;     // W panel row stride in doubles (bank-conflict free for the A fragments)
constexpr int KS_LD = 36;     // K* panel row stride in doubles (bank-conflict free for the B fragments)
constexpr int RT_MAX = 4;     // row tiles (of 8 rows) per warp: n <= 16 * 4 * 8 = 512
constexpr int MAXBIT = 30;

struct SweepArgs {
  pcabo_gp_model m;
  int b;                       // run (GP) of the batch to sweep against
  long long ncand, idx0;       // candidates in this launch and global index of the first one
  const double *X;             // optional explicit candidates [ncand][kmax] (raw space); nullptr -> Sobol
  const int *shift, *sstate;   // Sobol: [kmax] shifts and [kmax][30] scrambled direction numbers
  const double *lo, *hi;       // box [kmax]
  double *vals;                // optional [ncand]
  double *best_val;            // [gridDim.x]
  long long *best_idx;         // [gridDim.x] global indices
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Shared-memory carve-up (offsets in doubles from the dynamic shared base; buffers are addressed as base + offset so
// that the compiler keeps them in the shared state space: LDS / STS, no generic loads):
//   Ws[2][rows8][WS_LD] W panels | Ks[2][KB][KS_LD] K* panels | Zs[3][KB][zld] training rows (two panels ahead) |
//   As[3][KB] alpha | xn[TC][zld] normalised candidates | xx[TC] |xn|^2 | zz[rows16] |zn|^2 | red | sst (ints)

__host__ __device__ inline int sweep_zld(int kmax) { return ((kmax + 3) / 4) * 4 + 4; }

__host__ __device__ inline size_t sweep_smem_bytes(int n, int kmax) {
  const int rows8 = ((n + 7) / 8) * 8, rows16 = ((n + 15) / 16) * 16, zld = sweep_zld(kmax);
  size_t d = 2 * (size_t)rows8 * WS_LD + 2 * KB * KS_LD + 3 * (size_t)KB * zld + 3 * KB + (size_t)TC * zld + TC + rows16 +
             2 * NWARP * TC + 8;
  return d * sizeof(double) + (size_t)kmax * MAXBIT * sizeof(int);
}

__global__ void __launch_bounds__(SW_THREADS, 1) ei_sweep_kernel(SweepArgs a) {
  extern __shared__ __align__(16) double smd[];
  const pcabo_gp_model &m = a.m;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = m.n, kmax = m.kmax;
  const long ld = m.ld_n;
  const int b = a.b;
  const int kk = m.k ? m.k[b] : kmax;
  const int kk4 = (kk + 3) / 4;               // k-steps of the distance MMA
  const int zld = sweep_zld(kmax);
  const int rows8 = ((n + 7) / 8) * 8, rows16 = ((n + 15) / 16) * 16;
  const int nrt = rows8 / 8;
  const int ws_sz = rows8 * WS_LD, ks_sz = KB * KS_LD, zs_sz = KB * zld;
  double *const sWs = smd;
  double *const sKs = sWs + 2 * ws_sz;
  double *const sZs = sKs + 2 * ks_sz;
  double *const sAs = sZs + 3 * zs_sz;
  double *const sxn = sAs + 3 * KB;
  double *const sxx = sxn + TC * zld;
  double *const szz = sxx + TC;
  double *const sred = szz + rows16;
  int *const ssst = reinterpret_cast<int *>(sred + 2 * NWARP * TC + 8);
  const double *W = m.W + (size_t)b * ld * ld;
  const double *Zn = m.Zn + (size_t)b * ld * kmax;
  const double *alpha = m.alpha + (size_t)b * ld;
  const double *off = m.norm_off + (size_t)b * kmax, *coef = m.norm_coef + (size_t)b * kmax;
  const double inv_ell2 = 1.0 / (m.lengthscale * m.lengthscale);
  const double ystd = m.ystd[b], ymean = m.ymean[b], best_f = m.best_f[b], sgn = m.maximize ? 1.0 : -1.0;
  const double d0 = 1e-15;
  const double kss = (SQRT5 * d0 + 1.0 + (5.0 / 3.0) * (d0 * d0)) * exp(-SQRT5 * d0);

  if (a.X == nullptr)
    for (int i = tid; i < kk * MAXBIT; i += SW_THREADS) ssst[i] = a.sstate[i];
  // squared norms of the training rows (once per CTA) and zero padding of the staging buffers
  for (int j = tid; j < rows16; j += SW_THREADS) {
    double acc = 0.0;
    if (j < n)
      for (int dmm = 0; dmm < kk; ++dmm) { const double z = Zn[(size_t)j * kmax + dmm]; acc = fma(z, z, acc); }
    szz[j] = acc;
  }
  for (int i = tid; i < 3 * KB * zld; i += SW_THREADS) sZs[i] = 0.0;   // the three Zs buffers are contiguous
  for (int i = tid; i < TC * zld; i += SW_THREADS) sxn[i] = 0.0;
  double cta_best = -INFINITY;
  long long cta_idx = 0x7fffffffffffffffLL;
  const long long ntiles = (a.ncand + TC - 1) / TC;
  const int nchunks = (n + KB - 1) / KB;
  const int ar = lane >> 2, ak = lane & 3;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long c0 = tile * TC;
    __syncthreads();
    // ---- candidate coordinates, normalised: xn[c][j]
    for (int idx = tid; idx < TC * kk; idx += SW_THREADS) {
      const int c = idx / kk, j = idx % kk;
      const long long ci = c0 + c;
      double x = 0.0;
      if (ci < a.ncand) {
        if (a.X != nullptr) {
          x = a.X[(size_t)ci * kmax + j];
        } else {
          const unsigned long long g = (unsigned long long)(a.idx0 + ci);
          const int sh = a.shift[j];
          double raw;
          if (g == 0) {
            raw = (double)((float)sh) * (1.0 / 1073741824.0);
          } else {
            int q = sh;
            unsigned gray = (unsigned)(g ^ (g >> 1));
            while (gray) {
              const int l = __ffs(gray) - 1;
              q ^= ssst[j * MAXBIT + l];
              gray &= gray - 1;
            }
            raw = (double)q * (1.0 / 1073741824.0);
          }
          x = __dadd_rn(a.lo[j], __dmul_rn(__dsub_rn(a.hi[j], a.lo[j]), raw));
        }
      }
      sxn[c * zld + j] = (x - off[j]) / coef[j];
    }
    __syncthreads();
    if (tid < TC) {
      double acc = 0.0;
      for (int dmm = 0; dmm < kk; ++dmm) { const double x = sxn[tid * zld + dmm]; acc = fma(x, x, acc); }
      sxx[tid] = acc;
    }
    // ---- accumulators: warp owns row tiles rt = warp, warp + 16, ...; 4 column tiles each
    double acc[RT_MAX][CT][2];
#pragma unroll
    for (int t = 0; t < RT_MAX; ++t)
#pragma unroll
      for (int c = 0; c < CT; ++c) { acc[t][c][0] = 0.0; acc[t][c][1] = 0.0; }
    double mean_part[2] = {0.0, 0.0};  // partial alpha . k of this thread's two candidates (its K* tiles)

    // W panel `chunk` -> Ws[chunk & 1]; training rows / alpha of panel `chunk` -> Zs / As[chunk % 3]
    auto issue_w = [&](int chunk) {
      const int j0 = chunk * KB, buf = chunk & 1;
      const int i_lo = j0 & ~7;
      const int nrows = rows8 - i_lo;   // rows above i_lo are entirely above the diagonal for this panel
      for (int idx = tid; idx < nrows * (KB / 2); idx += SW_THREADS) {
        const int r = idx / (KB / 2), h = idx % (KB / 2);
        const int i = min(i_lo + r, n - 1);
        const int jc = min(j0 + 2 * h, (int)ld - 2);   // stay inside the (even) row; columns >= n meet zero rows of K*
        cp_async16(&sWs[buf * ws_sz + (i_lo + r) * WS_LD + 2 * h], &W[(size_t)i * ld + jc]);
      }
    };
    auto issue_z = [&](int chunk) {
      const int j0 = chunk * KB, zb = chunk % 3;
      for (int idx = tid; idx < KB * (kmax / 2); idx += SW_THREADS) {
        const int r = idx / (kmax / 2), h = idx % (kmax / 2);
        const int j = min(j0 + r, n - 1);
        cp_async16(&sZs[zb * zs_sz + r * zld + 2 * h], &Zn[(size_t)j * kmax + 2 * h]);
      }
      if (tid < KB / 2) {
        const int j = min(j0 + 2 * tid, (int)ld - 2);
        cp_async16(&sAs[zb * KB + 2 * tid], &alpha[j]);
      }
    };
    // K* tile of panel `chunk` by this warp (tile id = warp / 2): dot = Zs (8 x kk) xn^T (kk x 8) on the tensor cores,
    // sq = |z|^2 + |x|^2 - 2 dot, Ks[jj][c] = matern52(sqrt(sq) / ell); rows beyond n are zero
    auto kstar_tile = [&](int chunk) {
      const int j0 = chunk * KB, zb = chunk % 3, kb = chunk & 1;
      const int id = warp >> 1, tr = id / CT, tc = id % CT;
      double dt0 = 0.0, dt1 = 0.0;
      const double *zrow = sZs + zb * zs_sz + (tr * 8 + ar) * zld + ak;
      const double *xrow = sxn + (tc * 8 + ar) * zld + ak;
      for (int k4 = 0; k4 < kk4; ++k4) dmma_8x8x4(dt0, dt1, zrow[k4 * 4], xrow[k4 * 4]);
      const int jj = tr * 8 + ar;
      const double zzj = szz[j0 + jj], aj = sAs[zb * KB + jj];
      const bool on = j0 + jj < n;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = tc * 8 + ak * 2 + h;
        const double dot = h ? dt1 : dt0;
        const double sq = fmax((zzj + sxx[c]) - 2.0 * dot, 0.0);
        const double s2 = sq * inv_ell2;
        const double dist = sqrt(fmax(s2, 1e-30));
        double kv = (SQRT5 * dist + 1.0 + (5.0 / 3.0) * (dist * dist)) * exp(-SQRT5 * dist);
        if (!on) kv = 0.0;
        sKs[kb * ks_sz + jj * KS_LD + c] = kv;
        if (on) mean_part[h] = fma(kv, aj, mean_part[h]);
      }
    };

    // prologue: panel 0 (W, Z), Z of panel 1; K*(0)
    issue_w(0);
    issue_z(0);
    if (nchunks > 1) issue_z(1);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if ((warp & 1) == 0) kstar_tile(0);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int buf = chunk & 1;
      const int j0 = chunk * KB;
      cp_async_wait<0>();
      __syncthreads();   // K*(chunk), W(chunk), Z(chunk+1) are visible; every warp is past the MMAs of panel chunk - 1
      if (chunk + 1 < nchunks) issue_w(chunk + 1);
      if (chunk + 2 < nchunks) issue_z(chunk + 2);
      cp_async_commit();
      // the K* tile of the NEXT panel: warps alternate this duty so that others keep the tensor pipe busy meanwhile
      if (chunk + 1 < nchunks && ((warp + chunk + 1) & 1) == 0) kstar_tile(chunk + 1);
      // ---- T += Wpanel (rows x 16) * Kspanel (16 x 32) on the FP64 tensor cores
      {
        const double *Wsb = sWs + buf * ws_sz;
        const double *Ksb = sKs + buf * ks_sz;
#pragma unroll
        for (int k4 = 0; k4 < KB / 4; ++k4) {
          double bf[CT];
#pragma unroll
          for (int c = 0; c < CT; ++c) bf[c] = Ksb[(k4 * 4 + ak) * KS_LD + c * 8 + ar];
          const int jcol = j0 + k4 * 4 + ak;
#pragma unroll
          for (int t = 0; t < RT_MAX; ++t) {
            const int rt = warp + t * NWARP;
            if (rt < nrt && rt * 8 + 7 >= j0) {
              const int row = rt * 8 + ar;
              // entries above the diagonal and padding rows are not part of L^-1
              const double af = (row < n && jcol <= row) ? Wsb[row * WS_LD + k4 * 4 + ak] : 0.0;
#pragma unroll
              for (int c = 0; c < CT; ++c) dmma_8x8x4(acc[t][c][0], acc[t][c][1], af, bf[c]);
            }
          }
        }
      }
    }
    // ---- variance partials: thread holds T[row = rt*8 + lane/4][col = c*8 + (lane%4)*2 + {0,1}]
    double vpart[CT][2];
#pragma unroll
    for (int c = 0; c < CT; ++c) { vpart[c][0] = 0.0; vpart[c][1] = 0.0; }
#pragma unroll
    for (int t = 0; t < RT_MAX; ++t)
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        vpart[c][0] = fma(acc[t][c][0], acc[t][c][0], vpart[c][0]);
        vpart[c][1] = fma(acc[t][c][1], acc[t][c][1], vpart[c][1]);
      }
    // reduce over the 8 row groups of the warp (lanes with equal lane % 4): xor 4, 8, 16
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double v = vpart[c][h];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        vpart[c][h] = v;
      }
    // mean partials: thread holds candidates tc*8 + ak*2 + {0,1} for training rows tr*8 + ar of its panels
    {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double v = mean_part[h];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        mean_part[h] = v;
      }
    }
    __syncthreads();
    if (lane < 4) {
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        sred[warp * TC + c * 8 + lane * 2 + 0] = vpart[c][0];
        sred[warp * TC + c * 8 + lane * 2 + 1] = vpart[c][1];
      }
      sred[NWARP * TC + warp * 8 + lane * 2 + 0] = mean_part[0];   // [warp][8 candidates of tile column (warp/2) % CT]
      sred[NWARP * TC + warp * 8 + lane * 2 + 1] = mean_part[1];
    }
    __syncthreads();
    if (tid < TC) {
      const int c = tid;
      double v = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < NWARP; ++w8) v += sred[w8 * TC + c];
      double ms = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < NWARP; ++w8)
        if (((w8 >> 1) % CT) == c / 8) ms += sred[NWARP * TC + w8 * 8 + (c & 7)];
      const double mu = ymean + ystd * ms;
      const double var_raw = ystd * ystd * (kss - v);
      const double var = (var_raw >= 1e-10) ? var_raw : 1e-10;
      const double sg = sqrt(var);
      const double u = sgn * (mu - best_f) / sg;
      double lh, dlh;
      logei::log_h_and_grad(u, lh, dlh);
      double val = lh + log(sg);
      const long long ci = c0 + c;
      if (ci < a.ncand) {
        if (a.vals) a.vals[ci] = val;
      } else {
        val = -INFINITY;
      }
      // tile arg-max (first occurrence) over the 32 candidates of warp 0
      long long gi = a.idx0 + ci;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, val, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, gi, o);
        if (ov > val || (ov == val && oi < gi)) { val = ov; gi = oi; }
      }
      if (val > cta_best || (val == cta_best && gi < cta_idx)) { cta_best = val; cta_idx = gi; }
    }
  }
  if (tid == 0) { a.best_val[blockIdx.x] = cta_best; a.best_idx[blockIdx.x] = cta_idx; }
}

__global__ void sweep_final_kernel(int nparts, const double *pv, const long long *pi, double *best_val,
                                   long long *best_idx) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  double bv = -INFINITY;
  long long bi = 0x7fffffffffffffffLL;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    const double v = pv[i];
    const long long g = pi[i];
    if (v > bv || (v == bv && g < bi)) { bv = v; bi = g; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    bv = lane < nw ? sv[lane] : -INFINITY;
    bi = lane < nw ? si[lane] : 0x7fffffffffffffffLL;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { *best_val = bv; *best_idx = bi; }
  }
}

}  // namespace

extern "C" int pcabo_ei_sweep(const pcabo_gp_model *m, int run, long long ncand, long long idx0, const double *X,
                              const int *shift, const int *sstate, const double *lo, const double *hi, double *vals,
                              double *part_val, long long *part_idx, int nparts, double *best_val,
                              long long *best_idx, void *stream) {
  if (!m || run < 0 || run >= m->B || ncand <= 0 || !part_val || !part_idx || nparts <= 0 || !best_val || !best_idx ||
      !m->W || !m->Zn || !m->alpha || !m->norm_off || !m->norm_coef || !m->ymean || !m->ystd || !m->best_f)
    return PCABO_E_INVALID;
  if (X == nullptr && (!shift || !sstate || !lo || !hi)) return PCABO_E_INVALID;
  // n <= 512 (accumulators of the whole T tile live in registers), even ld_n and kmax (16-byte cp.async copies)
  if (m->n > 8 * NWARP * RT_MAX || m->n < 2 || (m->ld_n & 1) || (m->kmax & 1) || m->kmax > 104) return PCABO_E_UNSUPPORTED;
  const size_t smem = sweep_smem_bytes(m->n, m->kmax);
  if (smem > 227 * 1024) return PCABO_E_UNSUPPORTED;
  PCABO_CUDA_CHECK(cudaFuncSetAttribute(ei_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SweepArgs a;
  a.m = *m; a.b = run; a.ncand = ncand; a.idx0 = idx0; a.X = X; a.shift = shift; a.sstate = sstate; a.lo = lo; a.hi = hi;
  a.vals = vals; a.best_val = part_val; a.best_idx = part_idx;
  const long long ntiles = (ncand + TC - 1) / TC;
  const int grid = (int)(ntiles < nparts ? ntiles : nparts);
  cudaStream_t st = as_stream(stream);
  ei_sweep_kernel<<<grid, SW_THREADS, smem, st>>>(a);
  PCABO_LAUNCH_CHECK();
  sweep_final_kernel<<<1, 256, 0, st>>>(grid, part_val, part_idx, best_val, best_idx);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}
