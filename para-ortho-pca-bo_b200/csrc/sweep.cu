// K6b/K7: log-EI sweep over a large candidate set against one fitted GP with a fused arg-max (BASELINE config C4:
// n = 450 training points, r <= 40, 2^20 .. 2^26 scrambled-Sobol candidates in the 0.5-padded box).
//
// This is the raw-sample scoring pass of the reference's optimize_acqf (PCA_BO.py:607-614; botorch evaluates
// LogExpectedImprovement on raw_samples = 512 Sobol points in chunks of batch_limit = 5) scaled to millions of
// candidates, so unlike the 5-point evaluations of the BO loop it is a dense FP64 contraction and FP64-pipe bound:
//     T = W K*      W = L^-1 (n x n lower triangular),  K*[j][c] = matern52(|zn_j - xn_c| / ell)   (n x N_cand)
//     var_c = k** - sum_i T[i][c]^2 ,  mu_c = alpha . K*[:, c]  ->  log-EI_c  ->  arg-max
// One persistent CTA per SM owns tiles of 32 candidates and keeps the whole n x 32 tile of T in registers as FP64
// tensor-core accumulators (mma.sync.aligned.m8n8k4.f64, the only FP64 MMA shape on sm_100: SASS DMMA.8x8x4). The k
// dimension is streamed in panels of 16 training points.
//
// v5 pipeline (no block-wide barrier inside a tile):
//   * K* panels travel through an 8-deep shared-memory ring guarded by mbarriers (full: 4 producer warps, empty: 16
//     consumer warps). Every warp is a consumer of every panel and, on alternate panels, the producer of one 8 x 8 tile
//     of the panel FOUR panels ahead (distances on the tensor cores from the training rows read straight from L2, then
//     sqrt / exp on the FP64 pipe), so the Matern latency of one warp hides behind the DMMAs of the fifteen others and
//     production runs across tile boundaries (the ring never drains).
//   * W is staged per warp: a warp copies only the rows of its own row tiles (cp.async, double buffered, 16 B chunks)
//     and synchronises with nobody else; panels entirely above the diagonal of a row tile are skipped.
//   * candidates of the NEXT tile are generated warp by warp (scrambled Sobol, bit-identical to pcabo_sobol_points)
//     while the current tile is consumed, and the per-tile epilogue (variance / mean reduction, log-EI, arg-max) is
//     done by one warp on rotating duty behind an mbarrier: there is no block-wide barrier in the steady state.
#include "logei_core.cuh"

#ifdef SWEEP_PROF
__device__ long long g_sweep_prof[16][8];
#define SWP_T0() long long swp_t = clock64()
#define SWP(id) do { const long long swp_n = clock64(); if (blockIdx.x == 0 && lane == 0) g_sweep_prof[warp][id] += swp_n - swp_t; swp_t = swp_n; } while (0)
#else
#define SWP_T0()
#define SWP(id)
#endif

namespace {

using logei::SQRT5;

#ifndef SWEEP_WARPS
#define SWEEP_WARPS 16
#endif
constexpr int NWARP = SWEEP_WARPS;       // 16 (4 row tiles each, 128 registers) or 32 (2 row tiles each, 64 registers)
constexpr int SW_THREADS = NWARP * 32;
constexpr int DUTY = NWARP / 4;          // a warp produces a 16 x 8 strip of K* (two 8 x 8 tiles) every DUTY-th panel
constexpr int TC = 32;        // candidates per tile
constexpr int CT = TC / 8;    // 8-wide column tiles
constexpr int KB = 16;        // training points per k panel
constexpr int WS_LD = 16;     // W panel row stride in doubles; columns are XOR-swizzled by 4 * (row & 3) so that the A
                              // fragment loads (8 rows x 4 columns per warp) are bank-conflict free without padding
constexpr int KS_LD = 36;     // K* panel row stride in doubles (bank-conflict free for the B fragments)
constexpr int RT_MAX = 64 / NWARP;   // row tiles (of 8 rows) per warp: n <= NWARP * RT_MAX * 8 = 512
constexpr int RING = 8;       // K* ring depth (panels)
constexpr int AHEAD = 4;      // K* production distance (panels)
constexpr int MAXBIT = 30;

struct SweepArgs {
  pcabo_gp_model m;
  int b;                       // run (GP) of the batch to sweep against
  long long ncand, idx0;       // candidates in this launch and global index of the first one
  const double *X;             // optional explicit candidates [ncand][kmax] (raw space); nullptr -> Sobol
  const int *shift, *sstate;   // Sobol: [kmax] shifts and [kmax][30] scrambled direction numbers
  const double *lo, *hi;       // box [kmax]
  double *vals;                // optional [ncand]
  double *best_val;            // [gridDim.x] (nullptr in the multi-run form: values only)
  long long *best_idx;         // [gridDim.x] global indices
  int multi;                   // 1: one row of CTAs per run (run = blockIdx.y), X / vals advance by x_stride / ncand per run
  long long x_stride;          // doubles between the candidate sets of consecutive runs (multi-run form)
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// ---- mbarrier helpers (CTA scope, shared memory)
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(a) : "memory");
}
// arrival that publishes nothing (a consumer releasing a ring slot it has finished READING): relaxed, so that it
// does not wait for this warp's outstanding cp.async copies the way the default release arrival does
__device__ __forceinline__ void mbar_arrive_relaxed(unsigned long long *bar) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.relaxed.cta.shared::cta.b64 st, [%0];\n}\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "WAIT_LOOP:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra.uni WAIT_DONE;\n"
      " bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}

// Matern-5/2 of a squared scaled distance, k = (1 + a + a^2 / 3) exp(-a) with a = sqrt(5 s2), written out with a bare
// Newton square root and exponential (argument reduction by ln 2, degree-13 Taylor polynomial, exponent insertion):
// ~45 FP64 instructions instead of ~85 for the library sqrt() + exp(), which matters because on this part the scalar
// FP64 instructions share the execution pipe with the DMMAs. Relative error <= 4e-16 for s2 in [1e-30, 4e4]
// (tests/test_gpu_sweep.py checks the resulting log-EI against the oracle to 1e-8).
__device__ __forceinline__ double matern52_fast(double s2) {
  const double x = fmin(fmax(5.0 * s2, 5e-30), 4.0e5);      // a^2; a <= 632 keeps 2^n normal
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));    // ~20 good bits
  // two Newton steps on y ~ x^-1/2, then a = x y with one correction step
  double h = 0.5 * x;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  double a = x * y;
  a = fma(0.5 * y, fma(-a, a, x), a);
  // exp(-a): n = rint(-a log2 e), r = -a - n ln 2 (two-part), e^r by Taylor to degree 13 (|r| <= 0.3466)
  const double SHIFT = 6755399441055744.0;                   // 2^52 + 2^51
  double kd = fma(-a, 1.4426950408889634074, SHIFT);
  const int ki = __double2loint(kd);
  kd -= SHIFT;
  double r = fma(kd, -6.93147180369123816490e-01, -a);
  r = fma(kd, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;                         // 1 / 13!
  p = fma(p, r, 2.08767569878681e-09);
  p = fma(p, r, 2.505210838544172e-08);
  p = fma(p, r, 2.755731922398589e-07);
  p = fma(p, r, 2.7557319223985893e-06);
  p = fma(p, r, 2.48015873015873e-05);
  p = fma(p, r, 1.984126984126984e-04);
  p = fma(p, r, 1.388888888888889e-03);
  p = fma(p, r, 8.333333333333333e-03);
  p = fma(p, r, 4.1666666666666664e-02);
  p = fma(p, r, 1.6666666666666666e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double e = __hiloint2double(__double2hiint(p) + (ki << 20), __double2loint(p));
  return fma(a, fma(a, 1.0 / 3.0, 1.0), 1.0) * e;
}

// T[rows of this warp][32 candidates] += W panel (staged rows x 16) * K* panel (16 x 32) on the FP64 tensor cores.
// T0: first row tile of the warp that reaches down to the panel (compile time -> no per-tile tests in the loop);
// LAST: the panel may contain columns >= n (masked). No other masks: W holds explicit zeros above its diagonal inside
// the diagonal 32 x 32 blocks (written by the fit) and padding rows are dropped in the epilogue.
template <int T0, bool LAST>
__device__ __forceinline__ void mma_panel(double (&acc)[RT_MAX][CT][2], const double *__restrict__ Wsb,
                                          const double *__restrict__ Ksb, const int (&aoff)[KB / 4], int ntw,
                                          int col0, int n) {
#pragma unroll
  for (int k4 = 0; k4 < KB / 4; ++k4) {
    double bf[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) bf[c] = Ksb[k4 * 4 * KS_LD + c * 8];
#pragma unroll
    for (int t = T0; t < RT_MAX; ++t) {
      if (t < ntw) {
        double af = Wsb[t * 8 * WS_LD + aoff[k4]];
        if (LAST && !(col0 + k4 * 4 < n)) af = 0.0;
#pragma unroll
        for (int c = 0; c < CT; ++c) dmma_8x8x4(acc[t][c][0], acc[t][c][1], af, bf[c]);
      }
    }
  }
}

// Shared-memory carve-up (offsets in doubles from the dynamic shared base):
//   Ww[NWARP][2][RT_MAX * 8][WS_LD] per-warp W staging | Ks[RING][KB][KS_LD] K* ring | xn[2][TC][zld] normalised
//   candidates of the current / next tile | xx[2][TC] |xn|^2 | zz[rows16] |zn|^2 | al[rows16] alpha | red | mbarriers |
//   sst (ints)
__host__ __device__ inline int sweep_zld(int kmax) { return ((kmax + 3) / 4) * 4 + 4; }

__host__ __device__ inline size_t sweep_smem_bytes(int n, int kmax) {
  const int rows16 = ((n + 15) / 16) * 16, zld = sweep_zld(kmax);
  size_t d = (size_t)NWARP * 2 * RT_MAX * 8 * WS_LD + (size_t)RING * KB * KS_LD + 2 * (size_t)TC * zld + 2 * TC +
             2 * (size_t)rows16 + 4 * NWARP * TC + 2 * RING + 8;
  return d * sizeof(double) + (size_t)kmax * MAXBIT * sizeof(int);
}

// MULTI: one row of CTAs per run (run = blockIdx.y), values only -- the raw-sample scoring pass of the BO loop
template <bool MULTI>
__global__ void __launch_bounds__(SW_THREADS, 1) ei_sweep_kernel(SweepArgs a) {
  extern __shared__ __align__(16) double smd[];
  const pcabo_gp_model &m = a.m;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = m.n, kmax = m.kmax;
  const long ld = m.ld_n;
  const int b = MULTI ? (int)blockIdx.y : a.b;
  const double *const Xc = MULTI ? a.X + (size_t)b * a.x_stride : a.X;
  double *const valsc = MULTI ? a.vals + (size_t)b * a.ncand : a.vals;
  const int kk = m.k ? m.k[b] : kmax;
  const int kk4 = (kk + 3) / 4;               // k-steps of the distance MMA
  const int zld = sweep_zld(kmax);
  const int rows8 = ((n + 7) / 8) * 8, rows16 = ((n + 15) / 16) * 16;
  const int nrt = rows8 / 8;
  constexpr int ww_sz = RT_MAX * 8 * WS_LD, ks_sz = KB * KS_LD;
  double *const sWw = smd + (size_t)warp * 2 * ww_sz;          // this warp's two W buffers
  double *const sKs = smd + (size_t)NWARP * 2 * ww_sz;
  double *const sxn = sKs + RING * ks_sz;
  double *const sxx = sxn + 2 * TC * zld;
  double *const szz = sxx + 2 * TC;
  double *const sal = szz + rows16;
  double *const sred = sal + rows16;
  unsigned long long *const full_bar = reinterpret_cast<unsigned long long *>(sred + 4 * NWARP * TC);
  unsigned long long *const empty_bar = full_bar + RING;
  unsigned long long *const xn_bar = empty_bar + RING;        // [2] candidates of the even / odd tile are in place
  unsigned long long *const red_bar = xn_bar + 2;             // [2] partial sums of the even / odd tile are in place
  int *const ssst = reinterpret_cast<int *>(empty_bar + RING + 8);
  const double *W = m.W + (size_t)b * ld * ld;
  const double *Zn = m.Zn + (size_t)b * ld * kmax;
  const double *alpha = m.alpha + (size_t)b * ld;
  const double *off = m.norm_off + (size_t)b * kmax, *coef = m.norm_coef + (size_t)b * kmax;
  const double ell = m.lengthscales ? m.lengthscales[b] : m.lengthscale;
  const double inv_ell2 = 1.0 / (ell * ell);
  const double ystd = m.ystd[b], ymean = m.ymean[b], best_f = m.best_f[b], sgn = m.maximize ? 1.0 : -1.0;
  const double d0 = 1e-15;
  const double kss = (SQRT5 * d0 + 1.0 + (5.0 / 3.0) * (d0 * d0)) * exp(-SQRT5 * d0);

  if (Xc == nullptr)
    for (int i = tid; i < kk * MAXBIT; i += SW_THREADS) ssst[i] = a.sstate[i];
  // squared norms of the training rows and alpha (once per CTA); zero padding of the candidate buffers
  for (int j = tid; j < rows16; j += SW_THREADS) {
    double acc = 0.0;
    if (j < n)
      for (int dmm = 0; dmm < kk; ++dmm) { const double z = Zn[(size_t)j * kmax + dmm]; acc = fma(z, z, acc); }
    szz[j] = acc;
    sal[j] = (j < n) ? alpha[j] : 0.0;
  }
  for (int i = tid; i < 2 * TC * zld; i += SW_THREADS) sxn[i] = 0.0;
  if (tid < RING) { mbar_init(&full_bar[tid], 4); mbar_init(&empty_bar[tid], NWARP); }
  if (tid < 2) { mbar_init(&xn_bar[tid], NWARP); mbar_init(&red_bar[tid], NWARP); }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();   // zero padding and barriers are in place before any thread fills the candidate buffers

  double cta_best = -INFINITY;
  long long cta_idx = 0x7fffffffffffffffLL;
  const long long ntiles = (a.ncand + TC - 1) / TC;
  const int nchunks = (n + KB - 1) / KB;
  const int ar = lane >> 2, ak = lane & 3;
  // tiles of this CTA: blockIdx.x + k * gridDim.x, k = 0 .. ntk - 1
  const long long ntk = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long npanels = ntk * nchunks;     // global panel sequence of this CTA
  const int ahead = min(AHEAD, nchunks);       // production never runs further than into the next tile

  // ---- this warp's share (TC / NWARP candidates) of the normalised candidates of local tile k -> xn buffer k & 1,
  //      their squared norms, then one arrival on the tile-parity barrier: no block-wide barrier per tile
  auto gen_candidates = [&](long long k) {
    const long long c0 = ((long long)blockIdx.x + k * gridDim.x) * TC;
    const int par = (int)(k & 1);
    double *xb = sxn + (size_t)par * TC * zld;
    for (int cc = 0; cc < TC / NWARP; ++cc) {
      const int c = warp * (TC / NWARP) + cc;
      const long long ci = c0 + c;
      double ss = 0.0;
      for (int j = lane; j < kk; j += 32) {
        double x = 0.0;
        if (ci < a.ncand) {
          if (Xc != nullptr) {
            x = Xc[(size_t)ci * kmax + j];
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
        const double xv = (x - off[j]) / coef[j];
        xb[c * zld + j] = xv;
        ss = fma(xv, xv, ss);
      }
      ss = warp_sum(ss);
      if (lane == 0) sxx[par * TC + c] = ss;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&xn_bar[par]);
  };

  // partial alpha . k of this thread's two candidates (its K* tiles), one set per tile parity
  double mean_e0 = 0.0, mean_e1 = 0.0, mean_o0 = 0.0, mean_o1 = 0.0;   // even / odd tiles

  // ---- per-lane constants of the hot loops
  const int swz = ar & 3;
  int aoff[KB / 4];                                  // A fragment offsets inside a staged row tile (swizzled columns)
#pragma unroll
  for (int k4 = 0; k4 < KB / 4; ++k4) aoff[k4] = ar * WS_LD + (((k4 ^ swz) << 2) | ak);
  const int boff = ak * KS_LD + ar;                  // B fragment offset inside a K* panel
  int ntw = 0;                                       // row tiles owned by this warp: rt = warp + t * 16 < nrt
#pragma unroll
  for (int t = 0; t < RT_MAX; ++t) ntw += (warp + t * NWARP < nrt) ? 1 : 0;
  // W staging: lane copies chunk h (columns 2h, 2h+1) of rows r and r + 4 of each of its row tiles
  const int cr = lane >> 3, ch = lane & 7;
  const int wdst = cr * WS_LD + ((2 * ch) ^ (cr << 2));
  const double *wsrc = W + (size_t)(warp * 8 + cr) * ld + 2 * ch;
  const double *wlast = W + (size_t)(n - 1) * ld + 2 * ch;
  const size_t wsrc_t = (size_t)NWARP * 8 * ld, wsrc_q = 4 * (size_t)ld;
  unsigned rowok = 0;                                // bit 2 t + q: row (warp + 16 t) * 8 + cr + 4 q exists
#pragma unroll
  for (int t = 0; t < RT_MAX; ++t)
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if ((warp + t * NWARP) * 8 + cr + 4 * q < n) rowok |= 1u << (2 * t + q);
  // first row tile of this warp that reaches down to column j0 (tiles entirely above the diagonal are skipped)
  auto first_tile = [&](int j0) {
    const int need = j0 - 7 - warp * 8;              // rt * 8 + 7 >= j0  <=>  t * (8 NWARP) >= need
    return need <= 0 ? 0 : (need + 8 * NWARP - 1) / (8 * NWARP);
  };
  // K* production only touches whole 4-column groups of the training rows when they exist in memory
  const bool zfast = kk4 * 4 <= kmax;
  const int ptc = warp / DUTY;                       // candidate column tile this warp produces (both row tiles)

  // ---- producer: the 16 x 8 strip (both row tiles, column tile ptc) of K* panel `chunk` of the tile with parity `par`
  //      into ring slot `slot`. Four independent DMMA chains and four Matern evaluations per thread: the latency of
  //      the FP64 chains (which share the pipe with the other warps' DMMAs) overlaps instead of adding up.
  auto produce = [&](int chunk, int par, int slot, bool wait_empty, unsigned empty_parity, unsigned xn_parity) {
    const int j0 = chunk * KB;
    mbar_wait(&xn_bar[par], xn_parity);               // complete long ago except right after the prologue
    const int jr0 = min(j0 + ar, n - 1), jr1 = min(j0 + 8 + ar, n - 1);
    const bool on0 = j0 + ar < n, on1 = j0 + 8 + ar < n;
    double a00 = 0.0, a01 = 0.0, b00 = 0.0, b01 = 0.0;     // rows j0 + ar     : even / odd k-steps
    double a10 = 0.0, a11 = 0.0, b10 = 0.0, b11 = 0.0;     // rows j0 + 8 + ar
    const double *z0p = Zn + (size_t)jr0 * kmax + ak;
    const double *z1p = Zn + (size_t)jr1 * kmax + ak;
    const double *xrow = sxn + (size_t)par * TC * zld + (ptc * 8 + ar) * zld + ak;
    if (zfast) {
      int k4 = 0;
      for (; k4 + 2 <= kk4; k4 += 2) {
        const double u0 = __ldg(z0p + k4 * 4), u1 = __ldg(z0p + k4 * 4 + 4);
        const double v0 = __ldg(z1p + k4 * 4), v1 = __ldg(z1p + k4 * 4 + 4);
        const double x0 = xrow[k4 * 4], x1 = xrow[k4 * 4 + 4];
        dmma_8x8x4(a00, a01, u0, x0);
        dmma_8x8x4(a10, a11, v0, x0);
        dmma_8x8x4(b00, b01, u1, x1);
        dmma_8x8x4(b10, b11, v1, x1);
      }
      if (k4 < kk4) {
        const double x0 = xrow[k4 * 4];
        dmma_8x8x4(a00, a01, __ldg(z0p + k4 * 4), x0);
        dmma_8x8x4(a10, a11, __ldg(z1p + k4 * 4), x0);
      }
    } else {
      for (int k4 = 0; k4 < kk4; ++k4) {
        const bool cin = k4 * 4 + ak < kk;
        const double x0 = xrow[k4 * 4];
        dmma_8x8x4(a00, a01, cin ? __ldg(z0p + k4 * 4) : 0.0, x0);
        dmma_8x8x4(a10, a11, cin ? __ldg(z1p + k4 * 4) : 0.0, x0);
      }
    }
    const double zz0 = szz[j0 + ar], zz1 = szz[j0 + 8 + ar];
    const double al0 = sal[j0 + ar], al1 = sal[j0 + 8 + ar];
    const double *xx = sxx + par * TC + ptc * 8 + ak * 2;
    const double xx0 = xx[0], xx1 = xx[1];
    double k00 = matern52_fast(fmax((zz0 + xx0) - 2.0 * (a00 + b00), 0.0) * inv_ell2);
    double k01 = matern52_fast(fmax((zz0 + xx1) - 2.0 * (a01 + b01), 0.0) * inv_ell2);
    double k10 = matern52_fast(fmax((zz1 + xx0) - 2.0 * (a10 + b10), 0.0) * inv_ell2);
    double k11 = matern52_fast(fmax((zz1 + xx1) - 2.0 * (a11 + b11), 0.0) * inv_ell2);
    if (!on0) { k00 = 0.0; k01 = 0.0; }
    if (!on1) { k10 = 0.0; k11 = 0.0; }
    if (par) { mean_o0 = fma(k00, al0, fma(k10, al1, mean_o0)); mean_o1 = fma(k01, al0, fma(k11, al1, mean_o1)); }
    else { mean_e0 = fma(k00, al0, fma(k10, al1, mean_e0)); mean_e1 = fma(k01, al0, fma(k11, al1, mean_e1)); }
    // the slot is free once all warps consumed the panel that used it RING panels ago
    if (wait_empty) mbar_wait(&empty_bar[slot], empty_parity);
    double *ks = sKs + slot * ks_sz + ar * KS_LD + ptc * 8 + ak * 2;
    *reinterpret_cast<double2 *>(ks) = make_double2(k00, k01);
    *reinterpret_cast<double2 *>(ks + 8 * KS_LD) = make_double2(k10, k11);
    __syncwarp();
    if (lane == 0) mbar_arrive(&full_bar[slot]);
  };

  // ---- this warp's rows of W panel `chunk` -> its staging buffer `buf` (row tiles above the diagonal are skipped)
  auto issue_w = [&](int chunk, int buf) {
    const int j0 = chunk * KB;
    double *dst = sWw + buf * ww_sz + wdst;
    const int t0 = first_tile(j0);
    // last panel of a matrix whose row is shorter than the panel: stay inside the (even) row; the clamped columns are
    // >= n and are masked in the MMA loop of the last panel
    const int jc = (j0 + KB <= (int)ld) ? j0 : min(j0 + 2 * ch, (int)ld - 2) - 2 * ch;
    const double *src = wsrc + jc;
    const double *slast = wlast + jc;               // padding rows of the last row tile re-read row n - 1 (never used)
#pragma unroll
    for (int t = 0; t < RT_MAX; ++t)
      if (t >= t0 && t < ntw) {
        cp_async16(dst + t * 8 * WS_LD, (rowok & (1u << (2 * t))) ? src + t * wsrc_t : slast);
        cp_async16(dst + (t * 8 + 4) * WS_LD, (rowok & (2u << (2 * t))) ? src + t * wsrc_t + wsrc_q : slast);
      }
  };

  // ---- prologue: candidates of tile 0, first `ahead` panels of K*, W panel 0
  if (ntk > 0) gen_candidates(0);
  const int npan = (int)npanels;
  // producer cursor: panel pG = G + ahead of the global sequence, as (chunk, tile parity)
  int pG = 0, pchunk = 0, ptile = 0;
  auto advance_producer = [&]() {
    ++pG;
    if (++pchunk == nchunks) { pchunk = 0; ++ptile; }
  };
  for (; pG < ahead && pG < npan; advance_producer())
    if (((warp + pG) % DUTY) == 0) produce(pchunk, ptile & 1, pG & (RING - 1), false, 0u, (unsigned)((ptile >> 1) & 1));
  if (npan > 0) issue_w(0, 0);
  cp_async_commit();

  int G = 0;
  SWP_T0();
  for (long long k = 0; k < ntk; ++k) {
    const long long c0 = ((long long)blockIdx.x + k * gridDim.x) * TC;
    // candidates of the next tile (its K* production starts `ahead` panels before this tile ends)
    if (k + 1 < ntk) gen_candidates(k + 1);
    // accumulators: warp owns row tiles rt = warp, warp + 16, ...; 4 column tiles each
    double acc[RT_MAX][CT][2];
#pragma unroll
    for (int t = 0; t < RT_MAX; ++t)
#pragma unroll
      for (int c = 0; c < CT; ++c) { acc[t][c][0] = 0.0; acc[t][c][1] = 0.0; }

    for (int chunk = 0; chunk < nchunks; ++chunk, ++G) {
      const int wbuf = G & 1;
      SWP(0);
      // K* tile of the panel `ahead` panels further on: warps alternate this duty. (Before the W prefetch: the release
      // arrival at the end of produce() waits for this warp's outstanding cp.async copies.)
      if (pG < npan) {
        if (((warp + pG) % DUTY) == 0)
          produce(pchunk, ptile & 1, pG & (RING - 1), pG >= RING, (unsigned)(((pG / RING) - 1) & 1),
                  (unsigned)((ptile >> 1) & 1));
        advance_producer();
      }
      SWP(1);
      // W of the next panel of the sequence (wraps to panel 0 of the next tile)
      if (G + 1 < npan) issue_w(chunk + 1 < nchunks ? chunk + 1 : 0, wbuf ^ 1);
      cp_async_commit();
      SWP(2);
      cp_async_wait<1>();
      __syncwarp();
      SWP(3);
      const int slot = G & (RING - 1);
      mbar_wait(&full_bar[slot], (unsigned)((G / RING) & 1));
      SWP(4);
      // ---- T += Wpanel (rows x 16) * Kspanel (16 x 32) on the FP64 tensor cores
      {
        const double *Wsb = sWw + wbuf * ww_sz;
        const double *Ksb = sKs + slot * ks_sz + boff;
        const int t0 = first_tile(chunk * KB);
        if (chunk == nchunks - 1) {
          const int col0 = chunk * KB + ak;
          if (t0 == 0) mma_panel<0, true>(acc, Wsb, Ksb, aoff, ntw, col0, n);
          else if (t0 == 1) mma_panel<1 % RT_MAX, true>(acc, Wsb, Ksb, aoff, ntw, col0, n);
          else if (RT_MAX > 2 && t0 == 2) mma_panel<2 % RT_MAX, true>(acc, Wsb, Ksb, aoff, ntw, col0, n);
          else if (RT_MAX > 3 && t0 == 3) mma_panel<3 % RT_MAX, true>(acc, Wsb, Ksb, aoff, ntw, col0, n);
        } else {
          if (t0 == 0) mma_panel<0, false>(acc, Wsb, Ksb, aoff, ntw, 0, n);
          else if (t0 == 1) mma_panel<1 % RT_MAX, false>(acc, Wsb, Ksb, aoff, ntw, 0, n);
          else if (RT_MAX > 2 && t0 == 2) mma_panel<2 % RT_MAX, false>(acc, Wsb, Ksb, aoff, ntw, 0, n);
          else if (RT_MAX > 3 && t0 == 3) mma_panel<3 % RT_MAX, false>(acc, Wsb, Ksb, aoff, ntw, 0, n);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(&empty_bar[slot]);
      SWP(5);
    }
    SWP(0);
    // ---- epilogue of the tile. Variance partials: thread holds T[row = rt*8 + lane/4][col = c*8 + (lane%4)*2 + {0,1}]
    double vpart[CT][2];
#pragma unroll
    for (int c = 0; c < CT; ++c) { vpart[c][0] = 0.0; vpart[c][1] = 0.0; }
#pragma unroll
    for (int t = 0; t < RT_MAX; ++t)
      if ((warp + t * NWARP) * 8 + ar < n) {          // padding rows of the last row tile carry no L^-1 row
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          vpart[c][0] = fma(acc[t][c][0], acc[t][c][0], vpart[c][0]);
          vpart[c][1] = fma(acc[t][c][1], acc[t][c][1], vpart[c][1]);
        }
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
    double mp[2];
    if (k & 1) { mp[0] = mean_o0; mp[1] = mean_o1; mean_o0 = 0.0; mean_o1 = 0.0; }
    else { mp[0] = mean_e0; mp[1] = mean_e1; mean_e0 = 0.0; mean_e1 = 0.0; }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double v = mp[h];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      mp[h] = v;
    }
    // partial sums of this tile -> buffer of its parity; one arrival per warp. The warp on epilogue duty (they rotate)
    // waits for all sixteen, everybody else goes straight on to the next tile.
    const int epar = (int)(k & 1);
    double *rb = sred + (size_t)epar * 2 * NWARP * TC;
    if (lane < 4) {
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        rb[warp * TC + c * 8 + lane * 2 + 0] = vpart[c][0];
        rb[warp * TC + c * 8 + lane * 2 + 1] = vpart[c][1];
      }
      rb[NWARP * TC + warp * 8 + lane * 2 + 0] = mp[0];   // [warp][8 candidates of tile column warp / DUTY]
      rb[NWARP * TC + warp * 8 + lane * 2 + 1] = mp[1];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&red_bar[epar]);
    if (warp == (int)(k % NWARP)) {
      mbar_wait(&red_bar[epar], (unsigned)((k >> 1) & 1));
      const int c = lane;
      double v = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < NWARP; ++w8) v += rb[w8 * TC + c];
      double ms = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < NWARP; ++w8)
        if (w8 / DUTY == c / 8) ms += rb[NWARP * TC + w8 * 8 + (c & 7)];
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
        if (valsc) valsc[ci] = val;
      } else {
        val = -INFINITY;
      }
      // tile arg-max (first occurrence) over the 32 candidates
      long long gi = a.idx0 + ci;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, val, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, gi, o);
        if (ov > val || (ov == val && oi < gi)) { val = ov; gi = oi; }
      }
      if (val > cta_best || (val == cta_best && gi < cta_idx)) { cta_best = val; cta_idx = gi; }
    }
    // (the buffer of this parity is rewritten two tiles later, which nobody reaches before the duty warp, in program
    // order behind this epilogue, has consumed the panels in between)
    SWP(6);
  }
  cp_async_wait<0>();
  // every warp carries the best of the tiles it closed: combine
  __syncthreads();
  double *fin_v = sred;
  long long *fin_i = reinterpret_cast<long long *>(sred + NWARP);
  if (lane == 0) { fin_v[warp] = cta_best; fin_i[warp] = cta_idx; }
  __syncthreads();
  if (!MULTI && tid == 0) {
    double bv = -INFINITY;
    long long bi = 0x7fffffffffffffffLL;
    for (int w8 = 0; w8 < NWARP; ++w8)
      if (fin_v[w8] > bv || (fin_v[w8] == bv && fin_i[w8] < bi)) { bv = fin_v[w8]; bi = fin_i[w8]; }
    a.best_val[blockIdx.x] = bv;
    a.best_idx[blockIdx.x] = bi;
  }
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

#ifdef SWEEP_PROF
extern "C" int pcabo_sweep_prof_read(long long *out) {
  return cudaMemcpyFromSymbol(out, g_sweep_prof, sizeof(long long) * 16 * 8) == cudaSuccess ? 0 : -3;
}
#endif

extern "C" int pcabo_ei_sweep(const pcabo_gp_model *m, int run, long long ncand, long long idx0, const double *X,
                              const int *shift, const int *sstate, const double *lo, const double *hi, double *vals,
                              double *part_val, long long *part_idx, int nparts, double *best_val,
                              long long *best_idx, void *stream) {
  if (!m || run < 0 || run >= m->B || ncand <= 0 || !part_val || !part_idx || nparts <= 0 || !best_val || !best_idx ||
      !m->W || !m->Zn || !m->alpha || !m->norm_off || !m->norm_coef || !m->ymean || !m->ystd || !m->best_f)
    return PCABO_E_INVALID;
  if (X == nullptr && (!shift || !sstate || !lo || !hi)) return PCABO_E_INVALID;
  if (m->acq != PCABO_ACQ_LOG_EI) return PCABO_E_UNSUPPORTED;   // the sweep scores log-EI (BASELINE config C4)
  // n <= 512 (accumulators of the whole T tile live in registers), even ld_n and kmax (16-byte cp.async copies)
  if (m->n > 8 * NWARP * RT_MAX || m->n < 2 || (m->ld_n & 1) || (m->kmax & 1) || m->kmax > 104) return PCABO_E_UNSUPPORTED;
  const size_t smem = sweep_smem_bytes(m->n, m->kmax);
  if (smem > 227 * 1024) return PCABO_E_UNSUPPORTED;
  PCABO_CUDA_CHECK(cudaFuncSetAttribute(ei_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SweepArgs a;
  a.m = *m; a.b = run; a.ncand = ncand; a.idx0 = idx0; a.X = X; a.shift = shift; a.sstate = sstate; a.lo = lo; a.hi = hi;
  a.vals = vals; a.best_val = part_val; a.best_idx = part_idx; a.multi = 0; a.x_stride = 0;
  const long long ntiles = (ncand + TC - 1) / TC;
  const int grid = (int)(ntiles < nparts ? ntiles : nparts);
  cudaStream_t st = as_stream(stream);
  ei_sweep_kernel<false><<<grid, SW_THREADS, smem, st>>>(a);
  PCABO_LAUNCH_CHECK();
  sweep_final_kernel<<<1, 256, 0, st>>>(grid, part_val, part_idx, best_val, best_idx);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}

// Host-only: 1 when pcabo_logei_raw_runs covers a model of this shape (the tile pipeline keeps n x 32 accumulators in
// registers and stages candidates, K* panels and per-warp W tiles in shared memory), 0 otherwise.
extern "C" int pcabo_logei_raw_supported(int n, int kmax, int ld_n, int acq) {
  if (acq != PCABO_ACQ_LOG_EI) return 0;
  if (n > 8 * NWARP * RT_MAX || n < 2 || ld_n < n || (ld_n & 1) || (kmax & 1) || kmax < 2 || kmax > 104) return 0;
  return sweep_smem_bytes(n, kmax) <= 227 * 1024 ? 1 : 0;
}

// Multi-run form of the sweep: the raw-sample scoring pass of optimize_acqf (PCA_BO.py:607-614; 512 scrambled-Sobol
// points per run) for ALL runs of the batch in one launch, values out, no arg-max. Each run gets a row of `ctas_per_run`
// CTAs that share its candidate tiles (32 candidates each).
extern "C" int pcabo_logei_raw_runs(const pcabo_gp_model *m, int npts, const double *X, double *val, int ctas_per_run,
                                    void *stream) {
  if (!m || npts <= 0 || !X || !val || m->B <= 0 || !m->W || !m->Zn || !m->alpha || !m->norm_off || !m->norm_coef ||
      !m->ymean || !m->ystd || !m->best_f || (!m->lengthscales && !(m->lengthscale > 0.0)))
    return PCABO_E_INVALID;
  if (m->acq != PCABO_ACQ_LOG_EI) return PCABO_E_UNSUPPORTED;
  if (m->n > 8 * NWARP * RT_MAX || m->n < 2 || (m->ld_n & 1) || (m->kmax & 1) || m->kmax > 104 || m->B > 65535)
    return PCABO_E_UNSUPPORTED;
  const size_t smem = sweep_smem_bytes(m->n, m->kmax);
  if (smem > 227 * 1024) return PCABO_E_UNSUPPORTED;
  PCABO_CUDA_CHECK(cudaFuncSetAttribute(ei_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SweepArgs a;
  a.m = *m; a.b = 0; a.ncand = npts; a.idx0 = 0; a.X = X; a.shift = nullptr; a.sstate = nullptr; a.lo = nullptr;
  a.hi = nullptr; a.vals = val; a.best_val = nullptr; a.best_idx = nullptr; a.multi = 1;
  a.x_stride = (long long)npts * m->kmax;
  const int ntiles = (npts + TC - 1) / TC;
  if (ctas_per_run <= 0) {   // as many CTAs per run as keep the whole launch within one wave of the 148 SMs
    ctas_per_run = 1;
    while (2 * ctas_per_run <= ntiles && 2 * ctas_per_run * m->B <= 148) ctas_per_run *= 2;
  }
  if (ctas_per_run > ntiles) ctas_per_run = ntiles;
  dim3 grid(ctas_per_run, m->B);
  ei_sweep_kernel<true><<<grid, SW_THREADS, smem, as_stream(stream)>>>(a);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}
