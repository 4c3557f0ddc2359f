// log-EI value + gradient of up to 8 points of one run, as the device L-BFGS-B kernel (lbfgsb.cu) evaluates it
// 60-280 times per launch. Same mathematics as logei_core.cuh (botorch LogExpectedImprovement / ProbabilityOfImprovement
// on the reference's never-fitted Matern-5/2 GP: PCA_BO.py:199-203, 502-545, 607-614), different
// data movement:
//
//   * ONE sweep over W = L^-1 instead of a pass over W^T (t = W k) and a pass over W (s = W^T t). W is cut in row
//     panels of R rows x (columns up to the diagonal block); once panel I is in shared memory, t_I = W_I k is complete
//     (a row of t only needs the columns left of the diagonal) and is used at once for s += W_I^T t_I. Both products
//     run on the FP64 tensor cores (mma.sync m8n8k4, SASS DMMA.8x8x4; the 5 points of a chunk are padded to the 8
//     columns of the B operand), which also does the cross-lane reductions of the dot products.
//   * The panels are staged by the bulk-copy (TMA) engine: one elected thread issues cp.async.bulk row copies that
//     signal an mbarrier per panel with their byte count. The ring is a byte FIFO of variable-size panels; while a
//     panel is consumed the next ones are in flight, and after the last panel of an evaluation the first panels of the
//     NEXT evaluation are requested, so they land while the optimiser warp does its L-BFGS-B bookkeeping. If all
//     panels fit the ring (n up to ~200) they are loaded once per launch and every later evaluation runs from shared
//     memory only.
//   * s stays in the accumulator registers of the warps that own the column tiles and is turned into the gradient
//     coefficients c_j = dk_j/d(dist^2) (dlogEI/dmu alpha_j - 2 dlogEI/dvar s_j) straight from there.
#pragma once
#include "async_copy.cuh"
#include "logei_core.cuh"

namespace lbeval {

constexpr int NT = 256;         // threads per block
constexpr int NW = NT / 32;
constexpr int NBAR = 16;        // panels in flight (one mbarrier each)
constexpr int NPOFF = 32;       // ring offsets of the last NPOFF issued panels
constexpr int PRODUCER = NT - 1;

// smallest y >= x with y % 16 == 4: a row stride (in doubles) that makes both DMMA operand access patterns of the
// sweep (8 rows x 4 columns and 4 rows x 8 columns per warp) bank-conflict free, and keeps rows 16-byte aligned
__host__ __device__ inline int pad4mod16(int x) { return ((x + 11) / 16) * 16 + 4; }

__host__ __device__ inline int imin(int a, int b) { return a < b ? a : b; }

// row stride of panel I (rows I R .. I R + R - 1, columns 0 .. min((I + 1) R, n) - 1)
// (the column count is rounded up to a whole 8-column tile: the transposed product reads whole tiles)
__host__ __device__ inline int panel_rs(int I, int R, int n) { return pad4mod16((imin((I + 1) * R, n) + 7) & ~7); }

struct Layout {   // offsets in doubles from the start of the evaluation scratch (all even: 16-byte aligned)
  int S, R, npanels;
  int o_xn, o_kst, o_gf, o_alpha, o_part, o_red, o_bar, o_poff, o_ring;
  int max_panel;          // doubles of the largest panel
  long long all_panels;   // doubles of all panels together
};

__host__ __device__ inline Layout make_layout(int PT, int kmax, int n, int R) {
  Layout L;
  L.R = R;
  L.S = pad4mod16((n + 3) & ~3);
  L.npanels = (n + R - 1) / R;
  int o = 0;
  L.o_xn = o; o += (PT * kmax + 1) & ~1;
  L.o_kst = o; o += PT * L.S;
  L.o_gf = o; o += PT * L.S;
  L.o_alpha = o; o += L.S;
  const int part_a = 2 * NW * 8 * 12, part_g = NT * PT;   // double-buffered split-K partials of t (NSPLIT R = 8 NW rows of
  L.o_part = o; o += (part_a > part_g ? part_a : part_g);  // 12) / partials of the gradient (never live together)
  L.o_red = o; o += NW * PT + NW * 8 + 2 * PT + 2 + ((NW * PT + 2 * PT) & 1);
  L.o_bar = o; o += NBAR;
  L.o_poff = o; o += NPOFF / 2;
  L.o_ring = o;
  long long all = 0;
  int mx = 0;
  for (int I = 0; I < L.npanels; ++I) {
    const int sz = R * panel_rs(I, R, n);
    all += sz;
    if (sz > mx) mx = sz;
  }
  L.all_panels = all;
  L.max_panel = mx;
  return L;
}

// Producer-side state of the panel ring: a byte FIFO of variable-size panels (all offsets in doubles).
struct RingState {
  unsigned next, oldest;   // sequence numbers: next panel to request, oldest panel not yet released
  unsigned cons;           // next panel to consume
  int lo, hi;              // live region: [lo, hi), or [lo, end) + [0, hi) when wrapped
  bool wrapped;
};

// Place a panel of sz doubles behind the newest live panel (wrapping to offset 0 when it does not fit before the end
// of the ring). Returns false when there is no room until the oldest panel is released.
__host__ __device__ inline bool ring_place(RingState &r, int sz, int cap, int &off) {
  if (r.next == r.oldest) { r.lo = 0; r.hi = 0; r.wrapped = false; }
  if (!r.wrapped) {
    if (r.hi + sz <= cap) { off = r.hi; r.hi += sz; return true; }
    if (sz <= r.lo) { off = 0; r.hi = sz; r.wrapped = true; return true; }
    return false;
  }
  if (r.hi + sz <= r.lo) { off = r.hi; r.hi += sz; return true; }
  return false;
}

// new_lo: ring offset of the panel that becomes the oldest (ignored when none is left)
__host__ __device__ inline void ring_pop(RingState &r, int new_lo) {
  ++r.oldest;
  if (r.next != r.oldest) {
    if (new_lo < r.lo) r.wrapped = false;
    r.lo = new_lo;
  }
}

__host__ __device__ inline bool ring_may_request(const RingState &r, int npanels, bool resident) {
  if (resident ? (r.next >= (unsigned)npanels) : (r.next >= r.cons + (unsigned)npanels)) return false;   // one evaluation ahead
  return r.next - r.oldest < (unsigned)NBAR;
}

// Host replay of the request / release schedule of eval() for `evals` evaluations: true when a consumer never meets a
// panel that cannot have been requested (which would dead-lock the block). tight: the panel is released behind an
// extra block barrier right after its transposed product instead of at the first barrier of the next panel.
inline bool ring_schedule_ok(int n, int R, int cap, bool resident, bool tight, int evals = 3) {
  const int np = (n + R - 1) / R;
  RingState r{0, 0, 0, 0, 0, false};
  int poff[NPOFF];
  auto pump = [&]() {
    while (ring_may_request(r, np, resident)) {
      int off;
      if (!ring_place(r, R * panel_rs((int)(r.next % (unsigned)np), R, n), cap, off)) break;
      poff[r.next & (NPOFF - 1)] = off;
      ++r.next;
    }
  };
  auto pop = [&]() { ring_pop(r, poff[(r.oldest + 1) & (NPOFF - 1)]); };
  pump();
  for (int e = 0; e < evals; ++e) {
    for (int I = 0; I < np; ++I) {
      const unsigned seq = resident ? (unsigned)I : r.cons;
      if (seq >= r.next) return false;
      if (!resident && !tight && I > 0) pop();
      pump();
      ++r.cons;
      if (!resident && tight) { pop(); pump(); }
    }
    if (!resident && !tight) pop();
    pump();
  }
  return true;
}

struct Ctx {
  int n, kmax, kk, ld, acq;
  const double *ZnT, *Zn, *W, *alpha_g, *off, *coef;
  double inv_ell2, ymean, ystd, best_f, sgn;
  Layout L;
  double *base;        // shared-memory scratch
  int ring_cap;        // doubles
  bool resident;       // every panel stays in the ring for the whole launch
  bool tight;          // ring smaller than three of the largest panels: release each panel behind its own barrier
  bool loaded;         // resident panels have been waited for once
  RingState rs;        // cons is kept by every thread, the rest by the PRODUCER thread only
  long long *prof;     // optional cycle counters (slots 16..22, stamp in 31)
};

__device__ __forceinline__ void ev_tick(const Ctx &c, int id) {
  if (c.prof != nullptr && threadIdx.x == 0) {
    const long long t = clock64();
    if (id >= 0) c.prof[id] += t - c.prof[31];
    c.prof[31] = t;
  }
}

__device__ inline void load_ctx(Ctx &c, const pcabo_gp_model &m, int b) {
  c.n = m.n; c.kmax = m.kmax; c.ld = m.ld_n; c.acq = m.acq;
  c.kk = m.k ? m.k[b] : m.kmax;
  c.ZnT = m.ZnT + (size_t)b * m.kmax * m.ld_n;
  c.Zn = m.Zn + (size_t)b * m.ld_n * m.kmax;
  c.W = m.W + (size_t)b * m.ld_n * m.ld_n;
  c.alpha_g = m.alpha + (size_t)b * m.ld_n;
  c.off = m.norm_off + (size_t)b * m.kmax;
  c.coef = m.norm_coef + (size_t)b * m.kmax;
  c.inv_ell2 = 1.0 / (m.lengthscale * m.lengthscale);
  c.ymean = m.ymean[b]; c.ystd = m.ystd[b]; c.best_f = m.best_f[b];
  c.sgn = m.maximize ? 1.0 : -1.0;
  c.prof = nullptr;
}

// ---- producer side of the panel ring (thread PRODUCER only)
__device__ inline void ring_pump(Ctx &c) {
  const Layout &L = c.L;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(c.base + L.o_bar);
  int *poff = reinterpret_cast<int *>(c.base + L.o_poff);
  double *ring = c.base + L.o_ring;
  const int R = L.R, n = c.n, np = L.npanels;
  RingState &r = c.rs;
  while (ring_may_request(r, np, c.resident)) {
    const int I = (int)(r.next % (unsigned)np);
    const int RS = panel_rs(I, R, n);
    int off;
    if (!ring_place(r, R * RS, c.ring_cap, off)) break;
    poff[r.next & (NPOFF - 1)] = off;
    const int r0 = I * R, nrows = imin(R, n - r0);
    const unsigned bytes = (unsigned)(((imin(r0 + R, n) + 1) & ~1) * 8);
    unsigned long long *bar = &bars[r.next & (NBAR - 1)];
    acp::fence_proxy_async();
    acp::mbar_arrive_expect_tx(bar, (unsigned)nrows * bytes);
    const double *src = c.W + (size_t)r0 * c.ld;
    double *dst = ring + off;
    for (int q = 0; q < nrows; ++q) acp::bulk_g2s(dst + q * RS, src + (size_t)q * c.ld, bytes, bar);
    ++r.next;
  }
}

// the oldest panel has been consumed by every warp
__device__ inline void ring_release(Ctx &c) {
  const int *poff = reinterpret_cast<const int *>(c.base + c.L.o_poff);
  ring_pop(c.rs, poff[(c.rs.oldest + 1) & (NPOFF - 1)]);
}

// Called once per launch by all threads (ends with __syncthreads): barriers, zero padding, alpha, first requests.
__device__ inline void init(Ctx &c, int PT) {
  const Layout &L = c.L;
  const int tid = threadIdx.x;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(c.base + L.o_bar);
  if (tid == 0) {
    for (int i = 0; i < NBAR; ++i) acp::mbar_init(&bars[i], 1);
    acp::fence_barrier_init();
  }
  // kst / gf / alpha: entries >= n must be (and stay) zero: they meet the zero padding of the panels in the MMAs
  for (int i = tid; i < 2 * PT * L.S + L.S; i += NT) c.base[L.o_kst + i] = 0.0;
  for (int i = tid; i < c.ring_cap; i += NT) c.base[L.o_ring + i] = 0.0;   // never multiply uninitialised bits by zero
  __syncthreads();
  for (int i = tid; i < c.n; i += NT) c.base[L.o_alpha + i] = c.alpha_g[i];
  c.loaded = false;
  c.rs.next = 0; c.rs.oldest = 0; c.rs.cons = 0; c.rs.lo = 0; c.rs.hi = 0; c.rs.wrapped = false;
  if (tid == PRODUCER) ring_pump(c);
  __syncthreads();
}

// Called by all threads before the block exits (ends with __syncthreads): no bulk copy may still be in flight.
__device__ inline void drain(Ctx &c) {
  if (threadIdx.x == PRODUCER) {
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(c.base + c.L.o_bar);
    const unsigned first = c.resident ? (c.loaded ? c.rs.next : 0u) : c.rs.cons;
    for (unsigned s = first; s != c.rs.next; ++s) acp::mbar_wait(&bars[s & (NBAR - 1)], (s >> 4) & 1u);
  }
  __syncthreads();
}

// Value and gradient at pcount <= PT points. X: raw coordinates, point p at X + p * kk. val[p] <- acquisition value,
// grad[p * kk + j] <- gscale * d value / d x_j. All NT threads must call; ends with __syncthreads().
template <int PT, int R, int TI>
__device__ void eval(Ctx &c, int pcount, const double *X, double *val, double *grad, double gscale) {
  using logei::SQRT5;
  const Layout &L = c.L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = c.n, kmax = c.kmax, kk = c.kk, S = L.S;
  const long ld = c.ld;
  double *xn = c.base + L.o_xn, *kst = c.base + L.o_kst, *gf = c.base + L.o_gf, *alpha = c.base + L.o_alpha;
  double *part = c.base + L.o_part, *red = c.base + L.o_red, *ring = c.base + L.o_ring;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(c.base + L.o_bar);
  const int *poff = reinterpret_cast<const int *>(c.base + L.o_poff);
  const double inv_ell2 = c.inv_ell2;

  ev_tick(c, -1);
  for (int idx = tid; idx < PT * kmax; idx += NT) {
    const int p = idx / kmax, j = idx % kmax;
    xn[idx] = (p < pcount && j < kk) ? (X[p * kk + j] - c.off[j]) / c.coef[j] : 0.0;
  }
  __syncthreads();
  // ---- k vector and dk/d(dist^2) factor: thread per training point, coalesced ZnT, eight loads in flight
  double mpart[PT];
#pragma unroll
  for (int p = 0; p < PT; ++p) mpart[p] = 0.0;
  for (int i = tid; i < n; i += NT) {
    double sq[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) sq[p] = 0.0;
    int j = 0;
    for (; j + 8 <= kk; j += 8) {
      double z8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) z8[u] = c.ZnT[(size_t)(j + u) * ld + i];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int p = 0; p < PT; ++p) { const double d = z8[u] - xn[p * kmax + j + u]; sq[p] = fma(d, d, sq[p]); }
    }
    for (; j < kk; ++j) {
      const double z = c.ZnT[(size_t)j * ld + i];
#pragma unroll
      for (int p = 0; p < PT; ++p) { const double d = z - xn[p * kmax + j]; sq[p] = fma(d, d, sq[p]); }
    }
    const double al = alpha[i];
#pragma unroll
    for (int p = 0; p < PT; ++p) {
      const double s2 = sq[p] * inv_ell2;
      const double d = sqrt(fmax(s2, 1e-30));
      const double e = exp(-SQRT5 * d);
      const double kv = (SQRT5 * d + 1.0 + (5.0 / 3.0) * (d * d)) * e;
      kst[p * S + i] = kv;
      mpart[p] = fma(kv, al, mpart[p]);
      gf[p * S + i] = (s2 > 1e-30) ? -(5.0 / 3.0) * (1.0 + SQRT5 * d) * e * inv_ell2 : 0.0;
    }
  }
  __syncthreads();
  ev_tick(c, 16);

  // ---- the sweep over the row panels of W. Per panel: phase A (t_I = W_I k: warp -> row tile warp % RT and every
  //      NSPLIT-th k step, four independent accumulator chains), ONE block barrier, phase B (each lane folds the NSPLIT
  //      partials of the t values its B fragments need, then s_j += W_I^T t_I for the warp's column tiles, k step outer
  //      so that consecutive DMMAs feed different accumulators). Partials are double buffered: a warp that is done with
  //      phase B of panel I starts phase A of panel I + 1 without waiting for the others.
  double sf[TI][2];
#pragma unroll
  for (int ti = 0; ti < TI; ++ti) { sf[ti][0] = 0.0; sf[ti][1] = 0.0; }
  double vacc = 0.0;   // sum of t^2 (point lane >> 2) over the panels this warp accounts for
  constexpr int RT = R / 8, KQ = R / 4, NSPLIT = NW / RT, PLD = 12;   // PLD: row stride of the partials (conflict free)
  const int lr = lane >> 2, lc = lane & 3;
  const int wrt = warp % RT, wsp = warp / RT;
  for (int I = 0; I < L.npanels; ++I) {
    const unsigned seq = c.resident ? (unsigned)I : c.rs.cons;
    if (!c.resident || !c.loaded) acp::mbar_wait(&bars[seq & (NBAR - 1)], (seq >> 4) & 1u);
    const double *pan = ring + poff[seq & (NPOFF - 1)];
    const int r0 = I * R, c1 = imin(r0 + R, n), RS = panel_rs(I, R, n);
    double *pbuf = part + (I & 1) * (NSPLIT * R * PLD);
    {
      double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, c0 = 0.0, c1v = 0.0, d0 = 0.0, d1 = 0.0;
      const int ksteps = (c1 + 3) >> 2;
      const double *arow = pan + (wrt * 8 + lr) * RS + lc;
      const double *brow = kst + lr * S + lc;
      const bool bon = lr < PT;
      int ks = wsp;
      for (; ks + 3 * NSPLIT < ksteps; ks += 4 * NSPLIT) {
        const double w0 = arow[4 * ks], w1 = arow[4 * (ks + NSPLIT)], w2 = arow[4 * (ks + 2 * NSPLIT)],
                     w3 = arow[4 * (ks + 3 * NSPLIT)];
        const double k0 = bon ? brow[4 * ks] : 0.0, k1 = bon ? brow[4 * (ks + NSPLIT)] : 0.0,
                     k2 = bon ? brow[4 * (ks + 2 * NSPLIT)] : 0.0, k3 = bon ? brow[4 * (ks + 3 * NSPLIT)] : 0.0;
        acp::dmma_8x8x4(a0, a1, w0, k0);
        acp::dmma_8x8x4(b0, b1, w1, k1);
        acp::dmma_8x8x4(c0, c1v, w2, k2);
        acp::dmma_8x8x4(d0, d1, w3, k3);
      }
      for (; ks < ksteps; ks += NSPLIT) acp::dmma_8x8x4(a0, a1, arow[4 * ks], bon ? brow[4 * ks] : 0.0);
      *reinterpret_cast<double2 *>(pbuf + (wsp * R + wrt * 8 + lr) * PLD + 2 * lc) =
          make_double2((a0 + b0) + (c0 + d0), (a1 + b1) + (c1v + d1));
    }
    __syncthreads();
    if (tid == PRODUCER) {   // every warp is past phase B of the previous panel
      if (!c.resident && !c.tight && I > 0) ring_release(c);
      ring_pump(c);
    }
    {
      double bf[KQ];
#pragma unroll
      for (int kq = 0; kq < KQ; ++kq) {
        const int r = 4 * kq + lc;
        double t = 0.0;
#pragma unroll
        for (int sp = 0; sp < NSPLIT; ++sp) t += pbuf[(sp * R + r) * PLD + lr];
        bf[kq] = (r0 + r < n) ? t : 0.0;
      }
      if (warp == (I & (NW - 1))) {
#pragma unroll
        for (int kq = 0; kq < KQ; ++kq) vacc = fma(bf[kq], bf[kq], vacc);
      }
      const double *acol = pan + lc * RS + lr;
      const int ntile = (c1 + 7) >> 3;
#pragma unroll
      for (int kq = 0; kq < KQ; ++kq) {
#pragma unroll
        for (int ti = 0; ti < TI; ++ti) {
          const int ct = warp + NW * ti;
          if (ct < ntile) acp::dmma_8x8x4(sf[ti][0], sf[ti][1], acol[4 * kq * RS + ct * 8], bf[kq]);
        }
      }
    }
    ++c.rs.cons;
    if (c.tight) {   // small ring: hand the panel back at once so that the next large one can be requested
      __syncthreads();
      if (tid == PRODUCER) { ring_release(c); ring_pump(c); }
    }
  }
  c.loaded = true;
  // ---- reductions of the mean and variance partials
#pragma unroll
  for (int p = 0; p < PT; ++p) {
    const double ms = warp_sum(mpart[p]);
    if (lane == 0) red[warp * PT + p] = ms;
  }
  {
    // the four lanes lane & 3 of a point hold different rows: fold them, publish per (warp, point)
    double v = vacc;
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (lc == 0) red[NW * PT + warp * 8 + lr] = v;
  }
  __syncthreads();
  if (tid == PRODUCER) {
    if (!c.resident && !c.tight) ring_release(c);
    ring_pump(c);   // the first panels of the next evaluation travel while the optimiser warp works
  }
  ev_tick(c, 17);
  double *res = red + NW * PT + NW * 8;   // [2][PT]: d value / d mu_s, d value / d var_s
  if (tid < PT) {
    const int p = tid;
    double ms = 0.0, v = 0.0;
    for (int w8 = 0; w8 < NW; ++w8) ms += red[w8 * PT + p];
    for (int w8 = 0; w8 < NW; ++w8) v += red[NW * PT + w8 * 8 + p];
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
    double value, dmu, dvar;
    if (c.acq == PCABO_ACQ_PI) {
      value = 0.5 * erfc(-logei::INV_SQRT2 * u);
      const double phi = logei::INV_SQRT_2PI * exp(-0.5 * u * u);
      dmu = phi * c.sgn * ystd / sg;
      dvar = clamped ? 0.0 : -phi * u * ystd * ystd / (2.0 * var);
    } else {
      double lh, dlh;
      logei::log_h_and_grad(u, lh, dlh);
      value = lh + log(sg);
      dmu = dlh * c.sgn * ystd / sg;
      dvar = clamped ? 0.0 : (1.0 - u * dlh) * ystd * ystd / (2.0 * var);
    }
    if (p < pcount) val[p] = value;
    res[p] = dmu;
    res[PT + p] = dvar;
  }
  __syncthreads();
  ev_tick(c, 22);
  // ---- gradient coefficients straight from the accumulators: c_j = gf_j (dmu alpha_j - 2 dvar s_j), over the k vector
#pragma unroll
  for (int ti = 0; ti < TI; ++ti) {
    const int j = (warp + NW * ti) * 8 + lr;
    if (j < n) {
      const double al = alpha[j];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = 2 * lc + h;
        if (p < PT) kst[p * S + j] = gf[p * S + j] * (res[p] * al - 2.0 * res[PT + p] * sf[ti][h]);
      }
    }
  }
  __syncthreads();
  ev_tick(c, 19);
  // ---- grad[p][j] = gscale (1 / coef_j) sum_i c_i (xn_j - zn_ij): thread (ig, j), every IG-th training point
  {
    const int IG = NT / kk;   // >= 2 for kk <= 104
    const int ig = tid / kk, j = tid % kk;
    double acc[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) acc[p] = 0.0;
    if (ig < IG) {
      double xj[PT];
#pragma unroll
      for (int p = 0; p < PT; ++p) xj[p] = xn[p * kmax + j];
      int i = ig;
      for (; i + 7 * IG < n; i += 8 * IG) {
        double z8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) z8[u] = c.Zn[(size_t)(i + u * IG) * kmax + j];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int p = 0; p < PT; ++p) acc[p] = fma(kst[p * S + i + u * IG], xj[p] - z8[u], acc[p]);
      }
      for (; i < n; i += IG) {
        const double z = c.Zn[(size_t)i * kmax + j];
#pragma unroll
        for (int p = 0; p < PT; ++p) acc[p] = fma(kst[p * S + i], xj[p] - z, acc[p]);
      }
#pragma unroll
      for (int p = 0; p < PT; ++p) part[(ig * PT + p) * kk + j] = acc[p];
    }
    __syncthreads();
    for (int idx = tid; idx < pcount * kk; idx += NT) {
      const int p = idx / kk, jj = idx % kk;
      double sv = 0.0;
      for (int q = 0; q < IG; ++q) sv += part[(q * PT + p) * kk + jj];
      grad[idx] = gscale * sv / c.coef[jj];
    }
  }
  __syncthreads();
  ev_tick(c, 20);
}

}  // namespace lbeval
