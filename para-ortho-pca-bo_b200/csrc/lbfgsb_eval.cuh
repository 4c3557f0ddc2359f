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
//   * The panels are staged by the bulk-copy (TMA) engine: a dedicated producer warp (the ninth warp of the block; it
//     takes no part in the block barriers of the eight compute warps, which use a named barrier) issues cp.async.bulk
//     row copies that signal a "full" mbarrier per panel with their byte count; the compute warps hand a panel back
//     through an "empty" mbarrier (one arrival per warp) after its transposed product. The ring is a byte FIFO of
//     variable-size panels and the panel sequence is cyclic: after the last panel of an evaluation the first panels of
//     the NEXT evaluation are already requested and land while the optimiser warp does its L-BFGS-B bookkeeping. If
//     all panels fit the ring (n up to ~200) they are loaded once per launch and every later evaluation runs from
//     shared memory only.
//   * s stays in the accumulator registers of the warps that own the column tiles and is turned into the gradient
//     coefficients c_j = dk_j/d(dist^2) (dlogEI/dmu alpha_j - 2 dlogEI/dvar s_j) straight from there.
#pragma once
#include "async_copy.cuh"
#include "logei_core.cuh"
#include "w_packed.cuh"

namespace lbeval {

constexpr int NT = 256;         // threads per block
constexpr int NW = NT / 32;
constexpr int NBAR = 16;        // panels in flight (one mbarrier each)
constexpr int NPOFF = 32;       // ring offsets of the last NPOFF issued panels
constexpr int NT_ALL = NT + 32;  // + the producer warp

// block barrier of the NT compute threads (the producer warp never joins it)
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// smallest y >= x with y % 16 == 4: a row stride (in doubles) that makes both DMMA operand access patterns of the
// sweep (8 rows x 4 columns and 4 rows x 8 columns per warp) bank-conflict free, and keeps rows 16-byte aligned
__host__ __device__ inline int pad4mod16(int x) { return ((x + 11) / 16) * 16 + 4; }

__host__ __device__ inline int imin(int a, int b) { return a < b ? a : b; }

struct Layout {   // offsets in doubles from the start of the evaluation scratch (all even: 16-byte aligned)
  int S, R, npanels;
  int o_xn, o_kst, o_gf, o_alpha, o_part, o_red, o_bar, o_poff, o_ring;
  int max_panel;          // doubles of the largest panel
  long long all_panels;   // doubles of all panels together
};

__host__ __device__ inline Layout make_layout(int PT, int kmax, int n, int R) {
  Layout L;
  L.R = R;
  L.S = pad4mod16((n + 7) & ~7);   // whole 8-column tiles of the last group
  L.npanels = (n + R - 1) / R;
  int o = 0;
  L.o_xn = o; o += (PT * kmax + 1) & ~1;
  L.o_kst = o; o += PT * L.S;
  L.o_gf = o; o += PT * L.S;
  L.o_alpha = o; o += L.S;
  const int part_a = 2 * NW * 8 * 12, part_g = NT * PT;   // double-buffered split-K partials of t (NSPLIT R = 8 NW rows of
  L.o_part = o; o += (part_a > part_g ? part_a : part_g);  // 12) / partials of the gradient (never live together)
  L.o_red = o; o += NW * PT + NW * 8 + 2 * PT + 2 + ((NW * PT + 2 * PT) & 1);
  L.o_bar = o; o += 2 * NBAR + 2;   // full[NBAR], empty[NBAR], exit flag
  L.o_poff = o; o += NPOFF / 2;
  L.o_ring = o;
  const int G = R / 8, ng = ngroups(n);
  L.all_panels = packed_doubles(n);
  const int last_full = ((ng - 1) / G) * G;   // first group of the last panel
  long long mx = grp_off(ng) - grp_off(last_full);
  if (last_full >= G && grp_off(last_full) - grp_off(last_full - G) > mx) mx = grp_off(last_full) - grp_off(last_full - G);
  L.max_panel = (int)mx;
  return L;
}

// Producer-side state of the panel ring: a byte FIFO of variable-size panels (all offsets in doubles).
struct RingState {
  unsigned next, oldest;   // sequence numbers: next panel to request, oldest panel not yet released
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

struct Ctx {
  int n, kmax, kk, ld, acq;
  const double *ZnT, *Zn, *Wp, *alpha_g, *off, *coef;   // Wp: packed W of this run
  double inv_ell2, ymean, ystd, best_f, sgn;
  Layout L;
  double *base;        // shared-memory scratch
  int ring_cap;        // doubles
  bool resident;       // every panel stays in the ring for the whole launch
  bool loaded;         // resident panels have been waited for once
  unsigned cons;       // sequence number of the next panel to consume (all compute threads agree)
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
  c.Wp = m.Wp + (size_t)b * (size_t)m.wp_stride;
  c.alpha_g = m.alpha + (size_t)b * m.ld_n;
  c.off = m.norm_off + (size_t)b * m.kmax;
  c.coef = m.norm_coef + (size_t)b * m.kmax;
  { const double ell = m.lengthscales ? m.lengthscales[b] : m.lengthscale; c.inv_ell2 = 1.0 / (ell * ell); }
  c.ymean = m.ymean[b]; c.ystd = m.ystd[b]; c.best_f = m.best_f[b];
  c.sgn = m.maximize ? 1.0 : -1.0;
  c.prof = nullptr;
}

// ---- producer side of the panel ring: lane 0 of the producer warp runs this for the whole launch
__device__ inline void producer_loop(const Ctx &c) {
  const Layout &L = c.L;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(c.base + L.o_bar);
  unsigned long long *empty = full + NBAR;
  volatile int *exit_flag = reinterpret_cast<volatile int *>(c.base + L.o_bar + 2 * NBAR);
  int *poff = reinterpret_cast<int *>(c.base + L.o_poff);
  double *ring = c.base + L.o_ring;
  const int R = L.R, n = c.n, np = L.npanels;
  RingState r{0u, 0u, 0, 0, false};
  int I = 0;   // panel index of sequence number r.next (the sequence is cyclic over the panels)
  bool stop = false;
  while (!stop) {
    if (c.resident && r.next >= (unsigned)np) break;   // everything requested once: done
    const int g0 = I * (R / 8), g1 = imin(g0 + R / 8, ngroups(n));
    const long long src_off = grp_off(g0);
    const int sz = (int)(grp_off(g1) - src_off);
    int off = 0;
    // make room: hand-backs arrive in order, so the oldest live panel is the one to wait for
    while (r.next - r.oldest >= (unsigned)NBAR || !ring_place(r, sz, c.ring_cap, off)) {
      unsigned long long *eb = &empty[r.oldest & (NBAR - 1)];
      const unsigned par = (r.oldest >> 4) & 1u;
      while (!acp::mbar_try_wait(eb, par)) {
        if (*exit_flag) { stop = true; break; }
        __nanosleep(200);   // the optimiser warp shares this scheduler: do not spin on its issue slots
      }
      if (stop) break;
      ring_pop(r, poff[(r.oldest + 1) & (NPOFF - 1)]);
    }
    if (stop) break;
    poff[r.next & (NPOFF - 1)] = off;
    unsigned long long *bar = &full[r.next & (NBAR - 1)];
    acp::fence_proxy_async();
    acp::mbar_arrive_expect_tx(bar, (unsigned)sz * 8u);
    acp::bulk_g2s(ring + off, c.Wp + src_off, (unsigned)sz * 8u, bar);   // the whole panel: one contiguous copy
    ++r.next;
    if (++I == np) I = 0;
  }
  // no bulk copy may outlive the block: wait for everything that was requested and not handed back yet (resident
  // panels were all waited for by the first evaluation)
  if (!c.resident)
    for (unsigned s = r.oldest; s != r.next; ++s) acp::mbar_wait(&full[s & (NBAR - 1)], (s >> 4) & 1u);
}

// Called once per launch by ALL threads of the block (ends with __syncthreads): barriers, zero padding, alpha.
__device__ inline void init(Ctx &c, int PT) {
  const Layout &L = c.L;
  const int tid = threadIdx.x;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(c.base + L.o_bar);
  if (tid == 0) {
    for (int i = 0; i < NBAR; ++i) { acp::mbar_init(&bars[i], 1); acp::mbar_init(&bars[NBAR + i], NW); }
    *reinterpret_cast<int *>(c.base + L.o_bar + 2 * NBAR) = 0;
    acp::fence_barrier_init();
  }
  // kst / gf / alpha: entries >= n must be (and stay) zero: they meet the zero padding of the panels in the MMAs
  for (int i = tid; i < 2 * PT * L.S + L.S; i += NT_ALL) c.base[L.o_kst + i] = 0.0;
  __syncthreads();
  for (int i = tid; i < c.n; i += NT_ALL) c.base[L.o_alpha + i] = c.alpha_g[i];
  c.loaded = false;
  c.cons = 0;
  __syncthreads();
}

// The optimiser is done: tell the producer (thread 0 of the compute warps).
__device__ inline void signal_exit(const Ctx &c) {
  *reinterpret_cast<volatile int *>(c.base + c.L.o_bar + 2 * NBAR) = 1;
}

// Value and gradient at pcount <= PT points. X: raw coordinates, point p at X + p * kk. val[p] <- acquisition value,
// grad[p * kk + j] <- gscale * d value / d x_j. All NT compute threads must call; ends with their block barrier.
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
  unsigned long long *empty = bars + NBAR;
  const int *poff = reinterpret_cast<const int *>(c.base + L.o_poff);
  const double inv_ell2 = c.inv_ell2;

  ev_tick(c, -1);
  for (int idx = tid; idx < PT * kmax; idx += NT) {
    const int p = idx / kmax, j = idx % kmax;
    xn[idx] = (p < pcount && j < kk) ? (X[p * kk + j] - c.off[j]) / c.coef[j] : 0.0;
  }
  cta_sync();
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
  cta_sync();
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
    const unsigned seq = c.resident ? (unsigned)I : c.cons;
    if (!c.resident || !c.loaded) acp::mbar_wait(&bars[seq & (NBAR - 1)], (seq >> 4) & 1u);
    ev_tick(c, 27);   // (profiling builds only: wait for the panel)
    const double *pan = ring + poff[seq & (NPOFF - 1)];
    const int g0 = I * RT, ng = imin(RT, ngroups(n) - g0);    // groups of this panel (RT except at the end)
    // offsets of the panel's groups inside the panel and their row strides (32-bit; the groups are consecutive)
    int goff[RT], grs[RT];
    {
      int o = 0;
#pragma unroll
      for (int j = 0; j < RT; ++j) { goff[j] = o; grs[j] = grp_rs(g0 + j); o += 8 * grs[j]; }
    }
    const int r0 = I * R, c1 = imin(r0 + 8 * ng, n);
    double *pbuf = part + (I & 1) * (NSPLIT * R * PLD);
    {
      double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, c0 = 0.0, c1v = 0.0, d0 = 0.0, d1 = 0.0;
      if (wrt < ng) {   // this warp's row tile (= group g0 + wrt) exists
        const int ksteps = 2 * (g0 + wrt) + 2;                 // its columns: 0 .. 8 (g0 + wrt) + 7
        int my_off = 0, my_rs = 0;
#pragma unroll
        for (int j = 0; j < RT; ++j) if (j == wrt) { my_off = goff[j]; my_rs = grs[j]; }
        const double *pa = pan + my_off + lr * my_rs + lc + 4 * wsp;
        // (the 8 - PT padding columns of the B operand repeat real points: whatever they produce is never read)
        const double *pb = kst + (lr < PT ? lr : lr - PT) * S + lc + 4 * wsp;
        int left = (ksteps - wsp + NSPLIT - 1) / NSPLIT;       // k steps wsp, wsp + NSPLIT, ... of this warp
        constexpr int ST = 4 * NSPLIT;                         // doubles between them
        for (; left >= 4; left -= 4, pa += 4 * ST, pb += 4 * ST) {
          const double w0 = pa[0], w1 = pa[ST], w2 = pa[2 * ST], w3 = pa[3 * ST];
          const double k0 = pb[0], k1 = pb[ST], k2 = pb[2 * ST], k3 = pb[3 * ST];
          acp::dmma_8x8x4(a0, a1, w0, k0);
          acp::dmma_8x8x4(b0, b1, w1, k1);
          acp::dmma_8x8x4(c0, c1v, w2, k2);
          acp::dmma_8x8x4(d0, d1, w3, k3);
        }
        if (left >= 2) {
          const double w0 = pa[0], w1 = pa[ST], k0 = pb[0], k1 = pb[ST];
          acp::dmma_8x8x4(a0, a1, w0, k0);
          acp::dmma_8x8x4(b0, b1, w1, k1);
          pa += 2 * ST; pb += 2 * ST; left -= 2;
        }
        if (left >= 1) acp::dmma_8x8x4(c0, c1v, pa[0], pb[0]);
      }
      *reinterpret_cast<double2 *>(pbuf + (wsp * R + wrt * 8 + lr) * PLD + 2 * lc) =
          make_double2((a0 + b0) + (c0 + d0), (a1 + b1) + (c1v + d1));
    }
    ev_tick(c, 28);   // phase A
    cta_sync();
    ev_tick(c, 29);   // barrier
    {
      double bf[KQ];
#pragma unroll
      for (int kq = 0; kq < KQ; ++kq) {
        const int r = 4 * kq + lc;
        double tp[NSPLIT];
#pragma unroll
        for (int sp = 0; sp < NSPLIT; ++sp) tp[sp] = pbuf[(sp * R + r) * PLD + lr];
#pragma unroll
        for (int st = 1; st < NSPLIT; st *= 2)      // pairwise tree: log2(NSPLIT) dependent additions
#pragma unroll
          for (int sp = 0; sp + st < NSPLIT; sp += 2 * st) tp[sp] += tp[sp + st];
        bf[kq] = tp[0];   // (rows >= n are zero in the packed panel: t is exactly 0 there)
      }
      if (warp == (I & (NW - 1))) {
#pragma unroll
        for (int kq = 0; kq < KQ; ++kq) vacc = fma(bf[kq], bf[kq], vacc);
      }
      // A fragments of the transposed product: rows 4 kq + (lane & 3) of the panel live in group kq / 2
      // (a missing group at the end of the matrix has t = 0: the last existing group's address will do)
      int aoff[KQ];
#pragma unroll
      for (int kq = 0; kq < KQ; ++kq) {
        const int gj = (kq >> 1) < ng ? (kq >> 1) : ng - 1;
        int o = 0, r_ = 0;
#pragma unroll
        for (int j = 0; j < RT; ++j) if (j == gj) { o = goff[j]; r_ = grs[j]; }
        aoff[kq] = o + ((4 * kq & 7) + lc) * r_ + lr;
      }
      const int ntile = (c1 + 7) >> 3;
      // tile outer, k step inner: the tiles are independent accumulator chains, and a tile right of the diagonal block
      // is skipped by a uniform branch (predicated-off DMMAs measured slower than the chains they were meant to hide)
#pragma unroll
      for (int ti = 0; ti < TI; ti += 2) {
        const int ct = warp + NW * ti, ct2 = ct + NW;
        if (ti + 1 < TI && ct2 <= g0) {
          // two tiles left of every group's diagonal block: all k steps, operands loaded up front, the two accumulator
          // chains interleaved (a DMMA that depends on the previous one waits ~30 cycles)
          double av[KQ], aw[KQ];
#pragma unroll
          for (int kq = 0; kq < KQ; ++kq) { av[kq] = pan[aoff[kq] + ct * 8]; aw[kq] = pan[aoff[kq] + ct2 * 8]; }
#pragma unroll
          for (int kq = 0; kq < KQ; ++kq) {
            acp::dmma_8x8x4(sf[ti][0], sf[ti][1], av[kq], bf[kq]);
            acp::dmma_8x8x4(sf[ti + 1][0], sf[ti + 1][1], aw[kq], bf[kq]);
          }
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (ti + h < TI) {
              const int cth = ct + NW * h;
              if (cth <= g0) {          // left of every group's diagonal block: all k steps
                double av[KQ];
#pragma unroll
                for (int kq = 0; kq < KQ; ++kq) av[kq] = pan[aoff[kq] + cth * 8];
#pragma unroll
                for (int kq = 0; kq < KQ; ++kq) acp::dmma_8x8x4(sf[ti + h][0], sf[ti + h][1], av[kq], bf[kq]);
              } else if (cth < ntile) {  // a diagonal tile of the panel: only the groups at or below it hold something there
#pragma unroll
                for (int kq = 0; kq < KQ; ++kq)
                  if (cth <= g0 + (kq >> 1)) acp::dmma_8x8x4(sf[ti + h][0], sf[ti + h][1], pan[aoff[kq] + cth * 8], bf[kq]);
              }
            }
          }
        }
      }
    }
    ev_tick(c, 30);   // partial sums + phase B
    ++c.cons;
    if (!c.resident) {   // this warp is done with the panel
      __syncwarp();
      if (lane == 0) acp::mbar_arrive(&empty[seq & (NBAR - 1)]);
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
  // ---- the gradient is linear in the two head results (d value / d mu_s, d value / d var_s):
  //          grad_j = A sum_i gf_i alpha_i (xn_j - zn_ij) - 2 B sum_i gf_i s_i (xn_j - zn_ij),
  //      so its two sums do not wait for the head (erfc / erfcx / log chains of five lanes, ~4 k cycles): the factors
  //      gf alpha and gf s go over the k vector and gf straight from the accumulators, then warp 0 evaluates the head WHILE
  //      warps 1..7 accumulate the two sums; they meet at the barrier before the partials are combined.
#pragma unroll
  for (int ti = 0; ti < TI; ++ti) {
    const int j = (warp + NW * ti) * 8 + lr;
    if (j < n) {
      const double al = alpha[j];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = 2 * lc + h;
        if (p < PT) {
          const double g0 = gf[p * S + j];
          kst[p * S + j] = g0 * al;
          gf[p * S + j] = g0 * sf[ti][h];
        }
      }
    }
  }
  cta_sync();
  ev_tick(c, 17);
  double *res = red + NW * PT + NW * 8;   // [2][PT]: d value / d mu_s, d value / d var_s
  constexpr int NTG = NT - 32;            // threads of the gradient sums (warps 1..7)
  const int IG = NTG / kk;                // >= 2 for kk <= 104
  const int gt = tid - 32, ig = gt / kk, j = gt % kk;
  double acc1[PT], acc2[PT];
#pragma unroll
  for (int p = 0; p < PT; ++p) { acc1[p] = 0.0; acc2[p] = 0.0; }
  if (warp == 0) {
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
  } else if (ig < IG) {
    // thread (ig, j) of warps 1..7: every IG-th training point
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
        for (int p = 0; p < PT; ++p) {
          const double dz = xj[p] - z8[u];
          acc1[p] = fma(kst[p * S + i + u * IG], dz, acc1[p]);
          acc2[p] = fma(gf[p * S + i + u * IG], dz, acc2[p]);
        }
    }
    for (; i < n; i += IG) {
      const double z = c.Zn[(size_t)i * kmax + j];
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        const double dz = xj[p] - z;
        acc1[p] = fma(kst[p * S + i], dz, acc1[p]);
        acc2[p] = fma(gf[p * S + i], dz, acc2[p]);
      }
    }
  }
  cta_sync();
  ev_tick(c, 22);
  // ---- combine with the head results: partial of grad[p][j] over this thread's training points
  if (warp != 0 && ig < IG) {
#pragma unroll
    for (int p = 0; p < PT; ++p) part[(ig * PT + p) * kk + j] = res[p] * acc1[p] - 2.0 * res[PT + p] * acc2[p];
  }
  cta_sync();
  ev_tick(c, 19);
  for (int idx = tid; idx < pcount * kk; idx += NT) {
    const int p = idx / kk, jj = idx % kk;
    double sv = 0.0;
    for (int q = 0; q < IG; ++q) sv += part[(q * PT + p) * kk + jj];
    grad[idx] = gscale * sv / c.coef[jj];
  }
  cta_sync();
  ev_tick(c, 20);
}

}  // namespace lbeval
