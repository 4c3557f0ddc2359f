// L-BFGS-B (Byrd, Lu, Nocedal, Zhu 1995; subspace step of Morales & Nocedal 2011; More-Thuente line search) written as
// team-parallel code: the same source runs on one host thread (HostTeam, used by the CPU validation harness against
// scipy) and on one CUDA thread block (CudaTeam, one block per restart problem, all state in shared memory).
//
// It follows the published algorithm in the structure of L-BFGS-B 3.0 that scipy's `optimize.minimize(method=
// "L-BFGS-B")` executes for botorch's gen_candidates_scipy inside the reference's optimize_acqf call
// (PCA_BO.py:607-614): generalized Cauchy point -> free-variable subspace minimisation with the LEL^T factorisation
// of the 2m x 2m indefinite matrix -> projected Newton step -> dcsrch line search (ftol 1e-3, gtol 0.9, xtol 0.1) ->
// BFGS pair update with the curvature test -> tests pgtol (1e-5) and factr * epsmch (scipy ftol 2.22e-9).
// All variables are box constrained (nbd = 2), which is the only case optimize_acqf produces.
//
// Control flow is uniform across the team: every thread keeps the scalar state in registers (computed redundantly
// from team-reduced values); only arrays live in (shared) memory.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LB_HD __host__ __device__
// (the big phases stay inlined: as separate functions they take the Work / Ctx structs by reference through local
// memory, which measured 50 % slower on the B200)
#define LB_PHASE __host__ __device__
#else
#define LB_HD
#define LB_PHASE
#endif

#if defined(__CUDACC__) || defined(__GNUC__)
#define LB_RESTRICT __restrict__
#else
#define LB_RESTRICT
#endif

namespace lbfgsb {

constexpr int M = 10;       // history pairs (scipy maxcor)
constexpr int M2 = 2 * M;
constexpr double EPSMCH = 2.220446049250313e-16;
constexpr double FTOL_LS = 1e-3, GTOL_LS = 0.9, XTOL_LS = 0.1, STPMIN = 0.0, BIG = 1e10;
constexpr int MAXLS = 20;

enum Status { CONV_PGTOL = 0, CONV_FTOL = 1, STOP_MAXITER = 2, ABNORMAL = 3, RUNNING = 4 };

struct HostTeam {
  static constexpr int tid = 0, nt = 1, lane = 0, nl = 1;
  static constexpr bool w0 = true;
  static constexpr bool kWarpRegs = false;   // true: the team is one CUDA warp with register/shuffle small-matrix kernels
  // true: a helper team computes cmprlb and the first products of subsm WHILE this team factors K in formk (they do not
  // depend on the factor); the team then only joins the helper (join_helper) instead of running them itself
  static constexpr bool kAsyncHelper = false;
  void join_helper() {}
  void start_cauchy_pass1(const struct Work &, int, double) {}
  void sync() {}
  void wsync() {}
  double sum(double v) { return v; }
  double max(double v) { return v; }
  const double *sum_n(const double *v) { return v; }   // sums over the team of M2 values per thread
  void argmin(double &, int &) {}
  inline void formk_products(const struct Work &w, int col, double theta);
  inline void shift_pairs(const struct Work &w);
  void tick(int) {}
};

struct Work {
  int n;
  int ldp;                                       // row stride of ws / wy (pair_stride(n))
  double *x, *l, *u, *g, *z, *d, *t, *r, *xp;  // [n]
  double *ws, *wy;                               // [M][n]  stored pairs, oldest first
  double *sy, *ss, *wt, *jinv;                   // [M][M]
  double *wn;                                    // [M2][M2]
  double *p, *c, *wbp, *v, *wv;                  // [M2]
  double *tbrk;                                  // [n] breakpoints (negative: none / consumed)
  double *pre;                                   // [2] f1 and the breakpoint count of a Cauchy first pass computed ahead
  int *iwhere;                                   // [n]
  // reciprocals kept next to the factors (warp-register path): 1/sy_kk, 1/sqrt(sy_kk), 1/diag(wt), 1/diag(wn)
  double *rsy, *rsq, *wt_rd, *wn_rd;
};

// Row stride of the stored pairs: a multiple of 8 congruent 8 mod 16, so that consecutive rows start in opposite halves of
// the shared-memory banks (the block-wide products of formk read 8 consecutive entries of 4 rows per warp; with the
// stride n = 5 k a multiple of 16 that was a 4-way bank conflict: 12 k cycles per formk).
LB_HD inline int pair_stride(int n) { return ((n + 15) & ~15) + 8; }

// pairs_external: the BFGS pairs ws / wy (2 M pair_stride(n) doubles) live in a buffer of the caller's instead of the work area
LB_HD inline size_t work_doubles(int n, bool pairs_external = false) {
  return (size_t)9 * n + (pairs_external ? 0 : (size_t)2 * M * pair_stride(n)) + 4 * M * M + M2 * M2 + 5 * M2 + n + (n + 1) / 2 + 4;
}

// EXT (compile time): the pairs live in `pairs` (global memory) instead of the work area. A run-time choice would hide
// the address space of ws / wy from the compiler, which then issues generic loads (LD.E) for them everywhere instead of
// shared-memory loads (LDS): measured 3-4x slower phases.
template <bool EXT = false>
LB_HD inline void carve(Work &w, int n, double *base, double *pairs = nullptr) {
  w.n = n;
  double *q = base;
  w.x = q; q += n; w.l = q; q += n; w.u = q; q += n; w.g = q; q += n; w.z = q; q += n;
  w.d = q; q += n; w.t = q; q += n; w.r = q; q += n; w.xp = q; q += n;
  w.ldp = pair_stride(n);
  if (EXT) { w.ws = pairs; w.wy = pairs + M * w.ldp; }
  else { w.ws = q; q += M * w.ldp; w.wy = q; q += M * w.ldp; }
  w.sy = q; q += M * M; w.ss = q; q += M * M; w.wt = q; q += M * M; w.jinv = q; q += M * M;
  w.rsy = w.jinv; w.rsq = w.jinv + M; w.wt_rd = w.jinv + 2 * M; w.wn_rd = w.jinv + 3 * M;
  w.wn = q; q += M2 * M2;
  w.p = q; q += M2; w.c = q; q += M2; w.wbp = q; q += M2; w.v = q; q += M2; w.wv = q; q += M2;
  w.tbrk = q; q += n;
  w.pre = q; q += 2;     // (two of the four spare doubles of work_doubles)
  w.iwhere = reinterpret_cast<int *>(q);
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------ warp-register kernels
// Small dense kernels for a team that is exactly one CUDA warp: matrices live in registers (lane j owns column j or row
// j), pivots and right-hand sides travel by shuffles, divisions are replaced by stored reciprocal diagonals. The serial
// chain per elimination step is one shuffle + one multiply + one FMA instead of shared-memory round trips.
namespace dev {
constexpr unsigned FULL = 0xffffffffu;
__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL, v, src); }

// Upper triangle of a 20 x 20 matrix enumerated column by column: entry e -> i | (j << 8), i <= j (0xffff: padding).
__device__ const unsigned short kTri20[224] = {
    0, 256, 257, 512, 513, 514, 768, 769, 770, 771, 1024, 1025, 1026, 1027, 1028, 1280,
    1281, 1282, 1283, 1284, 1285, 1536, 1537, 1538, 1539, 1540, 1541, 1542, 1792, 1793, 1794, 1795,
    1796, 1797, 1798, 1799, 2048, 2049, 2050, 2051, 2052, 2053, 2054, 2055, 2056, 2304, 2305, 2306,
    2307, 2308, 2309, 2310, 2311, 2312, 2313, 2560, 2561, 2562, 2563, 2564, 2565, 2566, 2567, 2568,
    2569, 2570, 2816, 2817, 2818, 2819, 2820, 2821, 2822, 2823, 2824, 2825, 2826, 2827, 3072, 3073,
    3074, 3075, 3076, 3077, 3078, 3079, 3080, 3081, 3082, 3083, 3084, 3328, 3329, 3330, 3331, 3332,
    3333, 3334, 3335, 3336, 3337, 3338, 3339, 3340, 3341, 3584, 3585, 3586, 3587, 3588, 3589, 3590,
    3591, 3592, 3593, 3594, 3595, 3596, 3597, 3598, 3840, 3841, 3842, 3843, 3844, 3845, 3846, 3847,
    3848, 3849, 3850, 3851, 3852, 3853, 3854, 3855, 4096, 4097, 4098, 4099, 4100, 4101, 4102, 4103,
    4104, 4105, 4106, 4107, 4108, 4109, 4110, 4111, 4112, 4352, 4353, 4354, 4355, 4356, 4357, 4358,
    4359, 4360, 4361, 4362, 4363, 4364, 4365, 4366, 4367, 4368, 4369, 4608, 4609, 4610, 4611, 4612,
    4613, 4614, 4615, 4616, 4617, 4618, 4619, 4620, 4621, 4622, 4623, 4624, 4625, 4626, 4864, 4865,
    4866, 4867, 4868, 4869, 4870, 4871, 4872, 4873, 4874, 4875, 4876, 4877, 4878, 4879, 4880, 4881,
    4882, 4883, 65535, 65535, 65535, 65535, 65535, 65535, 65535, 65535, 65535, 65535, 65535, 65535, 65535, 65535,
};

// Right-looking Cholesky R^T R = A of the leading n x n block (n <= 20 = 2 m) of the upper-stored A (ld = lda) in shared
// memory. What is left in A is the UNIT triangular factor U = D^-1 R (U_ij = R_ij / R_ii for j > i; the diagonal is not
// used) together with rdiag[i] = 1 / R_ii, so that the triangular solves below have no multiplication on their serial
// chain. For pivots k < nsplit the rows/columns >= nsplit receive +a_ki a_kj / a_kk instead of -a_ki a_kj / a_kk: with
// nsplit = col on the 2 col x 2 col matrix of formk this is the LEL^T factorisation (Cholesky of the (1,1) block,
// R12 = R11^-T A12, Cholesky of A22 + R12^T R12) in one sweep.
// The 210 entries of the triangle are dealt flat to the 32 lanes (seven each, fixed for the whole factorisation). A
// pivot costs: one reciprocal, then for all seven entries at once (branch free, so their FP64 latencies of 23 cycles
// overlap) a_ij -= a_ki a_kj / a_kk from the UNSCALED pivot row, and one warp barrier. (Round 1 gave lane j the whole
// column j and published the scaled pivot row through three warp barriers per pivot: 18.8 k cycles for 20 x 20; this
// one takes 14.5 k in isolation, tools/micro/chol_warp.cu. Measured and rejected in round 2: the matrix in registers with
// shuffle-broadcast multipliers, fully unrolled (24.7 k), and computing the next pivot's reciprocal ahead of the updates
// (18.7 k).)
// Returns 0 or 1 + index of the first non-positive pivot.
__device__ inline int chol_upper_warp(double *A, int lda, int n, int nsplit, double *rdiag) {
  const int lane = threadIdx.x & 31;
  constexpr int EPL = 7;
  int off[EPL], oi[EPL], oj[EPL], row[EPL];
  bool sgn[EPL];
#pragma unroll
  for (int q = 0; q < EPL; ++q) {
    const unsigned v = kTri20[lane + 32 * q];
    const int i = v & 0xff, j = v >> 8;
    const bool on = j < n;              // (padding has j = 255)
    row[q] = on ? i : -1;               // -1: never updated
    off[q] = on ? i * lda + j : 0;      // a valid address either way: loads are unconditional
    oi[q] = on ? i : 0;
    oj[q] = on ? j : 0;
    sgn[q] = on && i >= nsplit;
  }
  int bad = 0;
  for (int k = 0; k < n; ++k) {
    const double *__restrict__ rk = A + k * lda;
    double piv = rk[k];
    double a[EPL], ri[EPL], rj[EPL];
#pragma unroll
    for (int q = 0; q < EPL; ++q) { a[q] = A[off[q]]; ri[q] = rk[oi[q]]; rj[q] = rk[oj[q]]; }
    if (!(piv > 0.0)) { if (bad == 0) bad = k + 1; piv = 1.0; }
    const double pinv = 1.0 / piv;
    if (lane == 0) rdiag[k] = piv;      // turned into 1 / sqrt(piv) after the sweep
    const bool plus_k = k < nsplit;
#pragma unroll
    for (int q = 0; q < EPL; ++q) {
      const double prod = (ri[q] * pinv) * rj[q];
      const double v = (plus_k && sgn[q]) ? a[q] + prod : a[q] - prod;
      if (row[q] > k) A[off[q]] = v;
    }
    __syncwarp();
  }
  // unit factor: U_ij = a_ij / a_ii (j > i)
  const double pv = (lane < n) ? rdiag[lane] : 1.0;
  __syncwarp();
#pragma unroll
  for (int q = 0; q < EPL; ++q) {
    if (row[q] >= 0 && oj[q] > oi[q]) A[off[q]] = A[off[q]] / rdiag[oi[q]];
  }
  __syncwarp();
  if (lane < n) rdiag[lane] = rsqrt(pv);
  __syncwarp();
  return bad;
}

// In place on b (shared memory, length n <= 32): trans ? R^T x = b : R x = b with R = D U from chol_upper_warp (U unit
// upper triangular in shared memory, rdiag = 1 / diag(R)). The right-hand side lives in one register per lane and
// travels by shuffles; the serial chain per step is one shuffle and one FMA (the diagonal scaling is done for all lanes
// at once, before or after).
__device__ inline void tri_solve_warp(const double *R, int ldr, int n, double *b, const double *rdiag, bool trans) {
  const int lane = threadIdx.x & 31;
  const bool on = lane < n;
  double bi = on ? b[lane] : 0.0;
  const double rd = on ? rdiag[lane] : 0.0;
  // (the factor entry of step j + 1 is loaded before the shuffle of step j: its latency stays off the serial chain)
  if (trans) {   // U^T y = b, x = D^-1 y
    double u = (lane > 0 && on) ? R[lane] : 0.0;
    for (int j = 0; j < n; ++j) {
      const double un = (lane > j + 1 && on && j + 1 < n) ? R[(j + 1) * ldr + lane] : 0.0;
      const double yj = shfl(bi, j);
      bi = fma(-u, yj, bi);
      u = un;
    }
    bi *= rd;
  } else {       // U x = D^-1 b
    bi *= rd;
    double u = (lane < n - 1) ? R[lane * ldr + n - 1] : 0.0;
    for (int j = n - 1; j >= 0; --j) {
      const double un = (j >= 1 && lane < j - 1) ? R[lane * ldr + j - 1] : 0.0;
      const double xj = shfl(bi, j);
      bi = fma(-u, xj, bi);
      u = un;
    }
  }
  if (on) b[lane] = bi;
  __syncwarp();
}
}  // namespace dev
#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------ small dense helpers
// Cholesky of the leading n x n block of an upper-stored symmetric matrix A (ld = lda): R^T R = A, R overwrites the
// upper triangle. Lane-parallel right-looking sweep executed by warp 0. Returns (uniformly, via smem flag) 0 / k+1.
template <class T>
LB_HD int chol_upper(T &tm, double *A, int lda, int n, int *flag) {
  if (tm.w0) {
    if (tm.lane == 0) *flag = 0;
    tm.wsync();
    for (int k = 0; k < n; ++k) {
      double akk = A[k * lda + k];
      if (!(akk > 0.0)) {
        if (tm.lane == 0 && *flag == 0) *flag = k + 1;
        akk = 1.0;
      }
      const double rkk = sqrt(akk);
      tm.wsync();
      if (tm.lane == 0) A[k * lda + k] = rkk;
      for (int j = k + 1 + tm.lane; j < n; j += tm.nl) A[k * lda + j] /= rkk;
      tm.wsync();
      for (int j = k + 1 + tm.lane; j < n; j += tm.nl) {
        const double rkj = A[k * lda + j];
        for (int i = k + 1; i <= j; ++i) A[i * lda + j] -= A[k * lda + i] * rkj;
      }
      tm.wsync();
    }
  }
  tm.sync();
  return *flag;
}

// Solve with the upper-triangular R (ld = ldr, order n) in place on b: trans ? R^T x = b : R x = b. Warp 0 only
// (callers synchronise the team afterwards).
template <class T>
LB_HD void tri_solve_w0(T &tm, const double *R, int ldr, int n, double *b, bool trans) {
  if (!tm.w0) return;
  if (trans) {
    for (int j = 0; j < n; ++j) {
      const double xj = b[j] / R[j * ldr + j];
      tm.wsync();
      if (tm.lane == 0) b[j] = xj;
      for (int i = j + 1 + tm.lane; i < n; i += tm.nl) b[i] -= R[j * ldr + i] * xj;
      tm.wsync();
    }
  } else {
    for (int j = n - 1; j >= 0; --j) {
      const double xj = b[j] / R[j * ldr + j];
      tm.wsync();
      if (tm.lane == 0) b[j] = xj;
      for (int i = tm.lane; i < j; i += tm.nl) b[i] -= R[i * ldr + j] * xj;
      tm.wsync();
    }
  }
}

// p = M v with the compact middle matrix (bmv of the reference implementation): v, p of length 2 col.
template <class T>
LB_HD void bmv(T &tm, const Work &w, int col, const double *v, double *p) {
  if (col == 0) return;
#ifdef __CUDA_ARCH__
  if constexpr (T::kWarpRegs) {
    const int i = tm.lane;
    const bool on = i < col;
    const double v1 = on ? v[i] : 0.0, v2 = on ? v[col + i] : 0.0;
    const double rsyi = on ? w.rsy[i] : 0.0, rsqi = on ? w.rsq[i] : 0.0;
    const double u = v1 * rsyi;
    double p2 = v2;
#pragma unroll
    for (int k = 0; k < M; ++k) {
      const double uk = dev::shfl(u, k);
      if (on && k < i) p2 = fma(w.sy[i * M + k], uk, p2);
    }
    if (on) p[col + i] = p2;
    __syncwarp();
    dev::tri_solve_warp(w.wt, M, col, p + col, w.wt_rd, true);
    dev::tri_solve_warp(w.wt, M, col, p + col, w.wt_rd, false);
    p2 = on ? p[col + i] : 0.0;
    double sacc = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) {
      const double pk = dev::shfl(p2, k);
      if (on && k > i && k < col) sacc = fma(w.sy[k * M + i], pk, sacc);
    }
    if (on) p[i] = -(v1 * rsqi) * rsqi + sacc * rsyi;
    __syncwarp();
    return;
  }
#endif
  if (tm.w0) {
    // p2 = v2 + L D^-1 v1
    for (int i = tm.lane; i < col; i += tm.nl) {
      double s = 0.0;
      for (int k = 0; k < i; ++k) s += w.sy[i * M + k] * v[k] / w.sy[k * M + k];
      p[col + i] = v[col + i] + s;
    }
    tm.wsync();
    tri_solve_w0(tm, w.wt, M, col, p + col, true);   // J p2 = ., J = wt^T
    for (int i = tm.lane; i < col; i += tm.nl) p[i] = v[i] / sqrt(w.sy[i * M + i]);
    tm.wsync();
    tri_solve_w0(tm, w.wt, M, col, p + col, false);  // J^T p2 = p2
    for (int i = tm.lane; i < col; i += tm.nl) {
      double s = 0.0;
      for (int k = i + 1; k < col; ++k) s += w.sy[k * M + i] * p[col + k] / w.sy[i * M + i];
      p[i] = -p[i] / sqrt(w.sy[i * M + i]) + s;
    }
    tm.wsync();
  }
  tm.sync();
}

template <class T>
LB_HD double dot(T &tm, const double *a, const double *b, int n) {
  double s = 0.0;
  for (int i = tm.tid; i < n; i += tm.nt) s += a[i] * b[i];
  return tm.sum(s);
}

template <class T>
LB_HD double projgr(T &tm, const Work &w) {
  double sb = 0.0;
  for (int i = tm.tid; i < w.n; i += tm.nt) {
    double gi = w.g[i];
    if (gi < 0.0) gi = fmax(w.x[i] - w.u[i], gi);
    else gi = fmin(w.x[i] - w.l[i], gi);
    sb = fmax(sb, fabs(gi));
  }
  return tm.max(sb);
}

// ------------------------------------------------------------------------------------------------ Cauchy point
// First pass of the generalized Cauchy point: variable classification, breakpoints, d = -g on the free variables, z = x,
// p = W'd and c = 0. Needs the stored pairs but NOT the factor of T, so on the device a helper warp runs it for the NEXT
// iteration while the optimiser warp forms and factors T (update_pairs). f1 = -d'd and the number of breakpoints are
// returned and left in w.pre.
template <class T>
LB_PHASE void cauchy_pass1(T &tm, Work &w, int col, double theta, double &f1_out, int &nbreak_out) {
  const int n = w.n;
  double f1 = 0.0;
  double part[M2];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int j = 0; j < M2; ++j) part[j] = 0.0;
  double nbrk_local = 0.0;
  {
    const double *LB_RESTRICT gx = w.g, *LB_RESTRICT xx = w.x, *LB_RESTRICT lx = w.l, *LB_RESTRICT ux = w.u;
    const double *LB_RESTRICT wyx = w.wy, *LB_RESTRICT wsx = w.ws;
    double *LB_RESTRICT dx = w.d, *LB_RESTRICT tb_ = w.tbrk, *LB_RESTRICT zx = w.z;
    int *LB_RESTRICT iwx = w.iwhere;
    for (int i = tm.tid; i < n; i += tm.nt) {
      const double neggi = -gx[i];
      const double xi = xx[i];
      const double tl = xi - lx[i], tu = ux[i] - xi;
      int iw = iwx[i];
      if (iw != 3 && iw != -1) {
        iw = 0;
        if (tl <= 0.0) { if (neggi <= 0.0) iw = 1; }
        else if (tu <= 0.0) { if (neggi >= 0.0) iw = 2; }
        else if (fabs(neggi) <= 0.0) iw = -3;
        iwx[i] = iw;
      }
      double tb = -1.0, di = 0.0;
      if (iw == 0 || iw == -1) {
        di = neggi;
        f1 -= neggi * neggi;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int j = 0; j < M; ++j) {   // (rows >= col hold stale but finite pairs: their sums are never read)
          part[j] += wyx[j * w.ldp + i] * neggi;
          part[M + j] += wsx[j * w.ldp + i] * neggi;
        }
        if (neggi < 0.0) tb = tl / (-neggi);
        else if (neggi > 0.0) tb = tu / neggi;
        if (tb >= 0.0) nbrk_local += 1.0;
      }
      dx[i] = di;
      tb_[i] = tb;
      zx[i] = xi;
    }
  }
  f1 = tm.sum(f1);
  const int nbreak = (int)(tm.sum(nbrk_local) + 0.5);
  {
    const double *S = tm.sum_n(part);
    if (tm.tid == 0) {
      for (int j = 0; j < col; ++j) { w.p[j] = S[j]; w.p[col + j] = theta * S[M + j]; }
      for (int j = 0; j < 2 * col; ++j) w.c[j] = 0.0;
      w.pre[0] = f1;
      w.pre[1] = (double)nbreak;
    }
  }
  tm.sync();
  f1_out = f1;
  nbreak_out = nbreak;
}

// pass1_ready: cauchy_pass1 has been run for this iterate already (w.pre holds its two scalars)
template <class T>
LB_PHASE void cauchy(T &tm, Work &w, int col, double theta, double sbgnrm, int &nseg, bool pass1_ready = false) {
  const int n = w.n;
  nseg = 0;
  if (sbgnrm <= 0.0) {
    for (int i = tm.tid; i < n; i += tm.nt) w.z[i] = w.x[i];
    tm.sync();
    return;
  }
  double f1 = 0.0;
  int nbreak = 0;
  if (pass1_ready) { f1 = w.pre[0]; nbreak = (int)(w.pre[1] + 0.5); }
  else cauchy_pass1(tm, w, col, theta, f1, nbreak);
  tm.tick(11);
  if (nbreak == 0) return;  // no variable can move: xcp = x (all boxed, so "nfree == n + 1" of the original holds)
  double f2 = -theta * f1;
  const double f2_org = f2;
  if (col > 0) {
    bmv(tm, w, col, w.p, w.v);
    double s = 0.0;
    for (int j = 0; j < 2 * col; ++j) s += w.v[j] * w.p[j];
    f2 -= s;
  }
  tm.tick(12);
  double dtm = -f1 / f2;
  double tsum = 0.0;
  nseg = 1;
  int nleft = nbreak;
  double tj = 0.0;
  bool all_fixed = false;
  for (int guard = 0; guard <= nbreak; ++guard) {
    const double tj0 = tj;
    double tmin = INFINITY;
    int ibp = 0x7fffffff;
    for (int i = tm.tid; i < n; i += tm.nt) {
      const double tb = w.tbrk[i];
      if (tb >= 0.0 && (tb < tmin || (tb == tmin && i < ibp))) { tmin = tb; ibp = i; }
    }
    tm.argmin(tmin, ibp);
    tj = tmin;
    const double dt = tj - tj0;
    if (dtm < dt) break;
    tsum += dt;
    --nleft;
    const double dibp = w.d[ibp];
    const double xi = w.x[ibp];
    double zibp;
    tm.sync();
    if (tm.tid == 0) {
      w.d[ibp] = 0.0;
      w.tbrk[ibp] = -1.0;
      if (dibp > 0.0) { w.z[ibp] = w.u[ibp]; w.iwhere[ibp] = 2; }
      else { w.z[ibp] = w.l[ibp]; w.iwhere[ibp] = 1; }
    }
    zibp = (dibp > 0.0) ? (w.u[ibp] - xi) : (w.l[ibp] - xi);
    tm.sync();
    if (nleft == 0 && nbreak == n) { dtm = dt; all_fixed = true; break; }
    ++nseg;
    const double dibp2 = dibp * dibp;
    f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp;
    f2 = f2 - theta * dibp2;
    if (col > 0) {
      for (int j = tm.tid; j < 2 * col; j += tm.nt) {
        w.c[j] += dt * w.p[j];
        w.wbp[j] = (j < col) ? w.wy[j * w.ldp + ibp] : theta * w.ws[(j - col) * w.ldp + ibp];
      }
      tm.sync();
      bmv(tm, w, col, w.wbp, w.v);
      double wmc = 0.0, wmp = 0.0, wmw = 0.0;
      for (int j = 0; j < 2 * col; ++j) { wmc += w.c[j] * w.v[j]; wmp += w.p[j] * w.v[j]; wmw += w.wbp[j] * w.v[j]; }
      tm.sync();
      for (int j = tm.tid; j < 2 * col; j += tm.nt) w.p[j] -= dibp * w.wbp[j];
      tm.sync();
      f1 += dibp * wmc;
      f2 += 2.0 * dibp * wmp - dibp2 * wmw;
    }
    f2 = fmax(EPSMCH * f2_org, f2);
    if (nleft > 0) { dtm = -f1 / f2; continue; }
    f1 = 0.0; f2 = 0.0; dtm = 0.0;   // all variables are bounded
    break;
  }
  if (!all_fixed) {
    if (dtm <= 0.0) dtm = 0.0;
    tsum += dtm;
    for (int i = tm.tid; i < n; i += tm.nt) w.z[i] += tsum * w.d[i];
  }
  if (col > 0)
    for (int j = tm.tid; j < 2 * col; j += tm.nt) w.c[j] += dtm * w.p[j];
  tm.sync();
}

// ------------------------------------------------------------------------------------------------ subspace minimisation
// Raw inner products of formk split over free (Z, iwhere <= 0) and active (A) variables; entries e = tid, tid + nt, ...
// wn upper triangle layout (2 col x 2 col, ld M2):
//   (1,1): D + Y'ZZ'Y / theta ; (1,2): -L_a' + R_z' ; (2,2): theta S'AA'S
LB_HD inline void formk_products_range(const Work &w, int col, double theta, int tid, int nt) {
  const int n = w.n;
  for (int e = tid; e < 3 * col * col; e += nt) {
    const int blk = e / (col * col), rem = e % (col * col);
    const int i = rem / col, j = rem % col;
    if (blk == 0) {          // (1,1), i <= j
      if (i > j) continue;
      double s = 0.0;
      for (int k = 0; k < n; ++k) if (w.iwhere[k] <= 0) s += w.wy[i * w.ldp + k] * w.wy[j * w.ldp + k];
      w.wn[i * M2 + j] = s / theta + ((i == j) ? w.sy[i * M + i] : 0.0);
    } else if (blk == 1) {   // (2,2), i <= j
      if (i > j) continue;
      double s = 0.0;
      for (int k = 0; k < n; ++k) if (w.iwhere[k] > 0) s += w.ws[i * w.ldp + k] * w.ws[j * w.ldp + k];
      w.wn[(col + i) * M2 + col + j] = s * theta;
    } else {                 // (1,2): row jy = i (Y index), column col + iy, iy = j (S index)
      const int jy = i, iy = j;
      double s = 0.0;
      if (jy < iy) {         // -L_a(iy, jy): active variables
        for (int k = 0; k < n; ++k) if (w.iwhere[k] > 0) s += w.ws[iy * w.ldp + k] * w.wy[jy * w.ldp + k];
        s = -s;
      } else {               // R_z(iy, jy): free variables
        for (int k = 0; k < n; ++k) if (w.iwhere[k] <= 0) s += w.ws[iy * w.ldp + k] * w.wy[jy * w.ldp + k];
      }
      w.wn[jy * M2 + col + iy] = s;
    }
  }
}

inline void HostTeam::formk_products(const Work &w, int col, double theta) { formk_products_range(w, col, theta, 0, 1); }
inline void HostTeam::shift_pairs(const Work &w) {
  for (int i = 0; i < w.n; ++i)
    for (int j = 0; j < M - 1; ++j) { w.ws[j * w.ldp + i] = w.ws[(j + 1) * w.ldp + i]; w.wy[j * w.ldp + i] = w.wy[(j + 1) * w.ldp + i]; }
}

// LEL^T factorisation of the indefinite K (formk). Free set: iwhere <= 0. Returns 0 or a failure code.
template <class T>
LB_PHASE int formk(T &tm, Work &w, int col, double theta, int *flag) {
  tm.formk_products(w, col, theta);   // the team may hand this O(m^2 n) part to more threads than itself
  tm.sync();
  tm.tick(10);
#ifdef __CUDA_ARCH__
  if constexpr (T::kWarpRegs) {
    // Cholesky of (1,1), R12 = R11^-T A12 and Cholesky of A22 + R12^T R12 in one in-register sweep
    return dev::chol_upper_warp(w.wn, M2, 2 * col, col, w.wn_rd) != 0 ? -1 : 0;
  }
#endif
  int info = chol_upper(tm, w.wn, M2, col, flag);
  if (info != 0) return -1;
  // (1,2) <- L^-1 (1,2): one column per lane of warp 0 (each a forward substitution with R^T)
  if (tm.w0) {
    for (int js = tm.lane; js < col; js += tm.nl) {
      double *b = w.wn + col + js;   // column, stride M2
      for (int r = 0; r < col; ++r) {
        double s = b[r * M2];
        for (int k = 0; k < r; ++k) s -= w.wn[k * M2 + r] * b[k * M2];
        b[r * M2] = s / w.wn[r * M2 + r];
      }
    }
  }
  tm.sync();
  // (2,2) += (1,2)^T (1,2)
  for (int e = tm.tid; e < col * col; e += tm.nt) {
    const int is = e / col, js = e % col;
    if (is > js) continue;
    double s = 0.0;
    for (int k = 0; k < col; ++k) s += w.wn[k * M2 + col + is] * w.wn[k * M2 + col + js];
    w.wn[(col + is) * M2 + col + js] += s;
  }
  tm.sync();
  info = chol_upper(tm, w.wn + col * M2 + col, M2, col, flag);
  return info != 0 ? -2 : 0;
}

// r = -Z'(B(xcp - xk) + g) for the free variables (full-length array, active entries unused)
template <class T>
LB_PHASE void cmprlb(T &tm, Work &w, int col, double theta) {
  const int n = w.n;
  bmv(tm, w, col, w.c, w.p);  // p = M c (p is free for reuse after cauchy)
  {
    const double *LB_RESTRICT wyx = w.wy, *LB_RESTRICT wsx = w.ws, *LB_RESTRICT px = w.p;
    const double *LB_RESTRICT zx = w.z, *LB_RESTRICT xx = w.x, *LB_RESTRICT gx = w.g;
    const int *LB_RESTRICT iwx = w.iwhere;
    double *LB_RESTRICT rx = w.r;
    for (int i = tm.tid; i < n; i += tm.nt) {
      double ri = 0.0;
      if (iwx[i] <= 0) {
        ri = -theta * (zx[i] - xx[i]) - gx[i];
        for (int j = 0; j < col; ++j) ri += wyx[j * w.ldp + i] * px[j] + wsx[j * w.ldp + i] * (theta * px[col + j]);
      }
      rx[i] = ri;
    }
  }
  tm.sync();
}

// Subspace minimisation with the projected (Morales-Nocedal) safeguard. On entry w.z = xcp, w.r = reduced gradient;
// on exit w.z holds the subspace minimiser. xk = w.x, gk = w.g.
// wv = (Y'Z r, theta S'Z r): the right-hand side of the subspace system (does not need the factor of K)
template <class T>
LB_PHASE void subsm_products(T &tm, Work &w, int col, double theta) {
  const int n = w.n;
  double part[M2];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int j = 0; j < M2; ++j) part[j] = 0.0;
  for (int i = tm.tid; i < n; i += tm.nt) {
    if (w.iwhere[i] <= 0) {
      const double di = w.r[i];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int j = 0; j < M; ++j) { part[j] += w.wy[j * w.ldp + i] * di; part[M + j] += w.ws[j * w.ldp + i] * di; }
    }
  }
  {
    const double *S = tm.sum_n(part);
    if (tm.tid == 0)
      for (int j = 0; j < col; ++j) { w.wv[j] = S[j]; w.wv[col + j] = theta * S[M + j]; }
  }
  tm.sync();
}

template <class T>
LB_PHASE void subsm(T &tm, Work &w, int col, double theta) {
  const int n = w.n;
  if constexpr (!T::kAsyncHelper) subsm_products(tm, w, col, theta);
  // wv <- K^-1 wv through the LEL^T factor
#ifdef __CUDA_ARCH__
  if constexpr (T::kWarpRegs) {
    dev::tri_solve_warp(w.wn, M2, 2 * col, w.wv, w.wn_rd, true);
    if (tm.lane < col) w.wv[tm.lane] = -w.wv[tm.lane];
    __syncwarp();
    dev::tri_solve_warp(w.wn, M2, 2 * col, w.wv, w.wn_rd, false);
  } else
#endif
  if (tm.w0) {
    tri_solve_w0(tm, w.wn, M2, 2 * col, w.wv, true);
    for (int j = tm.lane; j < col; j += tm.nl) w.wv[j] = -w.wv[j];
    tm.wsync();
    tri_solve_w0(tm, w.wn, M2, 2 * col, w.wv, false);
  }
  tm.sync();
  // d = (1/theta) r + (1/theta^2) Z'W wv ; projected step; remember xcp in xp
  double iword_l = 0.0, ddp_l = 0.0;
  const double rtheta = 1.0 / theta;   // one reciprocal instead of a division per (variable, pair)
  for (int i = tm.tid; i < n; i += tm.nt) {
    const double zi = w.z[i];
    w.xp[i] = zi;
    double xk = zi;
    if (w.iwhere[i] <= 0) {
      double di = w.r[i];
      for (int j = 0; j < col; ++j) di += w.wy[j * w.ldp + i] * w.wv[j] * rtheta + w.ws[j * w.ldp + i] * w.wv[col + j];
      di *= rtheta;
      w.r[i] = di;  // r now holds the Newton direction on the free variables
      xk = zi + di;
      xk = fmax(w.l[i], xk);
      xk = fmin(w.u[i], xk);
      w.z[i] = xk;
      if (xk == w.l[i] || xk == w.u[i]) iword_l = 1.0;
    }
    ddp_l += (xk - w.x[i]) * w.g[i];   // sign of the directional derivative of the projected point (same pass)
  }
  const double iword = tm.max(iword_l);
  const double ddp = tm.sum(ddp_l);
  if (iword == 0.0) { tm.sync(); return; }
  if (ddp > 0.0) {
    // fall back to the truncated (backtracked) step along d from xcp
    double alpha = 1.0;
    int ibd = 0x7fffffff;
    for (int i = tm.tid; i < n; i += tm.nt) {
      if (w.iwhere[i] <= 0) {
        const double dk = w.r[i];
        double temp1 = 1.0;
        bool cand = false;
        if (dk < 0.0) {
          const double temp2 = w.l[i] - w.xp[i];
          if (temp2 >= 0.0) { temp1 = 0.0; cand = true; }
          else if (dk * 1.0 < temp2) { temp1 = temp2 / dk; cand = true; }
        } else if (dk > 0.0) {
          const double temp2 = w.u[i] - w.xp[i];
          if (temp2 <= 0.0) { temp1 = 0.0; cand = true; }
          else if (dk * 1.0 > temp2) { temp1 = temp2 / dk; cand = true; }
        }
        if (cand && (temp1 < alpha || (temp1 == alpha && i < ibd))) { alpha = temp1; ibd = i; }
      }
    }
    tm.argmin(alpha, ibd);
    if (alpha > 1.0) alpha = 1.0;
    tm.sync();
    for (int i = tm.tid; i < n; i += tm.nt) {
      double xi = w.xp[i];
      if (w.iwhere[i] <= 0) {
        const double dk = w.r[i];
        if (alpha < 1.0 && i == ibd) xi = (dk > 0.0) ? w.u[i] : ((dk < 0.0) ? w.l[i] : xi);
        else xi += alpha * dk;
      }
      w.z[i] = xi;
    }
  }
  tm.sync();
}

// ------------------------------------------------------------------------------------------------ line search (dcsrch)
struct LineSearch {
  bool brackt;
  int stage;
  double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
};

LB_HD inline void dcstep(double &stx, double &fx, double &dx, double &sty, double &fy, double &dy, double &stp,
                         double fp, double dp, bool &brackt, double stpmin, double stpmax) {
  const double sgnd = dp * (dx / fabs(dx));
  double stpf;
  if (fp > fx) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    const double p = (gamma - dx) + theta, q = ((gamma - dx) + gamma) + dp, r = p / q;
    const double stpc = stx + r * (stp - stx);
    const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
    brackt = true;
  } else if (sgnd < 0.0) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dx, r = p / q;
    const double stpc = stp + r * (stx - stp);
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
    brackt = true;
  } else if (fabs(dp) < fabs(dx)) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta, q = (gamma + (dx - dp)) + gamma, r = p / q;
    double stpc;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
      if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
      else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
    } else {
      stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
      stpf = fmin(stpmax, stpf);
      stpf = fmax(stpmin, stpf);
    }
  } else {
    if (brackt) {
      const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      const double s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
      double gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dy, r = p / q;
      stpf = stp + r * (sty - stp);
    } else if (stp > stx) stpf = stpmax;
    else stpf = stpmin;
  }
  if (fp > fx) { sty = stp; fy = fp; dy = dp; }
  else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

// returns 0: evaluate f,g at the new stp; 1: converged; 2: warning (treated like convergence by lnsrlb); <0 error
LB_HD inline int dcsrch_start(LineSearch &ls, double f, double g, double &stp, double stpmax) {
  if (stp < STPMIN || stp > stpmax || g >= 0.0) return -1;
  ls.brackt = false;
  ls.stage = 1;
  ls.finit = f; ls.ginit = g; ls.gtest = FTOL_LS * g;
  ls.width = stpmax - STPMIN; ls.width1 = ls.width / 0.5;
  ls.stx = 0.0; ls.fx = f; ls.gx = g; ls.sty = 0.0; ls.fy = f; ls.gy = g;
  ls.stmin = 0.0; ls.stmax = stp + 4.0 * stp;
  return 0;
}

LB_HD inline int dcsrch_step(LineSearch &ls, double f, double g, double &stp, double stpmax) {
  const double ftest = ls.finit + stp * ls.gtest;
  if (ls.stage == 1 && f <= ftest && g >= 0.0) ls.stage = 2;
  int res = 0;
  if (ls.brackt && (stp <= ls.stmin || stp >= ls.stmax)) res = 2;
  if (ls.brackt && ls.stmax - ls.stmin <= XTOL_LS * ls.stmax) res = 2;
  if (stp == stpmax && f <= ftest && g <= ls.gtest) res = 2;
  if (stp == STPMIN && (f > ftest || g >= ls.gtest)) res = 2;
  if (f <= ftest && fabs(g) <= GTOL_LS * (-ls.ginit)) res = 1;
  if (res != 0) return res;
  if (ls.stage == 1 && f <= ls.fx && f > ftest) {
    const double fm = f - stp * ls.gtest;
    double fxm = ls.fx - ls.stx * ls.gtest, fym = ls.fy - ls.sty * ls.gtest;
    const double gm = g - ls.gtest;
    double gxm = ls.gx - ls.gtest, gym = ls.gy - ls.gtest;
    dcstep(ls.stx, fxm, gxm, ls.sty, fym, gym, stp, fm, gm, ls.brackt, ls.stmin, ls.stmax);
    ls.fx = fxm + ls.stx * ls.gtest; ls.fy = fym + ls.sty * ls.gtest;
    ls.gx = gxm + ls.gtest; ls.gy = gym + ls.gtest;
  } else {
    dcstep(ls.stx, ls.fx, ls.gx, ls.sty, ls.fy, ls.gy, stp, f, g, ls.brackt, ls.stmin, ls.stmax);
  }
  if (ls.brackt) {
    if (fabs(ls.sty - ls.stx) >= 0.66 * ls.width1) stp = ls.stx + 0.5 * (ls.sty - ls.stx);
    ls.width1 = ls.width;
    ls.width = fabs(ls.sty - ls.stx);
  }
  if (ls.brackt) { ls.stmin = fmin(ls.stx, ls.sty); ls.stmax = fmax(ls.stx, ls.sty); }
  else { ls.stmin = stp + 1.1 * (stp - ls.stx); ls.stmax = stp + 4.0 * (stp - ls.stx); }
  stp = fmax(stp, STPMIN);
  stp = fmin(stp, stpmax);
  if ((ls.brackt && (stp <= ls.stmin || stp >= ls.stmax)) || (ls.brackt && ls.stmax - ls.stmin <= XTOL_LS * ls.stmax))
    stp = ls.stx;
  return 0;
}

// ------------------------------------------------------------------------------------------------ BFGS matrices
// Append the pair (s = d, y = r) to the history (oldest first), update S'Y, S'S and theta (matupd), then form and
// factor T = theta S'S + L D^-1 L' (formt). Returns 0 or -3 if T is not positive definite.
template <class T>
LB_PHASE int update_pairs(T &tm, Work &w, int &col, double &theta, double rr, double dr, double stp, double dtd,
                       int *flag) {
  const int n = w.n;
  if (col == M) {  // drop the oldest pair
    tm.shift_pairs(w);   // rows 1..M-1 of S and Y move up by one (the team may hand this to more threads than itself)
    tm.sync();
#ifdef __CUDA_ARCH__
    if constexpr (T::kWarpRegs) {
      // S'Y and S'S move up-left by one, row by row (row i is rewritten from row i + 1 before row i + 1 is): lanes
      // 0..8 carry S'Y, lanes 16..24 carry S'S, one element of the row each
      const int j = tm.lane & 15;
      double *mat = (tm.lane < 16) ? w.sy : w.ss;
      for (int i = 0; i < M - 1; ++i) {
        double v = 0.0;
        if (j < M - 1) v = mat[(i + 1) * M + j + 1];
        tm.wsync();
        if (j < M - 1) mat[i * M + j] = v;
      }
    } else
#endif
    if (tm.tid == 0) {
      for (int i = 0; i < M - 1; ++i)
        for (int j = 0; j < M - 1; ++j) { w.sy[i * M + j] = w.sy[(i + 1) * M + j + 1]; w.ss[i * M + j] = w.ss[(i + 1) * M + j + 1]; }
    }
    tm.sync();
    col = M - 1;
  }
  const int c = col;
  tm.tick(13);
  theta = rr / dr;
  // the new pair goes into row c and, in the same pass, the new row of S'Y (s_new . y_j) and the new column of S'S
  // (s_j . s_new) are accumulated (all M rows: the sums of the rows >= c are never read)
  double part[M2];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int j = 0; j < M2; ++j) part[j] = 0.0;
  for (int i = tm.tid; i < n; i += tm.nt) {
    const double di = w.d[i];
    w.ws[c * w.ldp + i] = di;
    w.wy[c * w.ldp + i] = w.r[i];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = 0; j < M; ++j) { part[j] += di * w.wy[j * w.ldp + i]; part[M + j] += w.ws[j * w.ldp + i] * di; }
  }
  {
    const double *S = tm.sum_n(part);
    if (tm.tid == 0) {
      for (int j = 0; j < c; ++j) { w.sy[c * M + j] = S[j]; w.ss[j * M + c] = S[M + j]; }
      w.ss[c * M + c] = (stp == 1.0) ? dtd : stp * stp * dtd;
      w.sy[c * M + c] = dr;
    }
  }
  col = c + 1;
  tm.sync();
  tm.tick(14);
  // the pairs are in place: the first pass of the NEXT iteration's Cauchy point can start (helper team, if there is one)
  if constexpr (T::kAsyncHelper) tm.start_cauchy_pass1(w, col, theta);
  for (int k = tm.tid; k < col; k += tm.nt) { w.rsy[k] = 1.0 / w.sy[k * M + k]; w.rsq[k] = 1.0 / sqrt(w.sy[k * M + k]); }
  tm.sync();
  // formt: upper triangle of T
  for (int e = tm.tid; e < col * col; e += tm.nt) {
    const int i = e / col, j = e % col;
    if (i > j) continue;
    double s = 0.0;
    const int k1 = (i < j ? i : j);
    for (int k = 0; k < k1; ++k) s += w.sy[i * M + k] * w.sy[j * M + k] * w.rsy[k];   // rsy = 1 / diag(S'Y)
    w.wt[i * M + j] = s + theta * w.ss[i * M + j];
  }
  tm.sync();
#ifdef __CUDA_ARCH__
  if constexpr (T::kWarpRegs) {
    const int bad = dev::chol_upper_warp(w.wt, M, col, col, w.wt_rd);
    if constexpr (T::kAsyncHelper) tm.join_helper();
    return bad != 0 ? -3 : 0;
  }
#endif
  return chol_upper(tm, w.wt, M, col, flag) != 0 ? -3 : 0;
}

// ------------------------------------------------------------------------------------------------ driver
// Fn: functor  double operator()(Team&, const double* x, double* g)  -> f, fills g; called by the whole team.
struct Result {
  int status, nit, nfev;
  double f;
};

template <class T, class Fn>
LB_HD Result minimize(T &tm, Work &w, Fn &fn, int maxiter, double factr, double pgtol, int *flag) {
  const int n = w.n;
  Result res;
  res.nit = 0; res.nfev = 0; res.status = RUNNING;
  // project the start point, initialise iwhere
  for (int i = tm.tid; i < n; i += tm.nt) {
    double xi = w.x[i];
    if (xi <= w.l[i]) xi = w.l[i];
    else if (xi >= w.u[i]) xi = w.u[i];
    w.x[i] = xi;
    w.iwhere[i] = (w.u[i] - w.l[i] <= 0.0) ? 3 : 0;
  }
  tm.sync();
  // The objective has ONE call site (inside the line-search loop): the first pass of the main loop only evaluates f, g at
  // the start point. On the GPU the evaluation is inlined at its call site and the optimiser warp is bound by
  // instruction fetch, so every further copy of it costs.
  double f = 0.0;
  int col = 0;
  double theta = 1.0;
  bool updatd = false;
  (void)updatd;
  double sbgnrm = 0.0;
  int iter = 0;
  int passes = 0;
  bool first = true;
  bool pass1_ready = false;   // the helper team ran cauchy_pass1 for the current iterate during update_pairs
  while (true) {
    double dtd = 0.0, stpmx = BIG, stp = 1.0, fold = f, gd = 0.0, gdold = 0.0;
    int ifun = 0, iback = 0;
    LineSearch ls;
    bool ls_failed = false;
    if (!first) {
      // every pass either completes an iteration or discards the BFGS memory once; the guard only fires on NaN storms
      if (++passes > 3 * maxiter + 20) { res.status = ABNORMAL; break; }
      int nseg;
      tm.tick(0);
      cauchy(tm, w, col, theta, sbgnrm, nseg, pass1_ready);
      pass1_ready = false;
      tm.tick(1);
      // free variables at the Cauchy point
      double nfree_l = 0.0;
      for (int i = tm.tid; i < n; i += tm.nt) nfree_l += (w.iwhere[i] <= 0) ? 1.0 : 0.0;
      const int nfree = (int)(tm.sum(nfree_l) + 0.5);
      if (nfree != 0 && col != 0) {
        const int info = formk(tm, w, col, theta, flag);
        if constexpr (T::kAsyncHelper) tm.join_helper();   // cmprlb + subsm_products were computed meanwhile
        if (info != 0) {  // refresh the memory and restart the iteration
          col = 0; theta = 1.0; updatd = false;
          continue;
        }
        tm.tick(2);
        if constexpr (!T::kAsyncHelper) cmprlb(tm, w, col, theta);
        tm.tick(3);
        subsm(tm, w, col, theta);
        tm.tick(4);
      }
      // search direction and line-search set-up in ONE pass over the variables: d = z - x, d'd, the largest feasible
      // step, the saved iterate (t, r) = (x, g) and g'd; the three reductions are independent chains
      {
        double dtd_l = 0.0, gd_l = 0.0, smx = BIG;
        for (int i = tm.tid; i < n; i += tm.nt) {
          const double xi = w.x[i], gi = w.g[i];
          const double a1 = w.z[i] - xi;
          w.d[i] = a1;
          w.t[i] = xi;
          w.r[i] = gi;
          dtd_l += a1 * a1;
          gd_l += gi * a1;
          if (a1 < 0.0) {
            const double a2 = w.l[i] - xi;
            if (a2 >= 0.0) smx = 0.0;
            else if (a1 * smx < a2) smx = a2 / a1;
          } else if (a1 > 0.0) {
            const double a2 = w.u[i] - xi;
            if (a2 <= 0.0) smx = 0.0;
            else if (a1 * smx > a2) smx = a2 / a1;
          }
        }
        dtd = tm.sum(dtd_l);
        gd = tm.sum(gd_l);
        stpmx = (iter == 0) ? 1.0 : -tm.max(-smx);   // (boxed problem: the first step is 1 as well)
        tm.sync();
      }
      gdold = gd;
      if (gd >= 0.0) ls_failed = true;  // not a descent direction
      else if (dcsrch_start(ls, f, gd, stp, stpmx) < 0) ls_failed = true;
      tm.tick(5);
    }
    if (!ls_failed) {
      while (true) {
        if (!first) {
          ++ifun; ++res.nfev; iback = ifun - 1;
          if (stp == 1.0) { for (int i = tm.tid; i < n; i += tm.nt) w.x[i] = w.z[i]; }
          else { for (int i = tm.tid; i < n; i += tm.nt) w.x[i] = stp * w.d[i] + w.t[i]; }
          tm.sync();
        }
        tm.tick(6);
        f = fn(tm, w.x, w.g);
        tm.tick(7);
        if (first) break;
        gd = dot(tm, w.g, w.d, n);
        const int r = dcsrch_step(ls, f, gd, stp, stpmx);
        if (r != 0) break;
        if (iback + 1 >= MAXLS) { ls_failed = true; break; }
      }
    }
    if (first) {
      first = false;
      res.nfev = 1;
      sbgnrm = projgr(tm, w);
      if (sbgnrm <= pgtol) { res.status = CONV_PGTOL; break; }
      continue;
    }
    if (ls_failed) {
      // restore the previous iterate
      for (int i = tm.tid; i < n; i += tm.nt) { w.x[i] = w.t[i]; w.g[i] = w.r[i]; }
      tm.sync();
      f = fold;
      if (col == 0) { res.status = ABNORMAL; break; }
      col = 0; theta = 1.0; updatd = false;
      continue;
    }
    tm.tick(8);
    ++iter;
    // projected-gradient norm of the new iterate and, in the same pass, the BFGS pair: r = g - g_old, r'r, d scaled by
    // the accepted step (if one of the tests below ends the run, r and d are not looked at again)
    double rr;
    {
      double sb = 0.0, rr_l = 0.0;
      for (int i = tm.tid; i < n; i += tm.nt) {
        const double gi = w.g[i], xi = w.x[i];
        const double pg = (gi < 0.0) ? fmax(xi - w.u[i], gi) : fmin(xi - w.l[i], gi);
        sb = fmax(sb, fabs(pg));
        const double ri = gi - w.r[i];
        w.r[i] = ri;
        rr_l += ri * ri;
        if (stp != 1.0) w.d[i] *= stp;
      }
      sbgnrm = tm.max(sb);
      rr = tm.sum(rr_l);
    }
    res.nit = iter;
    if (iter >= maxiter) { res.status = STOP_MAXITER; break; }       // scipy driver stops at NEW_X before the tests
    if (sbgnrm <= pgtol) { res.status = CONV_PGTOL; break; }
    const double ddum0 = fmax(fabs(fold), fmax(fabs(f), 1.0));
    if ((fold - f) <= EPSMCH * factr * ddum0) { res.status = CONV_FTOL; break; }
    double dr, ddum;
    if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; }
    else { dr = (gd - gdold) * stp; ddum = -gdold * stp; }
    if (dr <= EPSMCH * ddum) { updatd = false; continue; }  // skip the update
    updatd = true;
    const int info = update_pairs(tm, w, col, theta, rr, dr, stp, dtd, flag);
    tm.tick(9);
    pass1_ready = T::kAsyncHelper && info == 0;   // (a failed factorisation discards the memory: col and theta change)
    if (info != 0) { col = 0; theta = 1.0; updatd = false; }
  }
  res.f = f;
  return res;
}

}  // namespace lbfgsb
