// Device-resident multi-start optimiser: one thread block per joint restart problem (batch_limit points of one run)
// runs the whole L-BFGS-B minimisation of -sum(log-EI) with the value+gradient evaluated in the same block, so an
// optimize_acqf call costs ONE launch instead of hundreds of host round trips.
//
// Replaces botorch gen_candidates_scipy -> scipy.optimize.minimize(method="L-BFGS-B", options={"maxiter": 200})
// inside the reference's optimize_acqf call (PCA_BO.py:607-614). The algorithm is
// lbfgsb_core.cuh (validated on the host against scipy, tests/test_lbfgsb_core.py); tolerances are scipy's defaults
// (factr = 2.22e-9 / eps = 1e7, pgtol = 1e-5, m = 10, maxls = 20).
//
// Block = 9 warps, one block per SM (up to 227 KB of shared memory): warp 0 is the optimiser ("brain": Cauchy point,
// subspace minimisation, line search, BFGS update; small dense kernels warp-synchronous in shared memory), warps 0-7
// evaluate log-EI value + gradient together (lbfgsb_eval.cuh: one sweep over the packed W on the FP64 tensor cores) and
// share formk's O(m^2 n) products and the pair shift, warp 8 is the bulk-copy (TMA) producer of the W panel ring. The
// compute warps synchronise on a named barrier the producer never joins; the brain posts commands through a mailbox.
#include "lbfgsb_core.cuh"
#include "lbfgsb_eval.cuh"

namespace {

constexpr int LB_THREADS = lbeval::NT;

enum Cmd { CMD_EVAL = 1, CMD_FORMK = 2, CMD_EXIT = 3, CMD_SHIFT = 4, CMD_CAUCHY1 = 5 };

// Raw inner products of formk (see lbfgsb::formk_products_range for the layout) by the whole block. 8 lanes share one
// ROW of a block of the matrix (k = seg, seg + 8, ...): rows 0..col-1 of Y'ZZ'Y (entries j >= i), of S'AA'S (j >= i) and
// of the mixed block (all jy for one iy), 3 col <= 30 row tasks = 240 threads, ONE pass. The row's own vector and the
// free / active indicator are loaded once per k for up to 10 entries (round 1 dealt single entries to the lane groups:
// three loads per FMA and seven passes, 14 k cycles; now 8 k). The free / active split is a 0/1 multiplier folded into
// the row's own value instead of a branch; the summation order per entry is a fixed xor-shuffle tree (deterministic).
// (force-inlined into the kernel so that the pointers of `w` keep their shared-memory provenance: as a separate function
// it reads them from the struct and every access becomes a generic load)
__device__ __forceinline__ void formk_products_block(const lbfgsb::Work &w, int col, double theta) {
  using lbfgsb::M;
  using lbfgsb::M2;
  const int n = w.n, ldp = w.ldp, tid = threadIdx.x, seg = tid & 7, task = tid >> 3;
  const double *__restrict__ wy = w.wy;
  const double *__restrict__ ws = w.ws;
  const int *__restrict__ iw = w.iwhere;
  const int blk = task / col, i = task % col;   // tasks >= 3 col: nothing to do (but they join the shuffles)
  double s[M];
#pragma unroll
  for (int j = 0; j < M; ++j) s[j] = 0.0;
  if (blk < 3) {
    // row vector a (indexed by i) and the vectors b_j it meets; entry j is used for j >= j0 (triangular blocks)
    const double *a = (blk == 0) ? wy + i * ldp : ws + i * ldp;
    const double *bb = (blk == 1) ? ws : wy;
    // Every entry j < M is accumulated, from a valid row, whether it is used or not (the unused ones are dropped when
    // the results are stored): no branch and no select per entry. The 0/1 indicator of the free / active split is
    // folded into the row's own value (a_k * m is exact), so an entry costs ONE FMA per k. The phase is bound by
    // instruction issue (8 warps on 4 schedulers): with a select and a separate multiplication per entry it took
    // 11.5 k cycles (ncu: 170 M instructions on that one source line).
    int roff[M];
#pragma unroll
    for (int j = 0; j < M; ++j) roff[j] = (j < col ? j : col - 1) * ldp;
    if (blk < 2) {   // Y'ZZ'Y (free variables) / S'AA'S (active variables): one indicator for the whole row
      for (int k = seg; k < n; k += 8) {
        const bool fr = iw[k] <= 0;
        const double akm = (fr == (blk == 0)) ? a[k] : 0.0;
        double bv[M];
#pragma unroll
        for (int j = 0; j < M; ++j) bv[j] = bb[roff[j] + k];
#pragma unroll
        for (int j = 0; j < M; ++j) s[j] = fma(akm, bv[j], s[j]);
      }
    } else {         // row iy = i of S against column jy = j of Y: free variables if jy >= iy, active ones otherwise
      for (int k = seg; k < n; k += 8) {
        const bool fr = iw[k] <= 0;
        const double ak = a[k];
        const double a_hi = fr ? ak : 0.0, a_lo = fr ? 0.0 : ak;
        double bv[M];
#pragma unroll
        for (int j = 0; j < M; ++j) bv[j] = bb[roff[j] + k];
#pragma unroll
        for (int j = 0; j < M; ++j) s[j] = fma((j < i) ? a_lo : a_hi, bv[j], s[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < M; ++j) {
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 1);
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 2);
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 4);
  }
  if (blk < 3 && seg == 0) {
#pragma unroll
    for (int j = 0; j < M; ++j) {
      if (j < col) {
        if (blk == 0) { if (j >= i) w.wn[i * M2 + j] = s[j] / theta + ((i == j) ? w.sy[i * M + i] : 0.0); }
        else if (blk == 1) { if (j >= i) w.wn[(col + i) * M2 + col + j] = s[j] * theta; }
        else w.wn[j * M2 + col + i] = (j < i) ? -s[j] : s[j];   // row jy = j, column col + iy
      }
    }
  }
}

// The oldest pair leaves the memory: rows 1..M-1 of S and Y move up by one, all threads of the block (a thread per
// variable; loads first, then stores).
__device__ __forceinline__ void shift_pairs_block(const lbfgsb::Work &w) {
  using lbfgsb::M;
  const int n = w.n, ldp = w.ldp;
  for (int i = threadIdx.x; i < n; i += LB_THREADS) {
    double a[M - 1], b[M - 1];
#pragma unroll
    for (int j = 0; j < M - 1; ++j) { a[j] = w.ws[(j + 1) * ldp + i]; b[j] = w.wy[(j + 1) * ldp + i]; }
#pragma unroll
    for (int j = 0; j < M - 1; ++j) { w.ws[j * ldp + i] = a[j]; w.wy[j * ldp + i] = b[j]; }
  }
}

// Block-level mailbox: warp 0 runs the optimiser and posts commands; warps 1..7 wait in a command loop and join the
// block-wide phases (log-EI evaluation, the O(m^2 n) formk products). Everything else is warp-synchronous in warp 0.
struct Mailbox {
  int cmd, col;
  double theta;
};

// The L-BFGS-B "team" on the device is ONE warp: reductions are shuffles, sync is __syncwarp (no block barriers).
struct CudaWarpTeam {
  static constexpr bool kWarpRegs = true;
  // warp 1 computes cmprlb and the first products of subsm while warp 0 factors K (they only need the Cauchy point and the
  // factor of T): a named barrier of the two warps joins them (6.6 k + 3 k cycles hidden behind the 11 k of the Cholesky)
  static constexpr bool kAsyncHelper = true;
  __device__ void join_helper() { asm volatile("bar.sync 2, 64;" ::: "memory"); }
  // wakes the block; warp 1 runs lbfgsb::cauchy_pass1 for the next iteration and joins at the end of update_pairs
  __device__ void start_cauchy_pass1(const lbfgsb::Work &, int col, double theta) {
    Mailbox *m = post_slot();
    if (lane == 0) { m->cmd = CMD_CAUCHY1; m->col = col; m->theta = theta; }
    lbeval::cta_sync();
  }
  int tid, nt, lane, nl;
  bool w0;
  Mailbox *mb;          // TWO entries used alternately: a command is never overwritten while a warp may still be reading it
  int par;              // entry of the next command (every warp of the block counts the command rounds the same way)
  __device__ Mailbox *post_slot() { Mailbox *m = mb + par; par ^= 1; return m; }
  const lbfgsb::Work *wk;
  long long *prof;      // optional [16] cycle counters of this block (lane 0), or nullptr
  double *red;          // shared-memory scratch of sum_n: [20][33] + [20]
  long long last;
  // attribute the cycles since the previous tick to the phase that just ended (id = phase that STARTS now)
  __device__ void tick(int id) {
    if (prof == nullptr || lane != 0) return;
    const long long t = clock64();
    prof[id] += t - last;
    last = t;
  }
  __device__ void sync() { __syncwarp(); }
  __device__ void wsync() { __syncwarp(); }
  __device__ double sum(double v) { return warp_sum(v); }
  __device__ double max(double v) { return warp_max(v); }
  // Sums over the warp of M2 = 20 values per lane, through shared memory: lane j adds column j of the 32 x 20 table
  // (row stride 33: conflict free both ways) in four chains. A butterfly over all 20 values costs 200 shuffle
  // instructions per call site, and the optimiser warp is bound by instruction fetch (the kernel is ~600 KB of code).
  // Returns the 20 sums (shared memory).
  __device__ const double *sum_n(const double *vals) {
#pragma unroll
    for (int j = 0; j < lbfgsb::M2; ++j) red[j * 33 + lane] = vals[j];
    __syncwarp();
    if (lane < lbfgsb::M2) {
      const double *r = red + lane * 33;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 2
      for (int i = 0; i < 32; i += 4) { a0 += r[i]; a1 += r[i + 1]; a2 += r[i + 2]; a3 += r[i + 3]; }
      red[lbfgsb::M2 * 33 + lane] = (a0 + a1) + (a2 + a3);
    }
    __syncwarp();
    return red + lbfgsb::M2 * 33;
  }
  __device__ void argmin(double &val, int &idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, val, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov < val || (ov == val && oi < idx)) { val = ov; idx = oi; }
    }
  }
  __device__ void formk_products(const lbfgsb::Work &w, int col, double theta) {
    tick(23);
    Mailbox *m = post_slot();
    if (lane == 0) { m->cmd = CMD_FORMK; m->col = col; m->theta = theta; }
    lbeval::cta_sync();
    tick(24);
    formk_products_block(w, col, theta);
    tick(25);
    lbeval::cta_sync();
    tick(26);
  }
  __device__ void shift_pairs(const lbfgsb::Work &w) {
    Mailbox *m = post_slot();
    if (lane == 0) m->cmd = CMD_SHIFT;
    lbeval::cta_sync();
    shift_pairs_block(w);
    lbeval::cta_sync();
  }
};

template <int PT, int R, int TI>
struct LogEiObjective {
  lbeval::Ctx ctx;
  int pcount, kk;
  double *vals;  // smem [8]
  // called by warp 0 (the team); posts the evaluation to the whole block
  __device__ double operator()(CudaWarpTeam &tm, const double *x, double *g) {
    Mailbox *m = tm.post_slot();
    if (tm.lane == 0) m->cmd = CMD_EVAL;
    lbeval::cta_sync();
    return run(x, g);
  }
  // executed by all LB_THREADS threads
  __device__ double run(const double *x, double *g) {
    lbeval::eval<PT, R, TI>(ctx, pcount, x, vals, g, -1.0);
    double f = 0.0;
    for (int p = 0; p < pcount; ++p) f -= vals[p];
    return f;
  }
};

struct LbArgs {
  pcabo_gp_model m;
  int num_restarts, batch_limit, maxiter;
  double factr, pgtol;
  const double *ics, *box_lo, *box_hi;
  double *cand, *fval;
  int *status, *nit, *nfev;
  long long *prof;   // optional [B][nchunks][32] cycle counters
  const int *run_idx;  // optional [gridDim.y] runs to optimise (nullptr: all runs, blockIdx.y)
  int k_bound;         // shared memory was sized for k <= k_bound
  int ring_cap;        // doubles of the W panel ring
  int resident;        // every panel fits the ring
  double *wsy;         // optional global home of the BFGS pairs [gridDim.y][gridDim.x][2 M nvar_max] (nullptr: shared memory)
  int nvar_max;
};

template <int PT, int R, int TI, bool PG>
__global__ void __launch_bounds__(lbeval::NT_ALL, 1) lbfgsb_logei_kernel(LbArgs a) {   // 9 warps: one scheduler hosts 3 -> 168 registers
  extern __shared__ __align__(16) double sm[];
  const int b = a.run_idx ? a.run_idx[blockIdx.y] : blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  const int nchunks = gridDim.x;
  const int p0 = chunk * a.batch_limit;
  const int pcount = min(a.batch_limit, a.num_restarts - p0);
  LogEiObjective<PT, R, TI> obj;
  lbeval::load_ctx(obj.ctx, a.m, b);
  const int kk = obj.ctx.kk, kmax = a.m.kmax;
  const int nvar = pcount * kk;
  if (kk > a.k_bound) {   // the caller's bound on k was wrong: refuse (status -1) instead of overrunning shared memory
    if (tid == 0) {
      const size_t o = (size_t)b * nchunks + chunk;
      a.status[o] = -1; a.nit[o] = 0; a.nfev[o] = 0; a.fval[o] = 0.0;
    }
    return;
  }
  double *q = sm;
  obj.ctx.L = lbeval::make_layout(PT, kmax, a.m.n, R);
  obj.ctx.base = q;
  obj.ctx.ring_cap = a.ring_cap;
  obj.ctx.resident = a.resident != 0;
  q += obj.ctx.L.o_ring + a.ring_cap;
  lbfgsb::Work w;
  double *wsy = PG ? a.wsy + ((size_t)blockIdx.y * nchunks + chunk) * (size_t)(2 * lbfgsb::M) * lbfgsb::pair_stride(a.nvar_max) : nullptr;
  lbfgsb::carve<PG>(w, nvar, q, wsy);   // PG: the BFGS pairs live in the caller's workspace (global memory)
  q += lbfgsb::work_doubles(nvar, PG);
  Mailbox *mb = reinterpret_cast<Mailbox *>(q); q += 4;   // two entries
  obj.vals = q; q += 8;
  int *flag = reinterpret_cast<int *>(q); q += 2;
  long long *sprof = reinterpret_cast<long long *>(q); q += 32;   // cycle counters live in shared memory, flushed at the end
  // scratch of the optimiser warp's 20-value reductions: the partial-sum area of the evaluation (2 x 8 x 8 x 12 doubles),
  // which is only live inside an evaluation
  double *tred = obj.ctx.base + obj.ctx.L.o_part;
  obj.pcount = pcount; obj.kk = kk;
  if (a.prof) {
    if (tid < 32) sprof[tid] = 0;
    obj.ctx.prof = sprof;
  }
  lbeval::init(obj.ctx, PT);   // mbarriers, zero padding (all threads; ends with a block barrier)
  const double *blo = a.box_lo + (size_t)b * kmax, *bhi = a.box_hi + (size_t)b * kmax;
  lbfgsb::Result r;
  if (tid >= LB_THREADS) {
    // ---- producer warp: keeps the panel ring of W full for the whole launch (lane 0), then waits at the final barrier
    if (tid == LB_THREADS) lbeval::producer_loop(obj.ctx);
    __syncwarp();
  } else {
    const double *ic = a.ics + ((size_t)b * a.num_restarts + p0) * kmax;
    // rows of S / Y beyond the pairs stored so far take part in the unguarded products (their sums are never read):
    // start them from zero rather than from whatever the memory held
    for (int i = tid; i < 2 * lbfgsb::M * w.ldp; i += LB_THREADS) w.ws[i] = 0.0;
    for (int i = tid; i < nvar; i += LB_THREADS) {
      const int p = i / kk, j = i % kk;
      const double lo = blo[j], hi = bhi[j];
      w.l[i] = lo; w.u[i] = hi;
      w.x[i] = fmax(fmin(ic[(size_t)p * kmax + j], hi), lo);   // columnwise_clamp of the initial conditions
    }
    lbeval::cta_sync();
    if (tid < 32) {
      CudaWarpTeam tm;
      tm.tid = tid; tm.nt = 32; tm.lane = tid; tm.nl = 32; tm.w0 = true; tm.mb = mb; tm.par = 0; tm.wk = &w;
      tm.prof = a.prof ? sprof : nullptr;
      tm.red = tred;
      tm.last = clock64();
      r = lbfgsb::minimize(tm, w, obj, a.maxiter, a.factr, a.pgtol, flag);
      {
        Mailbox *m = tm.post_slot();
        if (tid == 0) { m->cmd = CMD_EXIT; lbeval::signal_exit(obj.ctx); }
      }
      lbeval::cta_sync();
    } else {
      int par = 0;
      for (;;) {
        lbeval::cta_sync();
        const Mailbox *cm = mb + par;
        par ^= 1;
        const int c = cm->cmd;
        if (c == CMD_EVAL) {
          obj.run(w.x, w.g);
        } else if (c == CMD_FORMK) {
          const int fcol = cm->col;
          const double ftheta = cm->theta;
          formk_products_block(w, fcol, ftheta);
          lbeval::cta_sync();
          if (tid < 64) {   // warp 1: the helper team (see CudaWarpTeam::kAsyncHelper)
            CudaWarpTeam th;
            th.tid = tid - 32; th.nt = 32; th.lane = tid - 32; th.nl = 32; th.w0 = true; th.mb = mb; th.par = 0; th.wk = &w;
            th.prof = nullptr; th.red = tred + 720; th.last = 0;
            lbfgsb::cmprlb(th, w, fcol, ftheta);
            lbfgsb::subsm_products(th, w, fcol, ftheta);
            th.join_helper();
          }
        } else if (c == CMD_SHIFT) {
          shift_pairs_block(w);
          lbeval::cta_sync();
        } else if (c == CMD_CAUCHY1) {
          if (tid < 64) {   // warp 1: first pass of the next Cauchy point while warp 0 forms and factors T
            const int ccol = cm->col;
            const double ctheta = cm->theta;
            CudaWarpTeam th;
            th.tid = tid - 32; th.nt = 32; th.lane = tid - 32; th.nl = 32; th.w0 = true; th.mb = mb; th.par = 0; th.wk = &w;
            th.prof = nullptr; th.red = tred + 720; th.last = 0;
            double f1;
            int nbreak;
            lbfgsb::cauchy_pass1(th, w, ccol, ctheta, f1, nbreak);
            th.join_helper();
          }
        } else {
          break;
        }
      }
    }
  }
  __syncthreads();   // every bulk copy has landed (the producer drained the ring before it got here)
  double *out = a.cand + ((size_t)b * a.num_restarts + p0) * kmax;
  for (int i = tid; i < pcount * kmax; i += lbeval::NT_ALL) {
    const int p = i / kmax, j = i % kmax;
    out[i] = (j < kk) ? fmax(fmin(w.x[p * kk + j], bhi[j]), blo[j]) : 0.0;
  }
  if (tid == 0) {
    const size_t o = (size_t)b * nchunks + chunk;
    a.status[o] = r.status; a.nit[o] = r.nit; a.nfev[o] = r.nfev; a.fval[o] = r.f;
  }
  if (a.prof && tid < 32) a.prof[((size_t)b * nchunks + chunk) * 32 + tid] += sprof[tid];
}

constexpr size_t SMEM_LIMIT = 227 * 1024;
constexpr int TAIL_DOUBLES = 4 + 8 + 2 + 32 + 8;   // mailbox (two entries), values, flag, cycle counters, slack

// Shared-memory plan of one launch: panel height, ring capacity, where the BFGS pairs live.
struct Plan {
  int R, TI, ring_cap, resident, pairs_global;
  size_t smem;
};

inline bool make_plan(int PT, int n, int kmax, int nvar_max, bool have_ws, Plan &pl) {
  const int tiles = (n + 7) / 8;
  pl.TI = (tiles + lbeval::NW - 1) / lbeval::NW;
  // preference: pairs in shared memory before pairs in the workspace; everything resident (32-row panels) before
  // streaming; 16-row panels when two of the largest fit the ring, else 8-row panels (the ring must hold the largest).
  // The ring takes all the shared memory that is left (at most all panels).
  for (int pairs_global = 0; pairs_global <= (have_ws ? 1 : 0); ++pairs_global) {
    const size_t work = lbfgsb::work_doubles(nvar_max, pairs_global != 0) + TAIL_DOUBLES;
    for (int pass = 0; pass < 3; ++pass) {
      const int R = pass == 0 ? 32 : pass == 1 ? 16 : 8;
      if (R != 8 && (pl.TI > 8 || (pairs_global && R == 32))) continue;   // kernel variants that are not instantiated
      const lbeval::Layout L = lbeval::make_layout(PT, kmax, n, R);
      const size_t fixed = ((size_t)L.o_ring + work) * sizeof(double);
      if (fixed >= SMEM_LIMIT) continue;
      const long long avail = (long long)((SMEM_LIMIT - fixed) / sizeof(double)) & ~1LL;
      const bool resident = L.all_panels <= avail && L.npanels <= lbeval::NBAR;
      if (pass == 0 && !resident) continue;
      if (pass == 1 && !resident && avail < 2LL * L.max_panel) continue;
      if (avail < L.max_panel) continue;
      pl.R = R;
      pl.resident = resident ? 1 : 0;
      pl.ring_cap = (int)(avail < L.all_panels ? avail : L.all_panels);
      pl.pairs_global = pairs_global;
      pl.smem = ((size_t)L.o_ring + pl.ring_cap + work) * sizeof(double);
      return true;
    }
  }
  return false;
}

template <int PT, int R, int TI, bool PG>
int launch_one(const LbArgs &a, int nsel, size_t smem, cudaStream_t st) {
  const int nchunks = (a.num_restarts + a.batch_limit - 1) / a.batch_limit;
  PCABO_CUDA_CHECK(cudaFuncSetAttribute(lbfgsb_logei_kernel<PT, R, TI, PG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(nchunks, nsel);
  lbfgsb_logei_kernel<PT, R, TI, PG><<<grid, lbeval::NT_ALL, smem, st>>>(a);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}

int launch(LbArgs &a, int nsel, double *workspace, long long workspace_doubles, cudaStream_t st) {
  constexpr int PT = 5;
  const int nchunks = (a.num_restarts + a.batch_limit - 1) / a.batch_limit;
  a.nvar_max = a.batch_limit * a.k_bound;
  const long long ws_need = (long long)nsel * nchunks * 2 * lbfgsb::M * lbfgsb::pair_stride(a.nvar_max);
  Plan pl;
  if (!make_plan(PT, a.m.n, a.m.kmax, a.nvar_max, workspace != nullptr && workspace_doubles >= ws_need, pl))
    return PCABO_E_UNSUPPORTED;
  a.ring_cap = pl.ring_cap; a.resident = pl.resident; a.wsy = pl.pairs_global ? workspace : nullptr;
  // instantiated: shared-memory pairs with (R, TI) = (32, 8), (16, 8), (8, 8), (8, 17); workspace pairs with (16, 8),
  // (8, 8), (8, 17). make_plan only produces these (n > 512 always streams 8-row panels).
  if (!pl.pairs_global) {
    if (pl.TI <= 8) {
      if (pl.R == 32) return launch_one<PT, 32, 8, false>(a, nsel, pl.smem, st);
      if (pl.R == 16) return launch_one<PT, 16, 8, false>(a, nsel, pl.smem, st);
      return launch_one<PT, 8, 8, false>(a, nsel, pl.smem, st);
    }
    if (pl.TI <= 17 && pl.R == 8) return launch_one<PT, 8, 17, false>(a, nsel, pl.smem, st);
  } else {
    if (pl.TI <= 8 && pl.R == 16) return launch_one<PT, 16, 8, true>(a, nsel, pl.smem, st);
    if (pl.TI <= 8 && pl.R == 8) return launch_one<PT, 8, 8, true>(a, nsel, pl.smem, st);
    if (pl.TI <= 17 && pl.R == 8) return launch_one<PT, 8, 17, true>(a, nsel, pl.smem, st);
  }
  return PCABO_E_UNSUPPORTED;
}

}  // namespace

extern "C" long long pcabo_lbfgsb_workspace_doubles(int n, int kmax, int k_bound, int num_restarts, int batch_limit,
                                                    int nsel) {
  if (n < 1 || kmax < 1 || num_restarts < 1 || batch_limit < 1 || batch_limit > 5 || nsel < 1) return -1;
  if (k_bound <= 0 || k_bound > kmax) k_bound = kmax;
  Plan pl;
  const int nvar_max = batch_limit * k_bound;
  if (make_plan(5, n, kmax, nvar_max, false, pl)) return 0;          // everything fits shared memory
  if (!make_plan(5, n, kmax, nvar_max, true, pl)) return -1;         // does not fit either way
  const int nchunks = (num_restarts + batch_limit - 1) / batch_limit;
  return (long long)nsel * nchunks * 2 * lbfgsb::M * lbfgsb::pair_stride(nvar_max);
}

extern "C" int pcabo_lbfgsb_plan(int n, int kmax, int k_bound, int batch_limit, int have_workspace, int *plan6) {
  if (n < 1 || kmax < 1 || batch_limit < 1 || batch_limit > 5 || !plan6) return PCABO_E_INVALID;
  if (k_bound <= 0 || k_bound > kmax) k_bound = kmax;
  Plan pl;
  if (!make_plan(5, n, kmax, batch_limit * k_bound, have_workspace != 0, pl)) return PCABO_E_UNSUPPORTED;
  plan6[0] = pl.R; plan6[1] = pl.ring_cap; plan6[2] = pl.resident; plan6[3] = 0; plan6[4] = pl.pairs_global;
  plan6[5] = (int)pl.smem;
  return PCABO_OK;
}

extern "C" int pcabo_lbfgsb_logei_runs_ws(const pcabo_gp_model *m, int nsel, const int *run_idx, int k_bound,
                                          int num_restarts, int batch_limit, int maxiter, double factr, double pgtol,
                                          const double *ics, const double *box_lo, const double *box_hi, double *cand,
                                          double *fval, int *status, int *nit, int *nfev, long long *prof,
                                          double *workspace, long long workspace_doubles, void *stream) {
  if (!m || num_restarts <= 0 || batch_limit <= 0 || maxiter <= 0 || !ics || !box_lo || !box_hi || !cand || !fval ||
      !status || !nit || !nfev || m->B <= 0 || m->n < 1 || m->ld_n < m->n || !m->norm_off ||
      !m->norm_coef || !m->ZnT || !m->alpha || !m->Wp || (m->wp_stride & 1) ||
      m->wp_stride < lbeval::packed_doubles(m->n) || !m->Zn || !m->ymean || !m->ystd || !m->best_f ||
      (!m->lengthscales && !(m->lengthscale > 0.0)))
    return PCABO_E_INVALID;
  if (run_idx ? (nsel <= 0 || nsel > m->B) : (nsel != m->B)) return PCABO_E_INVALID;
  if (k_bound <= 0 || k_bound > m->kmax) k_bound = m->kmax;
  if (m->kmax > 104 || m->n > 1088 || batch_limit > 5) return PCABO_E_UNSUPPORTED;
  LbArgs a;
  a.m = *m; a.num_restarts = num_restarts; a.batch_limit = batch_limit; a.maxiter = maxiter;
  a.factr = factr; a.pgtol = pgtol; a.ics = ics; a.box_lo = box_lo; a.box_hi = box_hi;
  a.cand = cand; a.fval = fval; a.status = status; a.nit = nit; a.nfev = nfev; a.prof = prof;
  a.run_idx = run_idx; a.k_bound = k_bound;
  return launch(a, nsel, workspace, workspace_doubles, as_stream(stream));
}

extern "C" int pcabo_lbfgsb_logei_runs(const pcabo_gp_model *m, int nsel, const int *run_idx, int k_bound,
                                       int num_restarts, int batch_limit, int maxiter, double factr, double pgtol,
                                       const double *ics, const double *box_lo, const double *box_hi, double *cand,
                                       double *fval, int *status, int *nit, int *nfev, long long *prof, void *stream) {
  return pcabo_lbfgsb_logei_runs_ws(m, nsel, run_idx, k_bound, num_restarts, batch_limit, maxiter, factr, pgtol, ics,
                                    box_lo, box_hi, cand, fval, status, nit, nfev, prof, nullptr, 0, stream);
}

extern "C" int pcabo_lbfgsb_logei_prof(const pcabo_gp_model *m, int num_restarts, int batch_limit, int maxiter,
                                       double factr, double pgtol, const double *ics, const double *box_lo,
                                       const double *box_hi, double *cand, double *fval, int *status, int *nit,
                                       int *nfev, long long *prof, void *stream) {
  if (!m) return PCABO_E_INVALID;
  return pcabo_lbfgsb_logei_runs(m, m->B, nullptr, m->kmax, num_restarts, batch_limit, maxiter, factr, pgtol, ics,
                                 box_lo, box_hi, cand, fval, status, nit, nfev, prof, stream);
}

extern "C" int pcabo_lbfgsb_logei(const pcabo_gp_model *m, int num_restarts, int batch_limit, int maxiter, double factr,
                                  double pgtol, const double *ics, const double *box_lo, const double *box_hi,
                                  double *cand, double *fval, int *status, int *nit, int *nfev, void *stream) {
  return pcabo_lbfgsb_logei_prof(m, num_restarts, batch_limit, maxiter, factr, pgtol, ics, box_lo, box_hi, cand, fval,
                                 status, nit, nfev, nullptr, stream);
}
