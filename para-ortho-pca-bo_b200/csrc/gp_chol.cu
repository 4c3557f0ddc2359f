// Batched Cholesky L L^T = K, triangular inverse W = L^-1, alpha = W^T (W y) and W^T for the GP caches of
// PCA_BO._initialize_model (PCA_BO.py:502-545; gpytorch builds chol(K + sigma^2 I), K^-1 y and L^-T on the first posterior
// call), as a pipeline of tile kernels on the FP64 tensor cores (mma.sync m8n8k4, SASS DMMA.8x8x4).
//
// Round 1 ran the whole factorisation of a run in ONE thread block (90 of 148 SMs busy, scalar FMAs, 255 registers and
// 4.5 KB of local memory per thread: 3.7 ms for 90 x 450, 40 ms for 150 x 1050, 5 % of the DGEMM peak and slower than
// cuSOLVER's potrfBatched). Here every 64 x 64 tile of a block column is a thread block of its own:
//
//   for each block column j (left-looking):
//     chol_update_kernel  grid (tiles i >= j, runs): C = K[i][j] - sum_{k<j} L[i][k] L[j][k]^T  (64 x 64 x 64 DMMA steps,
//                         operands staged in shared memory, next step's tiles prefetched into registers); the diagonal
//                         block (i == j) is then factored in shared memory (right-looking, one barrier per column) and
//                         inverted, L[j][j] and W[j][j] = L[j][j]^-1 written
//     chol_solve_kernel   grid (tiles i > j, runs): L[i][j] = C[i][j] W[j][j]^T
//   tri_inverse_kernel    grid (block columns J, runs): W[i][J] = -W[i][i] sum_{k=J}^{i-1} L[i][k] W[k][J], i = J+1, ...
//                         (the block columns of L^-1 are independent of each other)
//   alpha_u / alpha_wt    grid (runs, 64-row blocks): u = W y_s, alpha = W^T u
//   transpose_kernel      grid (lower tiles, runs): WT = W^T
#include "async_copy.cuh"
#include "common.cuh"

namespace {

constexpr int T = 64;        // tile
constexpr int TLD = 68;      // shared-memory row stride: a multiple of 4, congruent 4 mod 16 -> both MMA operand patterns
                             // (8 rows x 4 columns, 4 rows x 8 columns) are bank-conflict free
constexpr int NTH = 256;
constexpr int PER = T * T / NTH;   // tile elements per thread

struct Tile { double v[PER]; };

// rows r0.., columns c0.. of a row-major matrix; entries with row >= n or column >= n read as zero
__device__ __forceinline__ void tile_load(Tile &t, const double *M, long ld, int r0, int c0, int n) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int idx = tid + q * NTH, r = idx >> 6, c = idx & 63;
    t.v[q] = (r0 + r < n && c0 + c < n) ? M[(long)(r0 + r) * ld + c0 + c] : 0.0;
  }
}
__device__ __forceinline__ void tile_store_smem(const Tile &t, double *S) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int idx = tid + q * NTH;
    S[(idx >> 6) * TLD + (idx & 63)] = t.v[q];
  }
}

// acc (8 rows of warp w x 64 columns, as 8 accumulator tiles) += A[64 x 64] * op(B):
//   NT: op(B)[k][c] = Bs[c][k] (both operands row-major with the contraction index fastest)
//   NN: op(B)[k][c] = Bs[k][c]
template <bool NT>
__device__ __forceinline__ void tile_mma(double (&acc)[8][2], const double *As, const double *Bs, int warp, int lane) {
  const int lr = lane >> 2, lc = lane & 3;
  const double *ap = As + (8 * warp + lr) * TLD + lc;
#pragma unroll 4
  for (int ks = 0; ks < T / 4; ++ks) {
    const double a = ap[4 * ks];
    double b[8];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) b[nt] = NT ? Bs[(8 * nt + lr) * TLD + 4 * ks + lc] : Bs[(4 * ks + lc) * TLD + 8 * nt + lr];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) acp::dmma_8x8x4(acc[nt][0], acc[nt][1], a, b[nt]);
  }
}

// Factorisation D = L L^T of the 64 x 64 diagonal block in shared memory (lower triangle of D, row stride TLD) and
// X = L^-1, by the 256 threads of the block. s_rdiag: 64 doubles of shared scratch; *s_info: first failed pivot (+ base).
// All threads call; ends with a block barrier; `lane` = threadIdx.x & 31.
__device__ __forceinline__ void diag_factor_invert(double *D, double *X, double *s_rdiag, int *s_info, int base) {
  const int tid = threadIdx.x, lane = tid & 31;
  // right-looking on the UNSCALED columns: d_rs -= d_rc d_sc / d_cc for c < s <= r; one barrier per column
  {
    const int tr = tid >> 4, tc = tid & 15;
    for (int c = 0; c < T - 1; ++c) {
      double d = D[c * TLD + c];
      if (!(d > 0.0)) {
        if (tid == 0 && *s_info == 0) *s_info = base + c + 1;
        d = 1e-300;
      }
      const double dinv = 1.0 / d;
      double colr[4], cols[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int r = c + 1 + tr + 16 * a, s = c + 1 + tc + 16 * a;
        colr[a] = (r < T) ? D[r * TLD + c] * dinv : 0.0;
        cols[a] = (s < T) ? D[s * TLD + c] : 0.0;
      }
      // (loading all 16 entries up front and storing conditionally -- branch free -- measured slower: 58.7 k against 55 k
      // cycles for the whole block, tools/micro/chol_diag.cu)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          const int r = c + 1 + tr + 16 * a, s = c + 1 + tc + 16 * bb;
          if (r < T && s <= r) D[r * TLD + s] = fma(-colr[a], cols[bb], D[r * TLD + s]);
        }
      __syncthreads();
    }
    {
      double d = D[(T - 1) * TLD + T - 1];
      if (!(d > 0.0) && tid == 0 && *s_info == 0) *s_info = base + T;
    }
  }
  // scale the columns: L_rc = d_rc / sqrt(d_cc) (a failed pivot is replaced by 1e-300 like the flag says)
  {
    const int c = tid & 63;
    double d = D[c * TLD + c];
    if (!(d > 0.0)) d = 1e-300;
    const double rs = rsqrt(d);
    __syncthreads();
    for (int r = tid >> 6; r < T; r += NTH / 64)
      if (r >= c) D[r * TLD + c] = (r == c) ? d * rs : D[r * TLD + c] * rs;
  }
  __syncthreads();
  // X = D^-1 by column sweeps, four threads per column: thread (c, q) keeps the entries b_r, r = 4 m + q, of the
  // right-hand side e_c in registers; step k: x_k = b_k / d_kk travels from its owner by a shuffle, b_r -= d_rk x_k for
  // r > k (independent FMAs). Fully unrolled: the serial chain per step is one shuffle and one multiplication.
  if (tid < T) s_rdiag[tid] = 1.0 / D[tid * TLD + tid];
  __syncthreads();
  {
    const int c = tid >> 2, q = tid & 3;
    double bv[T / 4];
#pragma unroll
    for (int m = 0; m < T / 4; ++m) bv[m] = (4 * m + q == c) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < T; ++k) {
      const int mk = k >> 2, qk = k & 3;
      const double xk = __shfl_sync(0xffffffffu, bv[mk], (lane & ~3) | qk) * s_rdiag[k];
      if (q == qk) bv[mk] = xk;
#pragma unroll
      for (int m = mk; m < T / 4; ++m) {
        const int r = 4 * m + q;
        if (m > mk || q > qk) bv[m] = fma(-D[r * TLD + k], xk, bv[m]);
      }
    }
#pragma unroll
    for (int m = 0; m < T / 4; ++m) X[(4 * m + q) * TLD + c] = bv[m];
  }
  __syncthreads();
}

struct CholP {
  int n, ld, j;
  double *L, *W;
  int *info;
};

// ---- block column j: update of the tiles (i, j), i >= j; factorisation + inverse of the diagonal block
__global__ void __launch_bounds__(NTH, 2) chol_update_kernel(CholP p) {
  extern __shared__ __align__(16) double sm[];
  double *As = sm, *Bs = sm + T * TLD;      // later: D = As (factor), X = Bs (inverse)
  // (the run index is the fastest grid dimension: the long diagonal-block CTAs (blockIdx.y = 0) of all runs are
  // scheduled first and spread one per SM, instead of sharing an SM with each other behind the short off-diagonal ones)
  const int b = blockIdx.x, j = p.j, i = j + blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const int n = p.n;
  const long ld = p.ld;
  double *L = p.L + (size_t)b * ld * ld;
  double acc[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = 0.0; acc[nt][1] = 0.0; }
  if (j > 0) {
    Tile ta, tb;
    tile_load(ta, L, ld, i * T, 0, n);
    tile_load(tb, L, ld, j * T, 0, n);
    for (int k = 0; k < j; ++k) {
      tile_store_smem(ta, As);
      tile_store_smem(tb, Bs);
      __syncthreads();
      if (k + 1 < j) {   // the next step's tiles travel while this step multiplies
        tile_load(ta, L, ld, i * T, (k + 1) * T, n);
        tile_load(tb, L, ld, j * T, (k + 1) * T, n);
      }
      tile_mma<true>(acc, As, Bs, warp, lane);
      __syncthreads();
    }
  }
  // C = K[i][j] - acc, in the accumulator layout: lane holds row 8 warp + lr, columns 8 nt + 2 lc + {0, 1}
  const int gr = i * T + 8 * warp + lr;
  if (i != j) {
    if (gr < n) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int gc = j * T + 8 * nt + 2 * lc;
        double *q = L + (long)gr * ld + gc;
        if (gc < n) q[0] -= acc[nt][0];
        if (gc + 1 < n) q[1] -= acc[nt][1];
      }
    }
    return;
  }
  // ---- diagonal block: D = lower part of C (identity outside the matrix), factor, invert
  double *D = As, *X = Bs;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = 8 * warp + lr, c = 8 * nt + 2 * lc + h;
      const int gc = j * T + c;
      double v = (r == c) ? 1.0 : 0.0;
      if (gr < n && gc < n && c <= r) v = L[(long)gr * ld + gc] - acc[nt][h];
      D[r * TLD + c] = (c <= r) ? v : 0.0;
    }
  __shared__ int s_info;
  if (tid == 0) s_info = 0;
  __syncthreads();
  __shared__ double s_rdiag[T];
  diag_factor_invert(D, X, s_rdiag, &s_info, j * T);
  double *W = p.W + (size_t)b * ld * ld;
  for (int idx = tid; idx < T * T; idx += NTH) {
    const int r = idx >> 6, c = idx & 63, g_r = j * T + r, g_c = j * T + c;
    if (g_r < n && g_c < n) {
      L[(long)g_r * ld + g_c] = (c <= r) ? D[r * TLD + c] : 0.0;
      W[(long)g_r * ld + g_c] = (c <= r) ? X[r * TLD + c] : 0.0;
    }
  }
  if (tid == 0 && s_info != 0) atomicCAS(&p.info[b], 0, s_info);
}

// ---- L[i][j] = C[i][j] W[j][j]^T for the tiles below the diagonal block
__global__ void __launch_bounds__(NTH, 2) chol_solve_kernel(CholP p) {
  extern __shared__ __align__(16) double sm[];
  double *As = sm, *Bs = sm + T * TLD;
  const int b = blockIdx.x, j = p.j, i = j + 1 + blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const int n = p.n;
  const long ld = p.ld;
  double *L = p.L + (size_t)b * ld * ld;
  const double *W = p.W + (size_t)b * ld * ld;
  Tile ta, tb;
  tile_load(ta, L, ld, i * T, j * T, n);
  tile_load(tb, W, ld, j * T, j * T, n);
  tile_store_smem(ta, As);
  tile_store_smem(tb, Bs);
  __syncthreads();
  double acc[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = 0.0; acc[nt][1] = 0.0; }
  tile_mma<true>(acc, As, Bs, warp, lane);
  const int gr = i * T + 8 * warp + lr;
  if (gr < n) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int gc = j * T + 8 * nt + 2 * lc;
      double *q = L + (long)gr * ld + gc;
      if (gc < n) q[0] = acc[nt][0];
      if (gc + 1 < n) q[1] = acc[nt][1];
    }
  }
}

// ---- block column J of W = L^-1 below its diagonal block
__global__ void __launch_bounds__(NTH, 2) tri_inverse_kernel(int n, int ld_n, const double *Lg, double *Wg) {
  extern __shared__ __align__(16) double sm[];
  double *As = sm, *Bs = sm + T * TLD;
  const int b = blockIdx.x, J = blockIdx.y;   // (J = 0, the longest block column, first)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const long ld = ld_n;
  const double *L = Lg + (size_t)b * ld * ld;
  double *W = Wg + (size_t)b * ld * ld;
  const int nb = (n + T - 1) / T;
  for (int i = J + 1; i < nb; ++i) {
    double acc[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = 0.0; acc[nt][1] = 0.0; }
    Tile ta, tb;
    tile_load(ta, L, ld, i * T, J * T, n);
    tile_load(tb, W, ld, J * T, J * T, n);
    for (int k = J; k < i; ++k) {
      tile_store_smem(ta, As);
      tile_store_smem(tb, Bs);
      __syncthreads();
      if (k + 1 < i) {
        tile_load(ta, L, ld, i * T, (k + 1) * T, n);
        tile_load(tb, W, ld, (k + 1) * T, J * T, n);
      } else {
        tile_load(ta, W, ld, i * T, i * T, n);     // W[i][i] for the closing product
      }
      tile_mma<false>(acc, As, Bs, warp, lane);
      __syncthreads();
    }
    // Bs <- S = sum_k L[i][k] W[k][J]; As <- W[i][i]; W[i][J] = -W[i][i] S
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      *reinterpret_cast<double2 *>(Bs + (8 * warp + lr) * TLD + 8 * nt + 2 * lc) = make_double2(acc[nt][0], acc[nt][1]);
    tile_store_smem(ta, As);
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = 0.0; acc[nt][1] = 0.0; }
    tile_mma<false>(acc, As, Bs, warp, lane);
    const int gr = i * T + 8 * warp + lr;
    if (gr < n) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int gc = J * T + 8 * nt + 2 * lc;
        double *q = W + (long)gr * ld + gc;
        if (gc < n) q[0] = -acc[nt][0];
        if (gc + 1 < n) q[1] = -acc[nt][1];
      }
    }
    __syncthreads();   // the tile just written is an operand of the next block row (same block: visible after the barrier)
  }
}

// ---- alpha = W^T (W y_s) in two kernels of (64-row / 64-column blocks) x runs (one block per run took 225 us at
//      90 x 450: a chain of L2 round trips; the 8 x 90 blocks of each kernel overlap theirs).
//      u = W y_s: warp per row, lanes over the columns j <= i; y_s is read from `alpha`, u goes to `u_out`.
__global__ void __launch_bounds__(NTH) alpha_u_kernel(int n, int ld_n, const double *Wg, const double *alpha, double *u_out) {
  const int b = blockIdx.x, i0 = blockIdx.y * T, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long ld = ld_n;
  const double *W = Wg + (size_t)b * ld * ld;
  const double *y = alpha + (size_t)b * ld;
  for (int i = i0 + warp; i < min(i0 + T, n); i += NTH / 32) {
    const double *row = W + (long)i * ld;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int jj = lane;
    for (; jj + 96 <= i; jj += 128) {
      const double w0 = row[jj], w1 = row[jj + 32], w2 = row[jj + 64], w3 = row[jj + 96];
      a0 = fma(w0, y[jj], a0); a1 = fma(w1, y[jj + 32], a1); a2 = fma(w2, y[jj + 64], a2); a3 = fma(w3, y[jj + 96], a3);
    }
    for (; jj <= i; jj += 32) a0 = fma(row[jj], y[jj], a0);
    const double s = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) u_out[(size_t)b * ld * ld + i] = s;   // first row of this run's scratch matrix
  }
}

//      alpha = W^T u: block = 64 columns; thread (q, c): column j0 + c, rows i = j0 + q, j0 + q + 4, ... (coalesced over
//      c), eight loads in flight; the four row classes are combined through shared memory.
__global__ void __launch_bounds__(NTH) alpha_wt_kernel(int n, int ld_n, const double *Wg, const double *u_in, double *alpha) {
  __shared__ double red[4][T];
  const int b = blockIdx.x, j0 = blockIdx.y * T, tid = threadIdx.x, c = tid & 63, q = tid >> 6;
  const long ld = ld_n;
  const double *W = Wg + (size_t)b * ld * ld;
  const double *u = u_in + (size_t)b * ld * ld;
  const int jj = j0 + c;
  double a0 = 0.0, a1 = 0.0;
  if (jj < n) {
    int i = j0 + q;     // rows above the diagonal of this column hold zeros inside the 64 x 64 diagonal tile
    for (; i + 28 < n; i += 32) {
      double w8[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) w8[t] = W[(long)(i + 4 * t) * ld + jj];
#pragma unroll
      for (int t = 0; t < 8; t += 2) { a0 = fma(w8[t], u[i + 4 * t], a0); a1 = fma(w8[t + 1], u[i + 4 * t + 4], a1); }
    }
    for (; i < n; i += 4) a0 = fma(W[(long)i * ld + jj], u[i], a0);
  }
  red[q][c] = a0 + a1;
  __syncthreads();
  if (q == 0 && jj < n) alpha[(size_t)b * ld + jj] = (red[0][c] + red[1][c]) + (red[2][c] + red[3][c]);
}

// ---- WT = W^T by 64 x 64 tiles (lower tiles of W -> upper tiles of WT); in a diagonal tile both triangles are written
//      (zeros opposite): the evaluation kernels read the 64 x 64 diagonal blocks of W and WT whole.
__global__ void __launch_bounds__(NTH) transpose_kernel(int n, int ld_n, double *Wg, double *WTg) {
  __shared__ double tl[T][T + 1];
  const int b = blockIdx.y;
  int t = blockIdx.x, ti = 0;
  while (t >= ti + 1) { t -= ti + 1; ++ti; }
  const int tj = t;
  const long ld = ld_n;
  double *W = Wg + (size_t)b * ld * ld;
  double *WT = WTg + (size_t)b * ld * ld;
  for (int idx = threadIdx.x; idx < T * T; idx += NTH) {
    const int r = idx >> 6, c = idx & 63, i = ti * T + r, jj = tj * T + c;
    tl[r][c] = (i < n && jj <= i) ? W[(long)i * ld + jj] : 0.0;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < T * T; idx += NTH) {
    const int r = idx >> 6, c = idx & 63;
    const int jj = tj * T + r, i = ti * T + c;   // WT[jj][i] = W[i][jj]
    if (i < n && jj < n) WT[(long)jj * ld + i] = tl[c][r];
    if (ti == tj) {
      const int gi = ti * T + r, gj = tj * T + c;
      if (gi < n && gj < n && gj > gi) W[(long)gi * ld + gj] = 0.0;   // strict upper part of W's diagonal tile
    }
  }
}

}  // namespace

// Called by pcabo_gp_fit after the kernel matrix (lower triangle of K + noise I) has been written into L.
int pcabo_chol_pipeline(int B, int n, int ld_n, double *L, double *W, double *WT, double *alpha, int *info, cudaStream_t st) {
  const int nb = (n + T - 1) / T;
  const size_t smem = (size_t)2 * T * TLD * sizeof(double);
  PCABO_CUDA_CHECK(cudaFuncSetAttribute(chol_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PCABO_CUDA_CHECK(cudaFuncSetAttribute(chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PCABO_CUDA_CHECK(cudaFuncSetAttribute(tri_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PCABO_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int) * B, st));
  CholP p;
  p.n = n; p.ld = ld_n; p.L = L; p.W = W; p.info = info;
  for (int j = 0; j < nb; ++j) {
    p.j = j;
    chol_update_kernel<<<dim3(B, j == 0 ? 1 : nb - j), NTH, smem, st>>>(p);   // (j = 0: only the diagonal block changes)
    PCABO_LAUNCH_CHECK();
    if (j + 1 < nb) {
      chol_solve_kernel<<<dim3(B, nb - j - 1), NTH, smem, st>>>(p);
      PCABO_LAUNCH_CHECK();
    }
  }
  if (nb > 1) {
    tri_inverse_kernel<<<dim3(B, nb - 1), NTH, smem, st>>>(n, ld_n, L, W);
    PCABO_LAUNCH_CHECK();
  }
  // u = W y_s lands in the first row of WT of each run (WT is only written by the transpose below), alpha = W^T u
  alpha_u_kernel<<<dim3(B, nb), NTH, 0, st>>>(n, ld_n, W, alpha, WT);
  PCABO_LAUNCH_CHECK();
  alpha_wt_kernel<<<dim3(B, nb), NTH, 0, st>>>(n, ld_n, W, WT, alpha);
  PCABO_LAUNCH_CHECK();
  transpose_kernel<<<dim3(nb * (nb + 1) / 2, B), NTH, 0, st>>>(n, ld_n, W, WT);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}
