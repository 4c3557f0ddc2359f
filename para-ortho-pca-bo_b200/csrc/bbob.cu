// Batched BBOB objectives on the device: f15 Rastrigin (rotated), f16 Weierstrass, f17 Schaffers F7 and
// f20 Schwefel, one thread block per (problem, point). Replaces the per-iteration host call `problem(x)` of the
// reference (ioh's C++ objective, PCA_BO.py:263 / AbstractBayesianOptimizer.py:163) for the synthetic BBOB workloads of
// BASELINE configs C1-C3; the instance data (x_opt, rotations, f_opt) are built on the host exactly like ioh / COCO
// build them (pcabo_b200/bbob.py) and passed in. One launch instead of ~100 small tensor operations per lock-step
// iteration: the host thread that drives the pipelined sub-batches is the bottleneck of the BO loop.
#include "common.cuh"

namespace {

constexpr double TWO_PI = 6.283185307179586476925286766559;

struct BbobArgs {
  int B, m, D;
  const int *fid;          // [B] 15, 16, 17 or 20
  const double *xopt;      // [B][D]
  const double *R;         // [B][D][D] first rotation
  const double *M;         // [B][D][D] second linear map (conditioning and rotation folded in)
  const double *sign;      // [B][D] f20
  const double *fopt;      // [B]
  const double *lam;       // [D] sqrt(10)^(i / (D - 1)) (f20)
  double f0;               // Weierstrass constant
  const double *X;         // [B][m][D]
  double *out;             // [B][m]
};

// COCO's oscillation transformation T_osz, as pcabo_b200/bbob.py::_t_osz evaluates it
__device__ __forceinline__ double t_osz(double x) {
  if (x == 0.0) return 0.0;
  const double t = log(fmax(fabs(x), 1e-300)) / 0.1;
  if (x > 0.0) return pow(exp(t + 0.49 * (sin(t) + sin(0.79 * t))), 0.1);
  return -pow(exp(t + 0.49 * (sin(0.55 * t) + sin(0.31 * t))), 0.1);
}

// COCO's asymmetry transformation T_asy^beta for component i of D
__device__ __forceinline__ double t_asy(double x, double beta, int i, int D) {
  if (!(x > 0.0)) return x;
  const double idx = (D > 1) ? (double)i / (D - 1.0) : 0.0;
  return pow(fmax(x, 1e-300), 1.0 + beta * idx * sqrt(x));
}

__global__ void __launch_bounds__(128) bbob_eval_kernel(BbobArgs a) {
  __shared__ double ys[128], zs[128], red[34];
  const int b = blockIdx.y, p = blockIdx.x, i = threadIdx.x, D = a.D;
  const int fid = a.fid[b];
  const double *x = a.X + ((size_t)b * a.m + p) * D;
  const double *xo = a.xopt + (size_t)b * D;
  const bool on = i < D;
  const double xi = on ? x[i] : 0.0;
  double f;
  if (fid == 20) {
    // Schwefel: x_hat = 2 sign x; z_hat_i = x_hat_i + 0.25 (x_hat_{i-1} - 2 |xopt_{i-1}|); z = 100 (Lambda (z_hat - 2|xopt|) + 2|xopt|)
    const double sg = on ? a.sign[(size_t)b * D + i] : 0.0;
    ys[i] = 2.0 * sg * xi;
    __syncthreads();
    double z = 0.0;
    if (on) {
      const double ax = fabs(xo[i]);
      double zh = ys[i];
      if (i > 0) zh = ys[i] + 0.25 * (ys[i - 1] - 2.0 * fabs(xo[i - 1]));
      z = 100.0 * (a.lam[i] * (zh - 2.0 * ax) + 2.0 * ax);
    }
    const double pe = on ? fmax(fabs(z) - 500.0, 0.0) : 0.0;
    const double pen = block_sum(pe * pe, red);
    const double s = block_sum(on ? z * sin(sqrt(fabs(z))) : 0.0, red);
    f = 0.01 * (pen + 418.9828872724339 - s / D);
  } else {
    const double *Rb = a.R + (size_t)b * D * D, *Mb = a.M + (size_t)b * D * D;
    ys[i] = on ? xi - xo[i] : 0.0;
    const double pe = on ? fmax(fabs(xi) - 5.0, 0.0) : 0.0;
    const double pen = block_sum(pe * pe, red);      // (contains the barrier that publishes ys)
    double z = 0.0;
    if (on)
      for (int j = 0; j < D; ++j) z = fma(Rb[(size_t)i * D + j], ys[j], z);
    double t = 0.0;
    if (on) {
      if (fid == 15) t = t_asy(t_osz(z), 0.2, i, D);
      else if (fid == 16) t = t_osz(z);
      else t = t_asy(z, 0.5, i, D);
    }
    zs[i] = t;
    __syncthreads();
    double z2 = 0.0;
    if (on)
      for (int j = 0; j < D; ++j) z2 = fma(Mb[(size_t)i * D + j], zs[j], z2);
    if (fid == 15) {
      const double c = block_sum(on ? cos(TWO_PI * z2) : 0.0, red);
      const double q = block_sum(on ? z2 * z2 : 0.0, red);
      f = 10.0 * (D - c) + q;
    } else if (fid == 16) {
      double s = 0.0;
      if (on) {
        double ak = 1.0, bk = 1.0;
        for (int k = 0; k < 12; ++k) {
          s += cos(TWO_PI * (z2 + 0.5) * bk) * ak;
          ak *= 0.5;
          bk *= 3.0;
        }
      }
      s = block_sum(s, red);
      const double d = s / D - a.f0;
      f = 10.0 * (d * d * d) + (10.0 / D) * pen;
    } else {
      __syncthreads();
      ys[i] = z2;
      __syncthreads();
      double r = 0.0;
      if (i < D - 1) {
        const double s = ys[i] * ys[i] + ys[i + 1] * ys[i + 1];
        const double sn = sin(50.0 * pow(s, 0.1));
        r = pow(s, 0.25) * (1.0 + sn * sn);
      }
      r = block_sum(r, red);
      const double q = r / (D - 1.0);
      f = q * q + 10.0 * pen;
    }
  }
  if (i == 0) a.out[(size_t)b * a.m + p] = f + a.fopt[b];
}

}  // namespace

extern "C" int pcabo_bbob_eval(int B, int m, int D, const int *fid, const double *xopt, const double *R, const double *M,
                               const double *sign, const double *fopt, const double *lam, double f0, const double *X,
                               double *out, void *stream) {
  if (B <= 0 || m <= 0 || D <= 0 || !fid || !xopt || !R || !M || !sign || !fopt || !lam || !X || !out) return PCABO_E_INVALID;
  if (D > 128 || m > 65535 * 32) return PCABO_E_UNSUPPORTED;
  BbobArgs a;
  a.B = B; a.m = m; a.D = D; a.fid = fid; a.xopt = xopt; a.R = R; a.M = M; a.sign = sign; a.fopt = fopt; a.lam = lam;
  a.f0 = f0; a.X = X; a.out = out;
  dim3 grid(m, B);
  bbob_eval_kernel<<<grid, 128, 0, as_stream(stream)>>>(a);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}
