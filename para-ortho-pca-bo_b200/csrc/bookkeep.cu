// Device-side bookkeeping of a BO iteration, so that the host never has to wait for the tail of optimize_acqf:
//   pick_best_kernel           torch.argmax over the restarts' final values + the winning candidate (optimize_acqf's return,
//                              PCA_BO.py:607-614 / 618-619)
//   append_observation_kernel  the out-of-bounds penalty, the append to the history and the best-so-far update of
//                              PCA_BO.__call__ (PCA_BO.py:239-281; AbstractBayesianOptimizer.assign_new_best :196-208),
//                              and best_f as botorch will store it in the next iteration (float32, PCA_BO.py:199-203)
#include "common.cuh"

namespace {

__global__ void pick_best_kernel(int R, int D, const double *vals, const double *cand, double *z_new, double *best_val,
                                 int *best_idx) {
  const int b = blockIdx.x;
  __shared__ int s_idx;
  if (threadIdx.x == 0) {
    const double *v = vals + (size_t)b * R;
    // first occurrence of the maximum (torch.argmax), NaNs skipped like pcabo_argmax does
    int bi = -1;
    double bv = -INFINITY;
    for (int r = 0; r < R; ++r) {
      const double x = v[r];
      if (x > bv || (bi < 0 && x == bv)) { bv = x; bi = r; }
    }
    if (bi < 0) { bi = 0; bv = v[0]; }
    s_idx = bi;
    best_val[b] = bv;
    if (best_idx) best_idx[b] = bi;
  }
  __syncthreads();
  const double *src = cand + ((size_t)b * R + s_idx) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) z_new[(size_t)b * D + d] = src[d];
}

__global__ void append_observation_kernel(int D, int ld_n, int n, const double *x_new, const double *z_new,
                                          const double *f_new, const int *oob, const double *acq, double penalty,
                                          int maximization, double *X, double *f, double *best, double *best_f,
                                          double *rec) {
  const int b = blockIdx.x;
  const bool out = oob != nullptr && oob[b] != 0;
  const double fv = out ? penalty : f_new[b];   // PCA_BO.py:259-263: an outside point is not evaluated
  const int rs = 2 * D + 3;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const double xv = x_new[(size_t)b * D + d];
    X[((size_t)b * ld_n + n) * D + d] = xv;
    if (rec) {
      rec[(size_t)b * rs + d] = xv;
      rec[(size_t)b * rs + D + d] = z_new ? z_new[(size_t)b * D + d] : xv;
    }
  }
  if (threadIdx.x == 0) {
    f[(size_t)b * ld_n + n] = fv;
    const double cur = best[b];
    const double nb = maximization ? (fv > cur ? fv : cur) : (fv < cur ? fv : cur);
    best[b] = nb;
    best_f[b] = (double)(float)nb;
    if (rec) {
      rec[(size_t)b * rs + 2 * D] = fv;
      rec[(size_t)b * rs + 2 * D + 1] = out ? 1.0 : 0.0;
      rec[(size_t)b * rs + 2 * D + 2] = acq ? acq[b] : 0.0;
    }
  }
}

}  // namespace

extern "C" int pcabo_pick_best(int B, int R, int D, const double *vals, const double *cand, double *z_new,
                               double *best_val, int *best_idx, void *stream) {
  if (B <= 0 || R <= 0 || D <= 0 || !vals || !cand || !z_new || !best_val) return PCABO_E_INVALID;
  pick_best_kernel<<<B, 64, 0, as_stream(stream)>>>(R, D, vals, cand, z_new, best_val, best_idx);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}

extern "C" int pcabo_append_observation(int B, int D, int ld_n, int n, const double *x_new, const double *z_new,
                                        const double *f_new, const int *oob, const double *acq, double penalty,
                                        int maximization, double *X, double *f, double *best, double *best_f,
                                        double *rec, void *stream) {
  if (B <= 0 || D <= 0 || n < 0 || n >= ld_n || !x_new || !f_new || !X || !f || !best || !best_f) return PCABO_E_INVALID;
  append_observation_kernel<<<B, 64, 0, as_stream(stream)>>>(D, ld_n, n, x_new, z_new, f_new, oob, acq, penalty,
                                                               maximization, X, f, best, best_f, rec);
  PCABO_LAUNCH_CHECK();
  return PCABO_OK;
}
