/*
 * pcabo_b200.h -- C ABI of the B200-native PCA-BO hot path (libpcabo_b200.so).
 *
 * Drop-in boundary for the inner loop of IvanBanny/para-ortho-pca-bo. Every entry point replaces one
 * Python-level operator seam of the reference (citations are relative to the reference repo root):
 *
 *   pcabo_weights_pca_fit   <- PCA_BO._calculate_weights                    Algorithms/BayesianOptimization/PCA_BO.py:316-341
 *                              PCA_BO._transform_points_to_reduced_space   PCA_BO.py:343-408  (sklearn PCA().fit/.transform)
 *   pcabo_inverse_map       <- PCA_BO._transform_point_to_original_space   PCA_BO.py:410-434  (+ the bounds test PCA_BO.py:248)
 *   pcabo_gp_fit            <- PCA_BO._initialize_model                    PCA_BO.py:502-545  (SingleTaskGP + Normalize + Standardize,
 *                              Vanilla_BO._initialise_model                Vanilla_BO.py:166-195; prediction caches of gpytorch)
 *                              and the search box of optimize_acqf_and_get_observation, PCA_BO.py:558-573
 *   pcabo_gp_mll            <- (north_star only; the reference never fits the GP)
 *   pcabo_logei             <- LogExpectedImprovement(model, best_f, maximize)(X)   PCA_BO.py:199-203, class map :679-686
 *   pcabo_argmax            <- torch.argmax over restarts / raw samples inside botorch.optimize_acqf, PCA_BO.py:607-614
 *   pcabo_lbfgsb_logei      <- botorch gen_candidates_scipy (scipy L-BFGS-B, maxiter 200) inside optimize_acqf, PCA_BO.py:607-614
 *   pcabo_ei_sweep          <- the raw-sample scoring pass of optimize_acqf scaled to 2^20..2^26 candidates (BASELINE config 4)
 *
 * Conventions
 *   - All array arguments are DEVICE pointers to row-major FP64 (or int32) unless a name ends in _host.
 *   - The batch dimension B (independent BO runs stepping in lock-step) is outermost. All runs of a batch share
 *     n (points evaluated so far) and D; the reduced dimension k[b] differs per run.
 *   - `ld_n` is the row capacity (leading dimension in points) of per-run arrays, ld_n >= n.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream). Calls are asynchronous.
 *   - Return value: 0 on success, a negative PCABO_E_* code otherwise. No exceptions cross the ABI; the callee
 *     allocates nothing that outlives the call.
 */
#ifndef PCABO_B200_H
#define PCABO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PCABO_ABI_VERSION 2   /* 2: pcabo_gp_model gained Wp / wp_stride (packed W for the bulk-copy engine) and lengthscales */

#define PCABO_OK 0
#define PCABO_E_INVALID (-1)      /* bad argument (null pointer, n < 2, ...) */
#define PCABO_E_UNSUPPORTED (-2)  /* size outside what the kernels support (D > 104, n > 2048, k > 104) */
#define PCABO_ACQ_LOG_EI 0
#define PCABO_ACQ_PI 1
#define PCABO_E_CUDA (-3)         /* a CUDA runtime call failed; see pcabo_last_cuda_error() */

int pcabo_abi_version(void);
const char *pcabo_error_string(int code);
const char *pcabo_last_cuda_error(void);

/* Fitted GP in eval mode (what botorch keeps inside SingleTaskGP + its prediction caches). */
typedef struct pcabo_gp_model {
  int B;            /* runs in the batch */
  int n;            /* training points per run */
  int kmax;         /* column capacity of per-run z arrays (= D) */
  int ld_n;         /* row capacity */
  int maximize;     /* acquisition(maximize=...) */
  int acq;          /* PCABO_ACQ_LOG_EI (0): botorch LogExpectedImprovement, the reference's "expected_improvement"
                       (PCA_BO.py:681-682); PCABO_ACQ_PI (1): ProbabilityOfImprovement (PCA_BO.py:683-684) */
  double lengthscale;       /* Matern-5/2 lengthscale (reference: softplus(0) = ln 2) */
  const int *k;             /* [B] active dimensions per run */
  const double *norm_off;   /* [B][kmax]  Normalize offset  (lo)        */
  const double *norm_coef;  /* [B][kmax]  Normalize coefficient (hi-lo) */
  const double *ZnT;        /* [B][kmax][ld_n] normalised training inputs, transposed */
  const double *alpha;      /* [B][ld_n]  K^-1 y_s */
  const double *W;          /* [B][ld_n][ld_n] L^-1, lower triangular, row-major */
  const double *ymean;      /* [B] Standardize mean */
  const double *ystd;       /* [B] Standardize std  */
  const double *best_f;     /* [B] incumbent, already rounded to float32 as botorch stores it */
  const double *WT;         /* [B][ld_n][ld_n] (L^-1)^T, upper triangular, row-major (column j of W contiguous) */
  const double *Zn;         /* [B][ld_n][kmax] normalised training inputs, row-major */
  const double *Wp;         /* [B][wp_stride] W packed in 8-row groups by pcabo_gp_pack_w (pcabo_lbfgsb_logei* only) */
  long long wp_stride;      /* doubles between the packed W of consecutive runs (even, >= pcabo_gp_packed_doubles(n)) */
  const double *lengthscales; /* [B] per-run Matern lengthscale (opt-in hyper-parameter fit) or NULL: `lengthscale` for all */
} pcabo_gp_model;

/* Rank weights + weighted PCA + projection, one CTA per run.
 *   X      [B][ld_n][D]   evaluated points            f [B][ld_n] objective values
 *   noise  [B][n][D]      the N(0,1e-8) matrix of PCA_BO.py:376 (host RNG replay)
 *   logtab [>= n+1]       logtab[i] = numpy.log(i) (host table, keeps the weights bit-identical to numpy)
 * outputs
 *   w [B][ld_n], ranks [B][ld_n] (1 = best, ties by index), xbar [B][D] (data_mean), m2 [B][D] (pca.mean_),
 *   V [B][D][D] (components_, rows, svd_flip'ed), lam [B][D] (explained_variance_), evr [B][D] (ratio),
 *   k [B] (reduced_space_dim_num), Z [B][ld_n][D] (first k[b] columns valid, rest zero)                    */
int pcabo_weights_pca_fit(int B, int n, int D, int ld_n,
                          const double *X, const double *f, const double *noise, const double *logtab,
                          double var_threshold, int n_components, int maximization,
                          double *w, int *ranks, double *xbar, double *m2, double *V, double *lam, double *evr,
                          int *k, double *Z, void *stream);

/* x = z V_k + m2 + xbar and the bounds test. z [B][kmax=D], x [B][D], oob [B] (1 = outside [lo, hi]). */
int pcabo_inverse_map(int B, int D, const double *z, const int *k, const double *V, const double *m2,
                      const double *xbar, const double *lo, const double *hi, double *x, int *oob, void *stream);

/* GP construction + caches.
 *   Z [B][ld_n][kmax], k [B] (NULL: all kmax), f [B][ld_n]
 *   mode 0 (PCA_BO): Normalize bounds [zmin-0.1r, zmax+0.1r]; search box [zmin-0.5r, zmax+0.5r], min width 0.1
 *   mode 1 (Vanilla_BO): no input normalisation; search box = (fix_lo, fix_hi) [kmax]
 * outputs
 *   norm_off, norm_coef, box_lo, box_hi [B][kmax]; ymean, ystd [B]; ZnT [B][kmax][ld_n], Zn [B][ld_n][kmax];
 *   L [B][ld_n][ld_n] (Cholesky factor of K + noise I, lower), W [B][ld_n][ld_n] (L^-1), WT (its transpose; inside
 *   the diagonal 64 x 64 blocks the other triangle of W / WT is written as zero: pcabo_logei / pcabo_lbfgsb_logei
 *   read those blocks whole),
 *   alpha [B][ld_n],
 *   info [B] (0 ok, i > 0: non-positive pivot at row i -> caller should add jitter)                            */
int pcabo_gp_fit(int B, int n, int kmax, int ld_n, int mode,
                 const double *Z, const int *k, const double *f, const double *fix_lo, const double *fix_hi,
                 double lengthscale, double noise,
                 double *norm_off, double *norm_coef, double *box_lo, double *box_hi,
                 double *ymean, double *ystd, double *ZnT, double *Zn, double *L, double *W, double *WT,
                 double *alpha, int *info, void *stream);

/* pcabo_gp_fit with a per-run diagonal jitter [B] (NULL: none) added to K + noise I before the factorisation: the retry
 * ladder of gpytorch's psd_safe_cholesky (1e-8, 1e-7, 1e-6; linear_operator, reached from the first posterior call of
 * the model built at PCA_BO.py:535-544) for the runs whose info came back non-zero.                                */
int pcabo_gp_fit_jitter(int B, int n, int kmax, int ld_n, int mode,
                        const double *Z, const int *k, const double *f, const double *fix_lo, const double *fix_hi,
                        double lengthscale, double noise, const double *jitter,
                        double *norm_off, double *norm_coef, double *box_lo, double *box_hi,
                        double *ymean, double *ystd, double *ZnT, double *Zn, double *L, double *W, double *WT,
                        double *alpha, int *info, void *stream);

/* pcabo_gp_fit_jitter with per-run hyper-parameters lengthscales [B] / noises [B] (either may be NULL: the scalar is
 * used). The reference never fits its GP; this is the evaluation step of the opt-in hyper-parameter fit
 * (pcabo_b200/hyperfit.py: maximise (MLL + log p_LogNormal(-4,1)(noise)) / n like botorch's fit_gpytorch_mll would). */
int pcabo_gp_fit_hyper(int B, int n, int kmax, int ld_n, int mode,
                       const double *Z, const int *k, const double *f, const double *fix_lo, const double *fix_hi,
                       double lengthscale, double noise, const double *lengthscales, const double *noises,
                       const double *jitter,
                       double *norm_off, double *norm_coef, double *box_lo, double *box_hi,
                       double *ymean, double *ystd, double *ZnT, double *Zn, double *L, double *W, double *WT,
                       double *alpha, int *info, void *stream);

/* W (from pcabo_gp_fit) repacked for the bulk-copy (TMA) engine: the lower triangle in groups of 8 rows, group g =
 * rows 8g..8g+7 x columns 0..8g+7 (zero right of the diagonal and below row n), row stride 8g + (g odd ? 12 : 20)
 * doubles (a multiple of 4 congruent 4 mod 16: bank-conflict free for both FP64 MMA operand patterns), groups back to
 * back, so that any run of consecutive groups is ONE contiguous cp.async.bulk. pcabo_gp_packed_doubles(n) doubles per
 * run; Wp [B][wp_stride]. Part of _initialize_model's cache construction (PCA_BO.py:502-545) for the device optimiser. */
long long pcabo_gp_packed_doubles(int n);
int pcabo_gp_pack_w(int B, int n, int ld_n, const double *W, double *Wp, long long wp_stride, void *stream);

/* Exact marginal log-likelihood of the standardised targets from the caches: mll [B]. ys = (f-ymean)/ystd. */
int pcabo_gp_mll(int B, int n, int ld_n, const double *f, const double *ymean, const double *ystd,
                 const double *L, const double *alpha, double *mll, void *stream);

/* Gradient of that marginal log-likelihood w.r.t. the Matern lengthscale and the noise variance, from the caches of
 * pcabo_gp_fit: grad [B][2] = (dMLL/dl, dMLL/dsigma^2) = 1/2 alpha^T dK alpha - 1/2 tr(K^-1 dK), K^-1 = W^T W built tile
 * by tile on the fly. The reference never fits its GP, so nothing on the BO path calls this; it is the
 * building block of an opt-in hyper-parameter fit (north_star).                                                   */
int pcabo_gp_mll_grad(int B, int n, int kmax, int ld_n, const int *k, const double *Zn, const double *W,
                      const double *alpha, double lengthscale, double *grad, void *stream);

/* pcabo_gp_mll_grad with per-run lengthscales [B] (NULL: the scalar). */
int pcabo_gp_mll_grad_hyper(int B, int n, int kmax, int ld_n, const int *k, const double *Zn, const double *W,
                            const double *alpha, double lengthscale, const double *lengthscales, double *grad,
                            void *stream);

/* log-EI (and optionally its gradient, posterior mean and sigma) at npts raw-space points per run.
 *   X [B][npts][kmax]; val [B][npts]; grad [B][npts][kmax] or NULL; mu, sigma [B][npts] or NULL.               */
int pcabo_logei(const pcabo_gp_model *m, int npts, const double *X, double *val, double *grad,
                double *mu, double *sigma, void *stream);

/* Row-wise argmax with first-occurrence ties (torch.argmax): v [B][m] -> best_val [B], best_idx [B]. */
int pcabo_argmax(int B, int m, const double *v, double *best_val, int *best_idx, void *stream);

/* Device-resident multi-start L-BFGS-B on -sum(log-EI): one thread block per chunk of `batch_limit` restarts of one
 * run (botorch gen_candidates_scipy semantics: one joint problem per chunk, scipy defaults factr = 1e7, pgtol = 1e-5,
 * m = 10, maxls = 20).  ics [B][num_restarts][kmax] raw-space initial conditions (clamped to the box on entry),
 * box_lo / box_hi [B][kmax].  Outputs: cand [B][num_restarts][kmax] (clamped optimisers), and per chunk
 * (nchunks = ceil(num_restarts / batch_limit)): fval [B][nchunks] (= -sum logEI), status [B][nchunks]
 * (0 |proj g| <= pgtol, 1 relative reduction <= factr * eps, 2 iteration limit, 3 abnormal line-search termination ->
 * botorch would raise OptimizationWarning and retry), nit, nfev [B][nchunks].                                       */
int pcabo_lbfgsb_logei(const pcabo_gp_model *m, int num_restarts, int batch_limit, int maxiter, double factr,
                       double pgtol, const double *ics, const double *box_lo, const double *box_hi,
                       double *cand, double *fval, int *status, int *nit, int *nfev, void *stream);

/* Same as pcabo_lbfgsb_logei with optional per-chunk cycle counters prof [B][nchunks][32] (clock64 ticks of the
 * optimiser warp per phase: 1 cauchy, 2 formk, 3 cmprlb, 4 subsm, 5 line-search setup, 6 step to next point,
 * 7 log-EI evaluation, 8 dcsrch, 9 update, 0 convergence tests, 10-14 sub-phases; phases inside the log-EI
 * evaluation: 16 k vector, 17 sweep over the panels of W (t = W k and s = W^T t), 22 posterior head, 19 gradient
 * coefficients, 20 gradient; 31 scratch); prof may be NULL. Tuning aid.                                              */
int pcabo_lbfgsb_logei_prof(const pcabo_gp_model *m, int num_restarts, int batch_limit, int maxiter, double factr,
                            double pgtol, const double *ics, const double *box_lo, const double *box_hi,
                            double *cand, double *fval, int *status, int *nit, int *nfev, long long *prof,
                            void *stream);

/* General form: optimise only the runs listed in run_idx [nsel] (device ints; NULL = all runs, nsel = B) and size the
 * per-block shared memory for k <= k_bound (0 = kmax; a run with k > k_bound gets status -1 and is left untouched).
 * Outputs are indexed by RUN (not by position in run_idx), so a retry of a subset updates the same arrays in place:
 * this is botorch's retry_on_optimization_warning (optimize_acqf re-optimises only the problem that warned).
 * W is read in its packed form m->Wp (pcabo_gp_pack_w), one bulk copy per panel; m->W / m->WT are not used here.     */
int pcabo_lbfgsb_logei_runs(const pcabo_gp_model *m, int nsel, const int *run_idx, int k_bound, int num_restarts,
                            int batch_limit, int maxiter, double factr, double pgtol, const double *ics,
                            const double *box_lo, const double *box_hi, double *cand, double *fval, int *status,
                            int *nit, int *nfev, long long *prof, void *stream);

/* Same as pcabo_lbfgsb_logei_runs for problems whose L-BFGS-B state does not fit the 227 KB of shared memory of a
 * block next to the evaluation scratch (config C5: n up to 1050, k ~ 90): the BFGS pairs (2 x 10 x batch_limit x
 * k_bound doubles per chunk) then live in `workspace` (device, caller-owned, L2-resident; nothing is allocated by the
 * callee). pcabo_lbfgsb_workspace_doubles returns the number of doubles needed for `nsel` runs (0: shared memory
 * suffices and workspace may be NULL; -1: the problem does not fit either way -> PCABO_E_UNSUPPORTED).              */
long long pcabo_lbfgsb_workspace_doubles(int n, int kmax, int k_bound, int num_restarts, int batch_limit, int nsel);
/* The shared-memory plan the launcher would choose (host only, no device work): plan6 = {panel rows (8 / 16 / 32), ring
 * capacity in doubles, 1 if every panel of W stays resident in the ring, 0 (reserved), 1 if the BFGS pairs live in the
 * workspace, dynamic shared memory in bytes}.                                                                         */
int pcabo_lbfgsb_plan(int n, int kmax, int k_bound, int batch_limit, int have_workspace, int *plan6);
int pcabo_lbfgsb_logei_runs_ws(const pcabo_gp_model *m, int nsel, const int *run_idx, int k_bound, int num_restarts,
                               int batch_limit, int maxiter, double factr, double pgtol, const double *ics,
                               const double *box_lo, const double *box_hi, double *cand, double *fval, int *status,
                               int *nit, int *nfev, long long *prof, double *workspace, long long workspace_doubles,
                               void *stream);

/* log-EI sweep with fused arg-max over `ncand` candidates against run `run` of the model (BASELINE config C4).
 * Candidates: explicit X [ncand][kmax] (raw space), or X = NULL and generated in-kernel as scrambled Sobol points
 * (global indices idx0 .. idx0 + ncand - 1; shift [kmax], sstate [kmax][30] = the engine's shift and SCRAMBLED direction
 * numbers) scaled to [lo, hi] [kmax] exactly like pcabo_sobol_points. vals [ncand] optional (NULL: arg-max only).
 * part_val / part_idx [nparts]: scratch for per-CTA partial maxima (nparts = number of CTAs to use, e.g. SM count).
 * best_val / best_idx (device scalars): maximum log-EI and its GLOBAL candidate index (ties: lowest index).
 * FP64 tensor cores (DMMA.8x8x4); needs n <= 512, even ld_n and kmax.                                                */
int pcabo_ei_sweep(const pcabo_gp_model *m, int run, long long ncand, long long idx0, const double *X,
                   const int *shift, const int *sstate, const double *lo, const double *hi, double *vals,
                   double *part_val, long long *part_idx, int nparts, double *best_val, long long *best_idx,
                   void *stream);

/* The raw-sample scoring pass of optimize_acqf for ALL runs of the batch in one launch (PCA_BO.py:607-614: botorch scores
 * raw_samples = 512 scrambled-Sobol points per optimize_acqf call with LogExpectedImprovement, 5 at a time): the tile
 * pipeline of pcabo_ei_sweep (FP64 tensor cores, DMMA.8x8x4) with one row of `ctas_per_run` CTAs per run (0: chosen so
 * that the launch is one wave of the SMs), values out, no arg-max.  X [B][npts][kmax] raw-space points, val [B][npts].
 * Same limits as pcabo_ei_sweep (log-EI only, n <= 512, even ld_n and kmax); PCABO_E_UNSUPPORTED otherwise -- the
 * caller then scores with pcabo_logei, which has no such limits.                                                     */
int pcabo_logei_raw_runs(const pcabo_gp_model *m, int npts, const double *X, double *val, int ctas_per_run,
                         void *stream);
/* Host only: 1 if pcabo_logei_raw_runs supports a model of this shape (n, kmax, ld_n, acq), else 0. */
int pcabo_logei_raw_supported(int n, int kmax, int ld_n, int acq);

/* Scrambled Sobol raw samples scaled to the search box, generated on the device from host-drawn scramble bits
 * (replaces botorch draw_sobol_samples / torch SobolEngine(k, scramble=True).draw(npts, float64) inside optimize_acqf,
 * PCA_BO.py:607-614; bit-identical to torch).  shift [B][kmax] (30-bit), ltm_dots [B][kmax][30] (rows of the random
 * lower-triangular matrices with unit diagonal, packed MSB-first), state0 [kmax][30] unscrambled direction numbers,
 * box_lo / box_hi [B][kmax] -> X [B][npts][kmax] (columns >= k[b] zero).                                              */
int pcabo_sobol_points(int B, int npts, int kmax, const int *k, const int *shift, const int *ltm_dots,
                       const int *state0, const double *box_lo, const double *box_hi, double *X, void *stream);

/* Device replay of the two torch.randint(2, ...) draws of SobolEngine._scramble (shapes (k, 30) and (k, 30, 30)) from
 * each run's MT19937 state: mt_state [B][624] (uint32 words) and mt_pos [B] (index of the next word, 624 = regenerate
 * first) are advanced in place exactly as torch's CPU generator would be; shift / ltm_dots as in pcabo_sobol_points. */
int pcabo_sobol_scramble_bits(int B, int kmax, const int *k, unsigned *mt_state, int *mt_pos, int *shift,
                              int *ltm_dots, void *stream);

/* out[b][r][:] = X[b][idx[b][r]][:]  (selection of the initial conditions among the raw samples).                    */
int pcabo_gather_rows(int B, int npts, int nsel, int kmax, const double *X, const int *idx, double *out, void *stream);

/* botorch initialize_q_batch (the Boltzmann pick of the num_restarts initial conditions inside optimize_acqf,
 * PCA_BO.py:607-614) on the device, one run per block: weights exp(eta (v - mean) / std) of the raw-sample values
 * vals [B][nraw], torch.multinomial without replacement replayed from each run's torch CPU generator (MT19937 state
 * mt_state [B][624] + position mt_pos [B], advanced in place exactly as torch advances it), arg-max forced in last.
 * Outputs: ics [B][nsel][kmax] = the picked rows of X [B][nraw][kmax], idx [B][nsel], flag [B] (1: degenerate values,
 * nothing drawn and nothing written for that run -- botorch takes a random permutation and warns; host path).   */
int pcabo_select_initial_conditions(int B, int nraw, int nsel, int kmax, double eta, const double *vals,
                                    unsigned *mt_state, int *mt_pos, const double *X, double *ics, int *idx,
                                    int *flag, void *stream);

/* optimize_acqf's return value (PCA_BO.py:607-614, 618-619): torch.argmax over the final log-EI values of the restarts
 * (first occurrence) and the winning candidate.  vals [B][R], cand [B][R][D] -> z_new [B][D], best_val [B], best_idx [B]
 * (may be NULL).                                                                                                    */
int pcabo_pick_best(int B, int R, int D, const double *vals, const double *cand, double *z_new, double *best_val,
                    int *best_idx, void *stream);

/* The bookkeeping tail of PCA_BO.__call__ (PCA_BO.py:239-281) on the device, so that the host never waits for it:
 * f = oob ? penalty : f_new (an outside point is not evaluated, :259-263), X[b][n] = x_new[b], f[b][n] = f,
 * best[b] = min / max(best[b], f) (AbstractBayesianOptimizer.assign_new_best :196-208) and best_f[b] = best[b] rounded
 * to float32 as botorch will store it in the next iteration's acquisition (:199-203).  oob, z_new, acq, rec may be NULL;
 * rec [B][2 D + 3] = (x_new, z_new, f, oob, acq) is the packed record the host reads back asynchronously.          */
int pcabo_append_observation(int B, int D, int ld_n, int n, const double *x_new, const double *z_new,
                             const double *f_new, const int *oob, const double *acq, double penalty, int maximization,
                             double *X, double *f, double *best, double *best_f, double *rec, void *stream);

/* Batched BBOB objectives f15 / f16 / f17 / f20 (the synthetic workloads of BASELINE configs C1-C3), one block per
 * (problem, point): replaces the per-iteration host call problem(x) of the reference (PCA_BO.py:263,
 * AbstractBayesianOptimizer.py:163; ioh's C++ objective) when the runs of a batch are BBOB problems.
 *   fid [B] (15, 16, 17, 20), xopt [B][D], R / M [B][D][D] (first rotation / second linear map), sign [B][D] (f20),
 *   fopt [B], lam [D] = sqrt(10)^(i/(D-1)), f0 = the Weierstrass constant: instance data built on the host like ioh / COCO.
 *   X [B][m][D] (m points per problem) -> out [B][m] = f(x) (f_opt included, like ioh's problem(x)).           */
int pcabo_bbob_eval(int B, int m, int D, const int *fid, const double *xopt, const double *R, const double *M,
                    const double *sign, const double *fopt, const double *lam, double f0, const double *X,
                    double *out, void *stream);

/* pcabo_bbob_eval plus the functions that need per-peak instance data: f21 / f22 (Gallagher's 101 / 21 Gaussian peaks),
 * and f23 Katsuura / f24 Lunacek bi-Rastrigin (which only need xopt, M and, for f24, sign = sign of xopt) -- the set
 * BASELINE config C5 runs (f20-f24). aux [B][aux_stride] per problem: P, heights [P], rotated peak centres [P][D], axis
 * scales [P][D] (zeros for non-Gallagher problems); R = the global rotation of f21 / f22.                          */
int pcabo_bbob_eval_aux(int B, int m, int D, const int *fid, const double *xopt, const double *R, const double *M,
                        const double *sign, const double *fopt, const double *lam, double f0, const double *aux,
                        int aux_stride, const double *X, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PCABO_B200_H */
