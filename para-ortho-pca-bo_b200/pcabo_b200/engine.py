"""Lock-step batched PCA-BO / vanilla-BO engine: B independent runs advance one BO iteration per `step()`.

One process drives one GPU. Every numerical stage of the reference's hot loop (PCA_BO.py:178-281) runs in the
sm_100a kernels behind the C ABI; the host keeps only what must stay on the CPU generators for RNG-stream parity
(LHS DoE and the 1e-8 PCA noise from numpy's legacy MT19937, Sobol scrambling and the Boltzmann multinomial from
torch's CPU generator -- one private generator pair per run, seeded exactly like the reference seeds its global
ones in AbstractAlgorithm.py:310-328) and, in `optimizer="scipy"` mode, scipy's own L-BFGS-B driven by reverse
communication with all runs' restart problems sharing one value+gradient kernel launch per round.
"""
import math
import os
from concurrent.futures import ThreadPoolExecutor
from time import perf_counter, sleep

import numpy as np
import torch

from . import ops
from .acq_host import (GeneratorStateIO, LbfgsbProblem, select_initial_conditions,
                       select_initial_conditions_nonneg, sobol_state0)
from .lib import PcaboError

NUM_RESTARTS = 10      # PCA_BO.py:107
RAW_SAMPLES = 512      # PCA_BO.py:108
BATCH_LIMIT = 5        # PCA_BO.py:613
MAXITER = 200          # PCA_BO.py:613


def lhs_center(rs: np.random.RandomState, dim: int, n_samples: int) -> np.ndarray:
    """pyDOE 0.3.8 lhs(criterion="center") on an explicit RandomState (AbstractBayesianOptimizer.py:30-51)."""
    cut = np.linspace(0, 1, n_samples + 1)
    rs.rand(n_samples, dim)  # pyDOE draws and discards this matrix
    center = (cut[:n_samples] + cut[1:n_samples + 1]) / 2
    H = np.zeros((n_samples, dim))
    for j in range(dim):
        H[:, j] = rs.permutation(center)
    return H


def rescale_lhs(raw: np.ndarray, bounds: np.ndarray) -> np.ndarray:
    """AbstractBayesianOptimizer._rescale_lhs_points (:178-194)."""
    out = np.empty_like(raw)
    for d in range(raw.shape[1]):
        out[:, d] = (bounds[d, 1] - bounds[d, 0]) * raw[:, d] + bounds[d, 0]
    return out


class BatchedBOEngine:
    def __init__(self, problems, dim, bounds, budget, n_doe, seeds, algorithm="pca", var_threshold=0.95,
                 n_components=0, maximization=False, device="cuda:0", optimizer="scipy",
                 num_restarts=NUM_RESTARTS, raw_samples=RAW_SAMPLES, batch_limit=BATCH_LIMIT, maxiter=MAXITER,
                 np_rngs=None, torch_gens=None, profile_timing=False, batch_objective=None, host_threads=None,
                 acquisition="expected_improvement", lengthscale=ops.LENGTHSCALE, noise=ops.NOISE,
                 fit_hyperparameters=False):
        if not torch.cuda.is_available():
            raise PcaboError("BatchedBOEngine needs a CUDA device: the PCA-BO hot path has no CPU fallback")
        if algorithm not in ("pca", "vanilla"):
            raise ValueError(f"Invalid algorithm name: '{algorithm}'")
        if optimizer not in ("scipy", "device"):
            raise ValueError(f"Invalid optimizer: '{optimizer}'")
        if acquisition not in ops.ACQ_KINDS:
            # "upper_confidence_bound" cannot run in the reference either: it builds UpperConfidenceBound(model=...,
            # best_f=..., maximize=...) (PCA_BO.py:199-203) and botorch's class has no best_f argument
            raise ValueError(f"acquisition '{acquisition}' is not available on the accelerated path "
                             f"(one of {sorted(ops.ACQ_KINDS)})")
        self.acquisition = acquisition
        # GP hyper-parameters: the reference never fits its GP, these are gpytorch's / botorch's initial values
        self.lengthscale, self.noise = float(lengthscale), float(noise)
        self.fit_hyperparameters = bool(fit_hyperparameters)
        self.last_hyper = None
        # `problems`: a list of B callables p(x) -> float, or one batched objective (e.g. pcabo_b200.bbob.BatchedBBOB)
        # exposing B, eval_many(b, X[m, D]) -> ndarray and __call__(X[B, D] device tensor) -> [B] device tensor
        self.batched_problem = problems if hasattr(problems, "eval_many") else None
        self.problems = None if self.batched_problem is not None else list(problems)
        self.B = self.batched_problem.B if self.batched_problem is not None else len(self.problems)
        self.D = int(dim)
        self.bounds = np.asarray(bounds, dtype=np.float64).reshape(self.D, 2)
        self.budget = int(budget)
        self.n_doe = int(n_doe) if n_doe and n_doe > 0 else self.D
        self.seeds = list(seeds)
        self.algorithm, self.optimizer = algorithm, optimizer
        self.var_threshold, self.n_components = float(var_threshold), int(n_components or 0)
        self.maximization = bool(maximization)
        self.num_restarts, self.raw_samples = int(num_restarts), int(raw_samples)
        self.batch_limit, self.maxiter = int(batch_limit), int(maxiter)
        self.device = torch.device(device)
        self.profile_timing = profile_timing
        self.batch_objective = batch_objective
        B, D, cap = self.B, self.D, self.budget
        self.cap = cap
        # host RNG work may run ahead under the optimiser kernel only when the streams are the engine's own (the drop-in
        # classes pass the process-global generators, which the user's objective may share)
        self.owns_rngs = np_rngs is None
        self.np_rngs = np_rngs if np_rngs is not None else [np.random.RandomState(s) for s in self.seeds]
        if torch_gens is not None:
            self.torch_gens = torch_gens
        else:
            self.torch_gens = []
            for s in self.seeds:
                g = torch.Generator()
                g.manual_seed(int(s))
                self.torch_gens.append(g)
        f64 = dict(dtype=torch.float64, device=self.device)
        self.X = torch.zeros(B, cap, D, **f64)
        self.f = torch.zeros(B, cap, **f64)
        self.logtab = torch.from_numpy(np.log(np.maximum(np.arange(cap + 2), 1).astype(np.float64))).to(self.device)
        self.lo = torch.from_numpy(self.bounds[:, 0].copy()).to(self.device)
        self.hi = torch.from_numpy(self.bounds[:, 1].copy()).to(self.device)
        self.gp = ops.GPBuffers(B, cap, D, self.device)
        self.gp.maximize = self.maximization
        self.gp.acq = ops.ACQ_KINDS[acquisition]
        self.pca = None
        self.d_noise = torch.zeros(B * cap * D, **f64)
        self.h_noise = torch.zeros(B * cap * D, dtype=torch.float64).pin_memory()
        self.gen_io = GeneratorStateIO(B)
        self.h_q = torch.zeros(B, self.raw_samples, dtype=torch.float64)
        self.h_info = torch.zeros(B, dtype=torch.int32).pin_memory()   # Cholesky status of the last GP fit
        self.d_mt_state_prev = self.d_mt_pos_prev = None
        self.state0 = sobol_state0(D).to(self.device)
        self.k_full = torch.full((B,), D, dtype=torch.int32, device=self.device)
        if self.owns_rngs:
            # host-side torch is RNG plumbing on tiny tensors: intra-op threads only add contention. Process-global, so
            # only when the engine is the application (bench, ExperimentRunner); the drop-in classes, which share the
            # process and its generators with user code, leave the setting alone
            torch.set_num_threads(1)
        # per-run host RNG work (numpy normal, torch randint / multinomial) releases the GIL: spread it over the cores
        self.pool = ThreadPoolExecutor(max_workers=max(1, min(B, host_threads or (os.cpu_count() or 1))))
        self.n = 0
        self.x_evals = [[] for _ in range(B)]
        self.f_evals = [[] for _ in range(B)]
        self.z_evals = [None] * B
        self.oob_evals = [[] for _ in range(B)]   # per point: True = out-of-bounds proposal, penalised, objective NOT called
        self.current_best = [(-math.inf if self.maximization else math.inf)] * B
        self.current_best_index = [0] * B
        self.k_host = np.full(B, D, dtype=np.int64)
        self.timing_logs = {"pca": [], "SingleTaskGP": [], "optimize_acqf": []}
        self.stats = dict(kernel_launches=0, lbfgsb_rounds=0, h2d_bytes=0, d2h_bytes=0, retries=0, jitter_refits=0,
                          wait_opt_s=0.0, wait_ic_s=0.0)   # host time blocked on the optimiser / raw-sample events
        self.last = {}
        self.noise_ready_n = -1     # the pinned noise buffer holds the draw for the iteration with this many points
        self.noise_copied = torch.cuda.Event()
        self.noise_futs = None
        self.overlap_hook = None
        self.lb_out = None
        self.lb_ws = None         # device workspace of the L-BFGS-B pairs when they do not fit shared memory (C5)
        self.h_status = self.h_nfev = self.h_nit = None
        self.h_vals = self.h_state_back = self.h_pos_back = self.h_k = self.h_flag = None
        self.d_mt_state = self.d_mt_pos = self.ic_out = None
        self.gens_on_host = True       # False: the device-resident MT19937 states are ahead of the torch CPU generators
        self.force_host_ic = False     # tests: always take the host path for the Boltzmann pick
        self.raw_scoring = "tensor"    # "scalar": always score the raw samples with pcabo_logei (tests, comparisons)
        # asynchronous tail (PipelinedEngines sets it): penalty / append / best-so-far run on the device
        # (pcabo_append_observation) and the host lists are brought up to date later from packed records read back
        # asynchronously (_drain), so step_end() returns without waiting for the GPU
        self.async_tail = False
        self.best_dev = torch.zeros(B, **f64)        # running best of every run (unrounded)
        self.best_dev_valid = False                  # best_dev / gp.best_f are maintained by the device tail
        self.rec_dev = torch.zeros(B, 2 * D + 3, **f64)
        self.pending_host = []                       # [(n, pinned record, event, k)] of iterations not yet in the host lists
        self.rec_free = []
        self.pick_out = self.im_out = None
        # optional low-priority side stream for the optimiser kernel (PipelinedEngines): its blocks hold a whole SM for
        # milliseconds, and the short kernels of the other sub-batches' model fits must not queue behind them
        self.opt_stream = None
        # longest-first dispatch of the optimiser kernel's blocks: 2 B chunk problems want more SMs than the GPU has (180
        # against 148 for 90 runs), the blocks dispatched last start when the first chunks finish, and the launch lasts as
        # long as its last block -- so the runs whose chunks took longest in the PREVIOUS iteration are dispatched first
        self.lpt_dispatch = True
        self.prev_cost = None          # [B] iterations + evaluations of each run's slowest chunk in the last full launch
        self.h_order = self.d_order = None
        self.d_ics_buf = None
        self.pending = None         # state of a step between step_begin() and step_end()
        self.phase_times = None     # set to {} to accumulate host wall time per phase (adds device syncs)
        self.kernel_events = None   # set to [] to record CUDA-event pairs around every value+grad log-EI launch

    # ------------------------------------------------------------------ bookkeeping (AbstractBayesianOptimizer.py:196-208)
    def _assign_new_best(self, b):
        fe = self.f_evals[b]
        self.current_best[b] = max(fe) if self.maximization else min(fe)
        self.current_best_index[b] = fe.index(self.current_best[b], self.current_best_index[b])

    def _update_best_with_last(self, b):
        """_assign_new_best after ONE appended value, without rescanning the history: a strictly better value is new and
        sits at the last index; otherwise min / max and `index(best, previous index)` are unchanged (ties keep the
        earlier position) -- identical to AbstractBayesianOptimizer.assign_new_best (:196-208)."""
        fe = self.f_evals[b]
        v = fe[-1]
        if (v > self.current_best[b]) if self.maximization else (v < self.current_best[b]):
            self.current_best[b] = v
            self.current_best_index[b] = len(fe) - 1

    def _evaluate(self, b, x):
        return float(self.problems[b](x))

    def initial_design(self):
        """AbstractBayesianOptimizer.__call__ (:142-168): LHS DoE, evaluation, best tracking."""
        B, D = self.B, self.D
        designs = [rescale_lhs(lhs_center(self.np_rngs[b], D, self.n_doe), self.bounds) for b in range(B)]
        if self.batched_problem is not None:
            vals = self.batched_problem.eval_all(np.stack(designs))
        for b in range(B):
            for j, p in enumerate(designs[b]):
                self.x_evals[b].append(p)
                self.f_evals[b].append(float(vals[b, j]) if self.batched_problem is not None else self._evaluate(b, p))
                self.oob_evals[b].append(False)
            self._assign_new_best(b)
        n0 = self.n_doe
        hx = torch.from_numpy(np.stack([np.stack(self.x_evals[b]) for b in range(B)]))
        hf = torch.from_numpy(np.array([self.f_evals[b] for b in range(B)], dtype=np.float64))
        self.X[:, :n0].copy_(hx)
        self.f[:, :n0].copy_(hf)
        self.stats["h2d_bytes"] += hx.numel() * 8 + hf.numel() * 8
        self.n = n0

    def attach_history(self, x_lists, f_lists):
        """Adopt already evaluated points (the caller's own list objects are kept and appended to)."""
        B = self.B
        self._drain(True)
        self.best_dev_valid = False
        n0 = len(f_lists[0])
        if any(len(f) != n0 for f in f_lists) or n0 > self.cap:
            raise PcaboError("attach_history: all runs of a batch must hold the same number of points (<= budget)")
        self.x_evals, self.f_evals = list(x_lists), list(f_lists)
        self.oob_evals = [[False] * n0 for _ in range(B)]
        hx = torch.from_numpy(np.stack([np.stack([np.asarray(x, dtype=np.float64) for x in x_lists[b]])
                                        for b in range(B)]))
        hf = torch.from_numpy(np.array([[float(v) for v in f_lists[b]] for b in range(B)], dtype=np.float64))
        self.X[:, :n0].copy_(hx)
        self.f[:, :n0].copy_(hf)
        self.stats["h2d_bytes"] += hx.numel() * 8 + hf.numel() * 8
        self.n = n0
        self.current_best_index = [0] * B
        for b in range(B):
            self._assign_new_best(b)

    # ------------------------------------------------------------------ one lock-step BO iteration
    def _sync_time(self):
        if self.profile_timing:
            torch.cuda.synchronize(self.device)
        return perf_counter()

    def _phase(self, name, t0):
        """Accumulate wall time since t0 under `name` when phase profiling is on; returns the new time origin."""
        if self.phase_times is None:
            return t0
        torch.cuda.synchronize(self.device)
        t1 = perf_counter()
        self.phase_times[name] = self.phase_times.get(name, 0.0) + (t1 - t0)
        return t1

    def _draw_noise(self, n, wait=True):
        """The 1e-8 PCA noise of the iteration that fits n points, drawn from each run's numpy stream (PCA_BO.py:376) on
        the thread pool (numpy fills the array without the GIL). wait=False returns with the draws in flight."""
        B, D = self.B, self.D
        hv_np = self.h_noise[:B * n * D].view(B, n, D).numpy()

        def _noise(b):
            hv_np[b] = self.np_rngs[b].normal(0, 1e-8, size=(n, D))

        self.noise_futs = [self.pool.submit(_noise, b) for b in range(B)]
        self.noise_ready_n = n
        if wait:
            self._wait_noise()

    def _wait_noise(self):
        if self.noise_futs:
            for fu in self.noise_futs:
                fu.result()
            self.noise_futs = None

    def _prefetch_noise(self):
        """Start drawing the NEXT iteration's noise while the optimiser kernel runs; the host thread does not wait
        for it (it is collected at the next model fit). The pinned buffer of this iteration has been consumed (its H2D
        copy precedes the kernel on the stream) and the numpy streams are used for nothing else between two iterations,
        so the draw order of every run is unchanged."""
        if self.owns_rngs and self.algorithm == "pca" and self.n + 1 < self.budget:
            self.noise_copied.synchronize()
            self._draw_noise(self.n + 1, wait=False)

    def _fit_pca(self):
        """_transform_points_to_reduced_space (PCA_BO.py:343-408): the 1e-8 noise from each run's numpy stream, then rank
        weights + weighted PCA + projection in one kernel."""
        B, D, n = self.B, self.D, self.n
        tp = perf_counter()
        hv = self.h_noise[:B * n * D].view(B, n, D)
        if self.noise_ready_n != n:
            self._wait_noise()
            self._draw_noise(n)
        self._wait_noise()
        self.noise_ready_n = -1
        self._phase("host_noise_rng", tp)
        dv = self.d_noise[:B * n * D].view(B, n, D)
        dv.copy_(hv, non_blocking=True)
        self.noise_copied.record()
        self.stats["h2d_bytes"] += hv.numel() * 8
        t0 = self._sync_time()
        self.pca = ops.weights_pca_fit(self.X, self.f, dv, self.logtab, n, self.var_threshold,
                                       self.n_components, self.maximization, out=self.pca)
        t1 = self._sync_time()
        self.stats["kernel_launches"] += 1
        self.timing_logs["pca"].append(t1 - t0)

    def _set_best_f(self):
        """best_f as botorch stores it: torch.as_tensor(python float) -> float32 (PCA_BO.py:199-203).
        Staged through pinned memory: a copy from pageable memory would block the host thread until the model fit
        enqueued just before it has finished (1.2 ms per call on the profile of round 2)."""
        if self.async_tail and self.best_dev_valid:
            return      # kept up to date on the device by pcabo_append_observation
        self._drain(True)
        if getattr(self, "h_best_f", None) is None:
            self.h_best_f = torch.zeros(2, self.B, dtype=torch.float64).pin_memory()
        cb = np.asarray(self.current_best, dtype=np.float64)
        self.h_best_f.numpy()[0] = cb.astype(np.float32).astype(np.float64)
        self.h_best_f.numpy()[1] = cb
        self.gp.best_f.copy_(self.h_best_f[0], non_blocking=True)
        self.stats["h2d_bytes"] += self.B * 8
        if self.async_tail:
            self.best_dev.copy_(self.h_best_f[1], non_blocking=True)
            self.best_dev_valid = True
            self.stats["h2d_bytes"] += self.B * 8

    def _fit_models(self):
        tp = perf_counter()
        if self.algorithm == "pca":
            self._fit_pca()
        t1 = self._sync_time()
        self._fit_gp()
        t2 = self._sync_time()
        self.stats["kernel_launches"] += ops.gp_fit_launches(self.n)
        self.timing_logs["SingleTaskGP"].append(t2 - t1)
        self._phase("pca_gp_kernels", tp)
        self._set_best_f()

    def _fit_gp(self, jitter=None):
        """_initialize_model (PCA_BO.py:502-545 / Vanilla_BO.py:166-195) + gpytorch's prediction caches; the Cholesky
        status of every run is read back asynchronously and examined in _propose_begin."""
        if self.fit_hyperparameters and jitter is None:
            # opt-in (the reference never fits its GP): per-run (lengthscale, noise) by maximum a posteriori
            # marginal likelihood, as botorch's fit_gpytorch_mll would do for this model; the caches are left at the fit
            from .hyperfit import fit_hyperparameters
            pca = self.algorithm == "pca"
            self.last_hyper = fit_hyperparameters(
                self.gp, self.pca["Z"] if pca else self.X, self.pca["k"] if pca else None, self.f, self.n,
                mode=0 if pca else 1, fix_lo=None if pca else self.lo, fix_hi=None if pca else self.hi)
            self.h_info.copy_(self.gp.info, non_blocking=True)
            self.stats["d2h_bytes"] += self.B * 4
            return
        hyper = dict(lengthscales=self.gp.lengthscales, noises=self.gp.noises) if self.fit_hyperparameters else {}
        if self.algorithm == "pca":
            ops.gp_fit(self.gp, self.pca["Z"], self.pca["k"], self.f, self.n, mode=0, lengthscale=self.lengthscale,
                       noise=self.noise, jitter=jitter, **hyper)
        else:
            ops.gp_fit(self.gp, self.X, None, self.f, self.n, mode=1, fix_lo=self.lo, fix_hi=self.hi,
                       lengthscale=self.lengthscale, noise=self.noise, jitter=jitter, **hyper)
        self.h_info.copy_(self.gp.info, non_blocking=True)
        self.stats["d2h_bytes"] += self.B * 4

    def _refit_failed_cholesky(self, st):
        """A non-positive pivot in some run's Cholesky: gpytorch's psd_safe_cholesky (linear_operator, reached from the
        first posterior call on the model of PCA_BO.py:535-544) retries with 1e-8, 1e-7, 1e-6 added to the diagonal and
        raises NotPSDError after that. The raw samples of this iteration were scored with the failed factor: the
        generators are put back to where they were and the initial conditions are drawn again from the refitted model,
        so the RNG streams advance exactly as in the reference (where the retry happens inside the factorisation)."""
        B = self.B
        bad = np.nonzero(self.h_info.numpy())[0]
        jitter = torch.zeros(B, dtype=torch.float64)
        for jit in (1e-8, 1e-7, 1e-6):
            jitter[torch.from_numpy(bad)] = jit
            self._fit_gp(jitter=jitter.to(self.device))
            torch.cuda.current_stream().synchronize()
            self.stats["jitter_refits"] += len(bad)
            self.stats["kernel_launches"] += ops.gp_fit_launches(self.n)
            bad = np.nonzero(self.h_info.numpy())[0]
            if len(bad) == 0:
                break
        else:
            raise PcaboError(f"Matrix not positive definite after repeatedly adding jitter up to 1.0e-06 (runs {bad.tolist()})")
        if st["ic"].get("fast"):
            self.d_mt_state.copy_(self.d_mt_state_prev)
            self.d_mt_pos.copy_(self.d_mt_pos_prev)
        st["ic"] = self._ic_launch(st["runs"], None, 1)
        st["ic"]["done"].synchronize()

    def _gens_to_host(self):
        """Bring the torch CPU generators up to date with the device-resident MT19937 states (fast path)."""
        if not self.gens_on_host:
            self.gen_io.restore(self.torch_gens, self.d_mt_state.cpu(), self.d_mt_pos.cpu())
            self.gens_on_host = True

    def _score_raw(self, x_raw):
        """log-EI of the raw Sobol samples of every run (botorch scores them 5 at a time, PCA_BO.py:607-614): the tile
        pipeline of the candidate sweep on the FP64 tensor cores when the model is within its limits (log-EI, n <= 512,
        even capacities), the general value kernel otherwise (config C5, PI, odd D)."""
        if self.raw_scoring == "tensor" and ops.logei_raw_supported(self.gp, self.n):
            return ops.logei_raw(self.gp, x_raw)["val"]
        return ops.logei(self.gp, x_raw)["val"]

    def _ic_launch_fast(self):
        """_ic_launch for the common case (all runs, 512 raw samples, log-EI, generators owned by the engine): the
        Boltzmann pick itself runs on the device too (pcabo_select_initial_conditions replays torch.multinomial from
        the device-resident generator states), so nothing but a per-run flag and k comes back to the host."""
        B, D, R = self.B, self.D, self.num_restarts
        if self.gens_on_host:
            st, pos = self.gen_io.export(self.torch_gens)
            self.d_mt_state = st.to(self.device, non_blocking=True)
            self.d_mt_pos = pos.to(self.device, non_blocking=True)
            self.gens_on_host = False
            self.stats["h2d_bytes"] += st.numel() * 4 + pos.numel() * 4
        k_dev = self.pca["k"] if self.algorithm == "pca" else self.k_full
        if self.d_mt_state_prev is None or self.d_mt_state_prev.shape != self.d_mt_state.shape:
            self.d_mt_state_prev, self.d_mt_pos_prev = torch.empty_like(self.d_mt_state), torch.empty_like(self.d_mt_pos)
        self.d_mt_state_prev.copy_(self.d_mt_state)      # kept for the (rare) jitter refit of a failed Cholesky
        self.d_mt_pos_prev.copy_(self.d_mt_pos)
        d_shift, d_dots = ops.sobol_scramble_bits(k_dev, self.d_mt_state, self.d_mt_pos, D)
        x_raw = ops.sobol_points(self.raw_samples, k_dev, d_shift, d_dots, self.state0, self.gp.box_lo, self.gp.box_hi)
        vals_dev = self._score_raw(x_raw)
        self.ic_out = ops.select_initial_conditions(vals_dev, self.d_mt_state, self.d_mt_pos, x_raw, R, out=self.ic_out)
        if self.h_flag is None:
            self.h_flag = torch.zeros(B, dtype=torch.int32).pin_memory()
            self.h_k = torch.zeros(B, dtype=torch.int32).pin_memory()
        self.h_flag.copy_(self.ic_out["flag"], non_blocking=True)
        if self.algorithm == "pca":
            self.h_k.copy_(self.pca["k"], non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        self.stats["kernel_launches"] += 4
        self.stats["d2h_bytes"] += 8 * B
        return dict(pending=list(range(B)), factor=1, x_raw=x_raw, vals_dev=vals_dev, done=done, fast=True)

    def _ic_launch(self, pending, k_host, factor):
        """Enqueue the device part of gen_batch_initial_conditions for the `pending` runs: Sobol scramble bits (each
        run's torch CPU generator, MT19937, replayed on the device), Sobol points scaled to the box, log-EI of the raw
        samples, and the asynchronous read-back of the values and generator states into pinned memory."""
        B, D = self.B, self.D
        if (factor == 1 and len(pending) == B and self.owns_rngs and self.gp.acq == ops.ACQ_LOG_EI
                and self.num_restarts <= 32 and self.raw_samples <= 1024 and not self.force_host_ic):
            return self._ic_launch_fast()
        self._gens_to_host()
        nraw = self.raw_samples * factor
        st, pos = self.gen_io.export(self.torch_gens)
        d_state = st.to(self.device, non_blocking=True)
        d_pos = pos.to(self.device, non_blocking=True)
        if len(pending) == B:
            k_dev = self.pca["k"] if self.algorithm == "pca" else self.k_full
        else:  # runs that are not pending draw nothing (k = 0)
            kk = torch.zeros(B, dtype=torch.int32)
            kk[pending] = torch.from_numpy(k_host[pending].astype(np.int32))
            k_dev = kk.to(self.device)
        d_shift, d_dots = ops.sobol_scramble_bits(k_dev, d_state, d_pos, D)
        x_raw = ops.sobol_points(nraw, k_dev, d_shift, d_dots, self.state0, self.gp.box_lo, self.gp.box_hi)
        vals_dev = self._score_raw(x_raw)
        if self.h_vals is None or self.h_vals.shape != vals_dev.shape:
            self.h_vals = torch.zeros(vals_dev.shape, dtype=torch.float64).pin_memory()
        if self.h_state_back is None:
            self.h_state_back = torch.zeros(B, 624, dtype=torch.int32).pin_memory()
            self.h_pos_back = torch.zeros(B, dtype=torch.int32).pin_memory()
            self.h_k = torch.zeros(B, dtype=torch.int32).pin_memory()
        self.h_state_back.copy_(d_state, non_blocking=True)
        self.h_pos_back.copy_(d_pos, non_blocking=True)
        self.h_vals.copy_(vals_dev, non_blocking=True)
        if self.algorithm == "pca":
            self.h_k.copy_(self.pca["k"], non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        self.stats["kernel_launches"] += 3
        self.stats["h2d_bytes"] += st.numel() * 4 + pos.numel() * 4
        self.stats["d2h_bytes"] += vals_dev.numel() * 8 + st.numel() * 4 + pos.numel() * 4
        return dict(pending=list(pending), factor=factor, x_raw=x_raw, done=done)

    def _ic_finish(self, h, d_ics):
        """Host part: generator states back, Boltzmann multinomial per run, gather of the selected rows into d_ics.
        Returns the runs whose raw values were degenerate (botorch retries those with more raw samples)."""
        B, R = self.B, self.num_restarts
        pending = h["pending"]
        h["done"].synchronize()
        if h.get("fast"):
            d_ics.copy_(self.ic_out["ics"])
            flagged = [b for b in pending if int(self.h_flag[b]) != 0]
            if not flagged:
                return []
            # degenerate raw values (botorch: random permutation pick + warning, then more raw samples): host path for
            # those runs, from the generator states the device left them in
            self._gens_to_host()
            vals = h["vals_dev"].cpu()
            idx, bad = select_initial_conditions(vals, R, self.torch_gens, flagged, None)
            d_idx = idx.to(torch.int32).to(self.device)
            picked = ops.gather_rows(h["x_raw"], d_idx)
            sel = torch.tensor(flagged, device=self.device)
            d_ics[sel] = picked[sel]
            return [b for b in flagged if bad[b]]
        self.gen_io.restore(self.torch_gens, self.h_state_back, self.h_pos_back)
        if self.gp.acq == ops.ACQ_PI:      # botorch picks initialize_q_batch_nonneg for non-negative acquisitions
            idx, bad = select_initial_conditions_nonneg(self.h_vals, R, self.torch_gens, pending)
        else:
            idx, bad = select_initial_conditions(self.h_vals, R, self.torch_gens, pending,
                                                 self.h_q if h["factor"] == 1 else None)
        d_idx = idx.to(torch.int32).to(self.device, non_blocking=True)
        picked = ops.gather_rows(h["x_raw"], d_idx)
        if len(pending) == B:
            d_ics.copy_(picked)
        else:
            sel = torch.tensor(pending, device=self.device)
            d_ics[sel] = picked[sel]
        self.stats["kernel_launches"] += 1
        self.stats["h2d_bytes"] += d_idx.numel() * 4
        return [b for b in pending if bad[b]]

    def _initial_conditions(self, runs, k_host, d_ics=None, first=None):
        """gen_batch_initial_conditions for the given runs.

        Host: only the Boltzmann multinomial from each run's own CPU generator. Device: scramble bits, Sobol points
        scaled to the box, log-EI of the raw samples, gather of the selected rows. `first`: an already launched
        _ic_launch handle for (runs, factor 1). Returns the device tensor of initial conditions [B, num_restarts, D]
        (rows of `runs` updated)."""
        B, D, R = self.B, self.D, self.num_restarts
        if d_ics is None:
            if self.d_ics_buf is None:
                self.d_ics_buf = torch.zeros(B, R, D, dtype=torch.float64, device=self.device)
            d_ics = self.d_ics_buf      # (every row is rewritten below; columns >= k stay zero)
        pending = list(runs)
        factor = 1
        while pending and factor < 5:
            tp = perf_counter()
            h = first if (first is not None and factor == 1) else self._ic_launch(pending, k_host, factor)
            if self.phase_times is not None:
                h["done"].synchronize()
            tp = self._phase("sobol_bits_points_logei_kernels", tp)
            pending = self._ic_finish(h, d_ics)
            self._phase("host_multinomial", tp)
            factor += 1
        return d_ics

    def _optimize_scipy(self, runs, d_ics, k_host, box_lo, box_hi):
        """gen_candidates_scipy for every chunk of batch_limit restarts of every run, in lock-step.

        Returns (cands {b: [num_restarts, k_b]}, warned {b: bool})."""
        B, D, R = self.B, self.D, self.num_restarts
        h_ics = d_ics.cpu().numpy()
        ics = {b: h_ics[b, :, :int(k_host[b])] for b in runs}
        chunks = [(s, min(s + self.batch_limit, R)) for s in range(0, R, self.batch_limit)]
        probs = {}
        for b in runs:
            kb = int(k_host[b])
            lo, hi = box_lo[b, :kb].numpy(), box_hi[b, :kb].numpy()
            for ci, (s, e) in enumerate(chunks):
                x0 = np.clip(ics[b][s:e], lo, hi).reshape(-1)                         # columnwise_clamp
                probs[(b, ci)] = LbfgsbProblem(x0, np.tile(lo, e - s), np.tile(hi, e - s), self.maxiter)
        h_x = torch.zeros(B, R, D, dtype=torch.float64).pin_memory()
        for b in runs:
            h_x[b, :, :int(k_host[b])] = torch.from_numpy(ics[b])
        active = {key for key, p in probs.items() if p.advance()}
        out = None
        while active:
            for (b, ci) in active:
                s, e = chunks[ci]
                kb = int(k_host[b])
                h_x[b, s:e, :kb] = torch.from_numpy(probs[(b, ci)].x.reshape(e - s, kb))
            d_x = h_x.to(self.device, non_blocking=True)
            if self.kernel_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = ops.logei(self.gp, d_x, want_grad=True, out=out)
                e1.record()
                n = self.n
                nbytes = sum(4 * n * n + 8 * n * int(kb) + 8 * n + R * (16 * int(kb) + 8) for kb in k_host)
                nflops = sum(R * (2 * n * n + 11 * n * int(kb) + 16 * n) for kb in k_host)
                self.kernel_events.append((e0, e1, nbytes, nflops))
            else:
                out = ops.logei(self.gp, d_x, want_grad=True, out=out)
            val = out["val"].cpu().numpy()
            grad = out["grad"].cpu().numpy()
            self.stats["kernel_launches"] += 1
            self.stats["lbfgsb_rounds"] += 1
            self.stats["h2d_bytes"] += h_x.numel() * 8
            self.stats["d2h_bytes"] += (val.size + grad.size) * 8
            nxt = set()
            for (b, ci) in active:
                s, e = chunks[ci]
                kb = int(k_host[b])
                p = probs[(b, ci)]
                p.supply(-float(val[b, s:e].sum()), -grad[b, s:e, :kb].reshape(-1))
                if p.advance():
                    nxt.add((b, ci))
            active = nxt
        cands, warned = {}, {}
        for b in runs:
            kb = int(k_host[b])
            lo, hi = box_lo[b, :kb].numpy(), box_hi[b, :kb].numpy()
            c = np.concatenate([probs[(b, ci)].x.reshape(-1, kb) for ci in range(len(chunks))], axis=0)
            cands[b] = np.clip(c, lo, hi)
            warned[b] = any(probs[(b, ci)].optimization_warning for ci in range(len(chunks)))
            self.last.setdefault("nit", {})[b] = [probs[(b, ci)].nit for ci in range(len(chunks))]
            self.last.setdefault("nfev", {})[b] = [probs[(b, ci)].nfev for ci in range(len(chunks))]
        return cands, warned

    def _launch_device(self, d_x, k_host, runs=None, out=None):
        """Enqueue gen_candidates_scipy for all (or the listed) runs and chunks: ONE launch of the device-resident
        L-BFGS-B kernel, then the asynchronous read-back of the per-chunk status words into pinned memory. With `runs`
        the rows of those runs in `out` are re-optimised in place (botorch's retry). Returns a handle for
        _finish_device; the host is free until then."""
        B = self.B
        sel = list(range(B)) if runs is None else list(runs)
        ev = None
        if self.kernel_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        run_idx = None if runs is None else torch.tensor(sel, dtype=torch.int32).to(self.device, non_blocking=True)
        if runs is None and self.lpt_dispatch and self.prev_cost is not None and B > 1:
            if self.h_order is None:
                self.h_order = torch.zeros(B, dtype=torch.int32).pin_memory()
                self.d_order = torch.zeros(B, dtype=torch.int32, device=self.device)
            self.h_order.numpy()[:] = np.argsort(-self.prev_cost, kind="stable")
            self.d_order.copy_(self.h_order, non_blocking=True)
            run_idx = self.d_order       # (results are indexed by run: the order changes nothing but the dispatch)
        k_bound = int(max(k_host[b] for b in sel))
        need = ops.lbfgsb_workspace_doubles(self.n, self.D, k_bound, self.num_restarts, self.batch_limit, len(sel))
        if need > 0 and (self.lb_ws is None or self.lb_ws.numel() < need):
            # sized once for the whole batch at the largest k (never reallocated inside the loop after that)
            full = ops.lbfgsb_workspace_doubles(self.n, self.D, self.D, self.num_restarts, self.batch_limit, B)
            self.lb_ws = torch.empty(max(need, full), dtype=torch.float64, device=self.device)
        side = self.opt_stream
        done = None
        if side is not None:
            side.wait_stream(torch.cuda.current_stream())
            main = torch.cuda.current_stream()
            torch.cuda.set_stream(side)
        try:
            if ev is not None and side is not None:
                ev[0].record()      # (time the kernel on the stream it runs on)
            out = ops.lbfgsb_logei(self.gp, d_x, self.gp.box_lo, self.gp.box_hi, self.batch_limit, self.maxiter, out=out,
                                   run_idx=run_idx, k_bound=k_bound, workspace=self.lb_ws)
            if ev is not None:
                ev[1].record()
            if self.h_status is None or self.h_status.shape != out["status"].shape:
                self.h_status = torch.zeros(out["status"].shape, dtype=torch.int32).pin_memory()
                self.h_nfev, self.h_nit = torch.zeros_like(self.h_status).pin_memory(), torch.zeros_like(self.h_status).pin_memory()
            self.h_status.copy_(out["status"], non_blocking=True)
            self.h_nfev.copy_(out["nfev"], non_blocking=True)
            self.h_nit.copy_(out["nit"], non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        finally:
            if side is not None:
                torch.cuda.set_stream(main)
                if done is not None:
                    main.wait_event(done)     # whatever follows on the sub-batch's own stream sees the candidates
        if self.overlap_hook is not None:        # host work that does not depend on the result runs under the kernel
            hook, self.overlap_hook = self.overlap_hook, None
            hook()
        # (run_idx stays referenced until the launch has been collected: it may have been used on the side stream)
        return dict(out=out, sel=sel, ev=ev, done=done, k_host=k_host, run_idx=run_idx)

    def _finish_device(self, h):
        """Wait for the optimiser kernel of `h`. Returns (the output dict with the device tensor of clamped candidates
        cand [B, R, D], warned {b: bool})."""
        tw = perf_counter()
        h["done"].synchronize()
        self.stats["wait_opt_s"] += perf_counter() - tw
        out, sel, ev, k_host = h["out"], h["sel"], h["ev"], h["k_host"]
        status, nfev, nit = self.h_status.numpy(), self.h_nfev.numpy(), self.h_nit.numpy()
        if (status[sel] < 0).any():
            raise PcaboError("pcabo_lbfgsb_logei_runs: k_bound below a run's k")
        if ev is not None:
            n = self.n
            # algorithmic bytes of one value+gradient evaluation of one chunk: the n x n lower-triangular FP64 cache
            # L^-1 once (4 n^2), the training set and alpha (8 n k + 8 n); times the evaluations actually done
            ks = np.asarray(k_host, dtype=np.int64)[sel]
            ev_run = nfev[sel].astype(np.int64)                       # [runs, chunks] evaluations
            nbytes = int((ev_run.sum(axis=1) * (4 * n * n + 8 * n * ks + 8 * n)).sum())
            # algorithmic flops: F_vg = 2 n^2 + 11 n k + 16 n per point and value+gradient evaluation
            pts = np.array([min(self.batch_limit, self.num_restarts - c * self.batch_limit) for c in range(nfev.shape[1])],
                           dtype=np.int64)
            nflops = int(((ev_run * pts[None, :]).sum(axis=1) * (2 * n * n + 11 * n * ks + 16 * n)).sum())
            self.kernel_events.append((ev[0], ev[1], nbytes, nflops))
        self.stats["kernel_launches"] += 1
        self.stats["lbfgsb_rounds"] += int(nfev[sel].max())
        self.stats["d2h_bytes"] += status.size * 12
        if len(sel) == self.B:
            # (a decaying maximum: a run that needed the full 200 iterations tends to need them again soon)
            cost = (nit.astype(np.int64) + nfev).max(axis=1).astype(np.float64)
            self.prev_cost = cost if self.prev_cost is None else np.maximum(cost, 0.7 * self.prev_cost)
        w3 = (status == 3).any(axis=1)
        warned = {b: bool(w3[b]) for b in sel}
        nl, fl = nit.tolist(), nfev.tolist()
        dn, df = self.last.setdefault("nit", {}), self.last.setdefault("nfev", {})
        for b in sel:
            dn[b] = nl[b]
            df[b] = fl[b]
        return out, warned

    def _optimize_device(self, d_x, k_host, runs=None, out=None):
        """_launch_device + _finish_device (blocking)."""
        return self._finish_device(self._launch_device(d_x, k_host, runs=runs, out=out))

    def _propose_launch(self):
        """optimize_acqf (PCA_BO.py:607-614), phase A: raw samples and their log-EI enqueued, nothing awaited."""
        runs = list(range(self.B))
        return dict(runs=runs, ic=self._ic_launch(runs, None, 1), handle=None)

    def _propose_begin(self, st):
        """Phase B: initial conditions selected (host multinomial), optimiser kernel in flight (device mode)."""
        B, D = self.B, self.D
        tw = perf_counter()
        st["ic"]["done"].synchronize()
        self.stats["wait_ic_s"] += perf_counter() - tw
        if self.h_info.numpy().any():
            self._refit_failed_cholesky(st)
        k_host = self.h_k.numpy().astype(np.int64) if self.algorithm == "pca" else np.full(B, D, np.int64)
        self.k_host = k_host
        self.stats["d2h_bytes"] += B * 4
        ics = self._initial_conditions(st["runs"], k_host, first=st["ic"])
        st.update(k_host=k_host, ics=ics)
        # the device-resident optimiser keeps the L-BFGS-B state of a chunk in shared memory and, past that size (config
        # C5: n up to 1050 with k ~ 90), the BFGS pairs in an L2-resident workspace; there is no host fallback
        st["use_device"] = self.optimizer == "device"
        if st["use_device"] and not ops.lbfgsb_fits(self.n, self.D, int(k_host.max()), self.batch_limit):
            raise PcaboError(f"device optimiser: n = {self.n}, k = {int(k_host.max())} is beyond what pcabo_lbfgsb_logei_runs_ws "
                             "supports; use optimizer='scipy'")
        if st["use_device"]:
            st["t_launch"] = perf_counter()
            self.overlap_hook = self._prefetch_noise
            st["handle"] = self._launch_device(ics, k_host, out=self.lb_out)
        return st

    def _propose_end(self, st):
        """The rest of optimize_acqf: returns device tensor z_new [B, D] and its log-EI values [B]."""
        B, D, R = self.B, self.D, self.num_restarts
        k_host, runs, ics = st["k_host"], st["runs"], st["ics"]
        if not st["use_device"]:
            # the host optimiser needs the search box; the device one reads it in HBM
            box_lo, box_hi = self.gp.box_lo.cpu(), self.gp.box_hi.cpu()
            self.stats["d2h_bytes"] += 2 * B * D * 8
            cands, warned = self._optimize_scipy(runs, ics, k_host, box_lo, box_hi)
            retry = [b for b in runs if warned[b]]
            if retry:  # botorch retry_on_optimization_warning: new initial conditions, one more attempt
                self.stats["retries"] += len(retry)
                ics = self._initial_conditions(retry, k_host, ics)
                c2, _ = self._optimize_scipy(retry, ics, k_host, box_lo, box_hi)
                cands.update(c2)
            h_c = torch.zeros(B, R, D, dtype=torch.float64).pin_memory()
            for b in runs:
                h_c[b, :, :int(k_host[b])] = torch.from_numpy(cands[b])
            d_c = h_c.to(self.device, non_blocking=True)
            self.stats["h2d_bytes"] += h_c.numel() * 8
        else:
            self.lb_out, warned = self._finish_device(st["handle"])
            self._phase("lbfgsb_kernel", st["t_launch"])
            retry = [b for b in runs if warned[b]]
            if retry:
                self.stats["retries"] += len(retry)
                ics = self._initial_conditions(retry, k_host, ics)
                tp = perf_counter()
                self.lb_out, _ = self._optimize_device(ics, k_host, runs=retry, out=self.lb_out)
                self._phase("lbfgsb_kernel_retry", tp)
            d_c = self.lb_out["cand"]
        vals = ops.logei(self.gp, d_c)["val"]                        # final no-grad evaluation of the candidates
        self.pick_out = ops.pick_best(vals, d_c, out=self.pick_out)  # torch.argmax over the restarts + the winner
        self.stats["kernel_launches"] += 2
        return self.pick_out["z"], self.pick_out["val"]

    def step_launch(self):
        """Phase A of a BO iteration: model fit and raw-sample scoring enqueued; returns without waiting for the GPU.
        Returns False when the budget is exhausted."""
        if self.n >= self.budget:
            self.pending = None
            return False
        self._fit_models()
        t0 = self._sync_time()
        self.pending = dict(t0=t0, st=self._propose_launch(), phase="A")
        return True

    def step_begin(self):
        """Phases A + B: everything up to the optimiser kernel in flight (device mode). step_end() completes the
        iteration; a caller can interleave several engines on different streams in between."""
        if self.pending is None or self.pending.get("phase") != "A":
            if not self.step_launch():
                return False
        self._propose_begin(self.pending["st"])
        self.pending["phase"] = "B"
        return True

    def wait_event(self):
        """The CUDA event the next phase of the pending iteration will block on (None if nothing is pending)."""
        if self.pending is None:
            return None
        st = self.pending["st"]
        if self.pending["phase"] == "A":
            return st["ic"]["done"]
        return st["handle"]["done"] if st.get("handle") is not None else None

    def step_end(self):
        """Last phase of a BO iteration (PCA_BO.py:239-281): candidates, inverse map, objective, bookkeeping."""
        if self.pending is None:
            return False
        if self.pending.get("phase") == "A":
            self.step_begin()
        B, D, n = self.B, self.D, self.n
        z_new, acq_val = self._propose_end(self.pending["st"])
        t1 = self._sync_time()
        self.timing_logs["optimize_acqf"].append(t1 - self.pending["t0"])
        self.pending = None
        tp = perf_counter()
        if self.async_tail and self.batched_problem is not None:
            self._step_end_async(z_new, acq_val)
            self._phase("inverse_map_objective_bookkeeping", tp)
            return True
        self._drain(True)
        self.best_dev_valid = False     # the host lists are the truth on this path
        if self.algorithm == "pca":
            im = ops.inverse_map(z_new, self.pca["k"], self.pca["V"], self.pca["m2"], self.pca["xbar"], self.lo, self.hi)
            self.stats["kernel_launches"] += 1
            x_dev = im["x"]
            oob = im["oob"].cpu().numpy().astype(bool)
        else:
            x_dev = z_new
            oob = np.zeros(B, dtype=bool)
        f_dev = self.batched_problem(x_dev).cpu().numpy() if self.batched_problem is not None else None
        x_new = x_dev.cpu().numpy()
        z_host = z_new.cpu().numpy()
        self.stats["d2h_bytes"] += 2 * B * D * 8 + B * 4
        new_f = np.empty(B, dtype=np.float64)
        for b in range(B):
            if oob[b]:
                new_f[b] = -1000.0 if self.maximization else 1000.0          # PCA_BO.py:260-261 (not evaluated)
            else:
                new_f[b] = f_dev[b] if f_dev is not None else self._evaluate(b, x_new[b])
            self.x_evals[b].append(x_new[b].copy())
            self.f_evals[b].append(float(new_f[b]))
            self.oob_evals[b].append(bool(oob[b]))
            self._update_best_with_last(b)
        self.X[:, n].copy_(torch.from_numpy(x_new))
        self.f[:, n].copy_(torch.from_numpy(new_f))
        self.stats["h2d_bytes"] += B * D * 8 + B * 8
        self.last.update(z_new=z_host, x_new=x_new, new_f=new_f, oob=oob, acq_val=acq_val.cpu().numpy(),
                         k=self.k_host.copy())
        self.n = n + 1
        self._phase("inverse_map_objective_bookkeeping", tp)
        return True


    def _step_end_async(self, z_new, acq_val):
        """The tail of PCA_BO.__call__ (:239-281) without a host round trip: inverse map, objective, penalty, append and
        best-so-far on the device; ONE packed record per iteration is read back asynchronously for the host lists."""
        B, D, n = self.B, self.D, self.n
        if not self.best_dev_valid:
            self._set_best_f()
        if self.algorithm == "pca":
            self.im_out = ops.inverse_map(z_new, self.pca["k"], self.pca["V"], self.pca["m2"], self.pca["xbar"], self.lo,
                                          self.hi, out=self.im_out)
            x_dev, oob = self.im_out["x"], self.im_out["oob"]
            self.stats["kernel_launches"] += 1
        else:
            x_dev, oob = z_new, None
        f_dev = self.batched_problem(x_dev)
        ops.append_observation(self.X, self.f, n, x_dev, z_new, f_dev, oob, acq_val,
                               -1000.0 if self.maximization else 1000.0, self.maximization,   # PCA_BO.py:260-261
                               self.best_dev, self.gp.best_f, self.rec_dev)
        rec = self.rec_free.pop() if self.rec_free else torch.zeros(B, 2 * D + 3, dtype=torch.float64).pin_memory()
        rec.copy_(self.rec_dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.pending_host.append((n, rec, ev, self.k_host.copy()))
        self.stats["kernel_launches"] += 2
        self.stats["d2h_bytes"] += rec.numel() * 8
        self.n = n + 1
        self._drain(False)

    def _drain(self, wait):
        """Bring the host lists (x_evals, f_evals, oob_evals, best tracking, `last`) up to date with the iterations whose
        packed records have arrived; wait=True blocks until all of them have."""
        D = self.D
        while self.pending_host:
            n, rec, ev, k = self.pending_host[0]
            if wait:
                ev.synchronize()
            elif not ev.query():
                break
            self.pending_host.pop(0)
            r = rec.numpy()
            for b in range(self.B):
                self.x_evals[b].append(r[b, :D].copy())
                self.f_evals[b].append(float(r[b, 2 * D]))
                self.oob_evals[b].append(bool(r[b, 2 * D + 1] != 0.0))
                self._update_best_with_last(b)
            self.last.update(z_new=r[:, D:2 * D].copy(), x_new=r[:, :D].copy(), new_f=r[:, 2 * D].copy(),
                             oob=r[:, 2 * D + 1] != 0.0, acq_val=r[:, 2 * D + 2].copy(), k=k)
            self.rec_free.append(rec)

    def step(self):
        """One BO iteration for all runs (PCA_BO.py:178-281)."""
        if not self.step_begin():
            return False
        ok = self.step_end()
        self._drain(True)
        return ok

    def run(self, max_iterations=None):
        if self.n == 0:
            self.initial_design()
        it = 0
        while self.n < self.budget and (max_iterations is None or it < max_iterations):
            self.step()
            it += 1
        return self


def schedule_sub_batches(items, k, context, pause=lambda: sleep(2e-5)):
    """The event-driven schedule of PipelinedEngines.run_steps (pure host logic: the engines only have to offer
    step_launch / step_begin / step_end / wait_event). items: [(engine, token)], `context(token)` a context manager entered
    around every engine call (the sub-batch's CUDA stream), `pause()` what to do when nothing has completed. Every engine
    gets exactly k iterations, each as A (step_launch) -> B (step_begin) -> E (step_end), B and E only once the event of
    the phase before has completed -- unless nothing at all is ready, then the loop pauses and looks again. Returns the
    seconds spent pausing."""
    remaining = {id(e): k for e, _ in items}
    a_wait, o_wait = [], []          # sub-batches with phase A / with the optimiser kernel in flight, oldest first
    idle = 0.0
    for item in items:
        with context(item[1]):
            item[0].step_launch()
        a_wait.append(item)

    def done(item):
        ev = item[0].wait_event()
        return ev is None or ev.query()

    while a_wait or o_wait:
        i = next((i for i, it in enumerate(a_wait) if done(it)), None)
        if i is not None:
            item = a_wait.pop(i)
            with context(item[1]):
                item[0].step_begin()
            o_wait.append(item)
            continue
        i = next((i for i, it in enumerate(o_wait) if done(it)), None)
        if i is not None:
            item = o_wait.pop(i)
            e = item[0]
            with context(item[1]):
                e.step_end()
                remaining[id(e)] -= 1
                if remaining[id(e)] > 0:
                    e.step_launch()
                    a_wait.append(item)
            continue
        # nothing has completed: poll (blocking on one particular event would leave a sub-batch that finishes earlier
        # than that one unattended)
        t0 = perf_counter()
        while not any(done(it) for it in a_wait) and not any(done(it) for it in o_wait):
            pause()
        idle += perf_counter() - t0
    return idle


class PipelinedEngines:
    """S independent sub-batches of runs, each a BatchedBOEngine on its own CUDA stream, interleaved by ONE host thread.

    Runs are independent, so splitting the per-GPU batch changes no result. The optimiser kernel is latency
    bound (a sub-batch of 30 runs takes as long as one of 90, and kernels of different streams share the SMs), so while
    the kernel of one sub-batch is in flight the host prepares the next sub-batch (model fit, raw samples, multinomial)
    and launches its kernel too: step_begin() of every engine returns with its kernel running, step_end() collects."""

    def __init__(self, engines, async_tail=True):
        self.engines = list(engines)
        # the sub-batches' own streams are high priority, the optimiser kernels go to low-priority side streams: when an
        # SM frees up, the short kernels of a model fit get it before the queued blocks of another optimiser kernel
        self.streams = [torch.cuda.Stream(device=e.device, priority=-1) for e in self.engines]
        for e in self.engines:
            e.opt_stream = torch.cuda.Stream(device=e.device, priority=0)
        # the engines' buffers were allocated and zero-filled on the stream that was current at construction: order the
        # side streams after that work before their first use
        for e, st in zip(self.engines, self.streams):
            st.wait_stream(torch.cuda.current_stream(e.device))
            # device objective + device optimiser: nothing in the tail of an iteration needs the host
            e.async_tail = bool(async_tail) and e.batched_problem is not None and e.optimizer == "device"
        self.B = sum(e.B for e in self.engines)
        self.idle_s = 0.0       # host time spent polling with nothing to do

    def initial_design(self):
        for e, st in zip(self.engines, self.streams):
            with torch.cuda.stream(st):
                e.initial_design()
        torch.cuda.synchronize(self.engines[0].device)

    def run_steps(self, k):
        """k lock-step iterations of every sub-batch, interleaved by ONE host thread. Per sub-batch and iteration the
        host does three short things: E (after its optimiser kernel: enqueue the tail of the iteration), A (enqueue the
        next iteration's model fit and raw-sample scoring) and B (read k / flags of phase A, launch the optimiser
        kernel). With the asynchronous tail none of them waits for the GPU, so the schedule is event driven: launch the
        optimiser of any sub-batch whose phase A has completed, else serve (E + A) any sub-batch whose optimiser has
        finished, else poll until one of the outstanding events completes (schedule_sub_batches)."""
        if k <= 0:
            return

        # (every call of a sub-batch runs with its own stream current)
        self.idle_s += schedule_sub_batches(list(zip(self.engines, self.streams)), k, torch.cuda.stream)
        for st in self.streams:
            st.synchronize()
        for e in self.engines:
            e._drain(True)

    def step(self):
        self.run_steps(1)
        return self.engines[0].n < self.engines[0].budget

    @property
    def n(self):
        return self.engines[0].n

    @property
    def stats(self):
        out = {}
        for e in self.engines:
            for k, v in e.stats.items():
                out[k] = out.get(k, 0) + v
        out["idle_poll_s"] = self.idle_s
        return out
