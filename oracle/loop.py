"""The PCA-BO / vanilla-BO outer loops restated on the CPU (oracle, test infrastructure).

Follows Algorithms/BayesianOptimization/PCA_BO.py:140-310 (`PCA_BO.__call__`),
Vanilla_BO.py:81-159, AbstractBayesianOptimizer.py:142-208 (DoE, best tracking) and
AbstractAlgorithm.py:310-328 (seeding). Like the reference it uses the GLOBAL numpy and torch
generators after `np.random.seed(seed); torch.manual_seed(seed)`; the draw order per run is
rand(n_DoE, D) (discarded) -> D permutations -> per iteration normal(0, 1e-8, (n, D)) [numpy]
and Sobol scramble randint's -> multinomial [torch].
"""
from time import perf_counter
import numpy as np
import torch

from . import doe, weights_pca, gp, acqf


class OracleRun:
    def __init__(self, problem, dim, bounds, budget, n_doe, seed, algorithm="pca", var_threshold=0.95,
                 n_components=0, maximization=False, num_restarts=acqf.NUM_RESTARTS,
                 raw_samples=acqf.RAW_SAMPLES, keep_trace=False, acquisition="expected_improvement"):
        self.problem, self.dim, self.bounds = problem, dim, np.asarray(bounds, dtype=np.float64)
        self.budget, self.n_doe, self.seed = budget, n_doe if n_doe > 0 else dim, seed
        self.algorithm, self.var_threshold, self.n_components = algorithm, var_threshold, n_components
        self.maximization = maximization
        # PCA_BO.py:679-686: name -> botorch class ("expected_improvement" is LogExpectedImprovement)
        self.acq_class = {"expected_improvement": gp.LogExpectedImprovement,
                          "probability_of_improvement": gp.ProbabilityOfImprovement}[acquisition]
        self.num_restarts, self.raw_samples = num_restarts, raw_samples
        self.x_evals, self.f_evals, self.z_evals = [], [], []
        self.current_best, self.current_best_index = (-np.inf if maximization else np.inf), 0
        self.timing = {"pca": [], "SingleTaskGP": [], "optimize_acqf": []}
        self.keep_trace, self.trace = keep_trace, []
        self.pca_state = None

    # AbstractBayesianOptimizer.py:196-208
    def assign_new_best(self):
        self.current_best = max(self.f_evals) if self.maximization else min(self.f_evals)
        self.current_best_index = self.f_evals.index(self.current_best, self.current_best_index)

    def initial_design(self):
        np.random.seed(self.seed)            # AbstractAlgorithm.py:320-324
        torch.manual_seed(self.seed)
        pts = doe.rescale(doe.lhs_center(self.dim, self.n_doe), self.bounds)
        for p in pts:
            self.x_evals.append(p)
            self.f_evals.append(self.problem(p))
        self.assign_new_best()

    def iteration(self):
        """One pass of the hot loop, PCA_BO.py:178-281 (or Vanilla_BO.py:103-141)."""
        rec = {} if self.keep_trace else None
        inner = [] if self.keep_trace else None
        if self.algorithm == "pca":
            X = np.vstack(self.x_evals)
            noise = np.random.normal(0, 1e-8, size=X.shape)             # PCA_BO.py:376
            t0 = perf_counter()
            st = weights_pca.weighted_pca_fit(X, self.f_evals, noise, self.var_threshold,
                                              self.n_components, self.maximization)
            self.timing["pca"].append(perf_counter() - t0)
            self.pca_state = st
            Z = st["Z"]
            self.z_evals = [Z[i, :] for i in range(Z.shape[0])]
            t0 = perf_counter()
            model = gp.GPModel(Z, self.f_evals, norm_bounds=acqf.normalize_box_pca(Z))
            self.timing["SingleTaskGP"].append(perf_counter() - t0)
            box = acqf.search_box_pca(Z)
        else:
            X = np.array(self.x_evals).reshape((-1, self.dim))
            t0 = perf_counter()
            model = gp.GPModel(X, self.f_evals, norm_bounds=None)
            self.timing["SingleTaskGP"].append(perf_counter() - t0)
            box = self.bounds.transpose().copy()
        acq = self.acq_class(model, best_f=self.current_best, maximize=self.maximization)
        t0 = perf_counter()
        cand, val = acqf.optimize_acqf(acq, box, self.num_restarts, self.raw_samples, trace=inner)
        self.timing["optimize_acqf"].append(perf_counter() - t0)
        new_z = cand.detach().numpy().ravel()
        if self.algorithm == "pca":
            st = self.pca_state
            new_x = weights_pca.inverse_map(new_z, st["components"][:st["k"]], st["pca_mean"], st["data_mean"])
            outside = (not np.all(new_x >= self.bounds[:, 0])) or (not np.all(new_x <= self.bounds[:, 1]))
            self.x_evals.append(new_x)
            self.z_evals.append(new_z)
            new_f = (-1000 if self.maximization else 1000) if outside else self.problem(new_x)   # :260-263
        else:
            new_x, outside = new_z, False
            self.x_evals.append(new_x)
            new_f = self.problem(new_x)
        self.f_evals.append(new_f)
        self.assign_new_best()
        if rec is not None:
            rec.update(n=len(self.f_evals) - 1, new_z=new_z.copy(), new_x=np.array(new_x).copy(), new_f=float(new_f),
                       outside=bool(outside), acq_value=float(val), box=box.copy(), inner=inner,
                       k=(self.pca_state["k"] if self.algorithm == "pca" else self.dim))
            self.trace.append(rec)
        return new_x, new_f

    def run(self, max_iterations=None):
        self.initial_design()
        it = 0
        while len(self.f_evals) < self.budget and (max_iterations is None or it < max_iterations):
            self.iteration()
            it += 1
        return self
