"""PCA-assisted Bayesian optimisation with the reference's API (Algorithms/BayesianOptimization/PCA_BO.py),
executed on the B200 hot path.

Same constructor, call protocol, attributes and errors as the reference class; the five operator seams of its hot
loop (`_calculate_weights`, `_transform_points_to_reduced_space`, `_initialize_model`,
`optimize_acqf_and_get_observation`, `_transform_point_to_original_space`; reference PCA_BO.py:183-246) run in the
sm_100a kernels behind include/pcabo_b200.h. There is no CPU fallback.
"""
import os
from typing import Callable, Optional

import numpy as np
import torch

from .AbstractBayesianOptimizer import AbstractBayesianOptimizer
from ._acquisition_names import (ACQUISITION_CLASSES, ALLOWED_ACQUISITION_FUNCTION_STRINGS,
                                 ALLOWED_SHORTHAND_ACQUISITION_FUNCTION_STRINGS, AnalyticAcquisitionFunction)


class FittedPCA:
    """Host view of the device-fitted PCA with sklearn's attribute names (the reference exposes `self.pca`)."""

    def __init__(self, components, mean, explained_variance, explained_variance_ratio, k):
        self.components_ = components[:k]
        self.mean_ = mean
        self.explained_variance_ = explained_variance
        self.explained_variance_ratio_ = explained_variance_ratio
        self.n_components_ = components.shape[0]

    def transform(self, X):
        X = np.asarray(X, dtype=np.float64)
        return X @ self.components_.T - self.mean_.reshape(1, -1) @ self.components_.T

    def inverse_transform(self, Z):
        return np.asarray(Z, dtype=np.float64) @ self.components_ + self.mean_


class PCA_BO(AbstractBayesianOptimizer):
    """PCA-BO (reference PCA_BO.py:48-720)."""

    TIME_PROFILES = ["SingleTaskGP", "optimize_acqf", "pca"]

    def __init__(self, budget: int, n_DoE: int = 0, n_components: int = 0, var_threshold: float = 0.95,
                 acquisition_function: str = "expected_improvement", random_seed: int = 43,
                 visualize: bool = False, **kwargs):
        super().__init__(budget, n_DoE, **kwargs)
        self.random_seed = random_seed
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        smoke_test = os.environ.get("SMOKE_TEST")
        self.__torch_config = {
            "device": device,
            "dtype": torch.double,
            "SMOKE_TEST": smoke_test,
            "BATCH_SIZE": 3 if not smoke_test else 2,
            "NUM_RESTARTS": 10 if not smoke_test else 2,
            "RAW_SAMPLES": 512 if not smoke_test else 32,
        }
        self.__acq_func_class = None
        self.__acq_func = None
        self.acquisition_function_name = acquisition_function
        self.n_components = n_components
        self.var_threshold = var_threshold
        self.data_mean = None
        self.pca = None
        self.component_matrix = None
        self.explained_variance_ratio = None
        self.reduced_space_dim_num = None
        self.visualize = visualize
        if self.visualize:
            raise NotImplementedError("visualize=True needs the reference's matplotlib Visualizer, which is outside "
                                      "the accelerated hot path")
        self.optimizer_backend = kwargs.get("optimizer_backend", "scipy")
        # extension (not in the reference, which never fits its GP): per-iteration MAP fit of (lengthscale, noise)
        self.fit_hyperparameters = bool(kwargs.get("fit_hyperparameters", False))
        self.__z_evals = []
        self._engine = None

    def __str__(self):
        return "This is an instance of a PCA-assisted BO Optimizer"

    def __call__(self, problem: Callable, dim: Optional[int] = -1, bounds: Optional[np.ndarray] = None, **kwargs) -> None:
        from pcabo_b200.engine import BatchedBOEngine
        if self.acquisition_function_name == "upper_confidence_bound":
            # the reference builds acquisition_function_class(model=..., best_f=..., maximize=...) for every choice
            # (PCA_BO.py:199-203); botorch's UpperConfidenceBound has no best_f argument, so this choice fails there too
            raise TypeError("UpperConfidenceBound.__init__() got an unexpected keyword argument 'best_f'")
        original_stdout = None
        if self._pbar is not None:
            from ..utils.tqdm_write_stream import redirect_stdout_to_tqdm
            original_stdout = redirect_stdout_to_tqdm(self._pbar)
        self.impose_random_seed()
        super().__call__(problem, dim, bounds, **kwargs)            # DoE on the global numpy generator
        if self._pbar is not None:
            self._pbar.update(self.n_DoE)
        eng = BatchedBOEngine([problem], self.dimension, self.bounds, self.budget, self.n_DoE, [self.random_seed],
                              algorithm="pca", var_threshold=self.var_threshold, n_components=self.n_components,
                              maximization=self.maximization, optimizer=self.optimizer_backend,
                              num_restarts=self.__torch_config["NUM_RESTARTS"],
                              raw_samples=self.__torch_config["RAW_SAMPLES"],
                              np_rngs=[np.random.mtrand._rand], torch_gens=[torch.default_generator],
                              profile_timing=True, acquisition=self.acquisition_function_name,
                              fit_hyperparameters=self.fit_hyperparameters)
        eng.attach_history([self.x_evals], [self.f_evals])
        self._engine = eng
        for _ in range(self.budget - self.n_DoE):
            if self.number_of_function_evaluations >= self.budget:
                break
            prev_best = self.current_best
            eng.step()                                              # PCA_BO.py:183-272 on the device
            self._pull_iteration_state(eng)
            self.number_of_function_evaluations += 1
            if self._pbar is not None:
                self._pbar.update(1)
            new_f = self.f_evals[-1]
            if self.verbose and ((self.maximization and new_f > prev_best) or
                                 (not self.maximization and new_f < prev_best)):
                print(f"Found better solution: {new_f}")
                print(f"At point: {self.x_evals[-1]}")
            self.assign_new_best()
            if self.verbose:
                print(f"Evaluations: {self.number_of_function_evaluations}/{self.budget}",
                      f"Best: x:{self.x_evals[self.current_best_index]} y:{self.current_best}", flush=True)
        for name in self.TIME_PROFILES:
            self.timing_logs[name] = list(eng.timing_logs[name])
        if self.verbose:
            print("Optimization Process finalized!")
        self.restore_random_states()
        if self._pbar is not None:
            from ..utils.tqdm_write_stream import restore_stdout
            restore_stdout(original_stdout)

    def _pull_iteration_state(self, eng) -> None:
        """Mirror the reference's per-iteration attributes (PCA_BO.py:364,386-399,408) on the host."""
        self._pull_pca_state(eng, int(eng.k_host[0]), len(self.f_evals) - 1)
        k = self.reduced_space_dim_num
        self.__z_evals = self.__z_evals + [eng.last["z_new"][0, :k].copy()]
        self.acquisition_function = self.acquisition_function_class(model=eng, best_f=self.current_best,
                                                                    maximize=self.maximization)

    def _pull_pca_state(self, eng, k: int, n_points: int) -> None:
        p = eng.pca
        comps = p["V"][0].cpu().numpy()
        self.data_mean = p["xbar"][0].cpu().numpy()
        self.component_matrix = np.copy(comps)
        self.explained_variance_ratio = p["evr"][0].cpu().numpy()
        self.reduced_space_dim_num = k
        self.pca = FittedPCA(comps, p["m2"][0].cpu().numpy(), p["lam"][0].cpu().numpy(), self.explained_variance_ratio, k)
        Z = p["Z"][0, :n_points, :k].cpu().numpy()
        self.__z_evals = [Z[i, :] for i in range(Z.shape[0])]
        if self.verbose:
            print(f"Using {k} principal components with "
                  f"{np.sum(self.explained_variance_ratio[:k]) * 100:.2f}% explained variance")

    def assign_new_best(self):
        super().assign_new_best()

    # ------------------------------------------------------------------ operator seams, callable on their own
    def _calculate_weights(self) -> np.ndarray:
        """Rank weights of the current evaluations (PCA_BO.py:316-341), computed by the PCA kernel."""
        st = self._device_pca_fit()
        return st["w"][0, :len(self.f_evals)].cpu().numpy()

    def _device_pca_fit(self):
        from pcabo_b200 import ops
        n, D = len(self.f_evals), self.dimension
        dev = torch.device("cuda:0")
        X = torch.from_numpy(np.vstack(self.x_evals)).reshape(1, n, D).to(dev)
        f = torch.from_numpy(np.asarray(self.f_evals, dtype=np.float64)).reshape(1, n).to(dev)
        noise = torch.zeros(1, n, D, dtype=torch.float64, device=dev)
        logtab = torch.from_numpy(np.log(np.maximum(np.arange(n + 2), 1).astype(np.float64))).to(dev)
        return ops.weights_pca_fit(X, f, noise, logtab, n, self.var_threshold, self.n_components, self.maximization)

    def _seam_engine(self):
        """A one-run engine holding the CURRENT x_evals / f_evals, for callers that drive the reference's operator seams
        one by one (PCA_BO.py:183, :186, :205) instead of through __call__. Draws from the process-global numpy / torch
        generators like the reference does."""
        from pcabo_b200.engine import BatchedBOEngine
        n = len(self.f_evals)
        if n < 2:
            raise ValueError("the accelerated operator seams need at least two evaluated points")
        eng = self._engine
        if eng is None or eng.x_evals[0] is not self.x_evals or eng.n != n or eng.budget < n + 1:
            eng = BatchedBOEngine([None], self.dimension, self.bounds, max(self.budget, n + 1), self.n_DoE,
                                  [self.random_seed], algorithm="pca", var_threshold=self.var_threshold,
                                  n_components=self.n_components, maximization=self.maximization,
                                  optimizer=self.optimizer_backend,
                                  num_restarts=self.torch_config["NUM_RESTARTS"],
                                  raw_samples=self.torch_config["RAW_SAMPLES"],
                                  np_rngs=[np.random.mtrand._rand], torch_gens=[torch.default_generator],
                                  profile_timing=True, acquisition=self.acquisition_function_name)
            eng.attach_history([self.x_evals], [self.f_evals])
            self._engine = eng
        return eng

    def _transform_points_to_reduced_space(self) -> None:
        """PCA_BO.py:343-408: rank weights, weighted PCA fit and projection of all evaluated points (one kernel);
        fills pca, data_mean, component_matrix, explained_variance_ratio, reduced_space_dim_num and z_evals."""
        eng = self._seam_engine()
        eng._fit_pca()
        k = int(eng.pca["k"][0].item())
        eng.k_host[:] = k
        self.timing_logs["pca"].append(eng.timing_logs["pca"][-1])
        self._pull_pca_state(eng, k, len(self.f_evals))

    def _initialize_model(self, **kwargs) -> None:
        """PCA_BO.py:502-545: SingleTaskGP(train_z, train_obj, Matern-5/2, Standardize, Normalize) and, eagerly, the
        prediction caches gpytorch would build on the first posterior call. Needs _transform_points_to_reduced_space()."""
        eng = self._seam_engine()
        if eng.pca is None:
            raise AttributeError("call _transform_points_to_reduced_space() before _initialize_model()")
        t0 = eng._sync_time()
        eng._fit_gp()
        self.timing_logs["SingleTaskGP"].append(eng._sync_time() - t0)
        self.assign_new_best()
        eng.current_best = [self.current_best]
        eng._set_best_f()
        self.acquisition_function = self.acquisition_function_class(model=eng, best_f=self.current_best,
                                                                    maximize=self.maximization)

    def optimize_acqf_and_get_observation(self) -> torch.Tensor:
        """PCA_BO.py:547-621: optimize_acqf(q=1, num_restarts, raw_samples, batch_limit 5, maxiter 200) over the padded
        box of the reduced space. Returns the new point as a CPU tensor of shape (1, k) like the reference."""
        eng = self._seam_engine()
        if eng.gp.n != len(self.f_evals):
            raise AttributeError("call _initialize_model() before optimize_acqf_and_get_observation()")
        t0 = eng._sync_time()
        st = eng._propose_begin(eng._propose_launch())
        z_new, _ = eng._propose_end(st)
        self.timing_logs["optimize_acqf"].append(eng._sync_time() - t0)
        k = int(eng.k_host[0])
        return z_new[:, :k].detach().cpu().reshape(1, k)

    def _transform_point_to_original_space(self, z: np.ndarray) -> np.ndarray:
        """PCA_BO.py:410-434."""
        if self.pca is None:
            return np.random.uniform(self.bounds[:, 0], self.bounds[:, 1])
        return (self.pca.inverse_transform(np.asarray(z).reshape(1, -1)) + self.data_mean).ravel()

    def __repr__(self):
        return super().__repr__()

    def reset(self):
        super().reset()
        self.__z_evals = []
        self.pca = None
        self.explained_variance_ratio = None
        self._engine = None

    @property
    def z_evals(self):
        return self.__z_evals

    @property
    def torch_config(self) -> dict:
        return self.__torch_config

    @property
    def acquisition_function_name(self) -> str:
        return self.__acquisition_function_name

    @acquisition_function_name.setter
    def acquisition_function_name(self, new_name: str) -> None:
        new_name = new_name.strip()
        if new_name in ALLOWED_SHORTHAND_ACQUISITION_FUNCTION_STRINGS:
            self.__acquisition_function_name = ALLOWED_SHORTHAND_ACQUISITION_FUNCTION_STRINGS[new_name]
        elif new_name.lower() in ALLOWED_ACQUISITION_FUNCTION_STRINGS:
            self.__acquisition_function_name = new_name
        else:
            raise ValueError("Oddly defined name")
        self.set_acquisition_function_subclass()

    def set_acquisition_function_subclass(self) -> None:
        self.__acq_func_class = ACQUISITION_CLASSES.get(self.__acquisition_function_name)

    @property
    def acquisition_function_class(self) -> Callable:
        return self.__acq_func_class

    @property
    def acquisition_function(self) -> AnalyticAcquisitionFunction:
        return self.__acq_func

    @acquisition_function.setter
    def acquisition_function(self, new_acquisition_function: AnalyticAcquisitionFunction) -> None:
        if issubclass(type(new_acquisition_function), AnalyticAcquisitionFunction):
            self.__acq_func = new_acquisition_function
        else:
            raise AttributeError("Acquisition function does not inherit from 'AnalyticAcquisitionFunction'",
                                 name="acquisition_function", obj=self.__acq_func)
