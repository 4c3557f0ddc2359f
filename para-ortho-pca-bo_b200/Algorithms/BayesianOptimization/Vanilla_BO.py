"""Vanilla Bayesian optimisation with the reference's API (Algorithms/BayesianOptimization/Vanilla_BO.py) on the
B200 hot path: the same GP / log-EI / optimize_acqf kernels as PCA_BO in the full D-dimensional space, without
input normalisation (Vanilla_BO.py:188-194) and with the problem bounds as the search box (:204)."""
import os
from typing import Callable, Optional

import numpy as np
import torch

from .AbstractBayesianOptimizer import AbstractBayesianOptimizer
from ._acquisition_names import (ACQUISITION_CLASSES, ALLOWED_ACQUISITION_FUNCTION_STRINGS,
                                 ALLOWED_SHORTHAND_ACQUISITION_FUNCTION_STRINGS, AnalyticAcquisitionFunction)


class Vanilla_BO(AbstractBayesianOptimizer):
    TIME_PROFILES = ["SingleTaskGP", "optimize_acqf"]

    def __init__(self, budget: int, n_DoE: int = 0, acquisition_function: str = "expected_improvement",
                 random_seed: int = 43, **kwargs):
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
        self.__acq_func = None
        self.acquisition_function_name = acquisition_function
        self.optimizer_backend = kwargs.get("optimizer_backend", "scipy")
        self._engine = None

    def __str__(self):
        return "This is an instance of a Vanilla BO Optimizer"

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
        super().__call__(problem, dim, bounds, **kwargs)
        if self._pbar is not None:
            self._pbar.update(self.n_DoE)
        eng = BatchedBOEngine([problem], self.dimension, self.bounds, self.budget, self.n_DoE, [self.random_seed],
                              algorithm="vanilla", maximization=self.maximization, optimizer=self.optimizer_backend,
                              num_restarts=self.__torch_config["NUM_RESTARTS"],
                              raw_samples=self.__torch_config["RAW_SAMPLES"],
                              np_rngs=[np.random.mtrand._rand], torch_gens=[torch.default_generator],
                              profile_timing=True, acquisition=self.acquisition_function_name)
        eng.attach_history([self.x_evals], [self.f_evals])
        self._engine = eng
        for _ in range(self.budget - self.n_DoE):
            if self.number_of_function_evaluations >= self.budget:
                break
            eng.step()
            self.acquisition_function = self.acquisition_function_class(model=eng, best_f=self.current_best,
                                                                        maximize=self.maximization)
            self.number_of_function_evaluations += 1
            if self._pbar is not None:
                self._pbar.update(1)
            self.assign_new_best()
            if self.verbose:
                print(f"Evaluations: {self.number_of_function_evaluations}/{self.budget}",
                      f"Best: x:{self.x_evals[self.current_best_index]} y:{self.current_best}", flush=True)
        for name in self.TIME_PROFILES:
            self.timing_logs[name] = list(eng.timing_logs[name])
        if self.verbose:
            print("Optimisation Process finalized!")
        self.restore_random_states()
        if self._pbar is not None:
            from ..utils.tqdm_write_stream import restore_stdout
            restore_stdout(original_stdout)

    def assign_new_best(self):
        super().assign_new_best()

    def __repr__(self):
        return super().__repr__()

    def reset(self):
        super().reset()
        self._engine = None

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
            raise ValueError(f"Oddly defined name {new_name}")
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
