"""Algorithm base class with the reference's public surface (Algorithms/AbstractAlgorithm.py).

Same attribute names, setter guards and error types as the reference; `ioh` is optional here (the reference
hard-imports it at AbstractAlgorithm.py:7-9): an ioh RealSingleObjective is recognised by duck typing
(`meta_data`, `bounds.lb/.ub`), any other callable needs `dim` and `bounds` like in the reference (:84-99).
"""
from abc import ABC, abstractmethod
from collections import defaultdict
from math import inf
from typing import Callable, Dict, List, Optional, Union

import numpy as np
import torch


def _is_ioh_problem(problem) -> bool:
    return hasattr(problem, "meta_data") and hasattr(problem, "bounds") and hasattr(problem.bounds, "lb")


class AbstractAlgorithm(ABC):
    TIME_PROFILES: List[str] = []

    @abstractmethod
    def __init__(self, **kwargs):
        self.__number_of_function_evaluations = 0
        verbose = kwargs.pop("verbose", False)
        maximization = kwargs.pop("maximization", False)
        self.__bounds = np.empty(shape=(1, 2))
        self.verbose = verbose
        self.__maximization = maximization
        self.__current_best_index = 0
        self.__current_best = -inf if self.__maximization else inf
        self.__problem = None
        self.__dimension = None
        self.__random_states = None
        self.random_seed = None
        self._pbar = kwargs.get("pbar")
        self.timing_logs = defaultdict(list)
        for name in self.TIME_PROFILES:
            self.timing_logs[name] = []

    @abstractmethod
    def __call__(self, problem: Callable, dim: Optional[int], bounds: Optional[np.ndarray], **kwargs):
        if _is_ioh_problem(problem):
            self.__problem = problem
            self.dimension = int(problem.meta_data.n_variables)
            opt_type = getattr(problem.meta_data, "optimization_type", 0)
            self.maximization = bool(getattr(opt_type, "value", opt_type))
            self.bounds = problem.bounds
        elif callable(problem):
            self.__problem = problem
            self.dimension = dim
            self.maximization = kwargs.pop("maximization", False)
            if isinstance(bounds, np.ndarray):
                self.bounds = bounds
            else:
                # the reference's isinstance(..., List[float]) check raises TypeError for anything but ndarray
                raise TypeError("The bounds for a callable problem must be given as a numpy.ndarray")
        else:
            raise AttributeError("The problem input is not well defined", name="problem", obj=problem)

    @abstractmethod
    def __str__(self):
        pass

    def __repr__(self):
        return super().__repr__()

    @abstractmethod
    def reset(self):
        self.number_of_function_evaluations = 0
        self.__current_best = -inf if self.maximization else inf
        self.__current_best_index = 0
        self.timing_logs = defaultdict(list)

    # ------------------------------------------------------------------ timing summaries (:127-140)
    @property
    def time_profile_names(self) -> List[str]:
        return list(self.timing_logs.keys())

    @property
    def average_times(self) -> Dict[str, float]:
        return {k: sum(v) / len(v) for k, v in self.timing_logs.items() if v}

    @property
    def total_times(self) -> Dict[str, float]:
        return {k: sum(v) for k, v in self.timing_logs.items()}

    # ------------------------------------------------------------------ guarded properties (:142-297)
    @property
    def number_of_function_evaluations(self) -> int:
        return self.__number_of_function_evaluations

    @number_of_function_evaluations.setter
    def number_of_function_evaluations(self, value: int) -> None:
        if isinstance(value, int) and value >= 0:
            self.__number_of_function_evaluations = value
        else:
            raise ValueError("The number of function evaluations must be a positive integer")

    @property
    def verbose(self) -> bool:
        return self.__verbose

    @verbose.setter
    def verbose(self, value: bool) -> None:
        self.__verbose = bool(value)

    @property
    def dimension(self) -> Union[int, None]:
        return self.__dimension

    @dimension.deleter
    def dimension(self) -> None:
        del self.__dimension

    @dimension.setter
    def dimension(self, value: Union[int, None]) -> None:
        if isinstance(value, int) and value > 0:
            self.__dimension = value
        elif value is None:
            self.__dimension = value
        else:
            raise ValueError("The new dimension is oddly set")

    @property
    def current_best(self) -> float:
        return self.__current_best

    @current_best.setter
    def current_best(self, value: float):
        improved = value >= self.__current_best if self.__maximization else value <= self.__current_best
        if not improved:
            raise ValueError("The assignment is incorrect")
        self.__current_best = value

    @property
    def current_best_index(self) -> int:
        return self.__current_best_index

    @current_best_index.setter
    def current_best_index(self, value: int) -> None:
        if isinstance(value, int) and value >= self.__current_best_index:
            self.__current_best_index = value
        else:
            raise ValueError("Something is wrong with this assignment")

    @property
    def maximization(self) -> bool:
        return self.__maximization

    @maximization.setter
    def maximization(self, value: bool) -> None:
        if self.__maximization != bool(value):
            self.__maximization = bool(value)
            self.__current_best = -inf if self.__maximization else inf

    @property
    def bounds(self) -> np.ndarray:
        return self.__bounds

    @bounds.setter
    def bounds(self, new_bounds):
        if hasattr(new_bounds, "lb") and hasattr(new_bounds, "ub"):      # ioh RealBounds
            lb, ub = new_bounds.lb, new_bounds.ub
            lb = np.full(self.dimension, lb, dtype=float) if np.ndim(lb) == 0 or np.size(lb) == 1 else np.asarray(lb, float)
            ub = np.full(self.dimension, ub, dtype=float) if np.ndim(ub) == 0 or np.size(ub) == 1 else np.asarray(ub, float)
            self.__bounds = np.hstack((lb.reshape(self.dimension, 1), ub.reshape(self.dimension, 1)))
            return
        arr = np.array(new_bounds)
        if arr.size == 2:
            lo, hi = arr.ravel()
            self.__bounds = np.hstack((np.ones((self.dimension, 1)) * lo, np.ones((self.dimension, 1)) * hi))
        elif arr.size % 2 == 0 and arr.size > 2:
            arr = arr.reshape((-1, 2))
            self.__bounds = np.hstack((arr[:, 0].reshape((-1, 1)), arr[:, 1].reshape((-1, 1))))
        else:
            raise AttributeError("The bounds should be a given in pairs", name="bounds")

    @property
    def random_seed(self) -> int:
        return self.__random_seed

    @random_seed.setter
    def random_seed(self, value: int) -> None:
        if isinstance(value, int) and value >= 0:
            self.__random_seed = value
        elif not hasattr(self, "_AbstractAlgorithm__random_seed"):
            self.__random_seed = None

    # ------------------------------------------------------------------ RNG handling (:299-360)
    def impose_random_seed(self) -> None:
        self.save_random_states()
        np.random.seed(self.__random_seed)
        torch.manual_seed(self.__random_seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed(self.__random_seed)
            torch.cuda.manual_seed_all(self.__random_seed)

    def save_random_states(self) -> None:
        self.__random_states = {"numpy_state": np.random.get_state(), "torch_cpu_state": torch.get_rng_state()}
        if torch.cuda.is_available():
            self.__random_states["torch_gpu_states"] = [torch.cuda.get_rng_state(i)
                                                        for i in range(torch.cuda.device_count())]

    def restore_random_states(self) -> None:
        # Faithful to the reference (AbstractAlgorithm.py:345): `hasattr(self, '__random_states')` is always False
        # because of name mangling, so the saved states are never restored and the consumed streams leak.
        if not hasattr(self, "__random_states"):
            return

    def compute_space_volume(self) -> float:
        return np.prod(self.bounds[:, 1] - self.bounds[:, 0])
