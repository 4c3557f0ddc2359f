"""BO base class + Latin-hypercube sampler with the reference's surface
(Algorithms/BayesianOptimization/AbstractBayesianOptimizer.py).

pyDOE 0.3.8 is not a dependency here: its `lhs` criteria are restated below on the GLOBAL legacy numpy
generator, so a seeded run draws exactly the reference's DoE (criterion "center" is pinned bit-exactly against
the reference's shipped IOHprofiler files; "maximin", "centermaximin" and "correlation" follow pyDOE's
published algorithm and are unpinned).
"""
from abc import abstractmethod
from typing import List, Optional, Union

import numpy as np

from ..AbstractAlgorithm import AbstractAlgorithm


def _lhs_classic(n: int, samples: int) -> np.ndarray:
    cut = np.linspace(0, 1, samples + 1)
    u = np.random.rand(samples, n)
    a, b = cut[:samples], cut[1:samples + 1]
    rd = np.zeros_like(u)
    for j in range(n):
        rd[:, j] = u[:, j] * (b - a) + a
    H = np.zeros_like(rd)
    for j in range(n):
        order = np.random.permutation(range(samples))
        H[:, j] = rd[order, j]
    return H


def _lhs_centered(n: int, samples: int) -> np.ndarray:
    cut = np.linspace(0, 1, samples + 1)
    np.random.rand(samples, n)  # pyDOE draws (and never uses) this matrix
    center = (cut[:samples] + cut[1:samples + 1]) / 2
    H = np.zeros((samples, n))
    for j in range(n):
        H[:, j] = np.random.permutation(center)
    return H


def _pdist(x: np.ndarray) -> np.ndarray:
    m = x.shape[0]
    return np.array([np.sqrt(np.sum((x[j] - x[i]) ** 2)) for i in range(m - 1) for j in range(i + 1, m)])


def lhs(n: int, samples: Optional[int] = None, criterion: Optional[str] = None, iterations: int = 5) -> np.ndarray:
    """pyDOE.lhs(n, samples, criterion, iterations)."""
    samples = n if samples is None else samples
    if criterion is None:
        return _lhs_classic(n, samples)
    crit = criterion.lower()
    if crit in ("center", "c"):
        return _lhs_centered(n, samples)
    if crit in ("maximin", "m", "centermaximin", "cm"):
        gen = _lhs_classic if crit in ("maximin", "m") else _lhs_centered
        maxdist, H = 0, None
        for _ in range(iterations):
            cand = gen(n, samples)
            d = _pdist(cand)
            if maxdist < np.min(d):
                maxdist = np.min(d)
                H = cand.copy()
        return H
    if crit in ("correlation", "corr"):
        mincorr, H = np.inf, None
        for _ in range(iterations):
            cand = _lhs_classic(n, samples)
            R = np.corrcoef(cand)
            if np.max(np.abs(R[R != 1])) < mincorr:
                mincorr = np.max(np.abs(R - np.eye(R.shape[0])))
                H = cand.copy()
        return H
    raise ValueError(f'Invalid value for "criterion": {criterion}')


class LHS_sampler:
    """Latin Hypercube Sampling wrapper (reference :8-103): same criteria names, shorthands and errors."""

    __default_criteria = ("center", "maximin", "centermaximin", "correlation")
    __reduced_criteria = {"c": "center", "m": "maximin", "cm": "centermaximin", "corr": "correlation"}

    def __init__(self, criterion: str = "correlation", iterations: int = 1000, sample_zero: bool = False):
        self.criterion = criterion
        self.iterations = iterations
        self.sample_zero = sample_zero

    def __call__(self, dim: int, n_samples: int) -> np.ndarray:
        points = lhs(n=dim, samples=n_samples, criterion=self.criterion, iterations=self.iterations)
        if self.sample_zero:
            points[0, :] = np.zeros_like(points[0, :])
        return points.reshape((n_samples, dim))

    @property
    def criterion(self) -> str:
        return self.__criterion

    @criterion.setter
    def criterion(self, new_criterion: str) -> None:
        if not isinstance(new_criterion, str):
            raise ValueError("The new criterion is not a string!")
        new_criterion = new_criterion.lower().strip()
        if new_criterion in self.__default_criteria:
            self.__criterion = new_criterion
        elif new_criterion in self.__reduced_criteria:
            self.__criterion = self.__reduced_criteria[new_criterion]
        else:
            raise ValueError("The criterion is not matching the set ones!")

    @property
    def iterations(self) -> int:
        return self.__iterations

    @iterations.setter
    def iterations(self, new_n_iter: int) -> None:
        if new_n_iter > 0:
            self.__iterations = int(new_n_iter)
        else:
            raise ValueError("Negative iterations not allowed")

    @property
    def sample_zero(self) -> bool:
        return self.__sample_zero

    @sample_zero.setter
    def sample_zero(self, new_change: bool) -> None:
        self.__sample_zero = new_change


class AbstractBayesianOptimizer(AbstractAlgorithm):
    """Shared BO state: budget, DoE, evaluated points, best tracking (reference :106-270)."""

    def __init__(self, budget: int, n_DoE: Optional[int] = 0, **kwargs):
        super().__init__(**kwargs)
        self.budget = budget
        self.n_DoE = n_DoE
        doe_parameters = None
        for key, item in kwargs.items():
            if key.lower().strip() == "doe_parameters":
                doe_parameters = item
        full = self.__build_LHS_parameters(doe_parameters)
        self.__lhs_sampler = LHS_sampler(criterion=full["criterion"], iterations=full["iterations"],
                                         sample_zero=full["sample_zero"])
        self.__x_evals = []
        self.__f_evals = []

    def __str__(self):
        pass

    def __call__(self, problem, dim: int, bounds: np.ndarray, **kwargs) -> None:
        """Initial design (reference :142-176): LHS on the global numpy generator, evaluation, best tracking."""
        super().__call__(problem, dim, bounds, **kwargs)
        if not isinstance(self.n_DoE, int) or self.n_DoE == 0:
            self.n_DoE = self.dimension
        init_points = self._rescale_lhs_points(self.lhs_sampler(self.dimension, self.n_DoE))
        for point in init_points:
            self.__x_evals.append(point)
            self.__f_evals.append(problem(point))
        self.assign_new_best()
        self.number_of_function_evaluations = self.n_DoE
        if self.verbose:
            print("After Initial sampling...",
                  f"Current Best: x:{self.__x_evals[self.current_best_index]} y:{self.current_best}", flush=True)

    def _rescale_lhs_points(self, raw_lhs_points: np.ndarray):
        out = np.empty_like(raw_lhs_points)
        for d in range(self.dimension):
            out[:, d] = (self.bounds[d, 1] - self.bounds[d, 0]) * raw_lhs_points[:, d] + self.bounds[d, 0]
        return out

    @abstractmethod
    def assign_new_best(self):
        self.current_best = max(self.__f_evals) if self.maximization else min(self.__f_evals)
        self.current_best_index = self.__f_evals.index(self.current_best, self.current_best_index)

    def __repr__(self):
        return super(AbstractAlgorithm, self).__repr__()

    def __build_LHS_parameters(self, params_dict: Union[dict, None]) -> dict:
        full = {"criterion": "center", "iterations": 1000, "sample_zero": False}
        if isinstance(params_dict, dict):
            full.update(params_dict)
        return full

    def reset(self) -> None:
        super().reset()
        self.__x_evals = []
        self.__f_evals = []

    @property
    def budget(self) -> int:
        return self.__budget

    @budget.setter
    def budget(self, new_budget: int) -> None:
        assert new_budget > 0
        self.__budget = int(new_budget)

    @property
    def n_DoE(self) -> int:
        return self.__n_DoE

    @n_DoE.setter
    def n_DoE(self, new_n_DOE: int) -> None:
        self.__n_DoE = int(new_n_DOE) if new_n_DOE >= 0 else None

    @property
    def lhs_sampler(self) -> LHS_sampler:
        return self.__lhs_sampler

    @property
    def f_evals(self) -> List[float]:
        return self.__f_evals

    @property
    def x_evals(self) -> List[np.ndarray]:
        return self.__x_evals
