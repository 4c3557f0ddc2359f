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


def _slot(name: str, convert=None, doc: str = ""):
    """A read/write attribute kept in `name` whose setter runs `convert` (validation + normalisation) first: the
    reference spells every one of these out as a property/setter pair; the behaviour (values accepted, exceptions raised)
    is the same."""
    def fget(self):
        return getattr(self, name)

    def fset(self, value):
        setattr(self, name, value if convert is None else convert(value))

    return property(fget, fset, doc=doc)


_LHS_CRITERIA = ("center", "maximin", "centermaximin", "correlation")
_LHS_SHORTHANDS = {"c": "center", "m": "maximin", "cm": "centermaximin", "corr": "correlation"}


def _as_criterion(value) -> str:
    if not isinstance(value, str):
        raise ValueError("The new criterion is not a string!")
    key = value.lower().strip()
    if key in _LHS_CRITERIA:
        return key
    if key in _LHS_SHORTHANDS:
        return _LHS_SHORTHANDS[key]
    raise ValueError("The criterion is not matching the set ones!")


def _as_iterations(value) -> int:
    if not value > 0:
        raise ValueError("Negative iterations not allowed")
    return int(value)


class LHS_sampler:
    """Latin Hypercube Sampling wrapper (reference :8-103): same criteria names, shorthands and errors."""

    criterion = _slot("_lhs_criterion", _as_criterion, "pyDOE criterion (full name)")
    iterations = _slot("_lhs_iterations", _as_iterations, "iterations of the maximin / correlation searches")
    sample_zero = _slot("_lhs_sample_zero", None, "replace the first design point by the origin")

    def __init__(self, criterion: str = "correlation", iterations: int = 1000, sample_zero: bool = False):
        self.criterion, self.iterations, self.sample_zero = criterion, iterations, sample_zero

    def __call__(self, dim: int, n_samples: int) -> np.ndarray:
        design = lhs(n=dim, samples=n_samples, criterion=self.criterion, iterations=self.iterations)
        if self.sample_zero:
            design[0, :] = 0.0
        return design.reshape((n_samples, dim))


def _as_budget(value) -> int:
    assert value > 0
    return int(value)


def _as_n_doe(value):
    return int(value) if value >= 0 else None


class AbstractBayesianOptimizer(AbstractAlgorithm):
    """Shared BO state: budget, DoE, evaluated points, best tracking (reference :106-270)."""

    budget = _slot("_bo_budget", _as_budget, "total number of objective evaluations")
    n_DoE = _slot("_bo_n_doe", _as_n_doe, "size of the initial design (0: the dimension)")
    lhs_sampler = property(lambda self: self._bo_sampler)
    f_evals = property(lambda self: self._bo_f)
    x_evals = property(lambda self: self._bo_x)

    def __init__(self, budget: int, n_DoE: Optional[int] = 0, **kwargs):
        super().__init__(**kwargs)
        self.budget, self.n_DoE = budget, n_DoE
        settings = {"criterion": "center", "iterations": 1000, "sample_zero": False}      # reference :121-133
        for key, item in kwargs.items():
            if key.lower().strip() == "doe_parameters" and isinstance(item, dict):
                settings.update(item)
        self._bo_sampler = LHS_sampler(**{k: settings[k] for k in ("criterion", "iterations", "sample_zero")})
        self._bo_x, self._bo_f = [], []

    def __str__(self):
        pass

    def __call__(self, problem, dim: int, bounds: np.ndarray, **kwargs) -> None:
        """Initial design (reference :142-176): LHS on the global numpy generator, evaluation, best tracking."""
        super().__call__(problem, dim, bounds, **kwargs)
        if not isinstance(self.n_DoE, int) or self.n_DoE == 0:
            self.n_DoE = self.dimension
        for point in self._rescale_lhs_points(self.lhs_sampler(self.dimension, self.n_DoE)):
            self._bo_x.append(point)
            self._bo_f.append(problem(point))
        self.assign_new_best()
        self.number_of_function_evaluations = self.n_DoE
        if self.verbose:
            print("After Initial sampling...",
                  f"Current Best: x:{self._bo_x[self.current_best_index]} y:{self.current_best}", flush=True)

    def _rescale_lhs_points(self, raw_lhs_points: np.ndarray):
        lo, hi = self.bounds[:self.dimension, 0], self.bounds[:self.dimension, 1]
        return (hi - lo)[None, :] * raw_lhs_points + lo[None, :]

    @abstractmethod
    def assign_new_best(self):
        pick = max if self.maximization else min
        self.current_best = pick(self._bo_f)
        self.current_best_index = self._bo_f.index(self.current_best, self.current_best_index)

    def __repr__(self):
        return super(AbstractAlgorithm, self).__repr__()

    def reset(self) -> None:
        super().reset()
        self._bo_x, self._bo_f = [], []
