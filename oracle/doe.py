"""Design of experiments: pyDOE 0.3.8 `lhs(criterion="center")` restated (oracle, test infrastructure).

Follows Algorithms/BayesianOptimization/AbstractBayesianOptimizer.py:30-51 (LHS_sampler.__call__),
:142-168 (DoE loop) and :178-194 (_rescale_lhs_points). pyDOE is not installed; its `_lhscentered`
is restated from the published source (doe_lhs.py): it draws and discards `np.random.rand(n, d)`,
then permutes the n bin centres independently per column with `np.random.permutation`.
Pinned bit-exactly against the reference's shipped IOHprofiler .dat files (tests/test_oracle_golden.py).
"""
import numpy as np


def lhs_center(dim: int, n_samples: int) -> np.ndarray:
    """pyDOE.lhs(n=dim, samples=n_samples, criterion='center') using the GLOBAL legacy numpy RNG."""
    cut = np.linspace(0, 1, n_samples + 1)
    _ = np.random.rand(n_samples, dim)  # drawn and discarded by pyDOE
    a = cut[:n_samples]
    b = cut[1:n_samples + 1]
    center = (a + b) / 2
    H = np.zeros((n_samples, dim))
    for j in range(dim):
        H[:, j] = np.random.permutation(center)
    return H


def rescale(raw: np.ndarray, bounds: np.ndarray) -> np.ndarray:
    """AbstractBayesianOptimizer.py:178-194: multiplier * raw + lower, column by column."""
    out = np.empty_like(raw)
    for d in range(raw.shape[1]):
        mult = bounds[d, 1] - bounds[d, 0]
        out[:, d] = mult * raw[:, d] + bounds[d, 0]
    return out


def run_seed(fid: int, dim: int, instance: int) -> int:
    """Algorithms/Experiment/ExperimentRunner.py:146."""
    return 1000 * fid + 10 * dim + instance


def budget_and_ndoe(dim: int, budget_factor: int = 10, doe_factor: float = 3.0):
    """Algorithms/Experiment/ExperimentRunner.py:144-145."""
    return budget_factor * dim + 50, int(doe_factor * dim)
