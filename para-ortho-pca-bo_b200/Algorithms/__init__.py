"""Drop-in mirror of the reference's `Algorithms` package (Algorithms/__init__.py:1-3) on the B200 hot path."""
from .BayesianOptimization.Vanilla_BO import Vanilla_BO
from .BayesianOptimization.PCA_BO import PCA_BO
from .Experiment.ExperimentRunner import ExperimentRunner

__all__ = ["Vanilla_BO", "PCA_BO", "ExperimentRunner"]
