"""Drop-in mirror of the reference's `Algorithms` package on the B200 hot path.

`from Algorithms import Vanilla_BO, PCA_BO, ExperimentRunner` works as it does with the reference (its
Algorithms/__init__.py:1-3 re-exports exactly these three names). The classes are resolved on first use, so importing the
package does not pull in torch or the CUDA library until an optimiser is actually needed.
"""
import importlib

_EXPORTS = {
    "Vanilla_BO": ".BayesianOptimization.Vanilla_BO",
    "PCA_BO": ".BayesianOptimization.PCA_BO",
    "ExperimentRunner": ".Experiment.ExperimentRunner",
}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    if name in _EXPORTS:
        value = getattr(importlib.import_module(_EXPORTS[name], __name__), name)
        globals()[name] = value
        return value
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
