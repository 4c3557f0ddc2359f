"""Acquisition-name plumbing shared by PCA_BO and Vanilla_BO (reference PCA_BO.py:34-45,643-720)."""

ALLOWED_ACQUISITION_FUNCTION_STRINGS = (
    "expected_improvement",
    "probability_of_improvement",
    "upper_confidence_bound",
)

ALLOWED_SHORTHAND_ACQUISITION_FUNCTION_STRINGS = {
    "EI": "expected_improvement",
    "PI": "probability_of_improvement",
    "UCB": "upper_confidence_bound",
}


class AnalyticAcquisitionFunction:
    """Marker base class standing in for botorch's AnalyticAcquisitionFunction in the `acquisition_function`
    setter check (reference PCA_BO.py:702-720)."""


class LogExpectedImprovement(AnalyticAcquisitionFunction):
    """Handle to the device-resident log-EI of the current iteration.

    `"expected_improvement"` maps to botorch's LogExpectedImprovement in the reference (PCA_BO.py:681-682);
    calling this object evaluates the sm_100a log-EI kernel on raw-space points `X` of shape (b, k) or (b, 1, k)
    and returns a CPU tensor of b values, like `acq(X)` does in botorch.
    """

    def __init__(self, model=None, best_f=None, maximize=False):
        self.model, self.best_f, self.maximize = model, best_f, maximize

    def __call__(self, X):
        import torch
        from pcabo_b200 import ops
        eng = self.model
        k = int(eng.k_host[0])
        X = torch.as_tensor(X, dtype=torch.float64).reshape(-1, k)
        pts = torch.zeros(1, X.shape[0], eng.D, dtype=torch.float64)
        pts[0, :, :k] = X
        return ops.logei(eng.gp, pts.to(eng.device))["val"][0].cpu()


class ProbabilityOfImprovement(LogExpectedImprovement):
    """Handle to the device-resident probability of improvement Phi(u) (botorch ProbabilityOfImprovement, the
    reference's "probability_of_improvement", PCA_BO.py:683-684); evaluated by the same kernel with acq = PCABO_ACQ_PI."""


class UpperConfidenceBound(AnalyticAcquisitionFunction):
    pass


ACQUISITION_CLASSES = {
    "expected_improvement": LogExpectedImprovement,
    "probability_of_improvement": ProbabilityOfImprovement,
    "upper_confidence_bound": UpperConfidenceBound,
}
