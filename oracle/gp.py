"""GP surrogate + analytic acquisition, restated on the CPU in torch FP64 (oracle, test infrastructure).

What the reference builds every iteration (Algorithms/BayesianOptimization/PCA_BO.py:502-545,
Vanilla_BO.py:166-195) is a botorch 0.13.0 `SingleTaskGP(train_z, train_obj,
covar_module=MaternKernel(2.5), outcome_transform=Standardize(m=1), input_transform=Normalize(d, bounds))`
that is NEVER fitted (no fit_gpytorch_mll anywhere) and is then queried through
`LogExpectedImprovement(model, best_f, maximize)` (PCA_BO.py:199-203, class map :679-686).
botorch / gpytorch / linear_operator are not installed and cannot be, so their published
semantics are restated here op-for-op:

  * Normalize            : Xn = (X - lo) / (hi - lo)                      (botorch transforms/input.py)
  * Standardize(m=1)     : y_s = (y - mean) / std (unbiased; std < 1e-8 -> 1) (transforms/outcome.py)
  * prior                : zero mean; Matern-5/2, one lengthscale softplus(0) = ln 2, no outputscale;
                           Gaussian noise sigma^2 = mode of LogNormal(-4, 1) = exp(-5)
  * K(Zn, Zn)            : gpytorch MaternKernel.forward: centre by the column mean of x1, divide by the
                           lengthscale, squared distance by the ||a||^2 + ||b||^2 - 2ab GEMM expansion,
                           diagonal zeroed, clamp_min(0), clamp_min(1e-30).sqrt()
  * caches (n <= 800)    : L = chol(K + sigma^2 I); alpha = K^-1 y_s (Cholesky solve);
                           R = (L^-1)^T via solve_triangular against the identity (fast_pred_var root)
  * k(x*, Zn) for q = 1  : x1 = x* is centred by its own mean (-> 0), so the expansion reduces to the
                           direct sum of squares of (Zn - x*) / ell
  * posterior            : mu_s = k* alpha ; v_s = k(x*, x*) - ||k* R||^2 ; mu = m + s mu_s ;
                           var = max(s^2 v_s, 1e-10) (gpytorch min_variance) then max(., 1e-12) (botorch)
  * best_f               : stored as torch.as_tensor(python float) -> float32
  * log-EI               : botorch analytic._log_ei_helper, three branches

Parity status: unpinned by the reference (it has no tests for these quantities).
"""
import math
import torch

LENGTHSCALE = math.log(2.0)          # softplus(0)
NOISE = math.exp(-5.0)               # LogNormalPrior(-4, 1).mode
SQRT5 = math.sqrt(5.0)
MIN_VAR_GPYTORCH = 1e-10
MIN_VAR_BOTORCH = 1e-12
_NEG_INV_SQRT2 = -(0.5 ** 0.5)
_INV_SQRT_2PI = 1.0 / math.sqrt(2.0 * math.pi)
_LOG_2PI = math.log(2.0 * math.pi)
_LOG_SQRT_PI_DIV_2 = math.log(math.pi / 2.0) / 2.0
_LOG2 = math.log(2.0)


# ----------------------------------------------------------------------------- kernel
def _sq_dist_expansion(x1: torch.Tensor, x2: torch.Tensor, x1_eq_x2: bool) -> torch.Tensor:
    """gpytorch.kernels.kernel.sq_dist."""
    adjustment = x1.mean(-2, keepdim=True)
    x1 = x1 - adjustment
    x2 = x1 if x1_eq_x2 else x2 - adjustment
    x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
    x1_pad = torch.ones_like(x1_norm)
    if x1_eq_x2:
        x2_norm, x2_pad = x1_norm, x1_pad
    else:
        x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
        x2_pad = torch.ones_like(x2_norm)
    x1_ = torch.cat([-2.0 * x1, x1_norm, x1_pad], dim=-1)
    x2_ = torch.cat([x2, x2_pad, x2_norm], dim=-1)
    res = x1_.matmul(x2_.transpose(-2, -1))
    if x1_eq_x2 and not x1.requires_grad and not x2.requires_grad:
        res.diagonal(dim1=-2, dim2=-1).fill_(0)
    return res.clamp_min(0)


def matern52(x1: torch.Tensor, x2: torch.Tensor, lengthscale: float = LENGTHSCALE,
             x1_eq_x2: bool = False) -> torch.Tensor:
    """gpytorch MaternKernel(nu=2.5).forward."""
    mean = x1.mean(dim=-2, keepdim=True)
    x1_ = (x1 - mean).div(lengthscale)
    x2_ = (x2 - mean).div(lengthscale)
    distance = _sq_dist_expansion(x1_, x2_, x1_eq_x2).clamp_min(1e-30).sqrt()
    exp_component = torch.exp(-SQRT5 * distance)
    constant_component = (SQRT5 * distance).add(1).add(5.0 / 3.0 * distance ** 2)
    return constant_component * exp_component


# ----------------------------------------------------------------------------- model
class GPModel:
    """The un-fitted SingleTaskGP of PCA_BO.py:535-544 / Vanilla_BO.py:183-195 in eval mode."""

    def __init__(self, train_x, train_obj, norm_bounds=None, lengthscale: float = LENGTHSCALE,
                 noise: float = NOISE):
        """train_x (n,k) and train_obj (n,) or (n,1): array-likes; norm_bounds (2,k) or None (Vanilla_BO,
        whose Normalize has transform_on_train/eval=False, Vanilla_BO.py:188-194)."""
        self.train_x = torch.as_tensor(train_x, dtype=torch.float64).reshape(len(train_x), -1)
        y = torch.as_tensor(train_obj, dtype=torch.float64).reshape(-1)
        self.n, self.k = self.train_x.shape
        self.lengthscale, self.noise = lengthscale, noise
        if norm_bounds is not None:
            b = torch.as_tensor(norm_bounds, dtype=torch.float64)
            self.offset = b[0].clone()
            self.coefficient = (b[1] - b[0]).clone()
            self.coefficient = torch.where(self.coefficient.abs() < 1e-12,
                                           torch.ones_like(self.coefficient), self.coefficient)
        else:
            self.offset = torch.zeros(self.k, dtype=torch.float64)
            self.coefficient = torch.ones(self.k, dtype=torch.float64)
        self.y_mean = y.mean()
        std = y.std() if self.n > 1 else torch.tensor(1.0, dtype=torch.float64)
        self.y_std = std if float(std) >= 1e-8 else torch.tensor(1.0, dtype=torch.float64)
        self.y_s = (y - self.y_mean) / self.y_std
        self.train_xn = self.normalize(self.train_x)
        self._caches = None

    def normalize(self, X: torch.Tensor) -> torch.Tensor:
        return (X - self.offset) / self.coefficient

    def caches(self):
        """K, L, alpha, Linv (built lazily on the first posterior call, as in gpytorch)."""
        if self._caches is None:
            K = matern52(self.train_xn, self.train_xn, self.lengthscale, x1_eq_x2=True)
            K = K + self.noise * torch.eye(self.n, dtype=torch.float64)
            L = torch.linalg.cholesky(K)
            alpha = torch.cholesky_solve(self.y_s.unsqueeze(-1), L).squeeze(-1)
            Linv = torch.linalg.solve_triangular(L, torch.eye(self.n, dtype=torch.float64), upper=False)
            self._caches = (K, L, alpha, Linv)
        return self._caches

    def kss(self) -> float:
        """k(x*, x*) as gpytorch evaluates it (distance clamped to sqrt(1e-30))."""
        d = torch.tensor(1e-30, dtype=torch.float64).sqrt()
        return float(torch.exp(-SQRT5 * d) * ((SQRT5 * d).add(1).add(5.0 / 3.0 * d ** 2)))

    def posterior(self, X: torch.Tensor):
        """X (b,k) raw-space points, each its own t-batch with q=1. Returns (mean, variance) (b,)."""
        _, _, alpha, Linv = self.caches()
        Xn = self.normalize(X)
        # q = 1: x1 = x* centred by itself -> 0, x2 = (Zn - x*) / ell; expansion == direct sum of squares
        diff = (self.train_xn.unsqueeze(0) - Xn.unsqueeze(1)).div(self.lengthscale)      # (b,n,k)
        sq = diff.pow(2).sum(-1).clamp_min(0)
        dist = sq.clamp_min(1e-30).sqrt()
        kstar = torch.exp(-SQRT5 * dist) * ((SQRT5 * dist).add(1).add(5.0 / 3.0 * dist ** 2))  # (b,n)
        mean_s = kstar @ alpha
        root = kstar @ Linv.T
        var_s = self.kss() - root.pow(2).sum(-1)
        mean = self.y_mean + self.y_std * mean_s
        var = (self.y_std ** 2 * var_s).clamp_min(MIN_VAR_GPYTORCH)
        return mean, var

    def mll(self) -> float:
        """Exact marginal log-likelihood of the standardised targets (north_star item, no reference call site)."""
        _, L, alpha, _ = self.caches()
        return float(-0.5 * (self.y_s @ alpha) - torch.log(torch.diagonal(L)).sum()
                     - 0.5 * self.n * _LOG_2PI)


def mll_and_grad(model: "GPModel"):
    """Exact marginal log-likelihood and its gradient w.r.t. (lengthscale, noise) by autograd (north_star item, no
    reference call site: the reference never fits its GP). Direct-difference distances, as the kernels use."""
    ls = torch.tensor(model.lengthscale, dtype=torch.float64, requires_grad=True)
    nz = torch.tensor(model.noise, dtype=torch.float64, requires_grad=True)
    X = model.train_xn
    sq = (X.unsqueeze(0) - X.unsqueeze(1)).pow(2).sum(-1) / ls ** 2
    d = sq.clamp_min(1e-30).sqrt()
    K = torch.exp(-SQRT5 * d) * ((SQRT5 * d).add(1).add(5.0 / 3.0 * d ** 2)) + nz * torch.eye(model.n, dtype=torch.float64)
    L = torch.linalg.cholesky(K)
    alpha = torch.cholesky_solve(model.y_s.unsqueeze(-1), L).squeeze(-1)
    mll = -0.5 * (model.y_s @ alpha) - torch.log(torch.diagonal(L)).sum() - 0.5 * model.n * _LOG_2PI
    g_ls, g_nz = torch.autograd.grad(mll, (ls, nz))
    return float(mll.detach()), float(g_ls), float(g_nz)


# ----------------------------------------------------------------------------- acquisition
def _phi(x):
    return _INV_SQRT_2PI * (-0.5 * x.square()).exp()


def _Phi(x):
    return 0.5 * torch.erfc(_NEG_INV_SQRT2 * x)


def _log1mexp(x):
    return torch.where(-_LOG2 < x, (-x.expm1()).log(), (-x.exp()).log1p())


def log_ei_helper(u: torch.Tensor) -> torch.Tensor:
    """botorch.acquisition.analytic._log_ei_helper (float64 constants)."""
    bound = -1
    u_upper = u.masked_fill(u < bound, bound)
    log_ei_upper = (_phi(u_upper) + u_upper * _Phi(u_upper)).log()
    neg_inv_sqrt_eps = -1e6
    u_lower = u.masked_fill(u > bound, bound)
    u_eps = u_lower.masked_fill(u < neg_inv_sqrt_eps, neg_inv_sqrt_eps)
    w = torch.log(torch.special.erfcx(_NEG_INV_SQRT2 * u_eps) * u_eps.abs()) + _LOG_SQRT_PI_DIV_2
    log_phi_u = -0.5 * (u.square() + _LOG_2PI)
    log_ei_lower = log_phi_u + torch.where(u > neg_inv_sqrt_eps, _log1mexp(w), -2 * u_lower.abs().log())
    return torch.where(u > bound, log_ei_upper, log_ei_lower)


def best_f_as_botorch(best_f: float) -> float:
    """`register_buffer("best_f", torch.as_tensor(best_f))` on a Python float -> float32."""
    return float(torch.as_tensor(float(best_f)))


class LogExpectedImprovement:
    """botorch LogExpectedImprovement(model, best_f, maximize) on the restated model."""

    def __init__(self, model: GPModel, best_f: float, maximize: bool = False):
        self.model = model
        self.best_f = best_f_as_botorch(best_f)
        self.maximize = maximize

    def mean_sigma(self, X):
        mean, var = self.model.posterior(X)
        return mean, var.clamp_min(MIN_VAR_BOTORCH).sqrt()

    def __call__(self, X: torch.Tensor) -> torch.Tensor:
        """X (b,k) or (b,1,k) -> (b,) log-EI values."""
        X = X.reshape(-1, self.model.k) if X.dim() == 3 else X
        mean, sigma = self.mean_sigma(X)
        u = (mean - self.best_f) / sigma
        u = u if self.maximize else -u
        return log_ei_helper(u) + sigma.log()

    def value_and_grad(self, X: torch.Tensor):
        """Values (b,) and d(sum values)/dX (b,k) by autograd, as botorch's gen_candidates_scipy does."""
        X = X.detach().clone().requires_grad_(True)
        vals = self(X)
        (grad,) = torch.autograd.grad(vals.sum(), X)
        return vals.detach(), grad


class ProbabilityOfImprovement(LogExpectedImprovement):
    """botorch ProbabilityOfImprovement(model, best_f, maximize): Phi(u) (the reference's
    "probability_of_improvement", PCA_BO.py:683-684). Registered as non-negative in botorch, which makes
    gen_batch_initial_conditions use initialize_q_batch_nonneg for it."""

    nonnegative = True

    def __call__(self, X):
        X = X.reshape(-1, self.model.k) if X.dim() == 3 else X
        mean, sigma = self.mean_sigma(X)
        u = (mean - self.best_f) / sigma
        u = u if self.maximize else -u
        return _Phi(u)


class ExpectedImprovement(LogExpectedImprovement):
    """exp(logEI) -- imported but unused by the reference (PCA_BO.py:13); kept for the EI sweep (config C4)."""

    def __call__(self, X):
        return super().__call__(X).exp()
