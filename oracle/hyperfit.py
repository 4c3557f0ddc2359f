"""CPU twin of pcabo_b200/hyperfit.py (TEST INFRASTRUCTURE, like everything under oracle/).

What botorch's fit_gpytorch_mll would do for the reference's SingleTaskGP (PCA_BO.py:535-544) -- the reference never
calls it, so there is no reference call site and nothing to pin against: parity unpinned, defined by this
restatement of SURVEY App. A.3: maximise (MLL + log p_LogNormal(-4,1)(noise)) / n over (raw_lengthscale via softplus, noise
>= 1e-4) with scipy.optimize.minimize(method="L-BFGS-B", jac=True)."""
import math

import numpy as np
from scipy.optimize import minimize

from . import gp

NOISE_LOWER = 1e-4


def _softplus(x):
    return math.log1p(math.exp(-abs(x))) + max(x, 0.0)


def fit(train_x, train_obj, norm_bounds=None, maxiter=50, init_lengthscale=gp.LENGTHSCALE, init_noise=gp.NOISE):
    """Returns (lengthscale, noise, objective, scipy result)."""
    n = len(train_obj)

    def fun(theta):
        ls, nz = _softplus(float(theta[0])), float(theta[1])
        model = gp.GPModel(train_x, train_obj, norm_bounds=norm_bounds, lengthscale=ls, noise=nz)
        mll, g_ls, g_nz = gp.mll_and_grad(model)
        lx = math.log(nz)
        logp = -lx - 0.5 * math.log(2.0 * math.pi) - 0.5 * (lx + 4.0) ** 2
        dlogp = -1.0 / nz - (lx + 4.0) / nz
        sig = 1.0 / (1.0 + math.exp(-float(theta[0])))
        return -(mll + logp) / n, np.array([-(g_ls * sig) / n, -(g_nz + dlogp) / n])

    x0 = np.array([init_lengthscale + math.log(-math.expm1(-init_lengthscale)), init_noise])
    res = minimize(fun, x0, jac=True, method="L-BFGS-B", bounds=[(None, None), (NOISE_LOWER, None)],
                   options={"maxiter": maxiter})
    return _softplus(float(res.x[0])), float(res.x[1]), float(res.fun), res
