"""Opt-in GP hyper-parameter fit for all runs of a batch (north_star item; SURVEY App. A.3 last bullet, N4).

The reference NEVER fits its surrogate: SingleTaskGP is built at PCA_BO.py:535-544 and queried with the prior values
(Matern-5/2 lengthscale ln 2, noise e^-5), so the default engine does not call this and stays
reference-identical. With `fit_hyperparameters=True` the engine does what botorch's `fit_gpytorch_mll(mll)` would do for
that model: maximise

    (MLL(y_s | lengthscale, noise) + log p_LogNormal(-4, 1)(noise)) / n

over the raw parameters (raw_lengthscale through softplus; the noise untransformed with lower bound 1e-4) with scipy's
L-BFGS-B. All B problems advance in lock-step through scipy's reverse-communication entry (acq_host.LbfgsbProblem);
every round costs ONE batched pcabo_gp_fit_hyper + pcabo_gp_mll + pcabo_gp_mll_grad_hyper on the device and one small
read-back ([B] values, [B, 2] gradients). The CPU twin used by the parity test is oracle/hyperfit.py.
"""
import math

import numpy as np
import torch

from . import ops
from .acq_host import LbfgsbProblem

NOISE_LOWER = 1e-4            # GaussianLikelihood noise constraint GreaterThan(1e-4)
PRIOR_LOC, PRIOR_SCALE = -4.0, 1.0


def softplus(x):
    return math.log1p(math.exp(-abs(x))) + max(x, 0.0)


def inv_softplus(y):
    return y + math.log(-math.expm1(-y))


def log_normal_logpdf_and_grad(x):
    """log p and d log p / dx of LogNormal(PRIOR_LOC, PRIOR_SCALE) at x > 0."""
    lx = math.log(x)
    z = (lx - PRIOR_LOC) / PRIOR_SCALE
    logp = -lx - math.log(PRIOR_SCALE) - 0.5 * math.log(2.0 * math.pi) - 0.5 * z * z
    dlogp = -1.0 / x - z / (PRIOR_SCALE * x)
    return logp, dlogp


def fit_hyperparameters(buf: ops.GPBuffers, Z, k, f, n, mode=0, fix_lo=None, fix_hi=None, maxiter=50,
                        init_lengthscale=ops.LENGTHSCALE, init_noise=ops.NOISE):
    """Fits (lengthscale, noise) of every run and leaves `buf` holding the caches AT the fitted values.
    Returns dict(lengthscale [B], noise [B], objective [B], nit [B], nfev [B]) (numpy)."""
    B = buf.B
    dev = Z.device
    probs = [LbfgsbProblem(np.array([inv_softplus(init_lengthscale), init_noise]), lo=[0.0, NOISE_LOWER], hi=[0.0, 0.0],
                           maxiter=maxiter, nbd=[0, 1]) for _ in range(B)]
    ls = np.full(B, init_lengthscale)
    nz = np.full(B, init_noise)
    d_ls = torch.empty(B, dtype=torch.float64, device=dev)
    d_nz = torch.empty(B, dtype=torch.float64, device=dev)
    obj = np.zeros(B)

    def evaluate():
        d_ls.copy_(torch.from_numpy(ls))
        d_nz.copy_(torch.from_numpy(nz))
        ops.gp_fit(buf, Z, k, f, n, mode=mode, fix_lo=fix_lo, fix_hi=fix_hi, lengthscales=d_ls, noises=d_nz)
        mll = ops.gp_mll(buf, f).cpu().numpy()
        grad = ops.gp_mll_grad(buf).cpu().numpy()
        return mll, grad

    active = [p.advance() for p in probs]
    while any(active):
        for b, p in enumerate(probs):
            if active[b]:
                ls[b], nz[b] = softplus(float(p.x[0])), float(p.x[1])
        mll, grad = evaluate()
        for b, p in enumerate(probs):
            if not active[b]:
                continue
            logp, dlogp = log_normal_logpdf_and_grad(nz[b])
            sig = 1.0 / (1.0 + math.exp(-float(p.x[0])))           # d softplus / d raw
            fval = -(mll[b] + logp) / n
            g = np.array([-(grad[b, 0] * sig) / n, -(grad[b, 1] + dlogp) / n])
            obj[b] = fval
            p.supply(fval, g)
            active[b] = p.advance()
    for b, p in enumerate(probs):
        ls[b], nz[b] = softplus(float(p.x[0])), float(p.x[1])
    evaluate()                                                      # caches at the fitted values
    return dict(lengthscale=ls.copy(), noise=nz.copy(), objective=obj, nit=np.array([p.nit for p in probs]),
                nfev=np.array([p.nfev for p in probs]))
