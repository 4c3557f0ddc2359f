"""`botorch.optim.optimize_acqf` as the reference calls it, restated on the CPU (oracle, test infrastructure).

Reference call site: Algorithms/BayesianOptimization/PCA_BO.py:607-614 (and Vanilla_BO.py:201-209):
    optimize_acqf(acq_function, bounds, q=1, num_restarts=10, raw_samples=512,
                  options={"batch_limit": 5, "maxiter": 200})
botorch 0.13.0 is not installed; its published control flow is restated:
  gen_batch_initial_conditions -> draw_sobol_samples (torch SobolEngine, scramble=True, GLOBAL torch RNG)
  -> acquisition on chunks of batch_limit points -> initialize_q_batch (Boltzmann torch.multinomial,
  argmax forced in) -> gen_candidates_scipy on each chunk of batch_limit restarts (ONE joint
  scipy L-BFGS-B problem per chunk, objective -sum(acq), autograd gradient, maxiter=200) -> clamp ->
  re-evaluate -> on OptimizationWarning regenerate the ICs and retry once -> argmax over restarts.
torch.quasirandom.SobolEngine, torch.multinomial and scipy.optimize.minimize are the real third-party
code and are called directly.
"""
import warnings
import numpy as np
import torch
from scipy.optimize import minimize

NUM_RESTARTS = 10      # PCA_BO.py:107
RAW_SAMPLES = 512      # PCA_BO.py:108
BATCH_LIMIT = 5        # PCA_BO.py:613
MAXITER = 200          # PCA_BO.py:613


class BadInitialCandidatesWarning(RuntimeWarning):
    pass


class OptimizationWarning(RuntimeWarning):
    pass


def search_box_pca(Z: np.ndarray) -> np.ndarray:
    """PCA_BO.py:558-573: [zmin - 0.5 range, zmax + 0.5 range], dims narrower than 0.1 widened. Returns (2,k)."""
    z_min = np.min(Z, axis=0)
    z_max = np.max(Z, axis=0)
    z_range = z_max - z_min
    zb = np.vstack([z_min - 0.5 * z_range, z_max + 0.5 * z_range]).T
    min_range = 0.1
    for i in range(zb.shape[0]):
        if zb[i, 1] - zb[i, 0] < min_range:
            mid = (zb[i, 1] + zb[i, 0]) / 2
            zb[i, 0] = mid - min_range / 2
            zb[i, 1] = mid + min_range / 2
    return zb.transpose().copy()


def normalize_box_pca(Z: np.ndarray) -> np.ndarray:
    """PCA_BO.py:512-518: Normalize bounds [zmin - 0.1 range, zmax + 0.1 range]. Returns (2,k)."""
    z_min = np.min(Z, axis=0)
    z_max = np.max(Z, axis=0)
    z_range = z_max - z_min
    return np.vstack([z_min - 0.1 * z_range, z_max + 0.1 * z_range])


def draw_sobol_samples(bounds: torch.Tensor, n: int, generator=None) -> torch.Tensor:
    """botorch.utils.sampling.draw_sobol_samples(bounds, n, q=1, seed=None) -> (n, k).

    With generator=None the scramble consumes the global torch CPU RNG exactly like
    SobolEngine(d, scramble=True, seed=None) does in the reference.
    """
    d = bounds.shape[-1]
    if generator is None:
        engine = torch.quasirandom.SobolEngine(d, scramble=True, seed=None)
        raw = engine.draw(n, dtype=torch.float64)
    else:
        raw = scrambled_sobol(d, n, generator)
    return bounds[0] + (bounds[1] - bounds[0]) * raw


def scrambled_sobol(d: int, n: int, generator: torch.Generator) -> torch.Tensor:
    """SobolEngine(d, scramble=True).draw(n, float64) with an explicit generator (torch quasirandom.py:_scramble)."""
    eng = torch.quasirandom.SobolEngine(d, scramble=False)
    MAXBIT = eng.MAXBIT
    shift_ints = torch.randint(2, (d, MAXBIT), generator=generator)
    eng.shift = torch.mv(shift_ints, torch.pow(2, torch.arange(0, MAXBIT)))
    ltm = torch.randint(2, (d, MAXBIT, MAXBIT), generator=generator).tril()
    torch._sobol_engine_scramble_(eng.sobolstate, ltm, d)
    eng.quasi = eng.shift.clone()
    eng._first_point = (eng.quasi / 2 ** MAXBIT).reshape(1, -1)
    eng._first_point = eng._first_point.to(torch.get_default_dtype())
    return eng.draw(n, dtype=torch.float64)


def initialize_q_batch(X: torch.Tensor, acq_vals: torch.Tensor, n: int, eta: float = 1.0, generator=None):
    """botorch.optim.initializers.initialize_q_batch for batch_shape == ()."""
    n_samples = X.shape[0]
    if n > n_samples:
        raise RuntimeError("n cannot be larger than the number of provided samples")
    if n == n_samples:
        return X, acq_vals
    Ystd = acq_vals.std(dim=0)
    if torch.any(Ystd == 0):
        warnings.warn("All acquisition values for raw samples points are the same.", BadInitialCandidatesWarning)
        perm = torch.randperm(n_samples, generator=generator)[:n]
        return X[perm], acq_vals[perm]
    max_val, max_idx = torch.max(acq_vals, dim=0)
    Zs = (acq_vals - acq_vals.mean(dim=0)) / Ystd
    etaZ = eta * Zs
    weights = torch.exp(etaZ)
    while torch.isinf(weights).any():
        etaZ *= 0.5
        weights = torch.exp(etaZ)
    idcs = torch.multinomial(weights, n, generator=generator)
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs], acq_vals[idcs]


def initialize_q_batch_nonneg(X: torch.Tensor, acq_vals: torch.Tensor, n: int, eta: float = 1.0, alpha: float = 1e-4,
                              generator=None):
    """botorch.optim.initializers.initialize_q_batch_nonneg [3P-mem] (used for acquisition functions botorch registers
    as non-negative: EI, PI, ...)."""
    n_samples = X.shape[0]
    if n > n_samples:
        raise RuntimeError("n cannot be larger than the number of provided samples")
    if n == n_samples:
        return X, acq_vals
    max_val, max_idx = torch.max(acq_vals, dim=0)
    if torch.any(max_val <= 0):
        warnings.warn("All acquisition values for raw sampled points are nonpositive, so initial conditions are being "
                      "selected randomly.", BadInitialCandidatesWarning)
        perm = torch.randperm(n_samples, generator=generator)[:n]
        return X[perm], acq_vals[perm]
    pos = acq_vals > 0
    num_pos = int(pos.sum())
    if num_pos < n:
        remaining = (~pos).nonzero(as_tuple=False).view(-1)
        rand = torch.randperm(remaining.shape[0], generator=generator)
        pos[remaining[rand[: n - num_pos]]] = 1
        return X[pos], acq_vals[pos]
    alpha_pos = acq_vals >= alpha * max_val
    while alpha_pos.sum() < n:
        alpha = 0.1 * alpha
        alpha_pos = acq_vals >= alpha * max_val
    alpha_pos_idcs = torch.arange(len(acq_vals))[alpha_pos]
    weights = torch.exp(eta * (acq_vals[alpha_pos] / max_val - 1))
    idcs = alpha_pos_idcs[torch.multinomial(weights, n, generator=generator)]
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs], acq_vals[idcs]


def gen_batch_initial_conditions(acq, bounds: torch.Tensor, num_restarts=NUM_RESTARTS, raw_samples=RAW_SAMPLES,
                                 batch_limit=BATCH_LIMIT, generator=None, trace=None):
    """botorch.optim.initializers.gen_batch_initial_conditions (q=1, no constraints)."""
    factor, max_factor = 1, 5
    ics = None
    while factor < max_factor:
        with warnings.catch_warnings(record=True) as ws:
            warnings.simplefilter("always", category=BadInitialCandidatesWarning)
            n = raw_samples * factor
            X_rnd = draw_sobol_samples(bounds, n, generator)
            with torch.no_grad():
                acq_vals = torch.cat([acq(x_) for x_ in X_rnd.split(batch_limit, dim=0)], dim=0)
            if trace is not None:
                trace.append(dict(kind="raw", X=X_rnd.clone(), vals=acq_vals.clone()))
            init = initialize_q_batch_nonneg if getattr(acq, "nonnegative", False) else initialize_q_batch
            ics, _ = init(X_rnd, acq_vals, num_restarts, generator=generator)
            if not any(issubclass(w.category, BadInitialCandidatesWarning) for w in ws):
                return ics
            if factor < max_factor:
                factor += 1
    warnings.warn("Unable to find non-zero acquisition function values.", BadInitialCandidatesWarning)
    return ics


def gen_candidates_scipy(ics: torch.Tensor, acq, bounds: torch.Tensor, maxiter=MAXITER, trace=None):
    """botorch.generation.gen.gen_candidates_scipy: one joint L-BFGS-B over all rows of `ics`."""
    lo, hi = bounds[0], bounds[1]
    clamped = torch.max(torch.min(ics, hi), lo)
    shapeX = clamped.shape
    x0 = clamped.reshape(-1).numpy().copy()
    sb = np.stack([lo.repeat(shapeX[0]).numpy(), hi.repeat(shapeX[0]).numpy()], axis=1)
    nfev = [0]

    def f_np(x):
        if np.isnan(x).any():
            raise RuntimeError("NaN in x")
        X = torch.from_numpy(x).reshape(shapeX).contiguous()
        vals, grad = acq.value_and_grad(X)
        nfev[0] += 1
        g = (-grad).reshape(-1).numpy().copy()
        if np.isnan(g).any():
            raise RuntimeError("NaN in gradient")
        return float(-vals.sum()), g

    res = minimize(f_np, x0, method="L-BFGS-B", jac=True, bounds=sb, options={"maxiter": maxiter})
    warned = False
    if not res.success:
        msg = res.message if isinstance(res.message, str) else res.message.decode()
        if not ("ITERATIONS REACHED LIMIT" in msg or "EVALUATIONS EXCEEDS LIMIT" in msg):
            warnings.warn(f"Optimization failed within scipy.optimize.minimize: {msg}", OptimizationWarning)
            warned = True
    cand = torch.from_numpy(res.x.copy()).reshape(shapeX)
    cand = torch.max(torch.min(cand, hi), lo)
    with torch.no_grad():
        vals = acq(cand)
    if trace is not None:
        trace.append(dict(kind="lbfgsb", x0=clamped.clone(), x=cand.clone(), vals=vals.clone(),
                          nit=res.nit, nfev=res.nfev, status=res.status, message=str(res.message)))
    return cand, vals, warned


def optimize_acqf(acq, bounds, num_restarts=NUM_RESTARTS, raw_samples=RAW_SAMPLES, batch_limit=BATCH_LIMIT,
                  maxiter=MAXITER, generator=None, trace=None):
    """optimize_acqf(..., q=1, return_best_only=True, retry_on_optimization_warning=True). Returns ((1,k), value)."""
    bounds = torch.as_tensor(bounds, dtype=torch.float64)

    def _optimize(ics):
        cands, vals, warned = [], [], False
        for chunk in ics.split(batch_limit):
            with warnings.catch_warnings(record=True) as ws:
                warnings.simplefilter("always", category=OptimizationWarning)
                c, v, _ = gen_candidates_scipy(chunk, acq, bounds, maxiter, trace)
            warned = warned or any(issubclass(w.category, OptimizationWarning) for w in ws)
            cands.append(c)
            vals.append(v)
        return torch.cat(cands), torch.cat(vals).flatten(), warned

    ics = gen_batch_initial_conditions(acq, bounds, num_restarts, raw_samples, batch_limit, generator, trace)
    cands, vals, warned = _optimize(ics)
    if warned:
        ics = gen_batch_initial_conditions(acq, bounds, num_restarts, raw_samples, batch_limit, generator, trace)
        cands, vals, warned = _optimize(ics)
    best = torch.argmax(vals.view(-1), dim=0)
    return cands[best].reshape(1, -1), vals[best]
