"""Shared builders of seeded synthetic PCA-BO states for the parity tests (oracle side is tests-only)."""
import numpy as np
import torch

from oracle import bbob, doe, weights_pca, gp, acqf


def synthetic_history(D, n, seed, fid=15, penalty_fraction=0.25, maximization=False):
    """A BO-like history: center-LHS DoE of ~n/2 points, then points clustered around the incumbent, a fraction of
    which received the reference's out-of-bounds penalty of +-1000 (PCA_BO.py:260-261) -> exact ties in f."""
    rng = np.random.RandomState(seed)
    prob = bbob.BBOBProblem(fid, D, seed % 7)
    n_doe = max(2, n // 2)
    st = np.random.get_state()
    np.random.seed(seed)
    X0 = doe.rescale(doe.lhs_center(D, n_doe), prob.bounds)
    np.random.set_state(st)
    f0 = prob.batch(X0)
    X, f = [x for x in X0], [float(v) for v in f0]
    while len(X) < n:
        best = X[int(np.argmin(f))]
        x = best + rng.normal(0, 1.5, size=D)
        if rng.rand() < penalty_fraction or np.any(np.abs(x) > 5):
            x = x + np.sign(x) * rng.rand(D) * 3
            fx = 1000.0
            if np.all(np.abs(x) <= 5):
                fx = float(prob.batch(x[None])[0])
        else:
            fx = float(prob.batch(x[None])[0])
        X.append(x)
        f.append(fx)
    X = np.array(X)
    f = np.array(f)
    if maximization:
        f = -f
    noise = rng.normal(0, 1e-8, size=(n, D))
    return X, f, noise


def oracle_state(X, f, noise, var_threshold=0.95, n_components=0, maximization=False):
    st = weights_pca.weighted_pca_fit(X, f, noise, var_threshold, n_components, maximization)
    Z = st["Z"]
    model = gp.GPModel(Z, f, norm_bounds=acqf.normalize_box_pca(Z))
    box = acqf.search_box_pca(Z)
    best = float(np.max(f) if maximization else np.min(f))
    acq = gp.LogExpectedImprovement(model, best, maximize=maximization)
    return st, model, box, acq


def batch_to_device(Xs, fs, noises, ld_n, device="cuda"):
    B = len(Xs)
    n, D = Xs[0].shape
    X = torch.zeros(B, ld_n, D, dtype=torch.float64)
    f = torch.zeros(B, ld_n, dtype=torch.float64)
    nz = torch.zeros(B, n, D, dtype=torch.float64)
    for b in range(B):
        X[b, :n] = torch.from_numpy(Xs[b])
        f[b, :n] = torch.from_numpy(fs[b])
        nz[b] = torch.from_numpy(noises[b])
    logtab = torch.from_numpy(np.log(np.maximum(np.arange(ld_n + 2), 1).astype(np.float64)))
    return X.to(device), f.to(device), nz.to(device), logtab.to(device)
