"""Rank-based weights and the weighted PCA of PCA-BO, restated on the CPU (oracle, test infrastructure).

Follows Algorithms/BayesianOptimization/PCA_BO.py:316-341 (_calculate_weights), :343-408
(_transform_points_to_reduced_space) and :410-434 (_transform_point_to_original_space).
`sklearn.decomposition.PCA` is the reference's real third-party dependency and IS installed here,
so `weighted_pca_fit` calls it exactly as the reference does; `weighted_pca_restated` is a numpy
restatement of what sklearn does (`_pca.py:_fit_full`, `_base.py:transform`, `extmath.svd_flip`)
that the CUDA kernel's structure follows and that is checked against sklearn in the tests.

Tie rule: the reference's `np.argsort(np.argsort(f))` uses numpy's default unstable
sort, so the order of exactly tied values is implementation defined. The oracle (and the kernel)
break ties by index (`kind="stable"`); on tie-free inputs this is identical to the reference.
"""
import numpy as np
from sklearn.decomposition import PCA


def rank_weights(f_evals, maximization: bool = False) -> tuple:
    """PCA_BO.py:325-341. Returns (weights float64[n], ranks int64[n]); rank 1 = best."""
    f = np.asarray(f_evals, dtype=np.float64)
    n = len(f)
    key = -f if maximization else f
    ranks = np.argsort(np.argsort(key, kind="stable"), kind="stable") + 1
    pre_weights = np.log(n) - np.log(ranks)
    weights = pre_weights / pre_weights.sum()
    return weights, ranks


def numpy_pairwise_sum(a) -> float:
    """Restatement of numpy's float64 pairwise summation (`DOUBLE_pairwise_sum`), SURVEY App. B.6.

    The CUDA kernel normalises the weights with exactly this order so that the weights are
    bit-identical to `pre_weights.sum()`.
    """
    a = np.asarray(a, dtype=np.float64)
    n = len(a)
    if n < 8:
        res = np.float64(-0.0) if n == 0 else a[0]
        for i in range(1, n):
            res = res + a[i]
        return float(res)
    if n <= 128:
        r = [a[i] for i in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + a[i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + a[i]
            i += 1
        return float(res)
    n2 = n // 2
    n2 -= n2 % 8
    return float(np.float64(numpy_pairwise_sum(a[:n2])) + np.float64(numpy_pairwise_sum(a[n2:])))


def select_components(ratio: np.ndarray, var_threshold: float, n_components: int = 0) -> int:
    """PCA_BO.py:389-394."""
    if n_components > 0:
        return int(n_components)
    cum = np.cumsum(ratio)
    k = int(np.sum(cum <= var_threshold) + 1)
    return max(1, min(k, len(ratio)))


def weighted_pca_fit(X, f_evals, noise, var_threshold=0.95, n_components=0, maximization=False) -> dict:
    """PCA_BO.py:358-408 with sklearn's PCA() exactly as the reference calls it.

    `noise` is the (n, D) array the reference draws with np.random.normal(0, 1e-8, size) at :376.
    """
    X = np.asarray(X, dtype=np.float64)
    weights, ranks = rank_weights(f_evals, maximization)
    data_mean = np.mean(X, axis=0)                                  # :364
    X_centered = X - data_mean                                      # :365
    weighted_X = X_centered * np.sqrt(weights[:, np.newaxis])       # :369
    weighted_X = weighted_X + noise                                 # :376-377
    pca = PCA()
    pca.fit(weighted_X)                                             # :383
    component_matrix = np.copy(pca.components_)                     # :386
    ratio = pca.explained_variance_ratio_
    k = select_components(ratio, var_threshold, n_components)       # :389-396
    pca.components_ = pca.components_[:k]                           # :399
    Z = pca.transform(X_centered)                                   # :407
    return dict(weights=weights, ranks=ranks, data_mean=data_mean, pca_mean=pca.mean_.copy(),
                components=component_matrix, explained_variance=pca.explained_variance_.copy(),
                explained_variance_ratio=ratio.copy(), k=k, Z=Z, pca=pca,
                solver=getattr(pca, "_fit_svd_solver", None))


def weighted_pca_restated(X, f_evals, noise, var_threshold=0.95, n_components=0, maximization=False,
                          method="eigh") -> dict:
    """numpy restatement of the same computation (the structure the CUDA kernel follows).

    method="eigh": eigen-decomposition of the ddof=1 covariance (sklearn's `covariance_eigh`);
    method="svd" : LAPACK SVD of the re-centred weighted matrix (sklearn's `full`).
    """
    X = np.asarray(X, dtype=np.float64)
    n, D = X.shape
    weights, ranks = rank_weights(f_evals, maximization)
    data_mean = X.mean(axis=0)
    Xc = X - data_mean
    Xw = Xc * np.sqrt(weights)[:, None] + noise
    m2 = Xw.mean(axis=0)
    A = Xw - m2
    if method == "svd":
        _, S, Vt = np.linalg.svd(A, full_matrices=False)
        lam = S ** 2 / (n - 1)
    else:
        C = (A.T @ A) / (n - 1)
        lam, V = np.linalg.eigh(C)
        lam = lam[::-1].copy()
        Vt = V[:, ::-1].T.copy()
        lam[lam < 0.0] = 0.0
    # svd_flip(u_based_decision=False): largest |entry| of each component is made positive
    idx = np.argmax(np.abs(Vt), axis=1)
    signs = np.sign(Vt[np.arange(Vt.shape[0]), idx])
    Vt = Vt * signs[:, None]
    r = min(n, D)
    lam, Vt = lam[:r], Vt[:r]
    ratio = lam / lam.sum()
    k = select_components(ratio, var_threshold, n_components)
    Vk = Vt[:k]
    Z = Xc @ Vk.T - m2 @ Vk.T
    return dict(weights=weights, ranks=ranks, data_mean=data_mean, pca_mean=m2, components=Vt,
                explained_variance=lam, explained_variance_ratio=ratio, k=k, Z=Z)


def inverse_map(z, components_k, pca_mean, data_mean) -> np.ndarray:
    """PCA_BO.py:427 -> sklearn inverse_transform: x = z @ V_k + m2, then + data_mean."""
    z = np.asarray(z, dtype=np.float64).reshape(1, -1)
    return ((z @ components_k + pca_mean) + data_mean).ravel()
