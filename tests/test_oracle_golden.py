"""CPU tests: the oracle against the reference's golden vectors (tests/golden/reference_kats.json, generated from the
reference's shipped IOHprofiler files by tests/golden/make_golden.py) and its own self-consistency checks."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import acqf, bbob, doe, gp, weights_pca

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
BOUNDS5 = np.tile([[-5.0, 5.0]], (5, 1))


@pytest.mark.parametrize("fid", [15, 20])
def test_doe_matches_reference_dat(fid):
    """Rows 1..10 of every run block = center-LHS DoE for seed 1000*fid + 50 + instance."""
    for inst, rows in enumerate(KAT["doe"][str(fid)]):
        np.random.seed(doe.run_seed(fid, 5, inst))
        X = doe.rescale(doe.lhs_center(5, 10), BOUNDS5)
        ref = np.array([r["x"] for r in rows])
        assert np.max(np.abs(X - ref)) < 5e-7          # .dat prints x with 6 decimals
        assert np.array_equal(np.round(X, 6), ref)


@pytest.mark.parametrize("fid", [15, 20])
def test_bbob_known_answers(fid):
    for inst, rows in enumerate(KAT["doe"][str(fid)]):
        p = bbob.BBOBProblem(fid, 5, inst)
        X = np.array([r["x"] for r in rows])
        y = np.array([r["raw_y"] for r in rows])
        np.testing.assert_allclose(p.batch(X) - p.fopt, y, rtol=2e-12)   # raw_y printed with 10 decimals
    for r in KAT["best"][str(fid)]:                                      # full-precision x and y
        p = bbob.BBOBProblem(fid, 5, r["instance"])
        assert abs(p(np.array(r["x"])) - p.fopt - r["y"]) <= 1e-12 * abs(r["y"])
    for r in KAT["rows"][str(fid)]:                                      # x rounded to 1e-6 -> loose
        p = bbob.BBOBProblem(fid, 5, r["instance"])
        assert abs(p(np.array(r["x"])) - p.fopt - r["raw_y"]) <= 1e-4 * max(1.0, abs(r["raw_y"]))


def test_best_x_of_doe_rows_is_bit_exact():
    """JSON `best.x` entries that are DoE points carry full precision (e.g. -1.4999999999999996)."""
    hits = 0
    for fid in (15, 20):
        for r in KAT["best"][str(fid)]:
            if r["evals"] <= 10:
                np.random.seed(doe.run_seed(fid, 5, r["instance"]))
                X = doe.rescale(doe.lhs_center(5, 10), BOUNDS5)
                assert np.array_equal(X[r["evals"] - 1], np.array(r["x"]))
                hits += 1
    assert hits >= 3


def test_pairwise_sum_is_numpy_sum():
    rng = np.random.default_rng(0)
    for n in list(range(1, 200)) + [255, 256, 257, 450, 1050, 2047]:
        v = rng.random(n) * rng.choice([1e-3, 1.0, 1e3])
        assert weights_pca.numpy_pairwise_sum(v) == float(v.sum())


def test_rank_weights_reference_formula():
    f = np.array([3.0, -1.0, 10.0, 2.5, 7.0])
    w, r = weights_pca.rank_weights(f)
    ref_r = np.argsort(np.argsort(f)) + 1                 # PCA_BO.py:333 on tie-free input
    assert np.array_equal(r, ref_r)
    pre = np.log(5) - np.log(ref_r)
    assert np.array_equal(w, pre / pre.sum())
    assert w[np.argmax(f)] == 0.0
    wm, rm = weights_pca.rank_weights(f, maximization=True)
    assert rm[np.argmax(f)] == 1
    # ties: index order (stable), best among ties is the earliest
    ft = np.array([1000.0, 5.0, 1000.0, 1000.0, 2.0])
    _, rt = weights_pca.rank_weights(ft)
    assert list(rt) == [3, 2, 4, 5, 1]


@pytest.mark.parametrize("D,n", [(5, 12), (10, 30), (10, 120), (20, 61), (40, 130)])
def test_restated_pca_matches_sklearn(D, n):
    from helpers import synthetic_history
    X, f, noise = synthetic_history(D, n, 3 * D + n)
    a = weights_pca.weighted_pca_fit(X, f, noise)
    for method in ("eigh", "svd"):
        b = weights_pca.weighted_pca_restated(X, f, noise, method=method)
        assert a["k"] == b["k"]
        lam = a["explained_variance"]
        big = lam > 1e-12 * lam[0]
        np.testing.assert_allclose(b["explained_variance"][big], lam[big], rtol=1e-9)
        np.testing.assert_allclose(b["Z"], a["Z"], atol=1e-9 * np.max(np.abs(a["Z"])))
        k = a["k"]
        np.testing.assert_allclose(b["components"][:k], a["components"][:k], atol=1e-8)
    x = weights_pca.inverse_map(a["Z"][3], a["components"][:a["k"]], a["pca_mean"], a["data_mean"])
    np.testing.assert_allclose(x, (a["pca"].inverse_transform(a["Z"][3].reshape(1, -1)) + a["data_mean"]).ravel())
    # solver switch of sklearn's "auto" 
    assert a["solver"] == ("covariance_eigh" if n >= 10 * D else "full")


def test_log_ei_helper_against_direct_formula():
    u = torch.tensor([-30.0, -8.0, -3.0, -1.0, -0.999, 0.0, 2.0], dtype=torch.float64)
    import math
    from scipy.special import log_ndtr, ndtr
    direct = []
    for v in u.tolist():
        if v > -5:
            direct.append(math.log(math.exp(-0.5 * v * v) / math.sqrt(2 * math.pi) + v * ndtr(v)))
        else:   # asymptotic series of h(u) = phi(u) (1/u^2 - 3/u^4 + 15/u^6 - 105/u^8)
            s = 1 / v ** 2 - 3 / v ** 4 + 15 / v ** 6 - 105 / v ** 8 + 945 / v ** 10
            direct.append(-0.5 * v * v - 0.5 * math.log(2 * math.pi) + math.log(s))
    got = gp.log_ei_helper(u).tolist()
    for g, d, v in zip(got, direct, u.tolist()):
        assert abs(g - d) < (1e-10 if v > -5 else 1e-6) * max(1.0, abs(d)), (v, g, d)
    assert gp.best_f_as_botorch(73.085663910096) == 73.08566284179688


def test_gp_model_identities():
    from helpers import synthetic_history, oracle_state
    X, f, noise = synthetic_history(8, 40, 11)
    st, model, box, acq = oracle_state(X, f, noise)
    K, L, alpha, Linv = model.caches()
    n = model.n
    assert torch.allclose(L @ L.T, K, atol=1e-12)
    assert torch.allclose(Linv @ L, torch.eye(n, dtype=torch.float64), atol=1e-10)
    assert torch.allclose(K @ alpha, model.y_s, atol=1e-9)
    # training data sit in [1/12, 11/12] after Normalize
    assert abs(float(model.train_xn.min()) - 1 / 12) < 1e-12 and abs(float(model.train_xn.max()) - 11 / 12) < 1e-12
    # posterior at a training point: small variance, mean close to the observation
    mu, var = model.posterior(model.train_x[:3])
    assert torch.all(var < model.y_std ** 2 * 0.02)
    # gradient by autograd == finite differences
    Xq = torch.from_numpy(box[0] + (box[1] - box[0]) * np.random.RandomState(0).rand(4, model.k))
    v, g = acq.value_and_grad(Xq)
    eps = 1e-6
    for j in range(model.k):
        e = torch.zeros_like(Xq)
        e[:, j] = eps
        fd = (acq(Xq + e) - acq(Xq - e)) / (2 * eps)
        assert torch.allclose(fd, g[:, j], rtol=1e-4, atol=1e-6)
    assert abs(model.mll() - float(torch.distributions.MultivariateNormal(
        torch.zeros(n, dtype=torch.float64), covariance_matrix=K).log_prob(model.y_s))) < 1e-8 * abs(model.mll())


def test_sobol_with_generator_equals_global_stream():
    w = torch.rand(100, dtype=torch.float64) + 0.1
    torch.manual_seed(7)
    a = torch.quasirandom.SobolEngine(6, scramble=True).draw(32, dtype=torch.float64)
    ia = torch.multinomial(w, 10)
    g = torch.Generator()
    g.manual_seed(7)
    b = acqf.scrambled_sobol(6, 32, g)
    ib = torch.multinomial(w, 10, generator=g)
    assert torch.equal(a, b) and torch.equal(ia, ib)


def test_oracle_loop_runs_and_penalises_out_of_bounds():
    from oracle import loop
    torch.set_num_threads(1)
    D = 5
    p = bbob.BBOBProblem(15, D, 0)
    r = loop.OracleRun(p, D, p.bounds, 10 * D + 50, 3 * D, doe.run_seed(15, D, 0), keep_trace=True,
                       num_restarts=4, raw_samples=64)
    r.run(max_iterations=4)
    assert len(r.f_evals) == 3 * D + 4
    for rec in r.trace:
        x = rec["new_x"]
        outside = (not np.all(x >= -5)) or (not np.all(x <= 5))
        assert rec["outside"] == outside
        if outside:
            assert rec["new_f"] == 1000      # PCA_BO.py:260-261
    assert r.current_best == min(r.f_evals)


def test_oracle_mll_gradient_matches_finite_differences():
    """oracle.gp.mll_and_grad (the autograd reference the MLL-gradient kernel is tested against) against central
    finite differences of the oracle's own marginal log-likelihood."""
    import numpy as np
    from oracle import gp
    rng = np.random.RandomState(1)
    Z, y = rng.rand(40, 3), rng.randn(40)
    m = gp.GPModel(Z, y, norm_bounds=None)
    mll, g_ls, g_nz = gp.mll_and_grad(m)
    # gpytorch's GEMM-form distances and the direct differences used here agree to rounding
    assert abs(mll - m.mll()) < 1e-9 * abs(mll)
    h = 1e-6
    f = lambda ls, nz: gp.GPModel(Z, y, norm_bounds=None, lengthscale=ls, noise=nz).mll()
    fd_ls = (f(m.lengthscale + h, m.noise) - f(m.lengthscale - h, m.noise)) / (2 * h)
    fd_nz = (f(m.lengthscale, m.noise + h * 1e-2) - f(m.lengthscale, m.noise - h * 1e-2)) / (2 * h * 1e-2)
    assert abs(g_ls - fd_ls) < 1e-5 * max(1.0, abs(fd_ls))
    assert abs(g_nz - fd_nz) < 1e-4 * max(1.0, abs(fd_nz))


def test_oracle_probability_of_improvement():
    """oracle.gp.ProbabilityOfImprovement = Phi(u) with botorch's sign convention, its autograd gradient against finite
    differences, and exp(logEI) / sigma = phi(u) + u Phi(u) on the same points (ties the two acquisitions together)."""
    import numpy as np
    import torch
    from oracle import gp
    rng = np.random.RandomState(2)
    Z, y = rng.rand(30, 4), rng.randn(30)
    m = gp.GPModel(Z, y, norm_bounds=None)
    best = float(y.min())
    for maximize in (False, True):
        b = float(y.max()) if maximize else best
        pi = gp.ProbabilityOfImprovement(m, b, maximize=maximize)
        ei = gp.LogExpectedImprovement(m, b, maximize=maximize)
        X = torch.from_numpy(rng.rand(6, 4))
        v, g = pi.value_and_grad(X)
        mean, sigma = pi.mean_sigma(X)
        u = (mean - pi.best_f) / sigma
        u = u if maximize else -u
        Phi = 0.5 * torch.erfc(-u / np.sqrt(2.0))
        assert torch.allclose(v, Phi, rtol=1e-12, atol=0)
        phi = torch.exp(-0.5 * u * u) / np.sqrt(2 * np.pi)
        assert torch.allclose(torch.exp(ei(X)) / sigma, phi + u * Phi, rtol=1e-9)
        h = 1e-6
        for j in range(4):
            Xp, Xm = X.clone(), X.clone()
            Xp[:, j] += h
            Xm[:, j] -= h
            fd = (pi(Xp) - pi(Xm)) / (2 * h)
            assert torch.allclose(g[:, j], fd, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("fid", [21, 22, 23, 24])
def test_bbob_f21_f24_definitions(fid):
    """f21-f24 (config C5) are restated from the COCO definitions and UNPINNED (the reference ships no vectors for them).
    What can be checked without ioh: the optimum value is attained at x_opt exactly, nothing in the box is below it,
    the boundary penalty is active outside, and the product's independent torch restatement agrees to 1e-11."""
    from oracle import bbob
    from pcabo_b200.bbob import BatchedBBOB
    for D in (2, 5, 20):
        for inst in (0, 7):
            p = bbob.BBOBProblem(fid, D, inst)
            assert p(p.xopt) == p.fopt
            assert np.all(np.abs(p.xopt) <= 5.0)
            rng = np.random.RandomState(fid * 100 + D + inst)
            X = rng.uniform(-5, 5, size=(300, D))
            y = p.batch(X)
            assert np.all(y >= p.fopt)
            Xo = X.copy()
            Xo[:, 0] = 6.5
            assert np.all(p.batch(Xo) >= p.fopt + (1e4 if fid == 24 else 1.0) * 1.5 ** 2 - 1e-9)   # boundary penalty
            got = BatchedBBOB([fid], [inst], D).eval_all(X[None])[0]
            assert np.max(np.abs(got - y) / np.maximum(1.0, np.abs(y))) < 1e-11


def test_oracle_hyperparameter_fit_is_a_stationary_point():
    """oracle/hyperfit.py (the CPU twin of the opt-in fit): the fitted (lengthscale, noise) improve the MAP objective of
    the prior values, respect the noise bound, and the objective's gradient vanishes there (finite differences)."""
    import math
    from helpers import synthetic_history, oracle_state
    from oracle import hyperfit, acqf, gp
    X, f, nz = synthetic_history(8, 40, 5)
    st, model, box, acq = oracle_state(X, f, nz)
    Z = st["Z"]
    nb = acqf.normalize_box_pca(Z)
    ls, noise, obj, res = hyperfit.fit(Z, f, norm_bounds=nb, maxiter=100)
    assert noise >= hyperfit.NOISE_LOWER and ls > 0

    def objective(l_, s_):
        m = gp.GPModel(Z, f, norm_bounds=nb, lengthscale=l_, noise=s_)
        lx = math.log(s_)
        return -(m.mll() - lx - 0.5 * math.log(2 * math.pi) - 0.5 * (lx + 4.0) ** 2) / len(f)

    assert abs(objective(ls, noise) - obj) < 1e-12
    assert obj <= objective(gp.LENGTHSCALE, gp.NOISE) + 1e-12
    h = 1e-5
    assert abs(objective(ls + h, noise) - objective(ls - h, noise)) / (2 * h) < 1e-3
    if noise > hyperfit.NOISE_LOWER * 1.01:
        assert abs(objective(ls, noise * (1 + h)) - objective(ls, noise * (1 - h))) / (2 * h * noise) < 1e-2
