"""GPU parity tests: every CUDA kernel, called through the C ABI (ctypes), against the CPU oracle on seeded inputs.

Bars (BASELINE.md section 4 / north_star): ranks, weights and selected component counts bit-exact; eigenvalues,
posterior mean / sigma, log-EI and gradients within 1e-8 relative in FP64.
"""
import numpy as np
import pytest
import torch

from helpers import synthetic_history, oracle_state, batch_to_device

pytestmark = pytest.mark.gpu

RTOL = 1e-8


def _ops():
    from pcabo_b200 import ops
    return ops


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


# (100, *): BASELINE config C5 (D = 100, n = 300 .. 1050, k ~ 80-90; sklearn: full SVD below n = 1000, eigh from there)
CASES = [(5, 10), (10, 30), (10, 77), (20, 61), (20, 250), (40, 120), (40, 399), (40, 450), (7, 33),
         (100, 300), (100, 801), (100, 1050)]


@pytest.mark.parametrize("D,n", CASES)
@pytest.mark.parametrize("maximization", [False, True])
def test_weights_pca_fit(D, n, maximization):
    ops = _ops()
    if D >= 100 and maximization:
        pytest.skip("config C5 sizes are covered in the minimisation sense (all BBOB problems are minimisation)")
    B = 3 if D < 100 else 2
    hist = [synthetic_history(D, n, 100 * D + n + b, maximization=maximization) for b in range(B)]
    ld_n = n + 5
    X, f, nz, logtab = batch_to_device([h[0] for h in hist], [h[1] for h in hist], [h[2] for h in hist], ld_n)
    out = ops.weights_pca_fit(X, f, nz, logtab, n, 0.95, 0, maximization)
    torch.cuda.synchronize()
    for b in range(B):
        Xb, fb, nb = hist[b]
        st, _, _, _ = oracle_state(Xb, fb, nb, maximization=maximization)
        # bit-exact integer / weight parity
        assert np.array_equal(out["ranks"][b, :n].cpu().numpy(), st["ranks"]), "ranks"
        assert np.array_equal(out["w"][b, :n].cpu().numpy(), st["weights"]), "weights must be bit-identical"
        assert int(out["k"][b]) == st["k"], "component count"
        k = st["k"]
        assert rel_err(out["xbar"][b].cpu().numpy(), st["data_mean"], 1e-12) < 1e-12
        assert np.max(np.abs(out["m2"][b].cpu().numpy() - st["pca_mean"])) < 1e-18 + 1e-10 * np.max(np.abs(st["pca_mean"]))
        lam = out["lam"][b].cpu().numpy()
        lam_ref = st["explained_variance"]
        r = len(lam_ref)
        # eigenvalues: 1e-8 relative for everything above the noise floor of the covariance
        big = lam_ref > 1e-12 * lam_ref[0]
        assert rel_err(lam[:r][big], lam_ref[big]) < RTOL, "eigenvalues"
        assert rel_err(out["evr"][b].cpu().numpy()[:r][big], st["explained_variance_ratio"][big]) < RTOL
        # components: compare each retained component where the relative eigen-gap is not tiny
        V = out["V"][b].cpu().numpy()
        Vr = st["components"]
        gaps = np.abs(np.diff(lam_ref)) / lam_ref[0]
        for c in range(min(k, r)):
            gl = gaps[c - 1] if c > 0 else 1.0
            gr = gaps[c] if c < r - 1 else 1.0
            if min(gl, gr) > 1e-6:
                assert np.max(np.abs(V[c] - Vr[c])) < 1e-8, f"component {c}"
                assert int(np.argmax(np.abs(V[c]))) == int(np.argmax(np.abs(Vr[c])))
        # orthonormality of the full basis
        assert np.max(np.abs(V @ V.T - np.eye(D))) < 1e-12
        Z = out["Z"][b, :n, :k].cpu().numpy()
        scale = np.max(np.abs(st["Z"]))
        assert np.max(np.abs(Z - st["Z"])) < RTOL * scale, "projection"
        assert float(out["Z"][b, :n, k:].abs().max()) == 0.0 if k < D else True


def test_weights_pca_fit_ncomponents_and_ties():
    """n_components > 0 bypasses the threshold (PCA_BO.py:389-396); heavy exact ties use the index tie rule."""
    ops = _ops()
    D, n = 6, 40
    Xb, fb, nb = synthetic_history(D, n, 7, penalty_fraction=0.6)
    X, f, nz, logtab = batch_to_device([Xb], [fb], [nb], n)
    out = ops.weights_pca_fit(X, f, nz, logtab, n, 0.95, 2, False)
    st, _, _, _ = oracle_state(Xb, fb, nb, n_components=2)
    assert int(out["k"][0]) == 2 == st["k"]
    assert np.array_equal(out["ranks"][0].cpu().numpy(), st["ranks"])
    assert np.array_equal(out["w"][0].cpu().numpy(), st["weights"])
    assert np.max(np.abs(out["Z"][0, :, :2].cpu().numpy() - st["Z"])) < 1e-8 * np.max(np.abs(st["Z"]))
    assert (fb == 1000.0).sum() > 5


def _fit_gp_batch(D, n, B, maximization=False, seed0=0, ld_n=None):
    ops = _ops()
    ld_n = n if ld_n is None else ld_n
    hist = [synthetic_history(D, n, seed0 + 31 * D + n + b, maximization=maximization) for b in range(B)]
    X, f, nz, logtab = batch_to_device([h[0] for h in hist], [h[1] for h in hist], [h[2] for h in hist], ld_n)
    pca = ops.weights_pca_fit(X, f, nz, logtab, n, 0.95, 0, maximization)
    buf = ops.GPBuffers(B, ld_n, D, X.device)
    ops.gp_fit(buf, pca["Z"], pca["k"], f, n, mode=0)
    torch.cuda.synchronize()
    return ops, hist, X, f, pca, buf


@pytest.mark.parametrize("D,n", [(5, 10), (10, 77), (20, 250), (40, 120), (40, 450), (100, 300), (100, 801), (100, 1050)])
def test_gp_fit(D, n):
    B = 2
    ops, hist, X, f, pca, buf = _fit_gp_batch(D, n, B)
    assert int(buf.info.abs().max()) == 0
    mll = ops.gp_mll(buf, f).cpu().numpy()
    for b in range(B):
        # oracle GP on the GPU's own Z (isolates the GP kernels from the PCA kernel)
        k = int(pca["k"][b])
        Z = pca["Z"][b, :n, :k].cpu().numpy()
        from oracle import gp, acqf
        model = gp.GPModel(Z, hist[b][1], norm_bounds=acqf.normalize_box_pca(Z))
        K, L, alpha, Linv = model.caches()
        nb = acqf.normalize_box_pca(Z)
        assert rel_err(buf.norm_off[b, :k].cpu().numpy(), nb[0], 1e-12) < 1e-13
        assert rel_err(buf.norm_coef[b, :k].cpu().numpy(), nb[1] - nb[0], 1e-12) < 1e-13
        box = acqf.search_box_pca(Z)
        assert rel_err(buf.box_lo[b, :k].cpu().numpy(), box[0], 1e-12) < 1e-13
        assert rel_err(buf.box_hi[b, :k].cpu().numpy(), box[1], 1e-12) < 1e-13
        assert abs(float(buf.ymean[b]) - float(model.y_mean)) <= 1e-12 * abs(float(model.y_mean)) + 1e-12
        assert abs(float(buf.ystd[b]) - float(model.y_std)) <= 1e-12 * float(model.y_std)
        Lg = torch.tril(buf.L[b, :n, :n]).cpu().numpy()
        assert np.max(np.abs(Lg - L.numpy())) < RTOL * np.max(np.abs(L.numpy())), "Cholesky factor"
        Wg = torch.tril(buf.W[b, :n, :n]).cpu().numpy()
        assert np.max(np.abs(Wg - Linv.numpy())) < RTOL * np.max(np.abs(Linv.numpy())), "L^-1"
        ag = buf.alpha[b, :n].cpu().numpy()
        assert np.max(np.abs(ag - alpha.numpy())) < RTOL * np.max(np.abs(alpha.numpy())), "alpha"
        assert abs(mll[b] - model.mll()) < RTOL * abs(model.mll()), "marginal log-likelihood"


# n values on both sides of every multiple of the 64-row block of the triangular passes, and block counts that are not
# a multiple of the 4 thread groups (257..448 lost row blocks once: regression)
@pytest.mark.parametrize("D,n,npts", [(5, 10, 7), (10, 77, 5), (20, 250, 16), (40, 120, 33), (40, 450, 10), (40, 450, 64),
                                      (12, 64, 5), (12, 65, 5), (12, 129, 7), (20, 193, 5), (20, 256, 5), (20, 257, 5),
                                      (40, 300, 5), (40, 330, 10), (40, 385, 6), (40, 448, 5), (40, 449, 5),
                                      (100, 300, 10), (100, 801, 7), (100, 1050, 12)])
@pytest.mark.parametrize("maximization", [False, True])
def test_logei_value_grad(D, n, npts, maximization):
    if D >= 100 and maximization:
        pytest.skip("config C5 sizes are covered in the minimisation sense")
    B = 2
    ops, hist, X, f, pca, buf = _fit_gp_batch(D, n, B, maximization)
    from oracle import gp, acqf
    rng = np.random.RandomState(5)
    pts = torch.zeros(B, npts, D, dtype=torch.float64)
    models = []
    for b in range(B):
        k = int(pca["k"][b])
        Z = pca["Z"][b, :n, :k].cpu().numpy()
        fb = hist[b][1]
        model = gp.GPModel(Z, fb, norm_bounds=acqf.normalize_box_pca(Z))
        best = float(np.max(fb) if maximization else np.min(fb))
        acq = gp.LogExpectedImprovement(model, best, maximize=maximization)
        models.append((k, acq))
        box = acqf.search_box_pca(Z)
        P = box[0] + (box[1] - box[0]) * rng.rand(npts, k)
        # a few points on / next to training points (tiny variance, u << -1) and on the box boundary
        P[0] = Z[rng.randint(n)]
        P[1] = Z[int(np.argmin(fb))] + 1e-3 * rng.randn(k)
        P[2] = box[0]
        pts[b, :, :k] = torch.from_numpy(P)
        buf.best_f[b] = gp.best_f_as_botorch(best)
    buf.maximize = maximization
    ptsd = pts.cuda()
    out = ops.logei(buf, ptsd, want_grad=True, want_posterior=True)
    out_v = ops.logei(buf, ptsd, want_grad=False)
    torch.cuda.synchronize()
    for b in range(B):
        k, acq = models[b]
        Xp = pts[b, :, :k]
        mean, sigma = acq.mean_sigma(Xp)
        vals, grad = acq.value_and_grad(Xp)
        assert rel_err(out["mu"][b].cpu().numpy(), mean.numpy(), 1e-9) < RTOL, "posterior mean"
        assert rel_err(out["sigma"][b].cpu().numpy(), sigma.numpy()) < RTOL, "posterior sigma"
        assert rel_err(out["val"][b].cpu().numpy(), vals.numpy(), 1e-6) < RTOL, "logEI"
        assert torch.equal(out["val"][b], out_v["val"][b]) or rel_err(out_v["val"][b].cpu().numpy(), vals.numpy(), 1e-6) < RTOL
        g = out["grad"][b, :, :k].cpu().numpy()
        gr = grad.numpy()
        scale = np.maximum(np.max(np.abs(gr), axis=1, keepdims=True), 1e-12)
        assert np.max(np.abs(g - gr) / scale) < 1e-7, "logEI gradient (per-point max-norm relative)"
        assert float(out["grad"][b, :, k:].abs().max()) == 0.0 if k < D else True


@pytest.mark.parametrize("D,n,npts,B", [(10, 77, 512, 3), (20, 250, 512, 2), (40, 120, 512, 4), (40, 331, 517, 3),
                                        (40, 450, 512, 5), (40, 450, 1024, 2), (12, 64, 40, 1)])
@pytest.mark.parametrize("maximization", [False, True])
def test_logei_raw_runs_matches_oracle(D, n, npts, B, maximization):
    """The raw-sample scoring pass of the BO loop on the tensor-core tile pipeline (pcabo_logei_raw_runs: all runs of the
    batch in one launch, per-run k, PCA-mode normalisation): every value within 1e-8 of the CPU oracle and of the scalar
    value kernel, for every way of dealing the tiles to CTAs."""
    ops, hist, X, f, pca, buf = _fit_gp_batch(D, n, B, maximization, ld_n=n + (n & 1))
    from oracle import gp, acqf
    rng = np.random.RandomState(11)
    pts = torch.zeros(B, npts, D, dtype=torch.float64)
    refs = []
    for b in range(B):
        k = int(pca["k"][b])
        Z = pca["Z"][b, :n, :k].cpu().numpy()
        fb = hist[b][1]
        model = gp.GPModel(Z, fb, norm_bounds=acqf.normalize_box_pca(Z))
        best = float(np.max(fb) if maximization else np.min(fb))
        acq = gp.LogExpectedImprovement(model, best, maximize=maximization)
        box = acqf.search_box_pca(Z)
        P = box[0] + (box[1] - box[0]) * rng.rand(npts, k)
        P[0] = Z[rng.randint(n)]                                   # on a training point
        P[1] = Z[int(np.argmin(fb))] + 1e-3 * rng.randn(k)          # next to the incumbent
        P[2] = box[0]
        pts[b, :, :k] = torch.from_numpy(P)
        buf.best_f[b] = gp.best_f_as_botorch(best)
        with torch.no_grad():
            refs.append(acq(torch.from_numpy(P)).numpy())
    buf.maximize = maximization
    assert ops.logei_raw_supported(buf, n)
    ptsd = pts.cuda()
    scalar = ops.logei(buf, ptsd)["val"].cpu().numpy()
    for cpr in (0, 1, 3, 16, 64):
        got = ops.logei_raw(buf, ptsd, ctas_per_run=cpr)["val"].cpu().numpy()
        for b in range(B):
            assert rel_err(got[b], refs[b], 1e-6) < RTOL, f"run {b}, {cpr} CTAs per run: log-EI vs oracle"
            assert rel_err(got[b], scalar[b], 1e-6) < RTOL, f"run {b}, {cpr} CTAs per run: log-EI vs pcabo_logei"


def test_logei_raw_runs_limits():
    """Outside the limits of the tile pipeline the entry point refuses (no silent fallback inside the library)."""
    ops, hist, X, f, pca, buf = _fit_gp_batch(5, 12, 1)             # odd column capacity
    from pcabo_b200.lib import PcaboError
    assert not ops.logei_raw_supported(buf, 12)
    with pytest.raises(PcaboError):
        ops.logei_raw(buf, torch.zeros(1, 32, 5, dtype=torch.float64, device="cuda"))


def test_logei_branches():
    """All three branches of log_h (u > -1, -1e6 < u <= -1, u <= -1e6) against the oracle's helper."""
    B, D, n = 1, 5, 12
    ops, hist, X, f, pca, buf = _fit_gp_batch(D, n, B)
    from oracle import gp, acqf
    k = int(pca["k"][0])
    Z = pca["Z"][0, :n, :k].cpu().numpy()
    model = gp.GPModel(Z, hist[0][1], norm_bounds=acqf.normalize_box_pca(Z))
    box = acqf.search_box_pca(Z)
    pts = torch.zeros(1, 6, D, dtype=torch.float64)
    pts[0, :, :k] = torch.from_numpy(box[0] + (box[1] - box[0]) * np.random.RandomState(0).rand(6, k))
    for best in [float(np.min(hist[0][1])), -1e5, -1e9, 1e4]:
        acq = gp.LogExpectedImprovement(model, best, maximize=False)
        buf.best_f[0] = gp.best_f_as_botorch(best)
        out = ops.logei(buf, pts.cuda(), want_grad=True)
        vals, grad = acq.value_and_grad(pts[0, :, :k])
        assert rel_err(out["val"][0].cpu().numpy(), vals.numpy(), 1e-6) < RTOL, best
        g, gr = out["grad"][0, :, :k].cpu().numpy(), grad.numpy()
        scale = np.maximum(np.max(np.abs(gr), axis=1, keepdims=True), 1e-12)
        assert np.max(np.abs(g - gr) / scale) < 1e-7, best


def test_inverse_map_and_argmax():
    ops = _ops()
    D, n, B = 10, 40, 3
    hist = [synthetic_history(D, n, 900 + b) for b in range(B)]
    X, f, nz, logtab = batch_to_device([h[0] for h in hist], [h[1] for h in hist], [h[2] for h in hist], n)
    pca = ops.weights_pca_fit(X, f, nz, logtab, n)
    rng = np.random.RandomState(1)
    z = torch.zeros(B, D, dtype=torch.float64)
    from oracle import weights_pca
    lo = torch.full((D,), -5.0, dtype=torch.float64).cuda()
    hi = torch.full((D,), 5.0, dtype=torch.float64).cuda()
    refs = []
    for b in range(B):
        k = int(pca["k"][b])
        zz = rng.randn(k) * (0.5 if b < 2 else 30.0)
        z[b, :k] = torch.from_numpy(zz)
        V = pca["V"][b].cpu().numpy()
        refs.append(weights_pca.inverse_map(zz, V[:k], pca["m2"][b].cpu().numpy(), pca["xbar"][b].cpu().numpy()))
    out = ops.inverse_map(z.cuda(), pca["k"], pca["V"], pca["m2"], pca["xbar"], lo, hi)
    for b in range(B):
        assert np.max(np.abs(out["x"][b].cpu().numpy() - refs[b])) < 1e-12 * max(1.0, np.max(np.abs(refs[b])))
        outside = (not np.all(refs[b] >= -5.0)) or (not np.all(refs[b] <= 5.0))
        assert bool(out["oob"][b]) == outside
    assert bool(out["oob"][2])
    v = torch.from_numpy(rng.randn(4, 513))
    v[1, 77] = v[1].max() + 1.0
    v[1, 300] = v[1, 77]          # tie -> first occurrence
    bv, bi = ops.argmax(v.cuda())
    assert torch.equal(bi.cpu().long(), torch.argmax(v, dim=1))
    assert torch.equal(bv.cpu(), v.max(dim=1).values)
    assert int(bi[1]) == 77


@pytest.mark.parametrize("maximization", [False, True])
def test_pick_best_and_append_observation(maximization):
    """The device-side tail of an iteration against its plain restatement: first-occurrence arg-max over the restarts and
    the winning candidate (optimize_acqf's return); out-of-bounds penalty instead of the objective value, append to the
    history, running best and its float32 rounding (PCA_BO.py:239-281, 199-203), packed record."""
    ops = _ops()
    rng = np.random.RandomState(3)
    B, R, D, ld_n, n = 7, 10, 6, 12, 9
    vals = rng.randn(B, R)
    vals[1, 4] = vals[1, 7] = vals[1].max() + 1.0          # a tie: the first occurrence wins
    vals[2, :] = -np.inf                                    # all equal
    vals[3, 2] = np.nan                                     # NaNs are skipped (like pcabo_argmax)
    cand = rng.randn(B, R, D)
    out = ops.pick_best(torch.from_numpy(vals).cuda(), torch.from_numpy(cand).cuda())
    idx = out["idx"].cpu().numpy()
    ref_idx = np.array([int(np.nanargmax(np.where(np.isnan(v), -np.inf, v))) for v in vals])
    assert np.array_equal(idx, ref_idx) and idx[1] == 4 and idx[2] == 0
    assert np.array_equal(out["z"].cpu().numpy(), cand[np.arange(B), ref_idx])
    assert np.array_equal(out["val"].cpu().numpy(), vals[np.arange(B), ref_idx])
    X = torch.from_numpy(rng.randn(B, ld_n, D)).cuda()
    f = torch.from_numpy(rng.randn(B, ld_n)).cuda()
    X0, f0 = X.clone(), f.clone()
    x_new, z_new, f_new, acq = rng.randn(B, D), rng.randn(B, D), 100.0 * rng.randn(B), rng.randn(B)
    oob = np.array([0, 1, 0, 0, 1, 0, 0], dtype=np.int32)
    best0 = 10.0 * rng.randn(B)
    best = torch.from_numpy(best0.copy()).cuda()
    best_f = torch.zeros(B, dtype=torch.float64, device="cuda")
    rec = torch.zeros(B, 2 * D + 3, dtype=torch.float64, device="cuda")
    pen = -1000.0 if maximization else 1000.0
    ops.append_observation(X, f, n, torch.from_numpy(x_new).cuda(), torch.from_numpy(z_new).cuda(),
                           torch.from_numpy(f_new).cuda(), torch.from_numpy(oob).cuda(), torch.from_numpy(acq).cuda(),
                           pen, maximization, best, best_f, rec)
    fv = np.where(oob != 0, pen, f_new)
    nb = np.maximum(best0, fv) if maximization else np.minimum(best0, fv)
    assert np.array_equal(X[:, n].cpu().numpy(), x_new) and np.array_equal(f[:, n].cpu().numpy(), fv)
    keep = [i for i in range(ld_n) if i != n]
    assert torch.equal(X[:, keep], X0[:, keep]) and torch.equal(f[:, keep], f0[:, keep])
    assert np.array_equal(best.cpu().numpy(), nb)
    assert np.array_equal(best_f.cpu().numpy(), nb.astype(np.float32).astype(np.float64))
    r = rec.cpu().numpy()
    assert np.array_equal(r[:, :D], x_new) and np.array_equal(r[:, D:2 * D], z_new) and np.array_equal(r[:, 2 * D], fv)
    assert np.array_equal(r[:, 2 * D + 1], oob.astype(np.float64)) and np.array_equal(r[:, 2 * D + 2], acq)
    # vanilla BO: no reduced space and no bounds test (z = x, nothing is ever outside)
    ops.append_observation(X, f, n + 1, torch.from_numpy(x_new).cuda(), None, torch.from_numpy(f_new).cuda(), None, None,
                           pen, maximization, best, best_f, rec)
    assert np.array_equal(f[:, n + 1].cpu().numpy(), f_new) and np.array_equal(rec.cpu().numpy()[:, D:2 * D], x_new)


def test_errors_are_loud():
    """The product path refuses CPU tensors and wrong dtypes instead of falling back."""
    ops = _ops()
    from pcabo_b200 import PcaboError
    with pytest.raises(PcaboError):
        ops.argmax(torch.zeros(2, 3, dtype=torch.float64))
    with pytest.raises(PcaboError):
        ops.argmax(torch.zeros(2, 3, dtype=torch.float32, device="cuda"))


# ------------------------------------------------------------------------------------------------ real late BO states
@pytest.mark.parametrize("name", ["late_state_0.npz", "late_state_39.npz", "late_state_30.npz"])
def test_late_trajectory_state_gp_and_logei(name):
    """GP caches, posterior, log-EI value + gradient and the device L-BFGS-B on states taken from real C3 trajectories at
    n = 330 (tests/golden/make_late_state_fixtures.py): 1000-penalties, clustered points, scales up to 1e12."""
    import os
    from oracle import gp, acqf
    ops = _ops()
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
    Z, f, best, ics = d["Z"], d["f"], float(d["best"]), d["ics"]
    n, k = Z.shape
    kmax, cap = 40, 450
    dev = torch.device("cuda")
    Zd = torch.zeros(1, cap, kmax, dtype=torch.float64)
    Zd[0, :n, :k] = torch.from_numpy(Z)
    fd = torch.zeros(1, cap, dtype=torch.float64)
    fd[0, :n] = torch.from_numpy(f)
    kd = torch.tensor([k], dtype=torch.int32, device=dev)
    buf = ops.GPBuffers(1, cap, kmax, dev)
    ops.gp_fit(buf, Zd.to(dev), kd, fd.to(dev), n, mode=0)
    buf.best_f[0] = gp.best_f_as_botorch(best)
    model = gp.GPModel(Z, list(f), norm_bounds=acqf.normalize_box_pca(Z))
    acq = gp.LogExpectedImprovement(model, best, maximize=False)
    _, L, alpha, Linv = model.caches()
    assert int(buf.info[0]) == 0
    Lg = torch.tril(buf.L[0, :n, :n]).cpu().numpy()
    assert np.max(np.abs(Lg - L.numpy())) < RTOL * np.max(np.abs(L.numpy())), "Cholesky factor"
    Wg = torch.tril(buf.W[0, :n, :n]).cpu().numpy()
    assert np.max(np.abs(Wg - Linv.numpy())) < RTOL * np.max(np.abs(Linv.numpy())), "L^-1"
    # alpha = K^-1 y_s with cond(K) ~ 4e4: two backward-stable solvers agree to cond * eps, not to 1e-8 of each entry
    ag = buf.alpha[0, :n].cpu().numpy()
    assert np.max(np.abs(ag - alpha.numpy())) < 1e-8 * np.max(np.abs(alpha.numpy())), "alpha"
    # points: the initial conditions of that iteration, some training points and points next to the incumbent
    rng = np.random.RandomState(3)
    pts = np.concatenate([ics, Z[rng.randint(n, size=3)], Z[int(np.argmin(f))][None] * (1 + 1e-9)], axis=0)
    npts = pts.shape[0]
    P = torch.zeros(1, npts, kmax, dtype=torch.float64)
    P[0, :, :k] = torch.from_numpy(pts)
    out = ops.logei(buf, P.to(dev), want_grad=True, want_posterior=True)
    Xp = torch.from_numpy(pts)
    mean, sigma = acq.mean_sigma(Xp)
    vals, grad = acq.value_and_grad(Xp)
    assert rel_err(out["mu"][0].cpu().numpy(), mean.numpy(), 1e-9) < RTOL, "posterior mean"
    assert rel_err(out["sigma"][0].cpu().numpy(), sigma.numpy()) < 1e-7, "posterior sigma"
    assert rel_err(out["val"][0].cpu().numpy(), vals.numpy(), 1e-6) < 1e-7, "logEI"
    g = out["grad"][0, :, :k].cpu().numpy()
    gr = grad.numpy()
    scale = np.maximum(np.max(np.abs(gr), axis=1, keepdims=True), 1e-300)
    assert np.max(np.abs(g - gr) / scale) < 1e-6, "logEI gradient (per-point max-norm relative)"
    # the device optimiser from the iteration's own initial conditions: no abnormal termination, and the joint optimum
    # of each chunk of 5 restarts matches scipy's L-BFGS-B on the oracle (the reference's gen_candidates_scipy)
    box = acqf.search_box_pca(Z)
    blo = torch.zeros(1, kmax, dtype=torch.float64)
    bhi = torch.zeros(1, kmax, dtype=torch.float64)
    blo[0, :k] = torch.from_numpy(np.asarray(box[0]))
    bhi[0, :k] = torch.from_numpy(np.asarray(box[1]))
    assert rel_err(buf.box_lo[0, :k].cpu().numpy(), blo[0, :k].numpy(), 1e-12) < 1e-12
    I = torch.zeros(1, 10, kmax, dtype=torch.float64)
    I[0, :, :k] = torch.from_numpy(ics)
    res = ops.lbfgsb_logei(buf, I.to(dev), buf.box_lo, buf.box_hi, k_bound=k)
    st = res["status"][0].cpu().numpy()
    assert ((st >= 0) & (st <= 2)).all(), st
    bounds = torch.tensor(np.stack([np.asarray(box[0]), np.asarray(box[1])]))
    cand = res["cand"][0, :, :k].cpu()
    for c in range(2):
        cs, vs, warned = acqf.gen_candidates_scipy(torch.from_numpy(ics[5 * c:5 * c + 5]), acq, bounds)
        assert not warned
        with torch.no_grad():
            vd = acq(cand[5 * c:5 * c + 5])
        fs, fdv = float(vs.sum()), float(vd.sum())
        assert abs(fs - fdv) <= 2e-4 * max(1.0, abs(fs)), (name, c, fs, fdv)


@pytest.mark.parametrize("D,n", [(100, 303), (100, 900), (100, 1050), (40, 450)])
def test_device_lbfgsb_large_configs(D, n):
    """pcabo_lbfgsb_logei at config C5 sizes (D = 100, k ~ 80-90, n = 303 / 900 / 1050) and at the end of C3: the kernel
    must accept the size (no host fallback), terminate normally, and reach the joint optimum scipy's L-BFGS-B reaches on
    the oracle (the reference's gen_candidates_scipy) from the same initial conditions."""
    from oracle import gp, acqf
    B = 2
    # the optimiser kernel stages rows of W with 16-byte bulk copies: the leading dimension must be even
    ops, hist, X, f, pca, buf = _fit_gp_batch(D, n, B, ld_n=n + (n & 1))
    rng = np.random.RandomState(11)
    ics = torch.zeros(B, 10, D, dtype=torch.float64)
    accs, boxes = [], []
    for b in range(B):
        k = int(pca["k"][b])
        Z = pca["Z"][b, :n, :k].cpu().numpy()
        fb = hist[b][1]
        model = gp.GPModel(Z, fb, norm_bounds=acqf.normalize_box_pca(Z))
        best = float(np.min(fb))
        accs.append((k, gp.LogExpectedImprovement(model, best, maximize=False)))
        box = acqf.search_box_pca(Z)
        boxes.append(box)
        P = box[0] + (box[1] - box[0]) * rng.rand(10, k)
        P[0] = Z[int(np.argmin(fb))] + 1e-2 * rng.randn(k)
        ics[b, :, :k] = torch.from_numpy(P)
        buf.best_f[b] = gp.best_f_as_botorch(best)
    res = ops.lbfgsb_logei(buf, ics.cuda(), buf.box_lo, buf.box_hi, k_bound=int(pca["k"].max()))
    st = res["status"].cpu().numpy()
    assert ((st >= 0) & (st <= 2)).all(), st
    gaps = []
    for b in range(B):
        k, acq = accs[b]
        bounds = torch.tensor(np.stack([np.asarray(boxes[b][0]), np.asarray(boxes[b][1])]))
        cand = res["cand"][b, :, :k].cpu()
        assert bool(((cand >= bounds[0]) & (cand <= bounds[1])).all())
        for c in range(2):
            cs, vs, warned = acqf.gen_candidates_scipy(ics[b, 5 * c:5 * c + 5, :k].clone(), acq, bounds)
            with torch.no_grad():
                vd = acq(cand[5 * c:5 * c + 5])
            fs, fdv = float(vs.sum()), float(vd.sum())
            # the device optimiser may stop at a slightly better or worse point of the same basin (maxiter 200 cuts
            # both off), and from these uniform-random starts a flipped line-search decision can end in another local
            # optimum: most chunks agree to 1e-3, none is far off
            gaps.append((fs - fdv) / max(1.0, abs(fs)))
            assert gaps[-1] <= 0.1, (D, n, b, c, fs, fdv)
    assert sum(g <= 1e-3 for g in gaps) >= len(gaps) - 1, (D, n, gaps)


@pytest.mark.parametrize("D,n,npts", [(5, 10, 7), (20, 250, 9), (40, 330, 10), (40, 450, 12)])
@pytest.mark.parametrize("maximization", [False, True])
def test_probability_of_improvement_value_grad(D, n, npts, maximization):
    """acq = PCABO_ACQ_PI: botorch ProbabilityOfImprovement (the reference's "probability_of_improvement",
    PCA_BO.py:683-684), value and autograd gradient against the oracle."""
    B = 2
    ops, hist, X, f, pca, buf = _fit_gp_batch(D, n, B, maximization, seed0=77)
    from oracle import gp, acqf
    rng = np.random.RandomState(6)
    pts = torch.zeros(B, npts, D, dtype=torch.float64)
    models = []
    for b in range(B):
        k = int(pca["k"][b])
        Z = pca["Z"][b, :n, :k].cpu().numpy()
        fb = hist[b][1]
        model = gp.GPModel(Z, fb, norm_bounds=acqf.normalize_box_pca(Z))
        best = float(np.max(fb) if maximization else np.min(fb))
        models.append((k, gp.ProbabilityOfImprovement(model, best, maximize=maximization)))
        box = acqf.search_box_pca(Z)
        P = box[0] + (box[1] - box[0]) * rng.rand(npts, k)
        P[0] = Z[int(np.argmin(fb))] + 1e-3 * rng.randn(k)
        P[1] = 0.5 * (Z[rng.randint(n)] + Z[rng.randint(n)])
        pts[b, :, :k] = torch.from_numpy(P)
        buf.best_f[b] = gp.best_f_as_botorch(best)
    buf.maximize = maximization
    buf.acq = ops.ACQ_PI
    out = ops.logei(buf, pts.cuda(), want_grad=True)
    torch.cuda.synchronize()
    for b in range(B):
        k, acq = models[b]
        vals, grad = acq.value_and_grad(pts[b, :, :k])
        v = out["val"][b].cpu().numpy()
        assert np.max(np.abs(v - vals.numpy())) < 1e-9 * max(1.0, float(vals.abs().max())), "PI"
        g = out["grad"][b, :, :k].cpu().numpy()
        gr = grad.numpy()
        scale = np.maximum(np.max(np.abs(gr), axis=1, keepdims=True), 1e-12)
        assert np.max(np.abs(g - gr) / scale) < 1e-7, "PI gradient (per-point max-norm relative)"


@pytest.mark.parametrize("D", [2, 5, 10, 40, 100])
def test_bbob_kernel_matches_oracle(D):
    """pcabo_bbob_eval_aux (f15-f17, f20-f24; several instances) against the CPU oracle's BBOB restatement, on points
    inside and outside the box (boundary penalty), one point per problem and many points per problem."""
    from pcabo_b200.bbob import BatchedBBOB
    from oracle import bbob
    fids = [15, 16, 17, 20, 15, 16, 17, 20, 21, 22, 23, 24, 21, 22, 23, 24]
    insts = [0, 0, 0, 0, 3, 7, 11, 29, 0, 0, 0, 0, 5, 13, 21, 29]
    prob = BatchedBBOB(fids, insts, D, device="cuda")
    rng = np.random.RandomState(D)
    m = 9
    X = rng.uniform(-5, 5, size=(len(fids), m, D))
    X[:, 0] = rng.uniform(-7, 7, size=(len(fids), D))          # outside the box
    X[:, 1] = 0.0
    got_many = prob.eval_all(X)
    got_one = prob(torch.from_numpy(np.ascontiguousarray(X[:, 2])).cuda()).cpu().numpy()
    for b, (fid, inst) in enumerate(zip(fids, insts)):
        p = bbob.BBOBProblem(fid, D, inst)
        ref = np.array([p(X[b, j]) for j in range(m)])
        assert np.max(np.abs(got_many[b] - ref) / np.maximum(1.0, np.abs(ref))) < 1e-11, (fid, inst)
        assert abs(got_one[b] - ref[2]) <= 1e-11 * max(1.0, abs(ref[2]))


@pytest.mark.parametrize("D,n", [(5, 10), (10, 77), (20, 250), (40, 330), (40, 450), (100, 300), (100, 1050)])
def test_gp_mll_gradient(D, n):
    """pcabo_gp_mll_grad (d MLL / d lengthscale, d MLL / d noise from the caches, K^-1 = W^T W built tile by tile)
    against autograd through the oracle's marginal log-likelihood."""
    B = 2
    ops, hist, X, f, pca, buf = _fit_gp_batch(D, n, B, seed0=5)
    from oracle import gp, acqf
    g = ops.gp_mll_grad(buf).cpu().numpy()
    for b in range(B):
        k = int(pca["k"][b])
        Z = pca["Z"][b, :n, :k].cpu().numpy()
        model = gp.GPModel(Z, hist[b][1], norm_bounds=acqf.normalize_box_pca(Z))
        mll, g_ls, g_nz = gp.mll_and_grad(model)
        assert abs(g[b, 0] - g_ls) <= 1e-8 * max(1.0, abs(g_ls)), ("d/dl", g[b, 0], g_ls)
        assert abs(g[b, 1] - g_nz) <= 1e-8 * max(1.0, abs(g_nz)), ("d/dnoise", g[b, 1], g_nz)


@pytest.mark.parametrize("D,n", [(10, 60), (20, 150)])
def test_hyperparameter_fit_matches_oracle(D, n):
    """Opt-in MAP fit of (lengthscale, noise) per run (pcabo_b200/hyperfit.py: lock-step L-BFGS-B driving
    pcabo_gp_fit_hyper + pcabo_gp_mll + pcabo_gp_mll_grad_hyper) against its CPU twin (oracle/hyperfit.py: scipy on the
    oracle's autograd MLL): fitted values to 1e-6, and the caches / log-EI at the fitted values to 1e-8. The reference
    never fits: parity here is against the restated botorch fit_gpytorch_mll objective (unpinned)."""
    from oracle import gp, acqf, hyperfit as ohf
    from pcabo_b200.hyperfit import fit_hyperparameters
    B = 3
    ops, hist, X, f, pca, buf = _fit_gp_batch(D, n, B)
    res = fit_hyperparameters(buf, pca["Z"], pca["k"], f, n, mode=0)
    torch.cuda.synchronize()
    for b in range(B):
        k = int(pca["k"][b])
        Z = pca["Z"][b, :n, :k].cpu().numpy()
        ls, nz, obj, r = ohf.fit(Z, hist[b][1], norm_bounds=acqf.normalize_box_pca(Z))
        assert abs(res["lengthscale"][b] - ls) <= 1e-6 * ls, (b, res["lengthscale"][b], ls)
        assert abs(res["noise"][b] - nz) <= 1e-6 * max(nz, 1e-4), (b, res["noise"][b], nz)
        assert abs(res["objective"][b] - obj) <= 1e-9 * max(1.0, abs(obj))
        assert res["objective"][b] <= -(gp.GPModel(Z, hist[b][1], norm_bounds=acqf.normalize_box_pca(Z)).mll()) / n + 1.0
        # caches and log-EI at the fitted hyper-parameters
        model = gp.GPModel(Z, hist[b][1], norm_bounds=acqf.normalize_box_pca(Z), lengthscale=float(res["lengthscale"][b]),
                           noise=float(res["noise"][b]))
        _, L, alpha, _ = model.caches()
        assert rel_err(buf.alpha[b, :n].cpu().numpy(), alpha.numpy(), 1e-300) < 1e-7
    # log-EI through the per-run lengthscales
    rng = np.random.RandomState(3)
    kmax = D
    Xq = torch.zeros(B, 4, kmax, dtype=torch.float64)
    refs = []
    for b in range(B):
        k = int(pca["k"][b])
        Z = pca["Z"][b, :n, :k].cpu().numpy()
        box = acqf.search_box_pca(Z)
        P = np.asarray(box[0]) + (np.asarray(box[1]) - np.asarray(box[0])) * rng.rand(4, k)
        Xq[b, :, :k] = torch.from_numpy(P)
        model = gp.GPModel(Z, hist[b][1], norm_bounds=acqf.normalize_box_pca(Z), lengthscale=float(res["lengthscale"][b]),
                           noise=float(res["noise"][b]))
        best = float(np.min(hist[b][1]))
        buf.best_f[b] = gp.best_f_as_botorch(best)
        acq = gp.LogExpectedImprovement(model, best, maximize=False)
        with torch.no_grad():
            refs.append(acq(torch.from_numpy(P)).numpy())
    out = ops.logei(buf, Xq.cuda())
    for b in range(B):
        assert rel_err(out["val"][b].cpu().numpy(), refs[b], 1e-300) < 1e-7
