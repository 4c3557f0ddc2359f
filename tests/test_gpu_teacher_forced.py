"""Teacher-forced proposal parity: every stored iteration of a CPU-oracle trajectory is replayed on the engine from
EXACTLY the oracle's state (history, numpy and torch generator states), so one iteration's L-BFGS-B decisions cannot
cascade into the next (free-running trajectories legitimately diverge once a line-search decision flips on a 1e-13
difference; profiles/r01_trajectory_parity_*.json). Compared per iteration: k (exact), out-of-bounds flag, proposed
z / x and the acquisition value, against the reference's hot loop as restated in oracle/loop.py (PCA_BO.py:178-281,
Vanilla_BO.py:103-141). Traces: tests/golden/teacher_*.npz, made by tests/golden/make_teacher_traces.py.

The per-iteration error distributions of both optimiser back ends are written to gpurun_out/ (copied to
profiles/r02_teacher_forced_parity.json)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
TRACES = ["c1_f15_d10_pca", "c2_f15_d20_pca", "c2_f20_d20_pca", "c2_f15_d20_vanilla", "c2_f20_d20_vanilla",
          "c3_f15_d40_pca"]


def _replay(name, optimizer):
    from pcabo_b200.bbob import BatchedBBOB
    from pcabo_b200.engine import BatchedBOEngine
    d = np.load(os.path.join(HERE, "golden", f"teacher_{name}.npz"))
    fid, D, inst, seed = int(d["fid"]), int(d["dim"]), int(d["inst"]), int(d["seed"])
    alg = str(d["algorithm"])
    prob = BatchedBBOB([fid], [inst], D, device="cuda")
    eng = BatchedBOEngine(prob, D, prob.bounds, int(d["budget"]), int(d["n_doe"]), [seed], algorithm=alg,
                          optimizer=optimizer)
    eng._prefetch_noise = lambda: None          # the numpy stream must advance exactly once per iteration
    np_at = {int(i): j for j, i in enumerate(d["np_idx"])}
    X, f = d["X"], d["f"]
    rows = []
    for t, n in enumerate(d["it_n"]):
        n = int(n)
        if t in np_at:
            j = np_at[t]
            eng.np_rngs[0].set_state(("MT19937", d["np_keys"][j], int(d["np_pos"][j]), int(d["np_has_gauss"][j]),
                                      float(d["np_cached"][j])))
        eng._gens_to_host()
        eng.gen_io.export(eng.torch_gens)         # fills the state template (seed / bookkeeping fields) before the words
        eng.gen_io.restore(eng.torch_gens, torch.from_numpy(d["torch_words"][t:t + 1].view(np.int32).copy()),
                           torch.tensor([int(d["torch_pos"][t])], dtype=torch.int32))
        eng.gens_on_host = True
        eng.attach_history([[x.copy() for x in X[:n]]], [[float(v) for v in f[:n]]])
        retries0 = eng.stats["retries"]
        eng.step()
        k = int(d["k"][t])
        zo, xo = d["new_z"][t, :k], d["new_x"][t]
        ze, xe = eng.last["z_new"][0, :k], eng.last["x_new"][0]
        rows.append(dict(
            n=n, k_oracle=k, k=int(eng.last["k"][0]) if alg == "pca" else D,
            oob_oracle=bool(d["outside"][t]), oob=bool(eng.last["oob"][0]),
            z_err=float(np.max(np.abs(ze - zo)) / max(1.0, float(np.max(np.abs(zo))))) if int(eng.last["k"][0]) == k or alg != "pca" else float("inf"),
            x_err=float(np.max(np.abs(xe - xo)) / max(1.0, float(np.max(np.abs(xo))))),
            acq_oracle=float(d["acq_value"][t]), acq=float(eng.last["acq_val"][0]),
            f_oracle=float(d["new_f"][t]), f=float(eng.last["new_f"][0]),
            nit_oracle=[int(v) for v in d["nit"][t]], nit=[int(v) for v in eng.last["nit"][0]],
            nfev_oracle=[int(v) for v in d["nfev"][t]], nfev=[int(v) for v in eng.last["nfev"][0]],
            retried_oracle=bool(d["retried"][t]), retried=eng.stats["retries"] > retries0))
    return rows


def _summary(rows):
    z = np.array([r["z_err"] for r in rows])
    a = np.array([abs(r["acq"] - r["acq_oracle"]) / max(1.0, abs(r["acq_oracle"])) for r in rows])
    q = lambda v, p: float(np.quantile(v[np.isfinite(v)], p)) if np.isfinite(v).any() else float("inf")  # noqa: E731
    return dict(iterations=len(rows), k_exact=int(sum(r["k"] == r["k_oracle"] for r in rows)),
                oob_equal=int(sum(r["oob"] == r["oob_oracle"] for r in rows)),
                z_err_le_1e8=int((z <= 1e-8).sum()), z_err_le_1e6=int((z <= 1e-6).sum()), z_err_le_1e4=int((z <= 1e-4).sum()),
                z_err_median=q(z, 0.5), z_err_p90=q(z, 0.9), z_err_max=float(z.max()),
                acq_err_le_1e8=int((a <= 1e-8).sum()), acq_err_median=q(a, 0.5), acq_err_p90=q(a, 0.9), acq_err_max=float(a.max()),
                acq_not_worse_1e6=int(sum(r["acq"] >= r["acq_oracle"] - 1e-6 * max(1.0, abs(r["acq_oracle"])) for r in rows)),
                nit_equal=int(sum(r["nit"] == r["nit_oracle"] for r in rows)),
                nfev_equal=int(sum(r["nfev"] == r["nfev_oracle"] for r in rows)),
                retried=int(sum(r["retried"] for r in rows)), retried_oracle=int(sum(r["retried_oracle"] for r in rows)),
                oob_fraction_oracle=float(np.mean([r["oob_oracle"] for r in rows])))


def _record(name, optimizer, rows, summ):
    out_dir = os.path.join(os.path.dirname(HERE), "gpurun_out")
    if not os.path.isdir(out_dir):
        return
    path = os.path.join(out_dir, "r02_teacher_forced_parity.json")
    allr = json.load(open(path)) if os.path.exists(path) else {}
    allr[f"{name}:{optimizer}"] = dict(summary=summ, z_err=[r["z_err"] for r in rows], n=[r["n"] for r in rows],
                                       nfev=[r["nfev"] for r in rows], nfev_oracle=[r["nfev_oracle"] for r in rows])
    json.dump(allr, open(path, "w"))


# Fraction of iterations whose proposal must agree with the oracle's to 1e-8 (measured on the B200, round 2:
# profiles/r02_teacher_forced_parity.json; BOTH back ends meet the same numbers). Where it does not, either a
# line-search decision flipped on a ~1e-13 difference of the acquisition value (nfev differs from the oracle's) or the
# same decisions were taken and 100-200 L-BFGS-B iterations amplified that difference along a flat direction of log-EI
# (z differs by up to ~2e-4 -- median below 1e-9 -- while the acquisition value still agrees to 1e-8).
Z_1E8_FRACTION = {"c1_f15_d10_pca": 0.85, "c2_f15_d20_pca": 0.5, "c2_f20_d20_pca": 0.5, "c2_f15_d20_vanilla": 1.0,
                  "c2_f20_d20_vanilla": 1.0, "c3_f15_d40_pca": 0.3}


def _check(name, rows, s):
    it = s["iterations"]
    assert s["k_exact"] == it, s                              # component count: exact, every iteration
    assert s["oob_equal"] >= 0.95 * it, s
    assert s["z_err_le_1e8"] >= Z_1E8_FRACTION[name] * it, s
    assert s["acq_err_le_1e8"] >= 0.8 * it, s                 # the optimum VALUE agrees to 1e-8 even where z moved
    assert s["acq_not_worse_1e6"] >= 0.9 * it, s
    same = [r for r in rows if r["nfev"] == r["nfev_oracle"]]
    assert len(same) >= 0.45 * it, s
    assert max(r["z_err"] for r in same) <= 1e-3, s          # same decisions => same proposal up to conditioning
    assert float(np.median([r["z_err"] for r in same])) <= 1e-8, s


@pytest.mark.parametrize("name", TRACES)
def test_teacher_forced_scipy_mode(name):
    """optimizer="scipy": scipy's own L-BFGS-B on the CUDA value+gradient kernel -- control flow identical to the
    reference; k / out-of-bounds exact, acquisition value 1e-8, proposals 1e-8 wherever the optimiser's decisions were
    not flipped or amplified (fractions above)."""
    rows = _replay(name, "scipy")
    s = _summary(rows)
    _record(name, "scipy", rows, s)
    _check(name, rows, s)


@pytest.mark.parametrize("name", TRACES)
def test_teacher_forced_device_mode(name):
    """optimizer="device" (the bench's back end): the from-scratch L-BFGS-B kernel meets the SAME bounds as scipy's
    L-BFGS-B driven on the CUDA value+gradient kernel."""
    rows = _replay(name, "device")
    s = _summary(rows)
    _record(name, "device", rows, s)
    _check(name, rows, s)
