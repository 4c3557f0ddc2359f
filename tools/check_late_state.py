"""Debugging aid (run on the B200): log-EI value / gradient parity against the CPU oracle on a LATE C3 state
(n ~ 330: clustered points, 1000-penalties, reduced-space scales up to 1e12), where the unit tests' random data never go."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "para-ortho-pca-bo_b200"))
import numpy as np
import torch
from pcabo_b200 import ops
from pcabo_b200.bbob import BatchedBBOB
from pcabo_b200.engine import BatchedBOEngine
from oracle import gp, acqf


def main():
    n_target = int(sys.argv[1]) if len(sys.argv) > 1 else 330
    D = 40
    fids = [f for f in (15, 16, 17) for _ in range(30)]
    insts = [i for _ in range(3) for i in range(30)]
    prob = BatchedBBOB(fids, insts, D, device="cuda")
    eng = BatchedBOEngine(prob, D, prob.bounds, 450, 120, [1000 * f + 400 + i for f, i in zip(fids, insts)], optimizer="device")
    eng.initial_design()
    while eng.n < n_target:
        eng.step()
    eng._fit_models()
    B, n = eng.B, eng.n
    k_host = eng.pca["k"].cpu().numpy().astype(np.int64)
    ics = eng._initial_conditions(list(range(B)), k_host)
    out = ops.logei(eng.gp, ics, want_grad=True, want_posterior=True)
    val, grad = out["val"].cpu().numpy(), out["grad"].cpu().numpy()
    mu, sg = out["mu"].cpu().numpy(), out["sigma"].cpu().numpy()
    Z = eng.pca["Z"].cpu().numpy()
    f = eng.f.cpu().numpy()
    h_ics = ics.cpu().numpy()
    rows = []
    for b in range(0, B, 3):
        kb = int(k_host[b])
        Zb = Z[b, :n, :kb]
        model = gp.GPModel(Zb, list(f[b, :n]), norm_bounds=acqf.normalize_box_pca(Zb))
        acq = gp.LogExpectedImprovement(model, best_f=float(eng.current_best[b]), maximize=False)
        X = torch.tensor(h_ics[b, :, :kb])
        v, g = acq.value_and_grad(X)
        v, g = v.numpy(), g.numpy()
        gn = np.abs(g).max(axis=1)
        rows.append(dict(b=b, k=kb, zscale=float(np.abs(Zb).max()),
                         val_rel=float(np.max(np.abs(val[b] - v) / np.maximum(1.0, np.abs(v)))),
                         grad_rel=float(np.max(np.abs(grad[b, :, :kb] - g).max(axis=1) / np.maximum(gn, 1e-300))),
                         grad_norm_oracle=float(gn.max()), grad_norm_dev=float(np.abs(grad[b, :, :kb]).max())))
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
