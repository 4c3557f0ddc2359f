"""A small pass through every kernel of libpcabo_b200.so for compute-sanitizer (memcheck / racecheck), sized so that it
finishes in minutes under the tool: two engine iterations at D = 10 (resident W panels), the model fit + log-EI + the
optimiser kernel in streaming mode at n = 300 (D = 20, maxiter 4), the C5-sized optimiser with its workspace at n = 520
(maxiter 2), a 4096-candidate sweep.
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "para-ortho-pca-bo_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from pcabo_b200 import ops
from pcabo_b200.bbob import BatchedBBOB
from pcabo_b200.engine import BatchedBOEngine


def fitted(D, n, B, seed=0):
    rng = np.random.RandomState(seed)
    ld = n + (n & 1)
    X = torch.zeros(B, ld, D, dtype=torch.float64)
    f = torch.zeros(B, ld, dtype=torch.float64)
    X[:, :n] = torch.from_numpy(rng.uniform(-5, 5, size=(B, n, D)))
    f[:, :n] = torch.from_numpy((X[:, :n] ** 2).sum(-1).numpy() + rng.randn(B, n))
    nz = torch.from_numpy(rng.normal(0, 1e-8, size=(B, n, D)))
    logtab = torch.from_numpy(np.log(np.maximum(np.arange(ld + 2), 1).astype(np.float64)))
    X, f, nz, logtab = X.cuda(), f.cuda(), nz.cuda(), logtab.cuda()
    pca = ops.weights_pca_fit(X, f, nz, logtab, n, 0.95, 0, False)
    buf = ops.GPBuffers(B, ld, D, X.device)
    ops.gp_fit(buf, pca["Z"], pca["k"], f, n, mode=0)
    buf.best_f.copy_(f[:, :n].min(dim=1).values)
    return pca, buf, f


def main():
    prob = BatchedBBOB([15, 20], [0, 1], 10, device="cuda")
    e = BatchedBOEngine(prob, 10, prob.bounds, 150, 30, [15100, 20101], optimizer="device")
    e.initial_design()
    e.step()
    e.step()
    torch.cuda.synchronize()
    print("engine ok", e.n)
    for D, n, B, maxiter in ((20, 300, 2, 4), (100, 520, 1, 2)):
        pca, buf, f = fitted(D, n, B)
        k = int(pca["k"].max())
        ics = torch.zeros(B, 10, D, dtype=torch.float64, device="cuda")
        ics[:, :, :k] = (buf.box_lo[:, None, :k] + (buf.box_hi - buf.box_lo)[:, None, :k] *
                         torch.rand(B, 10, k, dtype=torch.float64, device="cuda"))
        out = ops.logei(buf, ics, want_grad=True, want_posterior=True)
        res = ops.lbfgsb_logei(buf, ics, buf.box_lo, buf.box_hi, maxiter=maxiter, k_bound=k)
        g = ops.gp_mll_grad(buf)
        m = ops.gp_mll(buf, f)
        torch.cuda.synchronize()
        print("kernels ok", D, n, ops.lbfgsb_plan(n, D, k), res["status"].flatten().tolist(), float(out["val"].sum()), float(g.sum()), float(m.sum()))
    pca, buf, f = fitted(8, 64, 1)
    ws = ops.SweepWorkspace(torch.device("cuda"), nparts=4)
    Xc = buf.box_lo[0] + (buf.box_hi - buf.box_lo)[0] * torch.rand(4096, 8, dtype=torch.float64, device="cuda")
    bv, bi = ops.ei_sweep(buf, ws, 4096, 0, X=Xc.contiguous())
    torch.cuda.synchronize()
    print("sweep ok", float(bv), int(bi))


if __name__ == "__main__":
    main()
