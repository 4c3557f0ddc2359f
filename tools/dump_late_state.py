"""Debugging aid (run on the B200): dump the device GP caches of a few runs on a late C3 state for offline comparison."""
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


def main():
    n_target = int(sys.argv[1]) if len(sys.argv) > 1 else 330
    runs = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 39]
    D = 40
    fids = [f for f in (15, 16, 17) for _ in range(30)]
    insts = [i for _ in range(3) for i in range(30)]
    prob = BatchedBBOB(fids, insts, D, device="cuda")
    eng = BatchedBOEngine(prob, D, prob.bounds, 450, 120, [1000 * f + 400 + i for f, i in zip(fids, insts)], optimizer="device")
    eng.initial_design()
    infos = []
    while eng.n < n_target:
        eng.step()
        infos.append(int((eng.gp.info != 0).sum()))
    eng._fit_models()
    print("chol info nonzero per step:", infos, file=sys.stderr)
    print("info now:", eng.gp.info.cpu().numpy().tolist(), file=sys.stderr)
    B, n = eng.B, eng.n
    k_host = eng.pca["k"].cpu().numpy().astype(np.int64)
    ics = eng._initial_conditions(list(range(B)), k_host)
    out = ops.logei(eng.gp, ics, want_grad=True, want_posterior=True)
    g = eng.gp
    for b in runs:
        kb = int(k_host[b])
        np.savez_compressed(f"gpurun_out/late_state_{b}.npz", n=n, k=kb, Z=eng.pca["Z"][b, :n, :kb].cpu().numpy(),
                            f=eng.f[b, :n].cpu().numpy(), best=eng.current_best[b], best_f=float(g.best_f[b]),
                            L=g.L[b, :n, :n].cpu().numpy(), W=g.W[b, :n, :n].cpu().numpy(), WT=g.WT[b, :n, :n].cpu().numpy(),
                            alpha=g.alpha[b, :n].cpu().numpy(), ZnT=g.ZnT[b, :kb, :n].cpu().numpy(),
                            Zn=g.Zn[b, :n, :kb].cpu().numpy(), off=g.norm_off[b, :kb].cpu().numpy(),
                            coef=g.norm_coef[b, :kb].cpu().numpy(), ymean=float(g.ymean[b]), ystd=float(g.ystd[b]),
                            info=int(g.info[b]), ics=ics[b, :, :kb].cpu().numpy(), val=out["val"][b].cpu().numpy(),
                            grad=out["grad"][b, :, :kb].cpu().numpy(), mu=out["mu"][b].cpu().numpy(),
                            sigma=out["sigma"][b].cpu().numpy(), box_lo=g.box_lo[b, :kb].cpu().numpy(),
                            box_hi=g.box_hi[b, :kb].cpu().numpy())


if __name__ == "__main__":
    main()
