"""Tuning / debugging aid (run on the B200): advance the C3 batch to a given n, then run scipy's L-BFGS-B (driven by the
CUDA value+gradient kernel) and the device L-BFGS-B kernel from the SAME initial conditions and compare the
termination statuses; chunks where only one of them terminates abnormally are dumped for offline analysis."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "para-ortho-pca-bo_b200"))
import numpy as np
import torch
from pcabo_b200.bbob import BatchedBBOB
from pcabo_b200.engine import BatchedBOEngine


def main():
    n_target = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    out_path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/abnormal"
    D = 40
    fids = [f for f in (15, 16, 17) for _ in range(30)]
    insts = [i for _ in range(3) for i in range(30)]
    prob = BatchedBBOB(fids, insts, D, device="cuda")
    eng = BatchedBOEngine(prob, D, prob.bounds, 450, 120, [1000 * f + 400 + i for f, i in zip(fids, insts)], optimizer="device")
    eng.initial_design()
    while eng.n < n_target:
        eng.step()
    eng._fit_models()
    B = eng.B
    k_host = eng.pca["k"].cpu().numpy().astype(np.int64)
    box_lo, box_hi = eng.gp.box_lo.cpu(), eng.gp.box_hi.cpu()
    ics = eng._initial_conditions(list(range(B)), k_host)
    c_host, w_host = eng._optimize_scipy(list(range(B)), ics, k_host, box_lo, box_hi)
    nit_h, nfev_h = dict(eng.last["nit"]), dict(eng.last["nfev"])
    out, w_dev = eng._optimize_device(ics, k_host)
    st = out["status"].cpu().numpy()
    nit_d, nfev_d = out["nit"].cpu().numpy(), out["nfev"].cpu().numpy()
    res = {"n": eng.n, "k_mean": float(k_host.mean()), "scipy_warned_runs": int(sum(w_host.values())),
           "device_warned_runs": int(sum(w_dev.values())), "device_status_hist": np.bincount(st.reshape(-1), minlength=4).tolist()}
    Z = eng.pca["Z"].cpu().numpy()
    f = eng.f.cpu().numpy()
    h_ics = ics.cpu().numpy()
    dumped = 0
    cases = []
    for b in range(B):
        if w_dev[b] and not w_host[b]:
            kb = int(k_host[b])
            cases.append(dict(b=b, k=kb, status=st[b].tolist(), nit_dev=nit_d[b].tolist(), nfev_dev=nfev_d[b].tolist(),
                              nit_scipy=nit_h[b], nfev_scipy=nfev_h[b]))
            if dumped < 6:
                np.savez(f"{out_path}_{dumped}.npz", Z=Z[b, :eng.n, :kb], f=f[b, :eng.n], box_lo=box_lo[b, :kb].numpy(),
                         box_hi=box_hi[b, :kb].numpy(), ics=h_ics[b, :, :kb], best=eng.current_best[b],
                         status=st[b], cand_dev=out["cand"][b].cpu().numpy()[:, :kb], cand_scipy=c_host[b])
                dumped += 1
    res["cases"] = cases[:40]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
