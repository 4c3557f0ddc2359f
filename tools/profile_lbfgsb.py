"""Per-phase cycle breakdown of the device L-BFGS-B kernel on a C3-like state (tuning aid, run on the B200)."""
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

NAMES = ["tests", "cauchy_rest", "formk_chol", "cmprlb", "subsm", "ls_setup", "ls_step", "eval", "dcsrch", "update_formt_chol",
         "formk_products", "cauchy_pass1", "cauchy_bmv", "update_shift", "update_products", "x15",
         "ev_kvec", "ev_sweep", "x18", "ev_coeff", "ev_grad", "x21", "ev_head", "fk_pre", "fk_sync1", "fk_work", "fk_sync2", "sw_wait", "sw_phaseA", "sw_barrier", "sw_phaseB", "x31"]
NSLOT = 32


def main():
    D, steps = 40, int(sys.argv[1]) if len(sys.argv) > 1 else 3
    fids = [f for f in (15, 16, 17) for _ in range(30)]
    insts = [i for _ in range(3) for i in range(30)]
    prob = BatchedBBOB(fids, insts, D, device="cuda")
    eng = BatchedBOEngine(prob, D, prob.bounds, 450, 120, [1000 * f + 400 + i for f, i in zip(fids, insts)], optimizer="device")
    eng.initial_design()
    import time
    t0 = time.time()
    for s_ in range(steps):
        eng.step()
        if s_ % 20 == 0:
            print(f"step {s_} n={eng.n} t={time.time() - t0:.1f}s retries={eng.stats['retries']} "
                  f"nfev_max={max(max(v) for v in eng.last['nfev'].values())}", file=sys.stderr, flush=True)
    eng._fit_models()
    k_host = eng.pca["k"].cpu().numpy().astype(np.int64)
    ics = eng._initial_conditions(list(range(eng.B)), k_host)
    prof = torch.zeros(eng.B, 2, NSLOT, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.lbfgsb_logei(eng.gp, ics, eng.gp.box_lo, eng.gp.box_hi, prof=prof)
    e1.record()
    torch.cuda.synchronize()
    p = prof.cpu().numpy().reshape(-1, NSLOT).astype(np.float64)
    p[:, 31] = 0.0
    nit = out["nit"].cpu().numpy().reshape(-1)
    nfev = out["nfev"].cpu().numpy().reshape(-1)
    tot = p[:, :16].sum(1)
    worst = int(np.argmax(tot))
    res = {"kernel_ms": e0.elapsed_time(e1), "n": eng.n, "k_mean": float(k_host.mean()),
           "nit_hist_25": np.bincount(np.minimum(nit // 25, 8), minlength=9).tolist(),   # [0,25), [25,50) ... [175,200), 200
           "nit_mean": float(nit.mean()), "nit_max": int(nit.max()), "nfev_mean": float(nfev.mean()), "nfev_max": int(nfev.max()),
           "cycles_per_iteration_mean": {nm: float((p[:, i] / np.maximum(nit, 1)).mean()) for i, nm in enumerate(NAMES)},
           "share_mean": {nm: float(p[:, i].sum() / tot.sum()) for i, nm in enumerate(NAMES)},
           "worst_chunk": {"nit": int(nit[worst]), "nfev": int(nfev[worst]), "Mcycles": float(tot[worst] / 1e6),
                           "cycles_per_iteration": {nm: float(p[worst, i] / max(nit[worst], 1)) for i, nm in enumerate(NAMES)}}}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
