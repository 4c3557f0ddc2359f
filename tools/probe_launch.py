"""Where the time between the optimiser launch and its completion goes, as the host sees it (tuning aid)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "para-ortho-pca-bo_b200"))
import numpy as np, torch
from pcabo_b200.bbob import BatchedBBOB
from pcabo_b200.engine import BatchedBOEngine

def main():
    D = 40
    nruns = int(sys.argv[1]) if len(sys.argv) > 1 else 90
    prefetch = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
    fids = ([f for f in (15, 16, 17) for _ in range(30)])[:nruns]
    insts = ([i for _ in range(3) for i in range(30)])[:nruns]
    prob = BatchedBBOB(fids, insts, D, device="cuda")
    eng = BatchedBOEngine(prob, D, prob.bounds, 450, 120, [1000 * f + 400 + i for f, i in zip(fids, insts)], optimizer="device")
    eng.initial_design()
    eng.lpt_dispatch = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
    if not prefetch:
        eng._prefetch_noise = lambda: None
    rows = []
    for s in range(12):
        eng.kernel_events = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.step_launch()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        eng.step_begin()          # phase B: optimiser launched
        t2 = time.perf_counter()
        st = eng.pending["st"]
        st["handle"]["done"].synchronize()
        t3 = time.perf_counter()
        eng.step_end()
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        e0, e1 = eng.kernel_events[0][0], eng.kernel_events[0][1]
        nit = np.array([v for v in eng.last["nit"].values()])
        rows.append(dict(A_ms=(t1 - t0) * 1e3, B_host_ms=(t2 - t1) * 1e3, wait_ms=(t3 - t2) * 1e3, kernel_ev_ms=e0.elapsed_time(e1),
                         tail_ms=(t4 - t3) * 1e3, nit_max=int(nit.max()), nit_mean=float(nit.mean())))
    for r in rows[2:]:
        print(json.dumps({k: round(v, 2) for k, v in r.items()}))

main()
