"""Host / device timeline of PipelinedEngines over a few iterations (tuning aid, run on the B200)."""
import os, sys, time, json
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "para-ortho-pca-bo_b200"))
import numpy as np, torch
from pcabo_b200.bbob import BatchedBBOB
from pcabo_b200.engine import BatchedBOEngine, PipelinedEngines

def main():
    D, S = 40, int(sys.argv[1]) if len(sys.argv) > 1 else 3
    fids = [f for f in (15, 16, 17) for _ in range(30)]
    insts = [i for _ in range(3) for i in range(30)]
    seeds = [1000 * f + 400 + i for f, i in zip(fids, insts)]
    subs = []
    for s_ in range(S):
        sel = list(range(s_, 90, S))
        prob = BatchedBBOB([fids[i] for i in sel], [insts[i] for i in sel], D, device="cuda")
        subs.append(BatchedBOEngine(prob, D, prob.bounds, 450, 120, [seeds[i] for i in sel], optimizer="device",
                                    host_threads=max(1, (os.cpu_count() or 1))))
    pipe = PipelinedEngines(subs)
    pipe.initial_design()
    pipe.run_steps(8)
    torch.cuda.synchronize()
    trace = []
    base_ev = torch.cuda.Event(enable_timing=True)
    base_ev.record()
    torch.cuda.synchronize()
    t_base = time.perf_counter()
    gpu_marks = []
    for idx, (e, st) in enumerate(zip(pipe.engines, pipe.streams)):
        def wrap(name, fn, idx=idx, e=e, st=st):
            def f(*a, **k):
                ev0 = torch.cuda.Event(enable_timing=True); ev0.record()
                t0 = time.perf_counter()
                r = fn(*a, **k)
                t1 = time.perf_counter()
                ev1 = torch.cuda.Event(enable_timing=True); ev1.record()
                trace.append((idx, name, (t0 - t_base) * 1e3, (t1 - t_base) * 1e3))
                gpu_marks.append((idx, name, ev0, ev1))
                return r
            return f
        e.step_launch = wrap("A", e.step_launch)
        e._propose_begin = wrap("B", e._propose_begin)
        e.step_end = wrap("E", e.step_end)
        e.kernel_events = []
    t0 = time.perf_counter()
    pipe.run_steps(5)
    torch.cuda.synchronize()
    print("ms per step", (time.perf_counter() - t0) * 1e3 / 5)
    for (idx, name, h0, h1), (_, _, ev0, ev1) in zip(trace, gpu_marks):
        print(f"eng {idx} {name} host {h0:8.2f} -> {h1:8.2f} ({h1 - h0:5.2f})   gpu {base_ev.elapsed_time(ev0):8.2f} -> {base_ev.elapsed_time(ev1):8.2f}")
    for idx, e in enumerate(pipe.engines):
        for (k0, k1, _, _) in e.kernel_events:
            print(f"eng {idx} K gpu {base_ev.elapsed_time(k0):8.2f} -> {base_ev.elapsed_time(k1):8.2f} ({k0.elapsed_time(k1):5.2f})")

main()
