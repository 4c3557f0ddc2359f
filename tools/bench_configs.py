"""Throughput of the other BASELINE configs (C1, C2, C5) on one GPU: BO iterations/s over a few lock-step steps.

C1: PCA-BO f15 D=10, 1 run (budget 150, n_doe 30)                       -- the reference's own CPU-runnable case
C2: vanilla BO and PCA-BO on f15/f20 D=20, 30 runs each (budget 250, n_doe 60)
C5: PCA-BO D=100, budget 1050, n_doe 300, 30 runs x 5 functions (f20-f24); the device optimiser covers the whole run
    (past n ~ 500 its BFGS pairs live in an L2-resident workspace); both optimisers are reported at the start of the
    run, and --whole-c5 runs all 750 BO iterations of the 150 runs with the device optimiser (iterations/s per quarter)
Run on the GPU box: python tools/bench_configs.py [--steps 5] [--only C5] [--whole-c5]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "para-ortho-pca-bo_b200"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch


def measure(name, fids, insts, D, algorithm, optimizer, steps, warmup=2, n_start=None):
    from pcabo_b200.bbob import BatchedBBOB
    from pcabo_b200.engine import BatchedBOEngine
    prob = BatchedBBOB(fids, insts, D, device="cuda")
    budget, ndoe = 10 * D + 50, 3 * D
    eng = BatchedBOEngine(prob, D, prob.bounds, budget, ndoe, [1000 * f + 10 * D + i for f, i in zip(fids, insts)],
                          algorithm=algorithm, optimizer=optimizer)
    eng.initial_design()
    for _ in range(warmup):
        eng.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"config": name, "algorithm": algorithm, "optimizer": optimizer, "runs": eng.B, "dim": D, "n_points": eng.n,
           "iterations_per_s": eng.B * steps / dt, "ms_per_step": 1e3 * dt / steps,
           "lbfgsb_rounds_per_step": eng.stats["lbfgsb_rounds"] / (steps + warmup)}
    print(json.dumps(out), flush=True)
    return out


def whole_c5(fids, insts, D=100, S=1):
    """The whole C5 run on one GPU with the device optimiser: 150 runs x 750 BO iterations, n = 300 -> 1050. S > 1: the
    runs advance as S sub-batches (PipelinedEngines: asynchronous tail, event-driven schedule)."""
    from pcabo_b200.bbob import BatchedBBOB
    from pcabo_b200.engine import BatchedBOEngine, PipelinedEngines
    budget, ndoe = 10 * D + 50, 3 * D
    subs = []
    for s_ in range(S):
        sel = list(range(s_, len(fids), S))
        prob = BatchedBBOB([fids[i] for i in sel], [insts[i] for i in sel], D, device="cuda")
        subs.append(BatchedBOEngine(prob, D, prob.bounds, budget, ndoe,
                                    [1000 * fids[i] + 10 * D + insts[i] for i in sel], algorithm="pca", optimizer="device"))
    eng = subs[0] if S == 1 else PipelinedEngines(subs)
    eng.initial_design()
    torch.cuda.synchronize()
    marks, t0 = [], time.perf_counter()
    total = budget - ndoe
    if S == 1:
        for s_ in range(total):
            eng.step()
            if (s_ + 1) % (total // 4) == 0 or s_ + 1 == total:
                torch.cuda.synchronize()
                marks.append((eng.n, time.perf_counter() - t0))
    else:
        done = 0
        while done < total:
            q = min(total // 4, total - done)
            eng.run_steps(q)
            done += q
            torch.cuda.synchronize()
            marks.append((eng.n, time.perf_counter() - t0))

    class _All:      # the view of all runs the summary below needs
        B = len(fids)
        stats = eng.stats
        k_host = __import__("numpy").concatenate([e.k_host for e in subs])
        current_best = [v for e in subs for v in e.current_best]
    eng = _All
    k = eng.k_host if getattr(eng, "k_host", None) is not None else None
    out = {"config": "C5 whole run", "optimizer": "device", "sub_batches": S, "runs": eng.B, "dim": D, "bo_iterations": eng.B * total,
           "seconds": marks[-1][1], "iterations_per_s": eng.B * total / marks[-1][1],
           "by_quarter": [{"n_end": n, "iterations_per_s": eng.B * (total // 4) / (t - (marks[i - 1][1] if i else 0.0))}
                          for i, (n, t) in enumerate(marks[:4])],
           "retries": eng.stats["retries"], "jitter_refits": eng.stats.get("jitter_refits", 0),
           "k_final_mean": float(k.mean()) if k is not None else None, "k_final_max": int(k.max()) if k is not None else None,
           "best_f_median": float(sorted(eng.current_best)[len(eng.current_best) // 2]) if hasattr(eng, "current_best") else None}
    print(json.dumps(out), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--whole-c5", action="store_true")
    ap.add_argument("--sub-batches", type=int, default=1, help="--whole-c5: sub-batches (PipelinedEngines)")
    ap.add_argument("--skip-short", action="store_true", help="--whole-c5 only, without the short C5 measurements")
    a = ap.parse_args()
    want = set(a.only.split(",")) if a.only else None
    if not want or "C1" in want:
        for opt in ("scipy", "device"):
            measure("C1", [15], [0], 10, "pca", opt, a.steps)
    if not want or "C2" in want:
        fids = [f for f in (15, 20) for _ in range(30)]
        insts = [i for _ in range(2) for i in range(30)]
        for alg in ("vanilla", "pca"):
            measure("C2", fids, insts, 20, alg, "device", a.steps)
    if not want or "C5" in want:
        fids = [f for f in (20, 21, 22, 23, 24) for _ in range(30)]
        insts = [i for _ in range(5) for i in range(30)]
        if not a.skip_short:
            measure("C5", fids, insts, 100, "pca", "scipy", max(2, a.steps // 2), warmup=1)
            try:
                measure("C5", fids, insts, 100, "pca", "device", max(2, a.steps // 2), warmup=1)
            except Exception as e:
                print(json.dumps({"config": "C5", "optimizer": "device", "error": str(e)}), flush=True)
        if a.whole_c5:
            whole_c5(fids, insts, S=a.sub_batches)


if __name__ == "__main__":
    main()
