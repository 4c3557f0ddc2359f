#!/usr/bin/env python
"""bench.py -- PCA-BO iterations/sec on BASELINE config C3 (D=40, f15-f17 x 30 instances, budget 450, n_doe 120).

A "step" is ONE lock-step BO iteration of the whole per-GPU batch (90 runs: rank weights + weighted PCA fit, GP
caches, 512 raw log-EI samples, 10 restarts of L-BFGS-B on log-EI, inverse map, objective, bookkeeping), i.e. 90 BO
iterations. Runs are independent, so at N GPUs every rank owns its own 90 runs (instances 30*rank .. 30*rank+29 of
f15, f16, f17) and there is no data-path collective: scaling is "weak" and `value` is the whole-job
iterations/sec = N * 90 * K / (max over ranks of the device-synchronised time of K steps).

    python bench.py --gpus N --steps K --warmup W            # ours (torchrun launches it for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the CPU oracle on the host cores (reference arm)

The JSON line also carries `roofline` (dominant kernel, CUDA-event timed inside the timed region), `cpu_baseline`
(the CPU oracle on a bounded sample of the same workload, rank 0 at N=1), `e2e`, `clocks` and `gpu_launches`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "para-ortho-pca-bo_b200"))

DIM = 40
FIDS = (15, 16, 17)
RUNS_PER_FID = 30
BUDGET, N_DOE = 10 * DIM + 50, 3 * DIM
METRIC = "PCA-BO iterations/sec (30 runs x f15-f17, D=40)"
UNIT = "BO iterations/s"


def workload_config(n_start, steps, warmup, optimizer, extra=None):
    cfg = {"workload": "C3: PCA-BO D=40, f15-f17 x 30 instances (90 runs per GPU), budget 450, n_doe 120, log-EI, "
                       "num_restarts 10, raw_samples 512, batch_limit 5, maxiter 200, var_threshold 0.95",
           "runs_per_gpu": len(FIDS) * RUNS_PER_FID, "dim": DIM, "budget": BUDGET, "n_doe": N_DOE,
           "n_points_range": [n_start + warmup, n_start + warmup + steps],
           "optimizer": optimizer,
           "sub_batches": "the 90 runs of a GPU advance as S lock-step sub-batches on S CUDA streams, software-pipelined by "
                          "one host thread (S in the `sub_batches` key; results do not depend on S)",
           "l2_policy": "per-step inputs exceed L2: every step refits PCA/GP from new data and streams "
                        "90 x n^2/2 FP64 cache matrices; no result is reused between steps"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------ CPU oracle arm
def _oracle_worker(args):
    fid, inst, n_iter, budget_s = args
    import torch
    torch.set_num_threads(1)
    from oracle import bbob, doe, loop
    p = bbob.BBOBProblem(fid, DIM, inst)
    r = loop.OracleRun(p, DIM, p.bounds, BUDGET, N_DOE, doe.run_seed(fid, DIM, inst))
    r.initial_design()
    t0 = time.perf_counter()
    times = []
    for _ in range(n_iter):
        s = time.perf_counter()
        r.iteration()
        times.append(time.perf_counter() - s)
        if budget_s and time.perf_counter() - t0 > budget_s:
            break
    return time.perf_counter() - t0, times


def _oracle_worker_primed(args):
    """cpu_baseline sample: the oracle continues a run from a given history (x, f prefix) for n_iter BO iterations."""
    fid, inst, X, f, n_iter = args
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import bbob, doe, loop
    p = bbob.BBOBProblem(fid, DIM, inst)
    seed = doe.run_seed(fid, DIM, inst)
    r = loop.OracleRun(p, DIM, p.bounds, BUDGET, N_DOE, seed)
    np.random.seed(seed + len(f))
    torch.manual_seed(seed + len(f))
    r.x_evals = [np.asarray(x, dtype=np.float64) for x in X]
    r.f_evals = [float(v) for v in f]
    r.assign_new_best()
    times = []
    for _ in range(n_iter):
        s = time.perf_counter()
        r.iteration()
        times.append(time.perf_counter() - s)
    return times


def _pool_map(fn, jobs, cores):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    # one BLAS/OpenMP thread per worker: with 16 processes x 16 library threads the tiny operands thrash
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS")}
    for k in saved:
        os.environ[k] = "1"
    try:
        with ctx.Pool(cores) as pool:
            return pool.map(fn, jobs)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def oracle_sampled_throughput(histories, n_points, iters_per_point, cores=None):
    """cpu_baseline of the GPU arm: `cores` oracle runs (one per host core, 1 thread each) continue the histories the
    engine produced (histories[j] = (fid, inst, x_evals, f_evals)) from each n in n_points for iters_per_point BO
    iterations (the first one of each segment is warm-up). Returns (iterations/s, cores, seconds, description)."""
    cores = min(cores or os.cpu_count() or 1, len(histories))
    total_it, total_s = 0, 0.0
    for n0 in n_points:
        jobs = [(fid, inst, X[:n0], f[:n0], iters_per_point) for fid, inst, X, f in histories[:cores]]
        res = _pool_map(_oracle_worker_primed, jobs, cores)
        total_s += max(sum(t[1:]) for t in res)          # workers run concurrently: the slowest one bounds the segment
        total_it += cores * (iters_per_point - 1)
    return total_it / total_s, cores, total_s


def oracle_throughput(steps, warmup, cores=None, budget_s=None):
    """Reference CPU path: one oracle run per core (1 intra-op thread each), `cores` runs in parallel,
    each doing warmup+steps BO iterations from its DoE (cut short after `budget_s` seconds of wall time).
    Returns (iterations/s over the timed steps, cores, seconds, steps actually timed, per-step seconds)."""
    cores = cores or os.cpu_count() or 1
    steps = min(steps, BUDGET - N_DOE - warmup)
    jobs = [(FIDS[j % len(FIDS)], j // len(FIDS), warmup + steps, budget_s) for j in range(cores)]
    res = _pool_map(_oracle_worker, jobs, cores)
    done = min(len(r[1]) for r in res)          # iterations every worker finished inside the budget
    steps = max(1, done - warmup)
    # the workers run concurrently and independently (no lock-step on the CPU): the job's wall time for the timed
    # iterations is the slowest worker's total
    timed = max(sum(r[1][warmup:warmup + steps]) for r in res)
    per_step = [max(r[1][i] for r in res) for i in range(warmup + steps)]
    return cores * steps / timed, cores, timed, steps, per_step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    its, cores, secs, steps, per_step = oracle_throughput(args.steps, args.warmup, budget_s=args.ref_budget_s)
    sample = (f"{cores} of the 90 runs (one per host core, 1 torch thread each), BO iterations "
              f"{args.warmup}..{args.warmup + steps} from the DoE (n = {N_DOE + args.warmup}..{N_DOE + args.warmup + steps})"
              + ("" if steps == args.steps else f"; cut at the {args.ref_budget_s:.0f} s wall-time budget after {steps} of {args.steps} steps"))
    line = {"impl": "reference", "metric": METRIC, "value": its, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (BBOB f15-f17 restated)",
            "config": workload_config(N_DOE, steps, args.warmup, "scipy L-BFGS-B (CPU oracle)"),
            "cpu_baseline": {"value": its, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": its, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    if args.trace:
        json.dump({"n_first": N_DOE, "step_ms_max_over_workers": [1e3 * t for t in per_step]}, open(args.trace, "w"))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pcabo_b200.bbob import BatchedBBOB
    from pcabo_b200.engine import BatchedBOEngine, PipelinedEngines

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"

    fids = [f for f in FIDS for _ in range(RUNS_PER_FID)]
    insts = [RUNS_PER_FID * rank + i for _ in FIDS for i in range(RUNS_PER_FID)]
    seeds = [1000 * f + 10 * DIM + i for f, i in zip(fids, insts)]
    S = max(1, args.sub_batches)
    subs = []
    for s_ in range(S):
        sel = list(range(s_, len(fids), S))
        prob = BatchedBBOB([fids[i] for i in sel], [insts[i] for i in sel], DIM, device=dev)
        subs.append(BatchedBOEngine(prob, DIM, prob.bounds, BUDGET, N_DOE, [seeds[i] for i in sel], algorithm="pca",
                                    device=dev, optimizer=args.optimizer,
                                    host_threads=max(1, (os.cpu_count() or 1) // max(1, world))))
    eng = subs[0] if S == 1 else PipelinedEngines(subs)
    eng.initial_design()
    if args.n_start > N_DOE:                       # optional fast-forward (untimed) to a later point of the runs
        if S == 1:
            for _ in range(args.n_start - N_DOE):
                eng.step()
        else:
            eng.run_steps(args.n_start - N_DOE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def advance(k):
        if S == 1:
            for _ in range(k):
                eng.step()
        else:
            eng.run_steps(k)

    args.steps = max(1, min(args.steps, BUDGET - eng.n - args.warmup))
    advance(args.warmup)
    barrier()
    stats0 = dict(eng.stats)
    for e_ in subs:
        e_.kernel_events = []                       # CUDA-event pairs around every optimiser-kernel launch
        if args.phases:
            e_.phase_times = {}
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_first = eng.n
    t0 = time.perf_counter()
    ev0.record()
    step_ms, step_phases = [], []
    if args.trace and S == 1:
        for _ in range(args.steps):
            ts = time.perf_counter()
            p0 = dict(eng.phase_times) if args.phases else None
            eng.step()
            torch.cuda.synchronize()
            step_ms.append(1e3 * (time.perf_counter() - ts))
            if args.phases:
                step_phases.append({k_: round(1e3 * (v_ - p0.get(k_, 0.0)), 3) for k_, v_ in eng.phase_times.items()})
    else:
        advance(args.steps)
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_s = ev0.elapsed_time(ev1) * 1e-3
    clocks = sampler.stop() if rank == 0 else None
    secs = max(wall, dev_s)
    if world > 1:
        t = torch.tensor([secs], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t)
    B = sum(e_.B for e_ in subs)
    its = world * B * args.steps / secs
    stats1 = eng.stats
    launches = stats1["kernel_launches"] - stats0["kernel_launches"]
    h2d = (stats1["h2d_bytes"] - stats0["h2d_bytes"]) / args.steps
    d2h = (stats1["d2h_bytes"] - stats0["d2h_bytes"]) / args.steps

    # roofline of the dominant kernel: the value+gradient log-EI launch of the L-BFGS-B rounds (HBM/L2 bound:
    # per run it streams the n x n lower-triangular FP64 cache L^-1 for 10 points, 2.5 flop/byte)
    roof = None
    evs = [ev for e_ in subs for ev in (e_.kernel_events or [])]
    if evs:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        tot_ms = sum(a.elapsed_time(b) for a, b, _ in evs)
        tot_bytes = sum(nb for _, _, nb in evs)
        achieved = tot_bytes / (tot_ms * 1e-3) / 1e9
        roof = {"kernel": ("lbfgsb_logei_kernel<5> (device L-BFGS-B incl. log-EI value+grad)" if args.optimizer == "device" else "logei_kernel<16,true> (value+grad, L-BFGS-B rounds)"), "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650",
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at n = 122 from the ncu --set full capture
                # (profiles/r01_lbfgsb_ncu_raw_summary.csv): the n^2 caches are served from L2, DRAM sees ~2 % of
                # the algorithmic bytes of that launch (1.2 GB)
                "traffic": (26.124032e6 + 9.216e3) if args.optimizer == "device" else None,
                "traffic_note": "ncu capture of the launch at n = 122; algorithmic bytes of that launch 1.2e9 (L2-resident caches)",
                "launches": len(evs),
                # the S sub-batches' launches overlap on the GPU: `achieved` is per launch (one sub-batch) as defined;
                # the GPU-wide rate while they overlap is S times that
                "concurrent_launches": S,
                # what actually feeds the kernel is L2 (DRAM sees ~2 % of these bytes): both triangles (L^-1 and its
                # transpose) are streamed per evaluation, summed over the S overlapping launches, against the ~12.4 TB/s
                # (6300 B/clk) the L2 slices deliver chip-wide (B300_MICROARCH.md); run average -- during the triangular
                # passes at n = 400 the rate is 4.8 TB/s
                "l2": {"achieved_gbs": 2.0 * tot_bytes / secs / 1e9, "peak_gbs": 12400.0,
                       "frac": 2.0 * tot_bytes / secs / 1e9 / 12400.0}, "avg_launch_us": 1e3 * tot_ms / len(evs),
                "share_of_step": tot_ms * 1e-3 / secs / S}
    cpu_base = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        # the CPU oracle continues the first `cores` runs from histories this run produced, at points spread over the
        # n range that was timed (a bounded sample: ~10-30 s of host work)
        n_lo, n_hi = n_first, eng.n
        npts = [n_lo] if n_hi - n_lo < 30 else [n_lo, (n_lo + n_hi) // 2, max(n_lo, n_hi - args.cpu_steps)]
        allh = []
        for s_, e_ in enumerate(subs):            # sub-batch s_ holds runs s_, s_ + S, ... of the rank's list
            allh += [(fids[s_ + S * j], insts[s_ + S * j], e_.x_evals[j], e_.f_evals[j]) for j in range(e_.B)]
        hist = allh[::max(1, B // (os.cpu_count() or 1))]
        cits, cores, csecs = oracle_sampled_throughput(hist, npts, args.cpu_steps)
        cpu_base = {"value": cits, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{cores} of the 90 runs, one per host core (1 torch thread each), continuing this run's "
                              f"histories at n = {npts} for {args.cpu_steps - 1} timed BO iterations each (+1 warm-up), "
                              f"{csecs:.1f} s of wall time"}
    if rank == 0:
        line = {"metric": METRIC, "value": its, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (BBOB f15-f17 restated)",
                "config": workload_config(n_first - args.warmup, args.steps, args.warmup, args.optimizer,
                                          {"lbfgsb_rounds_per_step": (stats1["lbfgsb_rounds"] - stats0["lbfgsb_rounds"]) / args.steps,
                                           "retries": stats1["retries"] - stats0["retries"],
                                           "host_blocked_ms_per_step": {
                                               "optimiser_kernel": 1e3 * (stats1["wait_opt_s"] - stats0["wait_opt_s"]) / args.steps,
                                               "fit_and_raw_samples": 1e3 * (stats1["wait_ic_s"] - stats0["wait_ic_s"]) / args.steps}}),
                "roofline": roof, "cpu_baseline": cpu_base, "sub_batches": S,
                "e2e": {"value": its, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "note": "every step is driven through the public engine API with host-generated RNG inputs "
                                "(PCA noise, Sobol samples) copied in and proposals/objective values read back inside "
                                "the timed region"},
                "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if args.trace and rank == 0 and step_ms:
        json.dump({"n_first": n_first, "step_ms": step_ms, "k_mean_last": float(subs[0].k_host.mean()),
                   "step_phases_ms": step_phases},
                  open(args.trace, "w"))
    if args.phases and rank == 0:
        pt = {}
        for e_ in subs:
            for k_, v_ in e_.phase_times.items():
                pt[k_] = pt.get(k_, 0.0) + v_
        tot = sum(pt.values())
        print(json.dumps({"phase_ms_per_step": {k: 1e3 * v / args.steps for k, v in pt.items()},
                          "sum_ms_per_step": 1e3 * tot / args.steps}), file=sys.stderr, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=BUDGET - N_DOE - 3,
                    help="timed BO iterations per run (default: the whole C3 run after the warm-up steps)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--optimizer", default="device", choices=["scipy", "device"])
    ap.add_argument("--n-start", type=int, default=N_DOE, help="fast-forward (untimed) until n points before warm-up")
    ap.add_argument("--cpu-steps", type=int, default=6,
                    help="BO iterations per core and sample point of the cpu_baseline leg (the first is warm-up)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sub-batches", type=int, default=3,
                    help="split the per-GPU batch into S sub-batches on S CUDA streams, software-pipelined by one host "
                         "thread (the optimiser kernel of one overlaps the host work of the next)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="wall-time budget of the reference arm")
    ap.add_argument("--trace", default=None, help="write per-step wall times (ms) to this JSON file (adds a sync per step)")
    ap.add_argument("--phases", action="store_true", help="print a host wall-time breakdown per phase (adds syncs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
