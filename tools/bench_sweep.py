"""EI candidates/sec/GPU (BASELINE config C4): log-EI sweep with fused arg-max over 2^20 .. 2^26 scrambled-Sobol
candidates against a GP with n = 450, candidate-sharded over the ranks with ONE all_gather of (value, global index).

    python tools/bench_sweep.py [--log2n 24] [--r 40]            # 1 GPU
    torchrun --nproc-per-node N tools/bench_sweep.py ...           # N GPUs (NCCL)
Prints one JSON line per configuration: candidates/s (whole job), FP64 roofline fraction against the measured cuBLAS
DGEMM peak (gpurun_out/fp64_peaks.json: 35.5 TFLOP/s; F_val = n^2 + 3 n r + 16 n flop per candidate).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "para-ortho-pca-bo_b200"))
import numpy as np
import torch
import torch.distributed as dist

FP64_PEAK_TFLOPS = 35.47   # cuBLAS DGEMM 8192^3 measured on this pool (tools/measure_fp64.py)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, nargs="+", default=[20, 22, 24])
    ap.add_argument("--dims", dest="r", type=int, nargs="+", default=[8, 16, 32, 40])
    ap.add_argument("--n", type=int, default=450)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-baseline", action="store_true",
                    help="also time the CPU oracle (torch FP64, all host cores) on 2^14 of the same candidates")
    args = ap.parse_args()
    from pcabo_b200 import ops
    from pcabo_b200.parallel import candidate_range, global_argmax
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    torch.manual_seed(0)
    np.random.seed(0)
    n = args.n
    for r in args.r:
        kmax = r + (r & 1)
        rng = np.random.RandomState(0)
        Z = rng.uniform(1 / 12, 11 / 12, size=(n, r))
        y = rng.randn(n)
        buf = ops.GPBuffers(1, n + (n & 1), kmax, dev)
        Zd = torch.zeros(1, buf.ld_n, kmax, dtype=torch.float64)
        Zd[0, :n, :r] = torch.from_numpy(Z)
        fd = torch.zeros(1, buf.ld_n, dtype=torch.float64)
        fd[0, :n] = torch.from_numpy(y)
        lo = torch.full((kmax,), -1 / 3, dtype=torch.float64, device=dev)
        hi = torch.full((kmax,), 4 / 3, dtype=torch.float64, device=dev)
        ops.gp_fit(buf, Zd.to(dev), torch.tensor([r], dtype=torch.int32, device=dev), fd.to(dev), n, mode=1,
                   fix_lo=lo, fix_hi=hi)
        buf.best_f[0] = float(torch.as_tensor(float(y.min())))
        eng = torch.quasirandom.SobolEngine(r, scramble=True, seed=1234)
        shift = torch.zeros(kmax, dtype=torch.int32)
        sst = torch.zeros(kmax, 30, dtype=torch.int32)
        shift[:r] = eng.shift.to(torch.int32)
        sst[:r] = eng.sobolstate.to(torch.int32)
        shift, sst = shift.to(dev), sst.to(dev)
        ws = ops.SweepWorkspace(dev)
        if args.cpu_baseline and rank == 0:
            # reference-style evaluation (botorch LogExpectedImprovement on the gpytorch caches, restated in oracle/):
            # dense k* R products in torch FP64 on all host cores, chunks of 4096 candidates
            import time
            from oracle import gp as ogp
            ncpu = os.cpu_count() or 1
            torch.set_num_threads(ncpu)
            model = ogp.GPModel(Z, y, norm_bounds=None)
            acq = ogp.LogExpectedImprovement(model, float(y.min()), maximize=False)
            pts = -1 / 3 + (5 / 3) * torch.quasirandom.SobolEngine(r, scramble=True, seed=1234).draw(1 << 14, dtype=torch.float64)
            with torch.no_grad():
                acq(pts[:4096])
                t0 = time.perf_counter()
                vals = torch.cat([acq(c) for c in pts.split(4096)])
                dt = time.perf_counter() - t0
            print(json.dumps({"impl": "reference", "metric": "EI candidates/sec", "n": n, "r": r, "candidates": 1 << 14,
                              "candidates_per_s": (1 << 14) / dt, "cores": ncpu, "kind": "port",
                              "argmax": [float(vals.max()), int(vals.argmax())]}), flush=True)
            torch.set_num_threads(1)
        for l2 in args.log2n:
            N = 1 << l2
            a, b = candidate_range(N, rank, world)
            best = None
            times = []
            for rep in range(args.reps + 1):
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                bv, bi = ops.ei_sweep(buf, ws, b - a, a, sobol=(shift, sst, lo, hi))
                gv, gi = global_argmax(bv, bi)
                e1.record()
                torch.cuda.synchronize()
                t = e0.elapsed_time(e1) * 1e-3
                if world > 1:
                    tt = torch.tensor([t], dtype=torch.float64, device=dev)
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    t = float(tt)
                if rep > 0:
                    times.append(t)
                best = (float(gv), int(gi))
            t = min(times)
            flops = N * (n * n + 3 * n * r + 16 * n)
            if rank == 0:
                print(json.dumps({"metric": "EI candidates/sec", "n_gpus": world, "n": n, "r": r, "log2_candidates": l2,
                                  "candidates_per_s": N / t, "candidates_per_s_per_gpu": N / t / world,
                                  "ms": t * 1e3, "fp64_tflops": flops / t / 1e12 / world,
                                  "roofline_frac_fp64": flops / t / 1e12 / world / FP64_PEAK_TFLOPS,
                                  "argmax": best}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
