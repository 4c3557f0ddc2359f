"""Trajectory-level parity on BASELINE config C1 (PCA-BO, f15, D=10, budget 150, n_doe 30) and a D=20 case:
the lock-step engine (optimizer = scipy and device) against the CPU oracle loop on identical seeds.

For every instance: number of leading BO iterations whose proposals coincide with the oracle's (|dz| <= 1e-6 rel.),
the iteration of the first divergence (an L-BFGS-B line-search decision flipping on ~1e-13 kernel-vs-CPU rounding;
from there the runs are different-but-equally-valid BO trajectories), and the final best values.
Run on the GPU box:  python tools/trajectory_parity.py [--instances 8] [--dim 10] [--iters 120]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "para-ortho-pca-bo_b200"))
import numpy as np


def oracle_worker(a):
    fid, D, inst, iters, noise = a
    os.environ["OMP_NUM_THREADS"] = "1"
    import torch
    torch.set_num_threads(1)
    from oracle import bbob, doe, loop, gp
    if noise > 0:
        # control experiment: the oracle against ITSELF with a relative 1e-13 perturbation of every log-EI value
        # (the size of the kernel-vs-CPU rounding difference); shows how long two runs can stay together at all
        rs = np.random.RandomState(12345 + inst)
        base_call = gp.LogExpectedImprovement.__call__

        def noisy(self, X):
            v = base_call(self, X)
            return v * (1.0 + noise * torch.from_numpy(rs.standard_normal(v.shape)))

        gp.LogExpectedImprovement.__call__ = noisy
    p = bbob.BBOBProblem(fid, D, inst)
    budget, ndoe = doe.budget_and_ndoe(D)
    r = loop.OracleRun(p, D, p.bounds, budget, ndoe, doe.run_seed(fid, D, inst), keep_trace=True)
    r.run(max_iterations=iters)
    return dict(inst=inst, z=[t["new_z"].tolist() for t in r.trace], f=[t["new_f"] for t in r.trace],
                k=[t["k"] for t in r.trace], best=float(r.current_best), oob=[t["outside"] for t in r.trace])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fid", type=int, default=15)
    ap.add_argument("--dim", type=int, default=10)
    ap.add_argument("--instances", type=int, default=8)
    ap.add_argument("--iters", type=int, default=120)
    args = ap.parse_args()
    D, fid, insts = args.dim, args.fid, list(range(args.instances))
    ctx = mp.get_context("spawn")
    with ctx.Pool(min(len(insts), os.cpu_count() or 1)) as pool:
        ora = pool.map(oracle_worker, [(fid, D, i, args.iters, 0.0) for i in insts])
        ora_noisy = pool.map(oracle_worker, [(fid, D, i, args.iters, 1e-13) for i in insts])
    from pcabo_b200.bbob import BatchedBBOB
    from pcabo_b200.engine import BatchedBOEngine
    out = {"config": f"PCA-BO f{fid} D={D} budget {10 * D + 50} n_doe {3 * D}, {len(insts)} instances, {args.iters} iterations"}
    ctl, ctl_dev = [], []
    for o, p in zip(ora, ora_noisy):
        m = 0
        dev_at = None
        for it in range(args.iters):
            same = o["k"][it] == p["k"][it]
            if same:
                zr, zp = np.array(o["z"][it]), np.array(p["z"][it])
                dz = np.max(np.abs(zr - zp)) / max(1.0, np.max(np.abs(zr)))
                same = dz <= 1e-6
            if not same:
                dev_at = float(dz) if o["k"][it] == p["k"][it] else None
                break
            m += 1
        ctl.append(m)
        ctl_dev.append(dev_at)
    out["control_oracle_vs_oracle_with_1e-13_value_noise"] = {
        "matching_leading_iterations": ctl, "first_deviation_rel": ctl_dev,
        "median_best": float(np.median([p["best"] for p in ora_noisy]))}
    for opt in ("scipy", "device"):
        prob = BatchedBBOB([fid] * len(insts), insts, D, device="cuda")
        eng = BatchedBOEngine(prob, D, prob.bounds, 10 * D + 50, 3 * D, [1000 * fid + 10 * D + i for i in insts], optimizer=opt)
        eng.initial_design()
        match = [0] * len(insts)
        alive = [True] * len(insts)
        first_dev = [None] * len(insts)
        k_hist = []
        for it in range(args.iters):
            eng.step()
            k_hist.append(np.asarray(eng.last["k"], dtype=np.float64).copy())
            for b, o in enumerate(ora):
                if not alive[b]:
                    continue
                k = o["k"][it]
                ok = int(eng.last["k"][b]) == k
                dz = None
                if ok:
                    zr = np.array(o["z"][it])
                    dz = float(np.max(np.abs(eng.last["z_new"][b, :k] - zr)) / max(1.0, np.max(np.abs(zr))))
                    ok = dz <= 1e-6
                if ok:
                    match[b] += 1
                else:
                    alive[b] = False
                    first_dev[b] = dz
        out[opt] = {"matching_leading_iterations": match, "first_deviation_rel": first_dev,
                    "best_engine": [float(v) for v in eng.current_best],
                    "best_oracle": [o["best"] for o in ora],
                    "median_best_engine": float(np.median(eng.current_best)),
                    "median_best_oracle": float(np.median([o["best"] for o in ora])),
                    "oob_fraction_engine": float(np.mean([np.mean(np.array(eng.f_evals[b][3 * D:]) == 1000.0) for b in range(len(insts))])),
                    "oob_fraction_oracle": float(np.mean([np.mean(o["oob"]) for o in ora])),
                    "retries": eng.stats["retries"],
                    # distribution-level agreement once the trajectories have decorrelated: number of components k
                    "k_mean_engine_by_quarter": [float(np.mean(k_hist[q * args.iters // 4:(q + 1) * args.iters // 4]))
                                                 for q in range(4)],
                    "k_mean_oracle_by_quarter": [float(np.mean([o["k"][q * args.iters // 4:(q + 1) * args.iters // 4]
                                                                for o in ora])) for q in range(4)]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
