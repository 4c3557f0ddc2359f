"""Oracle-generated C3 histories for the stratified bench windows (bench.py --steps K and the reference arm).

Run in the build container (CPU only, ~20 min on 8 cores):   python tests/golden/make_bench_histories.py
The CPU oracle (oracle/loop.py) runs 16 of the 90 runs of BASELINE config C3 (f15 / f16 / f17 x instances 0..5, D = 40,
budget 450, n_doe 120, seed 1000 fid + 400 + instance) from their DoE to n = 436 points and the histories are stored
in tests/golden/bench_histories_c3.npz: X [16, 436, 40], f [16, 436], fid [16], inst [16]. The reference arm of
bench.py starts its timed windows from prefixes of these histories (n = 125, 285, 430), so a short --steps run samples
the whole n range of a C3 run instead of its cheapest end; nothing in it touches the CUDA library.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
N_END = 436
RUNS = [(fid, inst) for inst in range(6) for fid in (15, 16, 17)][:16]


def one(args):
    fid, inst = args
    import torch
    torch.set_num_threads(1)
    from oracle import bbob, doe, loop
    D = 40
    p = bbob.BBOBProblem(fid, D, inst)
    budget, ndoe = doe.budget_and_ndoe(D)
    r = loop.OracleRun(p, D, p.bounds, budget, ndoe, doe.run_seed(fid, D, inst))
    r.run(max_iterations=N_END - ndoe)
    return np.vstack(r.x_evals), np.asarray(r.f_evals, dtype=np.float64)


if __name__ == "__main__":
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = "1"
    with mp.get_context("spawn").Pool(int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)) as pool:
        res = pool.map(one, RUNS, chunksize=1)
    X = np.stack([r[0] for r in res])
    f = np.stack([r[1] for r in res])
    path = os.path.join(HERE, "bench_histories_c3.npz")
    np.savez_compressed(path, X=X, f=f, fid=np.array([r[0] for r in RUNS]), inst=np.array([r[1] for r in RUNS]))
    print("wrote", path, os.path.getsize(path), "bytes", X.shape)
