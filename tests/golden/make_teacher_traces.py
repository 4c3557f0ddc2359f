"""Oracle trajectories for the teacher-forced proposal-parity tests (tests/test_gpu_teacher_forced.py).

Run in the build container (CPU only; minutes):   python tests/golden/make_teacher_traces.py [name ...]
For each configuration the CPU oracle (oracle/loop.py: the reference's PCA_BO.__call__ / Vanilla_BO.__call__ restated,
PCA_BO.py:178-281, Vanilla_BO.py:103-141) runs from its seed, and for every stored iteration the file keeps what a
test needs to put the engine into EXACTLY the oracle's state before that iteration and to compare the outcome:
  X, f                  the run's full history (the iteration with n points sees X[:n], f[:n])
  it_n, k, new_z, new_x, new_f, outside, acq_value      the oracle's proposal of that iteration
  nit, nfev, status, retried                            scipy's L-BFGS-B counters of the two restart chunks
  torch_words, torch_pos     MT19937 state of the GLOBAL torch CPU generator before the iteration (Sobol scramble,
                             multinomial; AbstractAlgorithm.py:310-328 seeds it)
  np_idx, np_keys, np_pos, np_has_gauss, np_cached      numpy legacy generator state before the iterations np_idx
                             (the first stored one, and every one that does not directly follow its predecessor)
Configurations (BASELINE.json): C1 f15 D=10 instance 0, all 120 iterations; C2 f15 / f20 D=20 instance 0, PCA-BO and
vanilla BO, first 40 iterations; C3 f15 D=40 instance 0, every 5th of the 330 iterations (n = 120 .. 445).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import bbob, doe, loop  # noqa: E402

CONFIGS = {
    "c1_f15_d10_pca": dict(fid=15, dim=10, inst=0, algorithm="pca", iters=120, stride=1),
    "c2_f15_d20_pca": dict(fid=15, dim=20, inst=0, algorithm="pca", iters=40, stride=1),
    "c2_f20_d20_pca": dict(fid=20, dim=20, inst=0, algorithm="pca", iters=40, stride=1),
    "c2_f15_d20_vanilla": dict(fid=15, dim=20, inst=0, algorithm="vanilla", iters=40, stride=1),
    "c2_f20_d20_vanilla": dict(fid=20, dim=20, inst=0, algorithm="vanilla", iters=40, stride=1),
    "c3_f15_d40_pca": dict(fid=15, dim=40, inst=0, algorithm="pca", iters=330, stride=5),
}


def torch_mt_state():
    """(words[624] uint32, pos) of the global torch CPU generator (CPUGeneratorImplState: uint64 seed, int32 left,
    int32 seeded, uint64 next, uint64 state[624], ...; pos = 624 if left == 1 else next)."""
    buf = torch.get_rng_state().numpy()
    left = int(buf[8:12].view(np.int32)[0])
    nxt = int(buf[16:24].view(np.uint64)[0])
    words = buf[24:24 + 4992].view(np.uint64).astype(np.uint32)
    return words, (624 if left == 1 else nxt)


def make(name, cfg):
    torch.set_num_threads(1)
    fid, D, inst = cfg["fid"], cfg["dim"], cfg["inst"]
    p = bbob.BBOBProblem(fid, D, inst)
    budget, ndoe = doe.budget_and_ndoe(D)
    seed = doe.run_seed(fid, D, inst)
    r = loop.OracleRun(p, D, p.bounds, budget, ndoe, seed, algorithm=cfg["algorithm"], keep_trace=True)
    r.initial_design()
    rows = dict(it_n=[], k=[], new_z=[], new_x=[], new_f=[], outside=[], acq_value=[], nit=[], nfev=[], status=[],
                retried=[], torch_words=[], torch_pos=[], np_idx=[], np_keys=[], np_pos=[], np_has_gauss=[], np_cached=[])
    last_stored = None
    for t in range(cfg["iters"]):
        store = t % cfg["stride"] == 0
        if store:
            w, pos = torch_mt_state()
            if last_stored is None or last_stored != t - 1:
                st = np.random.get_state()
                rows["np_idx"].append(len(rows["it_n"]))
                rows["np_keys"].append(st[1].astype(np.uint32))
                rows["np_pos"].append(st[2]); rows["np_has_gauss"].append(st[3]); rows["np_cached"].append(st[4])
            last_stored = t
        r.iteration()
        if store:
            rec = r.trace[-1]
            lb = [e for e in rec["inner"] if e["kind"] == "lbfgsb"]
            z = np.zeros(D); z[:len(rec["new_z"])] = rec["new_z"]
            rows["it_n"].append(rec["n"]); rows["k"].append(rec["k"]); rows["new_z"].append(z)
            rows["new_x"].append(np.asarray(rec["new_x"], dtype=np.float64)); rows["new_f"].append(rec["new_f"])
            rows["outside"].append(rec["outside"]); rows["acq_value"].append(rec["acq_value"])
            rows["nit"].append([e["nit"] for e in lb[-2:]]); rows["nfev"].append([e["nfev"] for e in lb[-2:]])
            rows["status"].append([e["status"] for e in lb[-2:]]); rows["retried"].append(len(lb) > 2)
            rows["torch_words"].append(w); rows["torch_pos"].append(pos)
        r.trace[-1]["inner"] = None          # keep memory flat on the long run
    out = {k: np.asarray(v) for k, v in rows.items()}
    out.update(X=np.vstack(r.x_evals), f=np.asarray(r.f_evals, dtype=np.float64), fid=fid, dim=D, inst=inst, seed=seed,
               n_doe=ndoe, budget=budget, algorithm=cfg["algorithm"])
    path = os.path.join(HERE, f"teacher_{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "stored", len(out["it_n"]), "iterations,", os.path.getsize(path), "bytes; out-of-bounds",
          int(out["outside"].sum()), "retried", int(out["retried"].sum()), flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    for nm in names:
        make(nm, CONFIGS[nm])
