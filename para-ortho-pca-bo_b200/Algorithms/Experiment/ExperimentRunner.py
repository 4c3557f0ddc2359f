"""Experiment runner with the reference's constructor and `run_experiment()` (Algorithms/Experiment/ExperimentRunner.py).

Where the reference loops sequentially over (algorithm x problem x dimension x instance) (:93,:137), this runner
batches all instances of one (algorithm, problem, dimension) into one lock-step engine on the GPU; runs are
independent (own seed `1000*fid + 10*dim + instance`, own generators), so results per run equal a sequential
execution. Results are written in IOHprofiler format (`IOHprofiler_f<id>_<name>.json` +
`data_f<id>_<name>/IOHprofiler_f<id>_DIM<d>.dat`) without needing ioh.
"""
import json
import os
from time import time
from typing import List, Optional

import numpy as np


class ExperimentRunner:
    def __init__(self, algorithms: List[str], dimensions: List[int], problem_ids: List[int], num_runs: int = 30,
                 budget_factor: int = 10, doe_factor: float = 3.0, root_dir: str = os.getcwd(),
                 experiment_name: str = "experiment", acquisition_function: str = "expected_improvement",
                 pca_components: Optional[int] = None, var_threshold: float = 0.95, verbose: bool = False,
                 device: str = "cuda:0", optimizer_backend: str = "scipy", rank: int = 0, world_size: int = 1):
        self.algorithms = algorithms
        self.dimensions = dimensions
        self.problem_ids = problem_ids
        self.num_runs = num_runs
        self.budget_factor = budget_factor
        self.doe_factor = doe_factor
        self.root_dir = root_dir
        self.experiment_name = experiment_name
        self.acquisition_function = acquisition_function
        self.pca_components = pca_components
        self.var_threshold = var_threshold
        self.verbose = verbose
        self.instances = range(self.num_runs)
        self.doe_params = {"criterion": "center", "iterations": 1000}
        self.device, self.optimizer_backend = device, optimizer_backend
        self.rank, self.world_size = rank, world_size
        self.results = []
        # ioh keeps ONE problem object per (function, dimension, instance) for the whole suite and the reference attaches
        # every algorithm's logger to that same suite (ExperimentRunner.py:90,135): the evaluation counter and the
        # best-so-far of a problem carry over from one algorithm to the next (the shipped vanilla .dat files start at
        # evaluation 76 with the PCA run's raw_y_best). Same state here, keyed by (fid, dim, instance).
        self._problem_state = {}

    def _acquisition_long_name(self):
        """The runner hands `acquisition_function` to the optimiser classes (reference ExperimentRunner.py:154-170), whose
        setter accepts the shorthands "EI" / "PI" / "UCB" (PCA_BO.py:653-677)."""
        from ..BayesianOptimization._acquisition_names import (ALLOWED_ACQUISITION_FUNCTION_STRINGS,
                                                                ALLOWED_SHORTHAND_ACQUISITION_FUNCTION_STRINGS)
        name = ALLOWED_SHORTHAND_ACQUISITION_FUNCTION_STRINGS.get(self.acquisition_function, self.acquisition_function)
        if name not in ALLOWED_ACQUISITION_FUNCTION_STRINGS:
            raise ValueError("Oddly defined name")
        if name == "upper_confidence_bound":
            raise TypeError("UpperConfidenceBound.__init__() got an unexpected keyword argument 'best_f'")
        return name

    def shard(self, instances):
        """Run-level data parallelism: instances round-robin over ranks, no collective."""
        return [i for j, i in enumerate(instances) if j % self.world_size == self.rank]

    def run_experiment(self) -> None:
        from pcabo_b200.bbob import BatchedBBOB, NAMES
        from pcabo_b200.engine import BatchedBOEngine
        total = len(self.algorithms) * len(self.problem_ids) * len(self.dimensions) * self.num_runs
        print(f"\nRunning {total} experiments ({len(self.algorithms)} algorithms × {len(self.dimensions)} dimensions × "
              f"{len(self.problem_ids)} problems × {self.num_runs} runs)\n")
        for algorithm in self.algorithms:
            if algorithm not in ("vanilla", "pca"):
                raise ValueError(f"Invalid algorithm name: '{algorithm}'")
            for fid in self.problem_ids:
                for dim in self.dimensions:
                    insts = self.shard(list(self.instances))
                    if not insts:
                        continue
                    budget = self.budget_factor * dim + 50                       # reference :144
                    n_doe = int(self.doe_factor * dim)                           # :145
                    seeds = [1000 * fid + 10 * dim + i for i in insts]           # :146
                    prob = BatchedBBOB([fid] * len(insts), insts, dim, device=self.device)
                    eng = BatchedBOEngine(prob, dim, prob.bounds, budget, n_doe, seeds, algorithm=algorithm,
                                          # like the reference, `pca_components` is only logged as an experiment
                                          # attribute: PCA_BO is built without n_components (ExperimentRunner.py:167-177)
                                          var_threshold=self.var_threshold, n_components=0,
                                          maximization=False, device=self.device, optimizer=self.optimizer_backend,
                                          profile_timing=True, acquisition=self._acquisition_long_name())
                    t0 = time()
                    eng.run()
                    wall = time() - t0
                    self._write_iohprofiler(algorithm, fid, NAMES[fid], dim, insts, eng, prob, wall)
                    self.results.append(dict(algorithm=algorithm, fid=fid, dim=dim, instances=insts, wall=wall,
                                             best=[float(v) for v in eng.current_best]))
                    if self.verbose:
                        print(f"{algorithm} | {dim}-dim | F-{fid} | {len(insts)} runs in {wall:.1f}s")

    def _write_iohprofiler(self, algorithm, fid, name, dim, insts, eng, prob, wall):
        folder = os.path.join(self.root_dir, f"{algorithm}-{self.experiment_name}")
        data_dir = os.path.join(folder, f"data_f{fid}_{name}")
        os.makedirs(data_dir, exist_ok=True)
        suffix = "" if self.world_size == 1 else f".rank{self.rank}"
        dat_rel = f"data_f{fid}_{name}/IOHprofiler_f{fid}_DIM{dim}{suffix}.dat"
        profiles = ["SingleTaskGP", "optimize_acqf"] + (["pca"] if algorithm == "pca" else [])
        runs = []
        with open(os.path.join(folder, dat_rel), "w") as fh:
            for b, inst in enumerate(insts):
                fopt = float(prob.fopt[b])
                st = self._problem_state.setdefault((fid, dim, inst), {"evals": 0, "best": np.inf, "best_eval": 0,
                                                                         "best_x": None})
                fh.write("evaluations raw_y raw_y_best " + " ".join(f"x{j}" for j in range(dim)) + "\n")
                for x, y, oob in zip(eng.x_evals[b], eng.f_evals[b], eng.oob_evals[b]):
                    if oob:
                        # PCA_BO.py:259-263: an out-of-bounds proposal gets f = +-1000 WITHOUT calling problem(x), so
                        # ioh's logger never sees it: no row, no evaluation counted
                        continue
                    st["evals"] += 1
                    raw = y - fopt                       # ioh's raw_y is the untransformed objective
                    if raw < st["best"]:
                        st["best"], st["best_eval"], st["best_x"] = raw, st["evals"], x
                    fh.write(f"{st['evals']} {raw:.10f} {st['best']:.10f} " + " ".join(f"{v:.6f}" for v in x) + "\n")
                run = {"instance": inst, "evals": st["evals"],
                       "best": {"evals": st["best_eval"], "y": st["best"], "x": [float(v) for v in st["best_x"]]}}
                for p in profiles:   # per-batch kernel time attributed evenly to the runs of the batch
                    run[f"{p}_time"] = float(sum(eng.timing_logs[p])) / len(insts)
                run["time"] = wall / len(insts)
                runs.append(run)
        # ioh keeps the attributes in a sorted map (shipped JSON: alphabetical order)
        exp_attrs = [{"acquisition_function": f"{self.acquisition_function}"}, {"budget_factor": f"{self.budget_factor}"},
                     {"doe_factor": f"{self.doe_factor}"}]
        if algorithm == "pca":
            exp_attrs += [{"pca_components": f"{self.pca_components}"}, {"var_threshold": f"{self.var_threshold}"}]
        meta = {"version": "0.3.18", "suite": "BBOB", "function_id": fid, "function_name": name, "maximization": False,
                "algorithm": {"name": algorithm, "info": f"A {algorithm}-BO Implementation."},
                "experiment_attributes": exp_attrs,
                "run_attributes": [f"{p}_time" for p in profiles] + ["time"],
                "attributes": ["evaluations", "raw_y", "raw_y_best"],
                "scenarios": [{"dimension": dim, "path": dat_rel, "runs": runs}]}
        with open(os.path.join(folder, f"IOHprofiler_f{fid}_{name}{suffix}.json"), "w") as fh:
            json.dump(meta, fh, indent=1)
