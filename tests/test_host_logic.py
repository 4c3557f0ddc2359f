"""CPU tests of the product's host side: the C-ABI library loads and exports every declared symbol (no compute
without a GPU), the product fails loudly without CUDA, host mirrors of reference logic, and the world_size-2 gloo
path of run sharding / candidate-sharded argmax."""
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from pcabo_b200 import load_library, EXPORTED_SYMBOLS
    lib = load_library()
    header = open(os.path.join(ROOT, "include", "pcabo_b200.h")).read()
    declared = set(re.findall(r"\b(pcabo_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(EXPORTED_SYMBOLS), declared ^ set(EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.pcabo_abi_version() == 2
    assert lib.pcabo_error_string(-2).decode().startswith("size not supported")
    # argument validation happens before any CUDA call
    assert lib.pcabo_argmax(0, 0, None, None, None, None) == -1


def test_no_cpu_fallback():
    from pcabo_b200 import ops, PcaboError
    with pytest.raises(PcaboError, match="CUDA tensor"):
        ops.argmax(torch.zeros(2, 3, dtype=torch.float64))
    if not torch.cuda.is_available():
        from pcabo_b200.engine import BatchedBOEngine
        with pytest.raises(PcaboError, match="no CPU fallback"):
            BatchedBOEngine([lambda x: 0.0], 2, np.array([[-1, 1], [-1, 1.0]]), 10, 4, [1])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "para-ortho-pca-bo_b200")
    for base, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(base, fn)


def test_host_doe_and_bbob_match_reference_goldens():
    import json
    from pcabo_b200.engine import lhs_center, rescale_lhs
    from pcabo_b200.bbob import BatchedBBOB
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")))
    b5 = np.tile([[-5.0, 5.0]], (5, 1))
    for fid in (15, 20):
        prob = BatchedBBOB([fid] * 30, list(range(30)), 5)
        for inst, rows in enumerate(kat["doe"][str(fid)]):
            X = rescale_lhs(lhs_center(np.random.RandomState(1000 * fid + 50 + inst), 5, 10), b5)
            ref = np.array([r["x"] for r in rows])
            assert np.array_equal(np.round(X, 6), ref)
            y = prob.eval_many(inst, X) - float(prob.fopt[inst])
            np.testing.assert_allclose(y, [r["raw_y"] for r in rows], rtol=2e-12)


def test_lbfgsb_reverse_communication_equals_scipy_minimize():
    from scipy.optimize import minimize, rosen, rosen_der
    from pcabo_b200.acq_host import LbfgsbProblem
    rng = np.random.RandomState(0)
    for _ in range(4):
        n = 10
        lo, hi = np.full(n, -1.5), np.full(n, 1.2)
        x0 = np.clip(rng.uniform(-2, 2, n), lo, hi)
        res = minimize(lambda x: (rosen(x), rosen_der(x)), x0, jac=True, method="L-BFGS-B",
                       bounds=list(zip(lo, hi)), options={"maxiter": 200})
        p = LbfgsbProblem(x0, lo, hi, 200)
        while p.advance():
            p.supply(rosen(p.x), rosen_der(p.x))
        assert np.array_equal(res.x, p.x) and res.nit == p.nit and res.nfev == p.nfev
        assert (res.status == 0) == (p.warnflag == 0)


def test_initialize_q_batch_matches_oracle():
    from pcabo_b200.acq_host import initialize_q_batch, scrambled_sobol
    from oracle import acqf
    g1, g2 = torch.Generator(), torch.Generator()
    g1.manual_seed(3)
    g2.manual_seed(3)
    assert torch.equal(scrambled_sobol(7, 64, g1), acqf.scrambled_sobol(7, 64, g2))
    vals = torch.randn(64, dtype=torch.float64)
    X = torch.arange(64, dtype=torch.float64).reshape(64, 1)
    idx, bad = initialize_q_batch(vals, 10, g1)
    Xo, _ = acqf.initialize_q_batch(X, vals, 10, generator=g2)
    assert not bad and torch.equal(X[idx], Xo) and int(torch.argmax(vals)) in idx.tolist()
    idx, bad = initialize_q_batch(torch.zeros(64, dtype=torch.float64), 10, g1)
    assert bad and len(idx) == 10


def test_dropin_surface_without_gpu():
    from Algorithms import PCA_BO, Vanilla_BO, ExperimentRunner
    from Algorithms.BayesianOptimization.AbstractBayesianOptimizer import LHS_sampler
    opt = PCA_BO(budget=30, n_DoE=6, var_threshold=0.9, random_seed=5, maximization=True, verbose=False)
    assert opt.budget == 30 and opt.n_DoE == 6 and opt.maximization and opt.var_threshold == 0.9
    assert opt.current_best == -np.inf and opt.random_seed == 5
    opt.acquisition_function_name = " UCB "
    assert opt.acquisition_function_name == "upper_confidence_bound"
    with pytest.raises(ValueError):
        opt.acquisition_function_name = "nonsense"
    with pytest.raises(AttributeError):
        opt.acquisition_function = object()
    with pytest.raises(ValueError):
        opt.number_of_function_evaluations = -1
    with pytest.raises(AssertionError):
        PCA_BO(budget=0)
    with pytest.raises(ValueError):
        LHS_sampler(criterion="bogus")
    s = LHS_sampler(criterion="c", iterations=3)
    assert s.criterion == "center"
    np.random.seed(15050)
    H = s(5, 10)
    np.testing.assert_allclose(10 * H[0] - 5, [-1.5, -1.5, -4.5, 3.5, 0.5], atol=1e-12)
    for crit in ("maximin", "centermaximin", "correlation"):
        Hc = LHS_sampler(criterion=crit, iterations=3)(3, 6)
        assert Hc.shape == (6, 3) and np.all((Hc >= 0) & (Hc <= 1))
        assert all(sorted(np.floor(Hc[:, j] * 6).astype(int)) == list(range(6)) for j in range(3))
    v = Vanilla_BO(budget=20, acquisition_function="EI")
    assert v.TIME_PROFILES == ["SingleTaskGP", "optimize_acqf"] and v.torch_config["RAW_SAMPLES"] == 512
    r = ExperimentRunner(["pca"], [40], [15, 16, 17], num_runs=30, world_size=4, rank=1)
    assert r.shard(list(range(30))) == list(range(1, 30, 4))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world_size_2_gloo_sharding_and_argmax():
    """N > 1 host path on CPU: run sharding is disjoint and complete; the candidate-sharded argmax exchange
    (value, global index) picks the global first-occurrence maximum with one all_gather."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, script, str(r), "2", str(port)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "argmax ok" in outs[0]


def test_incremental_best_tracking_equals_rescan():
    """BatchedBOEngine._update_best_with_last == AbstractBayesianOptimizer.assign_new_best (:196-208) on sequences with
    ties, penalties and both optimisation senses (host logic only: no engine is constructed)."""
    from pcabo_b200.engine import BatchedBOEngine
    rng = np.random.RandomState(0)
    for maximization in (False, True):
        for trial in range(20):
            vals = rng.choice([1000.0, -1000.0, 3.5, 3.5, 7.25, -2.0, 0.0], size=60).tolist()
            vals += rng.randn(20).tolist()
            rng.shuffle(vals)
            a = BatchedBOEngine.__new__(BatchedBOEngine)
            b = BatchedBOEngine.__new__(BatchedBOEngine)
            for e in (a, b):
                e.maximization = maximization
                e.f_evals = [list(vals[:5])]
                e.current_best, e.current_best_index = [0.0], [0]
                BatchedBOEngine._assign_new_best(e, 0)
            for v in vals[5:]:
                a.f_evals[0].append(v)
                b.f_evals[0].append(v)
                a._assign_new_best(0)
                b._update_best_with_last(0)
                assert (a.current_best, a.current_best_index) == (b.current_best, b.current_best_index)


def test_initialize_q_batch_nonneg_matches_oracle():
    """The per-run non-negative initialiser (ProbabilityOfImprovement) against the oracle's restatement of botorch's
    initialize_q_batch_nonneg on the three branches: enough positive values, too few, none."""
    from pcabo_b200.acq_host import select_initial_conditions_nonneg
    from oracle import acqf
    rng = np.random.RandomState(4)
    cases = [np.abs(rng.randn(64)) * 0.3,                                  # all positive: alpha window + multinomial
             np.where(rng.rand(64) < 0.08, rng.rand(64), 0.0),             # fewer than n positive
             np.zeros(64)]                                                 # max <= 0: random pick, flagged bad
    cases[1][5] = 0.7
    X = torch.arange(64, dtype=torch.float64).reshape(64, 1)
    for ci, v in enumerate(cases):
        vals = torch.from_numpy(v)
        g1, g2 = torch.Generator(), torch.Generator()
        g1.manual_seed(11 + ci)
        g2.manual_seed(11 + ci)
        idx, bad = select_initial_conditions_nonneg(vals.reshape(1, -1).clone(), 10, [g1], [0])
        import warnings
        with warnings.catch_warnings(record=True) as ws:
            warnings.simplefilter("always")
            Xo, _ = acqf.initialize_q_batch_nonneg(X, vals.clone(), 10, generator=g2)
        assert torch.equal(X[idx[0]], Xo), ci
        assert bad[0] == any(issubclass(w.category, acqf.BadInitialCandidatesWarning) for w in ws)
        assert torch.equal(g1.get_state(), g2.get_state())                # same draws consumed


def test_dropin_acquisition_choices():
    """"probability_of_improvement" is accepted; "upper_confidence_bound" fails exactly where the reference fails
    (acquisition_function_class(model=, best_f=, maximize=) on a class without best_f, PCA_BO.py:199-203)."""
    from Algorithms import PCA_BO, Vanilla_BO, ExperimentRunner
    f = lambda x: float(np.sum(x * x))
    for cls in (PCA_BO, Vanilla_BO):
        opt = cls(budget=12, n_DoE=4, acquisition_function="UCB", verbose=False)
        with pytest.raises(TypeError):
            opt(f, dim=2, bounds=np.array([[-1.0, 1.0], [-1.0, 1.0]]))
        assert cls(budget=12, acquisition_function="PI").acquisition_function_name == "probability_of_improvement"
    with pytest.raises(TypeError):
        ExperimentRunner(["pca"], [5], [15], num_runs=1, acquisition_function="UCB")._acquisition_long_name()
    assert ExperimentRunner(["pca"], [5], [15], num_runs=1, acquisition_function="PI")._acquisition_long_name() == \
        "probability_of_improvement"


def test_iohprofiler_writer_matches_reference_layout(tmp_path):
    """The IOHprofiler writer against the layout of the reference's shipped pca-/vanilla-experiment files (fixture
    `ioh_format`, extracted by tests/golden/make_golden.py): key order, attribute names, header, and -- because the
    DoE, BBOB f15 and the run's best point are pinned -- the first data rows VERBATIM, including ioh's carry-over of the
    evaluation counter and best-so-far into the second algorithm's file (ExperimentRunner.py:90-135). Out-of-bounds
    proposals are penalised without calling problem(x) (PCA_BO.py:259-263), so they must not appear as rows."""
    import json as _json
    from types import SimpleNamespace
    from Algorithms.Experiment.ExperimentRunner import ExperimentRunner
    from pcabo_b200.bbob import BatchedBBOB, NAMES
    from pcabo_b200.engine import lhs_center, rescale_lhs
    kat = _json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")))
    fmt = kat["ioh_format"]
    fid, D, inst = 15, 5, 0
    prob = BatchedBBOB([fid], [inst], D)
    doe_x = rescale_lhs(lhs_center(np.random.RandomState(1000 * fid + 10 * D + inst), D, 10), prob.bounds)
    best_x = np.array(kat["best"][str(fid)][inst]["x"])
    rng = np.random.RandomState(0)
    runner = ExperimentRunner(["pca", "vanilla"], [D], [fid], num_runs=1, budget_factor=5, doe_factor=2.0,
                              root_dir=str(tmp_path), experiment_name="experiment", pca_components=0)

    def fake_engine(n_logged, with_oob):
        xs, oob = [x for x in doe_x], [False] * 10
        while len(xs) - sum(oob) < n_logged:
            logged = len(xs) - sum(oob)
            if with_oob and logged in (20, 21, 60) and not oob[-1]:
                xs.append(np.full(D, 7.0)); oob.append(True)              # penalised, never evaluated
            elif with_oob and logged == 47:
                xs.append(best_x); oob.append(False)                      # the reference run's best point, evaluation 48
            else:
                xs.append(rng.uniform(-5, 5, D) * 0 + np.array([5.0, -5.0, 5.0, -5.0, 5.0])); oob.append(False)
        fs = [1000.0 if o else float(prob.eval_many(0, x)[0]) for x, o in zip(xs, oob)]
        return SimpleNamespace(x_evals=[xs], f_evals=[fs], oob_evals=[oob],
                               timing_logs={"pca": [0.1], "SingleTaskGP": [0.2], "optimize_acqf": [0.3]})

    for alg in ("pca", "vanilla"):
        eng = fake_engine(75, with_oob=(alg == "pca"))
        runner._write_iohprofiler(alg, fid, NAMES[fid], D, [inst], eng, prob, 1.5)
        js = _json.load(open(tmp_path / f"{alg}-experiment" / f"IOHprofiler_f{fid}_{NAMES[fid]}.json"))
        ref = fmt[alg]
        assert list(js.keys()) == ref["top_keys"]
        assert js["algorithm"] == ref["algorithm"] and js["version"] == ref["version"] and js["suite"] == ref["suite"]
        assert js["run_attributes"] == ref["run_attributes"] and js["attributes"] == ref["attributes"]
        names = [list(a.keys())[0] for a in js["experiment_attributes"]]
        assert names == sorted(names) and set(ref["experiment_attribute_names"]) <= set(names)
        sc = js["scenarios"][0]
        assert list(sc.keys()) == ref["scenario_keys"] and sc["path"] == ref["path"] and sc["dimension"] == D
        run0 = sc["runs"][0]
        assert list(run0.keys()) == ref["run_keys"] and list(run0["best"].keys()) == ref["best_keys"]
        assert run0["evals"] == ref["run0_evals"]                          # 75, then 150 (ioh's counter carries over)
        lines = open(tmp_path / f"{alg}-experiment" / sc["path"]).read().splitlines()
        assert lines[:4] == ref["dat_first_lines"]                         # verbatim, incl. 76 ... 73.0856639101
        assert len(lines) == ref["dat_block_lines"]                        # header + 75 rows: OOB proposals not logged
        assert [int(l.split()[0]) for l in lines[1:]] == list(range(run0["evals"] - 74, run0["evals"] + 1))
        if alg == "pca":
            assert run0["best"]["evals"] == kat["best"][str(fid)][inst]["evals"]
            assert abs(run0["best"]["y"] - kat["best"][str(fid)][inst]["y"]) < 1e-9


def test_host_side_abi_planning_functions():
    """The ABI's host-only planning functions need no GPU: the packed-W size (w_packed.cuh), the optimiser kernel's
    shared-memory plan and its workspace need, for the configs of BASELINE (C1..C5)."""
    from pcabo_b200 import ops
    lib = ops.load_library()
    # packed W: groups of 8 rows, row stride 8 g + (g odd ? 12 : 20), back to back
    def packed(n):
        return sum(8 * (8 * g + (12 if g & 1 else 20)) for g in range((n + 7) // 8))
    for n in (1, 8, 9, 123, 150, 250, 450, 1050):
        assert lib.pcabo_gp_packed_doubles(n) == packed(n)
    assert lib.pcabo_gp_packed_doubles(0) == -1
    # plans: everything resident while it fits (32-row panels), streaming 16- then 8-row panels beyond; always <= 227 KB;
    # the BFGS pairs move to the caller's workspace only for the D = 100 stress config
    for (n, kmax, k) in [(31, 10, 8), (150, 10, 10), (123, 40, 32)]:
        p = ops.lbfgsb_plan(n, kmax, k)
        assert p["resident"] == 1 and p["panel_rows"] == 32 and p["pairs_in_workspace"] == 0
    for (n, kmax, k) in [(250, 20, 20), (300, 40, 40), (450, 40, 40), (303, 100, 90)]:
        p = ops.lbfgsb_plan(n, kmax, k)
        assert p["resident"] == 0 and p["panel_rows"] in (8, 16) and p["pairs_in_workspace"] == 0
        assert p["smem_bytes"] <= 227 * 1024 and p["ring_doubles"] >= 2 * p["panel_rows"] * (n + 20) // 2 or p["panel_rows"] == 8
        assert ops.lbfgsb_workspace_doubles(n, kmax, k) == 0
    for k in (90, 100):
        p = ops.lbfgsb_plan(1050, 100, k)
        assert p["pairs_in_workspace"] == 1 and p["panel_rows"] == 8 and p["smem_bytes"] <= 227 * 1024
        nvar = 5 * k
        assert ops.lbfgsb_workspace_doubles(1050, 100, k, 10, 5, 7) == 7 * 2 * 2 * 10 * (((nvar + 15) & ~15) + 8)
    assert ops.lbfgsb_workspace_doubles(450, 40, 40, 10, 8, 1) == -1        # batch_limit > 5 is not supported
    assert ops.gp_fit_launches(450) == 2 + 15 + 1 + 3 + 1


def test_sub_batch_schedule_with_simulated_engines():
    """PipelinedEngines' event-driven schedule as pure host logic, on simulated engines with a simulated clock: every
    sub-batch gets exactly k iterations, each as A -> B -> E with B / E only after the event of the phase before has
    completed, a sub-batch that finishes early is served before an older one that is still running, and the host pauses
    only when nothing is ready."""
    import contextlib
    from pcabo_b200.engine import schedule_sub_batches
    clock = [0.0]
    log = []

    class Ev:
        def __init__(self, t):
            self.t = t

        def query(self):
            return clock[0] >= self.t

    class Eng:
        def __init__(self, name, rng):
            self.name, self.rng, self.phase, self.ev, self.steps = name, rng, None, None, 0

        def step_launch(self):
            assert self.phase in (None, "E")
            self.phase, self.ev = "A", Ev(clock[0] + 1.0)                       # model fit: 1 ms
            log.append((self.name, "A", clock[0]))
            clock[0] += 0.3                                                     # host time of the call

        def step_begin(self):
            assert self.phase == "A" and self.ev.query(), "optimiser launched before the model fit completed"
            self.phase, self.ev = "B", Ev(clock[0] + self.rng.choice([4.0, 6.0, 13.0]))   # optimiser kernel
            log.append((self.name, "B", clock[0]))
            clock[0] += 0.3

        def step_end(self):
            assert self.phase == "B" and self.ev.query(), "tail enqueued before the optimiser finished"
            self.phase, self.ev = "E", None
            self.steps += 1
            log.append((self.name, "E", clock[0]))
            clock[0] += 0.3

        def wait_event(self):
            return self.ev

    def pause():
        clock[0] += 0.05

    rng = np.random.RandomState(0)
    engines = [Eng(i, rng) for i in range(5)]
    seen_ctx = []

    @contextlib.contextmanager
    def ctx(token):
        seen_ctx.append(token)
        yield

    idle = schedule_sub_batches([(e, 100 + e.name) for e in engines], 7, ctx, pause)
    assert all(e.steps == 7 and e.phase == "E" for e in engines)
    assert idle >= 0.0 and set(seen_ctx) == {100, 101, 102, 103, 104}
    for e in engines:                      # per engine: A B E A B E ... (the last iteration launches no further A)
        seq = "".join(p for n_, p, _ in log if n_ == e.name)
        assert seq == "ABE" * 7, seq
    # nothing is ever served while something that completed earlier in launch order of the same kind waits:
    # the host was never idle for longer than the shortest remaining event (it polls, it does not block on one event)
    total = clock[0]
    busy = 0.3 * len(log)
    assert total < busy + 7 * 14.0 + 5.0   # far below the serial sum of all kernels (5 engines x 7 x ~8 ms)
    # one iteration, one engine: degenerate sizes work too
    log.clear()
    one = Eng(0, rng)
    schedule_sub_batches([(one, 0)], 1, ctx, pause)
    assert one.steps == 1 and [p for _, p, _ in log] == ["A", "B", "E"]
    assert schedule_sub_batches([], 3, ctx, pause) == 0.0


def test_raw_scoring_support_predicate_is_host_only():
    """pcabo_logei_raw_supported (no device work): the tile pipeline covers log-EI models with n <= 512, even capacities
    and its shared memory within 227 KB -- config C3 yes, config C5 (D = 100) no, probability of improvement no."""
    from pcabo_b200.lib import load_library
    lib = load_library()
    yes = [(450, 40, 450), (123, 40, 450), (512, 32, 512), (150, 10, 150), (250, 20, 250), (2, 2, 2)]
    no = [(300, 100, 1050), (513, 32, 514), (450, 41, 450), (449, 40, 449), (450, 40, 448), (1, 40, 450), (450, 106, 450)]
    assert all(lib.pcabo_logei_raw_supported(n, k, ld, 0) == 1 for n, k, ld in yes)
    assert all(lib.pcabo_logei_raw_supported(n, k, ld, 0) == 0 for n, k, ld in no)
    assert lib.pcabo_logei_raw_supported(450, 40, 450, 1) == 0
