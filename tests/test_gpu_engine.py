"""GPU tests of the lock-step engine and the drop-in classes: trajectories against the CPU oracle loop."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle_run(fid, D, inst, iters, algorithm="pca"):
    from oracle import bbob, doe, loop
    torch.set_num_threads(1)
    p = bbob.BBOBProblem(fid, D, inst)
    budget, ndoe = doe.budget_and_ndoe(D)
    r = loop.OracleRun(p, D, p.bounds, budget, ndoe, doe.run_seed(fid, D, inst), algorithm=algorithm, keep_trace=True)
    r.run(max_iterations=iters)
    return r


def test_engine_trajectory_matches_oracle_c1():
    """Config C1 (f15, D=10, instance 0, seed 15100): first BO iterations, engine vs oracle on identical seeds."""
    from pcabo_b200.engine import BatchedBOEngine
    from pcabo_b200.bbob import BatchedBBOB
    fid, D, iters = 15, 10, 5
    insts = [0, 1]
    budget, ndoe = 10 * D + 50, 3 * D
    prob = BatchedBBOB([fid] * len(insts), insts, D, device="cuda")
    seeds = [1000 * fid + 10 * D + i for i in insts]
    eng = BatchedBOEngine(prob, D, prob.bounds, budget, ndoe, seeds, algorithm="pca")
    eng.initial_design()
    refs = [_oracle_run(fid, D, i, iters) for i in insts]
    for b, r in enumerate(refs):
        assert np.array_equal(np.array(eng.x_evals[b][:ndoe]), np.array(r.x_evals[:ndoe])), "DoE must be bit-identical"
        np.testing.assert_allclose(eng.f_evals[b][:ndoe], r.f_evals[:ndoe], rtol=1e-12)
    matched = [0] * len(insts)
    alive = [True] * len(insts)
    for it in range(iters):
        eng.step()
        for b, r in enumerate(refs):
            rec = r.trace[it]
            if not alive[b]:
                continue
            same_k = int(eng.last["k"][b]) == rec["k"]
            zerr = np.max(np.abs(eng.last["z_new"][b, :rec["k"]] - rec["new_z"])) / max(1.0, np.max(np.abs(rec["new_z"]))) \
                if same_k else np.inf
            if same_k and zerr < 1e-6 and bool(eng.last["oob"][b]) == rec["outside"]:
                matched[b] += 1
                assert abs(eng.last["acq_val"][b] - rec["acq_value"]) < 1e-6 * max(1.0, abs(rec["acq_value"]))
            else:
                alive[b] = False   # an L-BFGS-B line-search decision flipped: trajectories legitimately diverge
    # identical seeds, designs and candidate sets: the first iterations must coincide
    assert min(matched) >= 2, matched
    assert eng.stats["kernel_launches"] > 0 and eng.stats["lbfgsb_rounds"] > 0


def test_dropin_pca_bo_api():
    """`from Algorithms import PCA_BO` keeps the reference's call protocol and attributes."""
    from Algorithms import PCA_BO, Vanilla_BO, ExperimentRunner  # noqa: F401
    from pcabo_b200.bbob import BatchedBBOB
    D = 5
    problem = BatchedBBOB([15], [0], D).problem(0)
    opt = PCA_BO(budget=14, n_DoE=10, random_seed=15050, verbose=False,
                 DoE_parameters={"criterion": "center", "iterations": 1000})
    assert opt.TIME_PROFILES == ["SingleTaskGP", "optimize_acqf", "pca"]
    assert opt.acquisition_function_name == "expected_improvement"
    with pytest.raises(ValueError, match="Oddly defined name"):
        opt.acquisition_function_name = "thompson"
    opt.acquisition_function_name = "EI"
    opt(problem, dim=D, bounds=problem.bounds)
    assert opt.number_of_function_evaluations == 14 and len(opt.x_evals) == 14 and len(opt.f_evals) == 14
    # DoE identical to the reference's shipped data for seed 15050 (f15, D=5, instance 0)
    np.testing.assert_allclose(opt.x_evals[0], [-1.5, -1.5, -4.5, 3.5, 0.5], atol=1e-12)
    assert opt.current_best == min(opt.f_evals) and opt.f_evals.index(opt.current_best) == opt.current_best_index
    assert 1 <= opt.reduced_space_dim_num <= D and opt.component_matrix.shape == (D, D)
    assert opt.explained_variance_ratio.shape == (D,) and opt.data_mean.shape == (D,)
    for name in opt.TIME_PROFILES:
        assert len(opt.timing_logs[name]) == 4 and opt.total_times[name] > 0
    z = opt.z_evals[-1]
    x_back = opt._transform_point_to_original_space(z)
    np.testing.assert_allclose(x_back, opt.x_evals[-1], rtol=0, atol=1e-9)
    w = opt._calculate_weights()
    assert abs(w.sum() - 1.0) < 1e-12 and w.min() == 0.0
    opt.reset()
    assert opt.x_evals == [] and opt.pca is None

    van = Vanilla_BO(budget=13, n_DoE=10, random_seed=15050)
    van(problem, dim=D, bounds=problem.bounds)
    assert len(van.f_evals) == 13 and van.TIME_PROFILES == ["SingleTaskGP", "optimize_acqf"]
    X = np.array(van.x_evals)
    assert np.all(X >= -5 - 1e-12) and np.all(X <= 5 + 1e-12)
    with pytest.raises(TypeError):
        PCA_BO(budget=12, n_DoE=10)(lambda x: 0.0, dim=D, bounds=[[-5, 5]] * D)


def test_vanilla_engine_matches_oracle():
    from pcabo_b200.engine import BatchedBOEngine
    from pcabo_b200.bbob import BatchedBBOB
    fid, D, inst, iters = 20, 5, 3, 3
    prob = BatchedBBOB([fid], [inst], D, device="cuda")
    eng = BatchedBOEngine(prob, D, prob.bounds, 10 * D + 50, 3 * D, [1000 * fid + 10 * D + inst], algorithm="vanilla")
    eng.initial_design()
    r = _oracle_run(fid, D, inst, iters, algorithm="vanilla")
    eng.step()
    rec = r.trace[0]
    assert np.max(np.abs(eng.last["z_new"][0] - rec["new_z"])) < 1e-6 * max(1.0, np.max(np.abs(rec["new_z"])))
    assert abs(eng.last["new_f"][0] - rec["new_f"]) < 1e-6 * max(1.0, abs(rec["new_f"]))


def test_device_lbfgsb_matches_scipy_given_same_starts():
    """The device-resident L-BFGS-B against scipy's L-BFGS-B driven by the same CUDA value+gradient kernel, from the
    same initial conditions on the same GP: same optima (tolerance = L-BFGS-B's own stopping tolerance)."""
    from pcabo_b200 import ops
    from pcabo_b200.engine import BatchedBOEngine
    from pcabo_b200.bbob import BatchedBBOB
    fid, D = 15, 10
    insts = [0, 1, 2, 3]
    prob = BatchedBBOB([fid] * len(insts), insts, D, device="cuda")
    eng = BatchedBOEngine(prob, D, prob.bounds, 10 * D + 50, 3 * D, [1000 * fid + 10 * D + i for i in insts])
    eng.initial_design()
    for _ in range(3):
        eng.step()
    eng._fit_models()
    B = eng.B
    k_host = eng.pca["k"].cpu().numpy().astype(np.int64)
    box_lo, box_hi = eng.gp.box_lo.cpu(), eng.gp.box_hi.cpu()
    ics = eng._initial_conditions(list(range(B)), k_host)
    c_host, w_host = eng._optimize_scipy(list(range(B)), ics, k_host, box_lo, box_hi)
    nit_host = dict(eng.last["nit"])
    out_dev, w_dev = eng._optimize_device(ics, k_host)
    d_c = out_dev["cand"]
    nit_dev = dict(eng.last["nit"])
    vals_dev = ops.logei(eng.gp, d_c)["val"].cpu().numpy()
    h_c = torch.zeros(B, eng.num_restarts, D, dtype=torch.float64)
    for b in range(B):
        h_c[b, :, :int(k_host[b])] = torch.from_numpy(c_host[b])
    vals_host = ops.logei(eng.gp, h_c.cuda())["val"].cpu().numpy()
    same_iters = 0
    for b in range(B):
        assert not w_dev[b] and not w_host[b]
        # per chunk, the joint objective reached by both optimisers agrees to the solver tolerance
        for c in range(2):
            fh, fd = vals_host[b, 5 * c:5 * c + 5].sum(), vals_dev[b, 5 * c:5 * c + 5].sum()
            assert abs(fh - fd) <= 1e-5 * max(1.0, abs(fh)), (b, c, fh, fd, nit_host[b], nit_dev[b])
            same_iters += int(nit_host[b][c] == nit_dev[b][c])
        assert abs(vals_host[b].max() - vals_dev[b].max()) <= 1e-6 * max(1.0, abs(vals_host[b].max()))
    assert same_iters >= B   # most chunks follow scipy's iteration count exactly


def test_device_lbfgsb_subset_retry_in_place():
    """pcabo_lbfgsb_logei_runs: re-optimising a subset of runs (botorch's retry) with a tight k bound gives those runs
    bit-identical results to a full launch and leaves the rows of the other runs untouched; a wrong bound is refused."""
    from pcabo_b200 import ops
    from pcabo_b200.engine import BatchedBOEngine
    from pcabo_b200.bbob import BatchedBBOB
    fid, D = 15, 10
    insts = [0, 1, 2, 3, 4]
    prob = BatchedBBOB([fid] * len(insts), insts, D, device="cuda")
    eng = BatchedBOEngine(prob, D, prob.bounds, 10 * D + 50, 3 * D, [1000 * fid + 10 * D + i for i in insts])
    eng.initial_design()
    eng.step()
    eng._fit_models()
    B = eng.B
    k_host = eng.pca["k"].cpu().numpy().astype(np.int64)
    ics = eng._initial_conditions(list(range(B)), k_host)
    full = ops.lbfgsb_logei(eng.gp, ics, eng.gp.box_lo, eng.gp.box_hi)
    ref = {k: v.clone() for k, v in full.items()}
    sel = [3, 1]
    out = {k: torch.full_like(v, -7) for k, v in full.items()}
    ops.lbfgsb_logei(eng.gp, ics, eng.gp.box_lo, eng.gp.box_hi, out=out,
                     run_idx=torch.tensor(sel, dtype=torch.int32, device="cuda"), k_bound=int(k_host[sel].max()))
    for key in ("cand", "fval", "status", "nit", "nfev"):
        for b in range(B):
            if b in sel:
                assert torch.equal(out[key][b], ref[key][b]), (key, b)
            else:
                assert bool((out[key][b] == -7).all()), (key, b)
    if int(k_host.max()) > 1:
        bad = {k: torch.zeros_like(v) for k, v in full.items()}
        ops.lbfgsb_logei(eng.gp, ics, eng.gp.box_lo, eng.gp.box_hi, out=bad, k_bound=int(k_host.max()) - 1)
        st = bad["status"].cpu().numpy()
        assert (st[k_host == k_host.max()] == -1).all()


def test_engine_device_optimizer_trajectory():
    """optimizer='device': first iterations of config C1 against the oracle (same tolerance as the scipy mode)."""
    from pcabo_b200.engine import BatchedBOEngine
    from pcabo_b200.bbob import BatchedBBOB
    fid, D, inst, iters = 15, 10, 0, 3
    prob = BatchedBBOB([fid], [inst], D, device="cuda")
    eng = BatchedBOEngine(prob, D, prob.bounds, 10 * D + 50, 3 * D, [1000 * fid + 10 * D + inst], optimizer="device")
    eng.initial_design()
    r = _oracle_run(fid, D, inst, iters)
    eng.step()
    rec = r.trace[0]
    assert int(eng.last["k"][0]) == rec["k"]
    assert abs(eng.last["acq_val"][0] - rec["acq_value"]) < 1e-5 * max(1.0, abs(rec["acq_value"]))
    assert np.max(np.abs(eng.last["z_new"][0, :rec["k"]] - rec["new_z"])) < 1e-4 * max(1.0, np.max(np.abs(rec["new_z"])))
    for _ in range(iters - 1):
        eng.step()
    assert len(eng.f_evals[0]) == 3 * D + iters


def test_device_sobol_is_bit_identical_to_torch():
    """pcabo_sobol_points from host-drawn scramble bits == SobolEngine(k, scramble=True).draw(n, float64) scaled to the box."""
    from pcabo_b200 import ops
    from pcabo_b200.acq_host import draw_scramble_bits, sobol_state0
    from oracle import acqf
    B, D, n = 3, 40, 512
    ks = [31, 1, 40]
    shift = torch.zeros(B, D, dtype=torch.int64)
    dots = torch.zeros(B, D, 30, dtype=torch.int64)
    lo = torch.randn(B, D, dtype=torch.float64)
    hi = lo + torch.rand(B, D, dtype=torch.float64) * 3 + 0.1
    refs = []
    for b in range(B):
        g1, g2 = torch.Generator(), torch.Generator()
        g1.manual_seed(100 + b)
        g2.manual_seed(100 + b)
        draw_scramble_bits(ks[b], g1, shift[b], dots[b])
        raw = acqf.scrambled_sobol(ks[b], n, g2)
        refs.append(lo[b, :ks[b]] + (hi[b, :ks[b]] - lo[b, :ks[b]]) * raw)
        assert g1.get_state().equal(g2.get_state())          # identical stream consumption
    k = torch.tensor(ks, dtype=torch.int32).cuda()
    X = ops.sobol_points(n, k, shift.to(torch.int32).cuda(), dots.to(torch.int32).cuda(), sobol_state0(D).cuda(),
                         lo.cuda(), hi.cuda()).cpu()
    for b in range(B):
        assert torch.equal(X[b, :, :ks[b]], refs[b]), b
        assert float(X[b, :, ks[b]:].abs().max()) == 0.0 if ks[b] < D else True
    idx = torch.tensor([[5, 0, 511], [1, 1, 2], [100, 7, 3]], dtype=torch.int32)
    got = ops.gather_rows(X.cuda(), idx.cuda()).cpu()
    for b in range(B):
        assert torch.equal(got[b], X[b, idx[b].long()])


def test_device_mt19937_replay_matches_torch_randint():
    """pcabo_sobol_scramble_bits replays torch's CPU generator: same packed scramble bits as the two torch.randint
    draws of SobolEngine._scramble, and the generator ends in the state torch would have left it in."""
    from pcabo_b200 import ops
    from pcabo_b200.acq_host import GeneratorStateIO, draw_scramble_bits
    B, D = 5, 40
    ks = [31, 1, 40, 0, 17]
    g_dev = [torch.Generator() for _ in range(B)]
    g_ref = [torch.Generator() for _ in range(B)]
    for b in range(B):
        g_dev[b].manual_seed(500 + b)
        g_ref[b].manual_seed(500 + b)
        pre = [0, 3, 624, 1000, 12345][b]          # different positions inside / across the 624-word state
        if pre:
            torch.randint(2, (pre,), generator=g_dev[b])
            torch.randint(2, (pre,), generator=g_ref[b])
    io = GeneratorStateIO(B)
    for rep in range(2):                            # twice: the restored state must keep following torch
        st, pos = io.export(g_dev)
        d_state, d_pos = st.cuda(), pos.cuda()
        k = torch.tensor(ks, dtype=torch.int32).cuda()
        shift, dots = ops.sobol_scramble_bits(k, d_state, d_pos, D)
        io.restore(g_dev, d_state.cpu(), d_pos.cpu())
        for b in range(B):
            sh_ref = torch.zeros(D, dtype=torch.int64)
            dots_ref = torch.zeros(D, 30, dtype=torch.int64)
            if ks[b] > 0:
                draw_scramble_bits(ks[b], g_ref[b], sh_ref, dots_ref)
            assert torch.equal(shift[b].cpu().long(), sh_ref), (rep, b)
            assert torch.equal(dots[b].cpu().long(), dots_ref), (rep, b)
            a = torch.rand(5, generator=g_dev[b], dtype=torch.float64)
            r = torch.rand(5, generator=g_ref[b], dtype=torch.float64)
            assert torch.equal(a, r), (rep, b)


def test_pipelined_sub_batches_equal_one_batch():
    """PipelinedEngines (sub-batches on separate streams, step_begin / step_end interleaved by one host thread) gives
    every run the trajectory it has inside one big lock-step batch: runs are independent and the kernels treat them
    independently. (The batched objective goes through cuBLAS, whose kernel choice depends on the batch size, so f can
    differ in the last bits: compared to 1e-9 over the first iterations, before L-BFGS-B amplifies such bits.)"""
    from pcabo_b200.engine import BatchedBOEngine, PipelinedEngines
    from pcabo_b200.bbob import BatchedBBOB
    D, budget, ndoe, steps = 10, 150, 30, 2
    fids = [15, 15, 16, 16, 17, 17]
    insts = [0, 1, 0, 1, 0, 1]
    seeds = [1000 * f + 10 * D + i for f, i in zip(fids, insts)]

    def make(sel):
        prob = BatchedBBOB([fids[i] for i in sel], [insts[i] for i in sel], D, device="cuda")
        return BatchedBOEngine(prob, D, prob.bounds, budget, ndoe, [seeds[i] for i in sel], optimizer="device")

    whole = make(list(range(6)))
    whole.initial_design()
    for _ in range(steps):
        whole.step()
    S = 3
    subs = [make(list(range(s, 6, S))) for s in range(S)]
    pipe = PipelinedEngines(subs)
    pipe.initial_design()
    pipe.run_steps(steps)
    torch.cuda.synchronize()
    assert pipe.n == whole.n == ndoe + steps
    same = 0
    for s, e in enumerate(subs):
        for j in range(e.B):
            b = s + S * j
            fa, fb = np.array(e.f_evals[j]), np.array(whole.f_evals[b])
            xa, xb = np.array(e.x_evals[j]), np.array(whole.x_evals[b])
            assert np.allclose(fa[:ndoe + 1], fb[:ndoe + 1], rtol=1e-9, atol=1e-9), (s, j)
            assert np.allclose(xa[:ndoe + 1], xb[:ndoe + 1], rtol=1e-9, atol=1e-9), (s, j)
            same += int(np.allclose(xa, xb, rtol=1e-6, atol=1e-6))
    assert same >= 5
    # the asynchronous tail (device-side penalty / append / best-so-far, host lists filled from packed records) against
    # the synchronous one on the same sub-batches: every number identical
    assert all(e.async_tail for e in subs)
    subs2 = [make(list(range(s, 6, S))) for s in range(S)]
    pipe2 = PipelinedEngines(subs2, async_tail=False)
    pipe2.initial_design()
    pipe2.run_steps(steps)
    assert not any(e.async_tail for e in subs2)
    for e, e2 in zip(subs, subs2):
        assert not e.pending_host
        for j in range(e.B):
            assert e.f_evals[j] == e2.f_evals[j] and e.oob_evals[j] == e2.oob_evals[j], j
            assert np.array_equal(np.array(e.x_evals[j]), np.array(e2.x_evals[j])), j
        assert e.current_best == e2.current_best and e.current_best_index == e2.current_best_index
        assert np.array_equal(e.last["z_new"], e2.last["z_new"]) and np.array_equal(e.last["acq_val"], e2.last["acq_val"])
        assert torch.equal(e.X[:, :e.n], e2.X[:, :e2.n]) and torch.equal(e.f[:, :e.n], e2.f[:, :e2.n])
    # one more step through the plain API on an engine that used the asynchronous tail: the host lists are complete
    e = subs[0]
    with torch.cuda.stream(pipe.streams[0]):
        e.step()
    assert len(e.f_evals[0]) == e.n and not e.pending_host


def test_engine_probability_of_improvement_matches_oracle():
    """acquisition="probability_of_improvement": first iterations against the oracle loop (non-negative initialiser,
    Phi(u) values and gradients in the device optimiser)."""
    from pcabo_b200.engine import BatchedBOEngine
    from pcabo_b200.bbob import BatchedBBOB
    from oracle import bbob, doe, loop
    fid, D, inst, iters = 15, 10, 0, 3
    budget, ndoe = doe.budget_and_ndoe(D)
    torch.set_num_threads(1)
    p = bbob.BBOBProblem(fid, D, inst)
    r = loop.OracleRun(p, D, p.bounds, budget, ndoe, doe.run_seed(fid, D, inst), keep_trace=True,
                       acquisition="probability_of_improvement")
    r.run(max_iterations=iters)
    for optimizer in ("scipy", "device"):
        prob = BatchedBBOB([fid], [inst], D, device="cuda")
        eng = BatchedBOEngine(prob, D, prob.bounds, budget, ndoe, [1000 * fid + 10 * D + inst], optimizer=optimizer,
                              acquisition="probability_of_improvement")
        eng.initial_design()
        eng.step()
        rec = r.trace[0]
        assert int(eng.last["k"][0]) == rec["k"]
        tol = 1e-6 if optimizer == "scipy" else 1e-4
        assert abs(eng.last["acq_val"][0] - rec["acq_value"]) < tol * max(1.0, abs(rec["acq_value"])), optimizer
        assert np.max(np.abs(eng.last["z_new"][0, :rec["k"]] - rec["new_z"])) < \
            100 * tol * max(1.0, np.max(np.abs(rec["new_z"]))), optimizer
    with pytest.raises(ValueError):
        BatchedBOEngine(prob, D, prob.bounds, budget, ndoe, [1], acquisition="upper_confidence_bound")


def test_device_initial_condition_selection_equals_torch_multinomial():
    """pcabo_select_initial_conditions against the host restatement of botorch's initialize_q_batch (which uses
    torch's own exponential_ / topk on the run's CPU generator): same picks, same generator state afterwards; runs with
    degenerate values are flagged and left untouched."""
    from pcabo_b200 import ops
    from pcabo_b200.acq_host import GeneratorStateIO, select_initial_conditions
    B, nraw, kmax, R = 6, 512, 8, 10
    rng = np.random.RandomState(2)
    vals = torch.from_numpy(rng.randn(B, nraw) * np.array([1.0, 5.0, 40.0, 300.0, 1e-3, 1.0])[:, None])
    vals[5] = 0.25                                   # degenerate: all values equal
    vals[3, 7] = 2000.0                              # weights overflow: eta is halved
    X = torch.from_numpy(rng.rand(B, nraw, kmax))
    gens_a, gens_b = [], []
    for b in range(B):
        for gl in (gens_a, gens_b):
            g = torch.Generator()
            g.manual_seed(100 + b)
            _ = torch.randint(2, (b * 37 + 5,), generator=g)      # move every generator to a different position
            gl.append(g)
    io = GeneratorStateIO(B)
    st, pos = io.export(gens_a)
    d_st, d_pos = st.cuda(), pos.cuda()
    out = ops.select_initial_conditions(vals.cuda(), d_st, d_pos, X.cuda(), R)
    torch.cuda.synchronize()
    io.restore(gens_a, d_st.cpu(), d_pos.cpu())
    idx_h, bad = select_initial_conditions(vals.clone(), R, gens_b, list(range(B - 1)))
    flag = out["flag"].cpu().numpy()
    assert flag.tolist() == [0, 0, 0, 0, 0, 1] and not any(bad[:5])
    idx_d = out["idx"].cpu().numpy()
    for b in range(B - 1):
        assert sorted(idx_d[b].tolist()) == sorted(idx_h[b].tolist()), b
        assert idx_d[b, -1] == idx_h[b, -1] or int(torch.argmax(vals[b])) in idx_d[b].tolist()
        assert torch.equal(gens_a[b].get_state(), gens_b[b].get_state()), b
        picked = out["ics"][b].cpu()
        assert torch.equal(picked, X[b][torch.from_numpy(idx_d[b]).long()])
    g5 = torch.Generator()
    g5.manual_seed(105)
    _ = torch.randint(2, (5 * 37 + 5,), generator=g5)
    assert torch.equal(gens_a[5].get_state(), g5.get_state())     # flagged run: generator untouched


def test_engine_device_and_host_initial_conditions_agree():
    """The same runs with the Boltzmann pick on the device (default) and on the host (force_host_ic): identical first
    iterations (same picks from the same generator streams)."""
    from pcabo_b200.engine import BatchedBOEngine
    from pcabo_b200.bbob import BatchedBBOB
    D, budget, ndoe = 10, 150, 30
    fids, insts = [15, 16, 17, 15], [0, 1, 2, 3]
    seeds = [1000 * f + 10 * D + i for f, i in zip(fids, insts)]
    engs = []
    for force in (False, True):
        prob = BatchedBBOB(fids, insts, D, device="cuda")
        e = BatchedBOEngine(prob, D, prob.bounds, budget, ndoe, seeds, optimizer="device")
        e.force_host_ic = force
        e.initial_design()
        for _ in range(3):
            e.step()
        engs.append(e)
    a, b = engs
    for r in range(len(fids)):
        assert np.allclose(np.array(a.x_evals[r]), np.array(b.x_evals[r]), rtol=1e-9, atol=1e-9), r
    a._gens_to_host()
    for r in range(len(fids)):
        assert torch.equal(a.torch_gens[r].get_state(), b.torch_gens[r].get_state()), r


def test_device_initial_conditions_fall_back_for_flagged_runs():
    """A run the device pick flags as degenerate goes through the host path (generators synchronised from the device
    states first): its initial conditions are the host pick's rows, the other runs keep the device pick."""
    from pcabo_b200 import ops
    from pcabo_b200.engine import BatchedBOEngine
    from pcabo_b200.bbob import BatchedBBOB
    D, budget, ndoe = 10, 150, 30
    fids, insts = [15, 16, 17], [0, 1, 2]
    prob = BatchedBBOB(fids, insts, D, device="cuda")
    e = BatchedBOEngine(prob, D, prob.bounds, budget, ndoe, [1000 * f + 10 * D + i for f, i in zip(fids, insts)],
                        optimizer="device")
    e.initial_design()
    e.step()
    e._fit_models()
    h = e._ic_launch(list(range(3)), None, 1)
    assert h.get("fast")
    h["done"].synchronize()
    dev_pick = e.ic_out["ics"].clone()
    e.h_flag[1] = 1                                   # pretend run 1 had degenerate raw values
    d_ics = torch.zeros(3, e.num_restarts, D, dtype=torch.float64, device="cuda")
    still = e._ic_finish(h, d_ics)
    torch.cuda.synchronize()
    assert still == [] and e.gens_on_host
    assert torch.equal(d_ics[0], dev_pick[0]) and torch.equal(d_ics[2], dev_pick[2])
    # run 1: rows of its raw samples chosen by the host path (a valid pick containing the arg-max)
    vals = h["vals_dev"][1].cpu()
    rows = h["x_raw"][1].cpu()
    amax_row = rows[int(torch.argmax(vals))]
    assert any(torch.equal(d_ics[1, r].cpu(), amax_row) for r in range(e.num_restarts))
    # the engine keeps going from the synchronised generators
    e.step()
    assert e.n == ndoe + 2


def test_device_optimizer_workspace_and_no_host_fallback(monkeypatch):
    """The device-resident L-BFGS-B keeps its state in shared memory; where that does not fit (config C5: n up to 1050,
    k ~ 90) the BFGS pairs move to a caller-owned L2-resident workspace (pcabo_lbfgsb_logei_runs_ws). There is no host
    fallback: a size the library cannot take raises."""
    from pcabo_b200 import ops
    from pcabo_b200.engine import BatchedBOEngine
    from pcabo_b200.bbob import BatchedBBOB
    from pcabo_b200.lib import PcaboError
    assert ops.lbfgsb_workspace_doubles(450, 40, 40) == 0 and ops.lbfgsb_workspace_doubles(150, 10, 10) == 0
    need = ops.lbfgsb_workspace_doubles(1050, 100, 90, 10, 5, 3)
    assert need == 3 * 2 * 2 * 10 * (((5 * 90 + 15) & ~15) + 8)   # runs x chunks x (S, Y) x m x padded row of nvar
    assert ops.lbfgsb_fits(1050, 100, 90) and ops.lbfgsb_fits(1050, 100, 100)
    prob = BatchedBBOB([15, 16], [0, 1], 10, device="cuda")
    e = BatchedBOEngine(prob, 10, prob.bounds, 150, 30, [15100, 16101], optimizer="device")
    e.initial_design()
    e.step()
    monkeypatch.setattr(ops, "lbfgsb_fits", lambda *a, **k: False)
    with pytest.raises(PcaboError, match="beyond"):
        e.step()


def test_engine_with_hyperparameter_fit_opt_in():
    """fit_hyperparameters=True (extension; the reference never fits): the engine fits per-run (lengthscale, noise) every
    iteration and proposes from the fitted model; default stays the reference's fixed prior values."""
    from pcabo_b200.engine import BatchedBOEngine
    from pcabo_b200.bbob import BatchedBBOB
    prob = BatchedBBOB([15, 20], [0, 1], 10, device="cuda")
    e = BatchedBOEngine(prob, 10, prob.bounds, 150, 30, [15100, 20101], optimizer="device", fit_hyperparameters=True)
    e.initial_design()
    e.step()
    e.step()
    assert e.n == 32 and e.last_hyper is not None
    ls, nz = e.last_hyper["lengthscale"], e.last_hyper["noise"]
    assert (ls > 0).all() and (nz >= 1e-4).all() and (e.last_hyper["nfev"] >= 1).all()
    assert e.gp.lengthscales is not None and abs(float(e.gp.lengthscales[0]) - ls[0]) < 1e-12
    d = BatchedBOEngine(prob, 10, prob.bounds, 150, 30, [15100, 20101], optimizer="device")
    d.initial_design()
    d.step()
    assert d.last_hyper is None and d.gp.lengthscales is None
