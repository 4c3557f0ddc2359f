"""Host-side pieces of botorch's optimize_acqf that must stay on the CPU generators for RNG-stream parity
(SURVEY H6): scrambled Sobol raw samples, the Boltzmann selection of initial conditions, and a lock-step
reverse-communication driver of scipy's own L-BFGS-B (`scipy.optimize._lbfgsb.setulb`) that lets B x 2 joint
restart problems share one batched CUDA value+gradient launch per round.

Reference call site: PCA_BO.optimize_acqf_and_get_observation, PCA_BO.py:547-621 (optimize_acqf at :607-614 with
num_restarts=10, raw_samples=512, options={"batch_limit": 5, "maxiter": 200}); semantics in SURVEY App. A.5.
The acquisition values and gradients themselves always come from the CUDA kernels (pcabo_logei).
"""
import numpy as np
import torch
from scipy.optimize import _lbfgsb

# `setulb` is a private scipy entry point: the Fortran wrapper of scipy <= 1.14 took (..., task, iprint, csave, lsave,
# isave, dsave, maxls) with a byte-string `task`, the C translation of scipy >= 1.15 takes integer task words and an
# `ln_task` array. LbfgsbProblem is written against the latter; any other signature must fail at import, not corrupt
# memory in the middle of a run.
_SETULB_SIGNATURE = "setulb(m, x, l, u, nbd, f, g, factr, pgtol, wa, iwa, task, lsave, isave, dsave, maxls, ln_task)"


def _check_setulb_signature():
    doc = (getattr(_lbfgsb.setulb, "__doc__", None) or "").strip().splitlines()
    got = doc[0].strip() if doc else ""
    if got != _SETULB_SIGNATURE:
        import scipy
        raise ImportError(f"pcabo_b200.acq_host drives scipy's L-BFGS-B through scipy.optimize._lbfgsb.setulb with the "
                          f"signature of scipy >= 1.15 ({_SETULB_SIGNATURE}); scipy {scipy.__version__} exposes "
                          f"'{got or 'no signature'}'. Use optimizer='device' or a matching scipy.")


_check_setulb_signature()

MAXCOR = 10
FTOL = 2.2204460492503131e-09
GTOL = 1e-5
MAXFUN = 15000
MAXLS = 20


def scrambled_sobol(d: int, n: int, generator: torch.Generator) -> torch.Tensor:
    """torch.quasirandom.SobolEngine(d, scramble=True).draw(n, dtype=float64) with an explicit CPU generator.

    Mirrors SobolEngine._scramble (two randint draws from the generator), so a generator seeded like the reference's
    global torch RNG reproduces botorch.utils.sampling.draw_sobol_samples bit for bit.
    """
    eng = torch.quasirandom.SobolEngine(d, scramble=False)
    maxbit = eng.MAXBIT
    shift_ints = torch.randint(2, (d, maxbit), generator=generator)
    eng.shift = torch.mv(shift_ints, torch.pow(2, torch.arange(0, maxbit)))
    ltm = torch.randint(2, (d, maxbit, maxbit), generator=generator).tril()
    torch._sobol_engine_scramble_(eng.sobolstate, ltm, d)
    eng.quasi = eng.shift.clone(memory_format=torch.contiguous_format)
    eng._first_point = (eng.quasi / 2 ** maxbit).reshape(1, -1)
    return eng.draw(n, dtype=torch.float64)


def initialize_q_batch(acq_vals: torch.Tensor, n: int, generator: torch.Generator, eta: float = 1.0):
    """botorch initialize_q_batch: indices of the n initial conditions, or None when all values coincide
    (botorch then warns BadInitialCandidatesWarning and the caller draws more raw samples)."""
    n_samples = acq_vals.shape[0]
    if n >= n_samples:
        return torch.arange(n_samples), False
    ystd = acq_vals.std(dim=0)
    if bool(ystd == 0) or not bool(torch.isfinite(ystd)):
        perm = torch.randperm(n_samples, generator=generator)[:n]
        return perm, True
    max_idx = torch.argmax(acq_vals)
    z = (acq_vals - acq_vals.mean(dim=0)) / ystd
    etaz = eta * z
    weights = torch.exp(etaz)
    while torch.isinf(weights).any():
        etaz *= 0.5
        weights = torch.exp(etaz)
    idcs = torch.multinomial(weights, n, generator=generator)
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return idcs, False


def select_initial_conditions(vals: torch.Tensor, n: int, gens, runs, q_buf=None, eta: float = 1.0):
    """botorch initialize_q_batch for many runs at once: vals [B, nraw] (CPU) -> (idx [B, n] int64, bad [B] bool).

    Boltzmann weights exp(eta (v - mean) / std) are computed batched; the sampling without replacement is torch's own
    algorithm restated (ATen multinomial: q ~ Exp(1) drawn with `exponential_` from the run's generator, top-n of
    weights / q), which consumes each generator exactly like torch.multinomial(weights, n, generator=g) and returns the
    same indices; the arg-max of the raw values is forced in as the last pick when it was not drawn. Runs whose values
    all coincide get a random permutation pick and are flagged `bad` (BadInitialCandidatesWarning in botorch)."""
    B, nraw = vals.shape
    idx = torch.zeros(B, n, dtype=torch.int64)
    bad = [False] * B
    if n >= nraw:
        idx[:] = torch.arange(n)
        return idx, bad
    std = vals.std(dim=1)
    mean = vals.mean(dim=1)
    etaz = eta * (vals - mean[:, None]) / std[:, None]
    w = torch.exp(etaz)
    inf_rows = torch.isinf(w).any(dim=1)
    for _ in range(64):        # botorch halves eta until the weights are finite; bounded here (inf input values)
        if not bool(inf_rows.any()):
            break
        etaz[inf_rows] *= 0.5
        w[inf_rows] = torch.exp(etaz[inf_rows])
        inf_rows = torch.isinf(w).any(dim=1)
    q = q_buf if q_buf is not None and q_buf.shape == vals.shape else torch.empty_like(vals)
    degenerate = ~(torch.isfinite(std) & (std != 0))
    for b in runs:
        if bool(degenerate[b]):
            idx[b] = torch.randperm(nraw, generator=gens[b])[:n]
            bad[b] = True
            q[b] = 1.0
            w[b] = 1.0
        else:
            q[b].exponential_(1, generator=gens[b])
    top = torch.topk(w / q, n, dim=1).indices
    amax = vals.argmax(dim=1)
    missing = ~(top == amax[:, None]).any(dim=1)
    top[missing, -1] = amax[missing]
    for b in runs:
        if not bad[b]:
            idx[b] = top[b]
    return idx, bad


def select_initial_conditions_nonneg(vals: torch.Tensor, n: int, gens, runs, eta: float = 1.0, alpha: float = 1e-4):
    """botorch initialize_q_batch_nonneg, which gen_batch_initial_conditions uses for acquisition functions registered
    as non-negative (ProbabilityOfImprovement among them), per run: vals [B, nraw] (CPU) -> (idx [B, n], bad [B]).

    max <= 0: random permutation pick + BadInitialCandidatesWarning (flagged `bad`, botorch retries with more raw
    samples); fewer than n positive values: all of them plus randomly chosen others (in index order, as boolean
    masking returns them); otherwise Boltzmann weights exp(eta (v / max - 1)) over the points within alpha of the
    maximum (alpha shrinking by 10 until there are n), multinomial without replacement, arg-max forced in last.
    Every random draw comes from the run's own generator, in botorch's order."""
    B, nraw = vals.shape
    idx = torch.zeros(B, n, dtype=torch.int64)
    bad = [False] * B
    if n >= nraw:
        idx[:] = torch.arange(n)
        return idx, bad
    for b in runs:
        v = vals[b]
        max_val, max_idx = torch.max(v, dim=0)
        if bool(max_val <= 0):
            idx[b] = torch.randperm(nraw, generator=gens[b])[:n]
            bad[b] = True
            continue
        pos = v > 0
        num_pos = int(pos.sum())
        if num_pos < n:
            remaining = (~pos).nonzero(as_tuple=False).view(-1)
            rand = torch.randperm(remaining.shape[0], generator=gens[b])
            pos[remaining[rand[: n - num_pos]]] = True
            idx[b] = pos.nonzero(as_tuple=False).view(-1)
            continue
        a = alpha
        alpha_pos = v >= a * max_val
        while int(alpha_pos.sum()) < n:
            a = 0.1 * a
            alpha_pos = v >= a * max_val
        alpha_idcs = torch.arange(nraw)[alpha_pos]
        weights = torch.exp(eta * (v[alpha_pos] / max_val - 1))
        pick = alpha_idcs[torch.multinomial(weights, n, generator=gens[b])]
        if int(max_idx) not in pick.tolist():
            pick[-1] = max_idx
        idx[b] = pick
    return idx, bad


MAXBIT = 30
_POW2_DESC = None
_STATE0 = {}


def sobol_state0(kmax: int) -> torch.Tensor:
    """Unscrambled direction numbers of torch's SobolEngine for `kmax` dimensions, [kmax, 30] int32 (cached)."""
    if kmax not in _STATE0:
        _STATE0[kmax] = torch.quasirandom.SobolEngine(kmax, scramble=False).sobolstate.to(torch.int32).contiguous()
    return _STATE0[kmax]


def draw_scramble_bits(d: int, generator: torch.Generator, shift_out: torch.Tensor, dots_out: torch.Tensor):
    """The two randint draws of SobolEngine._scramble for `d` dimensions from `generator`, reduced to what the device
    kernel needs: shift_out[:d] (30-bit shift per dimension) and dots_out[:d, :30] (rows of the lower-triangular
    matrices with unit diagonal packed MSB-first, ATen `cdot_pow2`). Consumes exactly the stream torch consumes."""
    global _POW2_DESC
    if _POW2_DESC is None:
        _POW2_DESC = torch.pow(2, torch.arange(MAXBIT - 1, -1, -1))
    shift_ints = torch.randint(2, (d, MAXBIT), generator=generator)
    ltm = torch.randint(2, (d, MAXBIT, MAXBIT), generator=generator).tril_()
    shift_out[:d] = torch.mv(shift_ints, torch.pow(2, torch.arange(0, MAXBIT)))
    ltm.diagonal(0, -2, -1).fill_(1)
    dots_out[:d] = (ltm * _POW2_DESC).sum(-1)
    shift_out[d:] = 0
    dots_out[d:] = 0


class GeneratorStateIO:
    """Moves the MT19937 state of B torch CPU generators to / from packed arrays for the device replay kernel.

    Layout of torch's CPU generator state (CPUGeneratorImplState, 5056 bytes): uint64 seed, int32 left, int32 seeded,
    uint64 next, uint64 state[624] (32-bit words), normal-cache fields. `left`/`next` relate to a plain
    "index of the next word" pos by pos = 624 if left == 1 else next, and back by next = pos, left = 625 - pos."""

    def __init__(self, B: int):
        self.B = B
        self.buf = np.zeros((B, 5056), dtype=np.uint8)
        self.tbuf = [torch.from_numpy(self.buf[b]) for b in range(B)]
        self.left = self.buf[:, 8:12].view(np.int32)[:, 0]
        self.next = self.buf[:, 16:24].view(np.uint64)[:, 0]
        self.words = self.buf[:, 24:24 + 4992].view(np.uint64)
        self.h_state = torch.zeros(B, 624, dtype=torch.int32)
        self.h_pos = torch.zeros(B, dtype=torch.int32)
        if torch.cuda.is_available():
            self.h_state, self.h_pos = self.h_state.pin_memory(), self.h_pos.pin_memory()

    def export(self, gens):
        for b, g in enumerate(gens):
            self.tbuf[b].copy_(g.get_state())
        self.h_state.numpy().view(np.uint32)[:] = self.words.astype(np.uint32)
        self.h_pos.numpy()[:] = np.where(self.left == 1, 624, self.next.astype(np.int64)).astype(np.int32)
        return self.h_state, self.h_pos

    def restore(self, gens, state: torch.Tensor, pos: torch.Tensor):
        st = state.numpy().view(np.uint32)
        ps = pos.numpy().astype(np.int64)
        self.words[:] = st.astype(np.uint64)
        self.left[:] = (625 - ps).astype(np.int32)
        self.next[:] = ps.astype(np.uint64)
        for b, g in enumerate(gens):
            g.set_state(self.tbuf[b])


class LbfgsbProblem:
    """One scipy L-BFGS-B minimisation driven by reverse communication (mirrors scipy _minimize_lbfgsb)."""

    __slots__ = ("n", "x", "f", "g", "lo", "hi", "nbd", "wa", "iwa", "task", "ln_task", "lsave", "isave", "dsave",
                 "nit", "nfev", "maxiter", "done", "warnflag")

    def __init__(self, x0, lo, hi, maxiter=200, nbd=None):
        """nbd: scipy's bound codes per variable (0 none, 1 lower, 2 both, 3 upper); default: all boxed."""
        x0 = np.asarray(x0, dtype=np.float64)
        n = x0.size
        m = MAXCOR
        self.n = n
        self.x = np.array(x0, dtype=np.float64)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros(n, dtype=np.float64)
        self.lo = np.array(lo, dtype=np.float64)
        self.hi = np.array(hi, dtype=np.float64)
        self.nbd = np.full(n, 2, dtype=np.int32) if nbd is None else np.asarray(nbd, dtype=np.int32).copy()
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self.iwa = np.zeros(3 * n, dtype=np.int32)
        self.task = np.zeros(2, dtype=np.int32)
        self.ln_task = np.zeros(2, dtype=np.int32)
        self.lsave = np.zeros(4, dtype=np.int32)
        self.isave = np.zeros(44, dtype=np.int32)
        self.dsave = np.zeros(29, dtype=np.float64)
        self.nit, self.nfev, self.maxiter = 0, 0, maxiter
        self.done, self.warnflag = False, 0

    def advance(self) -> bool:
        """Run setulb until it asks for f,g at self.x (returns True) or terminates (returns False)."""
        factr = FTOL / np.finfo(float).eps
        while True:
            _lbfgsb.setulb(MAXCOR, self.x, self.lo, self.hi, self.nbd, self.f, self.g, factr, GTOL, self.wa, self.iwa,
                           self.task, self.lsave, self.isave, self.dsave, MAXLS, self.ln_task)
            t = self.task[0]
            if t == 3:
                return True
            if t == 1:
                self.nit += 1
                if self.nit >= self.maxiter:
                    self.task[0], self.task[1] = 5, 504
                elif self.nfev > MAXFUN:
                    self.task[0], self.task[1] = 5, 502
                continue
            break
        self.done = True
        if self.task[0] == 4:
            self.warnflag = 0
        elif self.nfev > MAXFUN or self.nit >= self.maxiter:
            self.warnflag = 1
        else:
            self.warnflag = 2
        return False

    def supply(self, f, g):
        self.f = np.array(f, dtype=np.float64)
        self.g[:] = g
        self.nfev += 1

    @property
    def optimization_warning(self) -> bool:
        """botorch _process_scipy_result: failure other than the iteration / evaluation limit raises OptimizationWarning."""
        return self.warnflag == 2
