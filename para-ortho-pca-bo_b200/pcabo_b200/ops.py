"""Torch-tensor front end of the C ABI: validates device/dtype/layout, passes raw device pointers and the current stream."""
import ctypes
import math

import torch

from .lib import GpModelStruct, PcaboError, check, load_library

LENGTHSCALE = math.log(2.0)   # softplus(0): gpytorch MaternKernel default raw_lengthscale = 0
NOISE = math.exp(-5.0)        # mode of botorch's LogNormalPrior(-4, 1) noise prior


def _dev_ptr(t, dtype, name):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise PcaboError(f"{name}: expected a CUDA tensor (the PCA-BO hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise PcaboError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise PcaboError(f"{name}: expected a contiguous tensor")
    return ctypes.c_void_p(t.data_ptr())


def _d(t, name):
    return _dev_ptr(t, torch.float64, name)


def _i(t, name):
    return _dev_ptr(t, torch.int32, name)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    # (torch.cuda.current_stream() builds a Stream object: ~10 us per call, a dozen calls per sub-batch and iteration)
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def weights_pca_fit(X, f, noise, logtab, n, var_threshold=0.95, n_components=0, maximization=False, out=None):
    """X [B, ld_n, D], f [B, ld_n], noise [B, n, D], logtab [>= n + 1]. Returns dict of device tensors."""
    lib = load_library()
    B, ld_n, D = X.shape
    dev = X.device
    if out is None:
        out = dict(
            w=torch.zeros(B, ld_n, dtype=torch.float64, device=dev),
            ranks=torch.zeros(B, ld_n, dtype=torch.int32, device=dev),
            xbar=torch.empty(B, D, dtype=torch.float64, device=dev),
            m2=torch.empty(B, D, dtype=torch.float64, device=dev),
            V=torch.empty(B, D, D, dtype=torch.float64, device=dev),
            lam=torch.empty(B, D, dtype=torch.float64, device=dev),
            evr=torch.empty(B, D, dtype=torch.float64, device=dev),
            k=torch.empty(B, dtype=torch.int32, device=dev),
            Z=torch.zeros(B, ld_n, D, dtype=torch.float64, device=dev),
        )
    if noise.shape != (B, n, D):
        raise PcaboError(f"noise: expected shape {(B, n, D)}, got {tuple(noise.shape)}")
    if logtab.numel() < n + 1:
        raise PcaboError("logtab: needs at least n + 1 entries")
    check(lib.pcabo_weights_pca_fit(B, n, D, ld_n, _d(X, "X"), _d(f, "f"), _d(noise, "noise"), _d(logtab, "logtab"),
                                    float(var_threshold), int(n_components), int(bool(maximization)),
                                    _d(out["w"], "w"), _i(out["ranks"], "ranks"), _d(out["xbar"], "xbar"),
                                    _d(out["m2"], "m2"), _d(out["V"], "V"), _d(out["lam"], "lam"),
                                    _d(out["evr"], "evr"), _i(out["k"], "k"), _d(out["Z"], "Z"), _stream()),
          "pcabo_weights_pca_fit")
    return out


def inverse_map(z, k, V, m2, xbar, lo, hi, out=None):
    """z [B, D] (first k[b] entries used) -> x [B, D], oob [B]."""
    lib = load_library()
    B, D = z.shape
    if out is None:
        out = dict(x=torch.empty(B, D, dtype=torch.float64, device=z.device),
                   oob=torch.empty(B, dtype=torch.int32, device=z.device))
    check(lib.pcabo_inverse_map(B, D, _d(z, "z"), _i(k, "k"), _d(V, "V"), _d(m2, "m2"), _d(xbar, "xbar"),
                                _d(lo, "lo"), _d(hi, "hi"), _d(out["x"], "x"), _i(out["oob"], "oob"), _stream()),
          "pcabo_inverse_map")
    return out


ACQ_LOG_EI, ACQ_PI = 0, 1
ACQ_KINDS = {"expected_improvement": ACQ_LOG_EI, "probability_of_improvement": ACQ_PI}


class GPBuffers:
    """Device storage for the batched GP caches (allocated once per engine, reused every iteration)."""

    def __init__(self, B, ld_n, kmax, device):
        f64 = dict(dtype=torch.float64, device=device)
        self.B, self.ld_n, self.kmax = B, ld_n, kmax
        self.norm_off = torch.zeros(B, kmax, **f64)
        self.norm_coef = torch.ones(B, kmax, **f64)
        self.box_lo = torch.zeros(B, kmax, **f64)
        self.box_hi = torch.zeros(B, kmax, **f64)
        self.ymean = torch.zeros(B, **f64)
        self.ystd = torch.ones(B, **f64)
        self.ZnT = torch.zeros(B, kmax, ld_n, **f64)
        self.Zn = torch.zeros(B, ld_n, kmax, **f64)
        self.L = torch.zeros(B, ld_n, ld_n, **f64)
        self.W = torch.zeros(B, ld_n, ld_n, **f64)
        self.WT = torch.zeros(B, ld_n, ld_n, **f64)
        # W packed in 8-row groups for the optimiser kernel's bulk copies (pcabo_gp_pack_w; ~ld_n^2 / 2 doubles per run)
        self.wp_stride = int(load_library().pcabo_gp_packed_doubles(int(ld_n)))
        self.Wp = torch.zeros(B, self.wp_stride, **f64)
        self.alpha = torch.zeros(B, ld_n, **f64)
        self.info = torch.zeros(B, dtype=torch.int32, device=device)
        self.best_f = torch.zeros(B, **f64)
        self.n = 0
        self.k = None
        self.maximize = False
        self.acq = ACQ_LOG_EI       # PCABO_ACQ_LOG_EI / PCABO_ACQ_PI (include/pcabo_b200.h)
        self.lengthscale = LENGTHSCALE
        self.lengthscales = None    # [B] per-run lengthscales after an opt-in hyper-parameter fit (hyperfit.py)
        self.noises = None          # [B] per-run noise variances, ditto


def gp_fit(buf: GPBuffers, Z, k, f, n, mode=0, fix_lo=None, fix_hi=None, lengthscale=LENGTHSCALE, noise=NOISE,
           jitter=None, lengthscales=None, noises=None):
    """Z [B, ld_n, kmax], k [B] int32 or None, f [B, ld_n]. Fills `buf`. jitter: optional [B] device tensor added to the
    diagonal of K + noise I (psd_safe_cholesky's retry ladder for runs whose buf.info came back non-zero).
    lengthscales / noises: optional [B] device tensors of per-run hyper-parameters (opt-in fit; the reference uses the
    fixed prior values for every run)."""
    lib = load_library()
    B, ld_n, kmax = Z.shape
    if (B, ld_n, kmax) != (buf.B, buf.ld_n, buf.kmax):
        raise PcaboError("gp_fit: buffer shape mismatch")
    check(lib.pcabo_gp_fit_hyper(B, n, kmax, ld_n, int(mode), _d(Z, "Z"), _i(k, "k"), _d(f, "f"), _d(fix_lo, "fix_lo"),
                                 _d(fix_hi, "fix_hi"), float(lengthscale), float(noise), _d(lengthscales, "lengthscales"),
                                 _d(noises, "noises"), _d(jitter, "jitter"),
                                 _d(buf.norm_off, "norm_off"), _d(buf.norm_coef, "norm_coef"), _d(buf.box_lo, "box_lo"),
                                 _d(buf.box_hi, "box_hi"), _d(buf.ymean, "ymean"), _d(buf.ystd, "ystd"), _d(buf.ZnT, "ZnT"),
                                 _d(buf.Zn, "Zn"), _d(buf.L, "L"), _d(buf.W, "W"), _d(buf.WT, "WT"), _d(buf.alpha, "alpha"),
                                 _i(buf.info, "info"), _stream()),
          "pcabo_gp_fit")
    check(lib.pcabo_gp_pack_w(B, n, ld_n, _d(buf.W, "W"), _d(buf.Wp, "Wp"), buf.wp_stride, _stream()), "pcabo_gp_pack_w")
    buf.n, buf.k, buf.lengthscale = n, k, lengthscale
    buf.lengthscales, buf.noises = lengthscales, noises
    return buf


def gp_fit_launches(n):
    """Kernel launches of one pcabo_gp_fit + pcabo_gp_pack_w call: prep, kernel matrix, per 64-column block an update (+
    a solve below the last), the triangular inverse, alpha, the transpose, the packing."""
    nb = (int(n) + 63) // 64
    return 2 + (2 * nb - 1) + (1 if nb > 1 else 0) + 3 + 1


def gp_mll(buf: GPBuffers, f):
    lib = load_library()
    out = torch.empty(buf.B, dtype=torch.float64, device=f.device)
    check(lib.pcabo_gp_mll(buf.B, buf.n, buf.ld_n, _d(f, "f"), _d(buf.ymean, "ymean"), _d(buf.ystd, "ystd"),
                           _d(buf.L, "L"), _d(buf.alpha, "alpha"), _d(out, "mll"), _stream()), "pcabo_gp_mll")
    return out


def gp_mll_grad(buf: GPBuffers):
    """(dMLL/dlengthscale, dMLL/dnoise) of the standardised targets from the caches: [B, 2] (device)."""
    lib = load_library()
    out = torch.empty(buf.B, 2, dtype=torch.float64, device=buf.W.device)
    check(lib.pcabo_gp_mll_grad_hyper(buf.B, buf.n, buf.kmax, buf.ld_n, _i(buf.k, "k") if buf.k is not None else None,
                                      _d(buf.Zn, "Zn"), _d(buf.W, "W"), _d(buf.alpha, "alpha"), float(buf.lengthscale),
                                      _d(buf.lengthscales, "lengthscales"), _d(out, "grad"), _stream()),
          "pcabo_gp_mll_grad")
    return out


def model_struct(buf: GPBuffers) -> GpModelStruct:
    # (built three or four times per sub-batch and iteration from the same fields: cached on those fields)
    key = (buf.n, None if buf.k is None else buf.k.data_ptr(), int(buf.acq), bool(buf.maximize), float(buf.lengthscale),
           None if buf.lengthscales is None else buf.lengthscales.data_ptr(), buf.best_f.data_ptr(), buf.Wp.data_ptr(),
           buf.W.data_ptr(), buf.alpha.data_ptr())
    if getattr(buf, "_ms_key", None) == key:
        return buf._ms
    m = _model_struct(buf)
    buf._ms_key, buf._ms = key, m
    return m


def _model_struct(buf: GPBuffers) -> GpModelStruct:
    m = GpModelStruct()
    m.B, m.n, m.kmax, m.ld_n = buf.B, buf.n, buf.kmax, buf.ld_n
    m.maximize = int(bool(buf.maximize))
    m.acq = int(buf.acq)
    m.lengthscale = float(buf.lengthscale)
    m.k = _i(buf.k, "k") if buf.k is not None else None
    m.norm_off, m.norm_coef = _d(buf.norm_off, "norm_off"), _d(buf.norm_coef, "norm_coef")
    m.ZnT, m.alpha, m.W = _d(buf.ZnT, "ZnT"), _d(buf.alpha, "alpha"), _d(buf.W, "W")
    m.ymean, m.ystd, m.best_f = _d(buf.ymean, "ymean"), _d(buf.ystd, "ystd"), _d(buf.best_f, "best_f")
    m.WT, m.Zn = _d(buf.WT, "WT"), _d(buf.Zn, "Zn")
    m.Wp, m.wp_stride = _d(buf.Wp, "Wp"), int(buf.wp_stride)
    m.lengthscales = _d(buf.lengthscales, "lengthscales") if buf.lengthscales is not None else None
    return m


def logei(buf: GPBuffers, X, want_grad=False, want_posterior=False, out=None):
    """X [B, npts, kmax] raw-space points -> dict(val [B, npts], grad [B, npts, kmax], mu, sigma)."""
    lib = load_library()
    B, npts, kmax = X.shape
    if B != buf.B or kmax != buf.kmax:
        raise PcaboError("logei: shape mismatch with the GP buffers")
    dev = X.device
    if out is None:
        out = dict(val=torch.empty(B, npts, dtype=torch.float64, device=dev))
        if want_grad:
            out["grad"] = torch.empty(B, npts, kmax, dtype=torch.float64, device=dev)
        if want_posterior:
            out["mu"] = torch.empty(B, npts, dtype=torch.float64, device=dev)
            out["sigma"] = torch.empty(B, npts, dtype=torch.float64, device=dev)
    m = model_struct(buf)
    check(lib.pcabo_logei(ctypes.byref(m), npts, _d(X, "X"), _d(out["val"], "val"), _d(out.get("grad"), "grad"),
                          _d(out.get("mu"), "mu"), _d(out.get("sigma"), "sigma"), _stream()), "pcabo_logei")
    return out


def logei_raw_supported(buf: GPBuffers, n=None) -> bool:
    """True when pcabo_logei_raw_runs (the tensor-core tile pipeline) covers this model: log-EI, n <= 512, even row
    capacity and column capacity, shared memory of the pipeline within 227 KB (the library decides)."""
    n = buf.n if n is None else n
    return bool(load_library().pcabo_logei_raw_supported(int(n), int(buf.kmax), int(buf.ld_n), int(buf.acq)))


def logei_raw(buf: GPBuffers, X, out=None, ctas_per_run=0):
    """Value-only log-EI of the raw samples X [B, npts, kmax] of every run in one launch of the sweep pipeline on the FP64
    tensor cores (pcabo_logei_raw_runs) -> dict(val [B, npts])."""
    lib = load_library()
    B, npts, kmax = X.shape
    if B != buf.B or kmax != buf.kmax:
        raise PcaboError("logei_raw: shape mismatch with the GP buffers")
    if out is None:
        out = dict(val=torch.empty(B, npts, dtype=torch.float64, device=X.device))
    m = model_struct(buf)
    check(lib.pcabo_logei_raw_runs(ctypes.byref(m), npts, _d(X, "X"), _d(out["val"], "val"), int(ctas_per_run), _stream()),
          "pcabo_logei_raw_runs")
    return out


def argmax(v):
    """v [B, m] -> (best_val [B], best_idx [B] int32), first occurrence on ties."""
    lib = load_library()
    B, m = v.shape
    bv = torch.empty(B, dtype=torch.float64, device=v.device)
    bi = torch.empty(B, dtype=torch.int32, device=v.device)
    check(lib.pcabo_argmax(B, m, _d(v, "v"), _d(bv, "best_val"), _i(bi, "best_idx"), _stream()), "pcabo_argmax")
    return bv, bi


def pick_best(vals, cand, out=None):
    """vals [B, R], cand [B, R, D] -> dict(z [B, D], val [B], idx [B] int32): first-occurrence arg-max over the restarts
    and the winning candidate (what optimize_acqf returns)."""
    lib = load_library()
    B, R, D = cand.shape
    if out is None:
        out = dict(z=torch.empty(B, D, dtype=torch.float64, device=cand.device),
                   val=torch.empty(B, dtype=torch.float64, device=cand.device),
                   idx=torch.empty(B, dtype=torch.int32, device=cand.device))
    check(lib.pcabo_pick_best(B, R, D, _d(vals, "vals"), _d(cand, "cand"), _d(out["z"], "z"), _d(out["val"], "val"),
                              _i(out["idx"], "idx"), _stream()), "pcabo_pick_best")
    return out


def append_observation(X, f, n, x_new, z_new, f_new, oob, acq, penalty, maximization, best, best_f, rec):
    """Device-side tail of a BO iteration (pcabo_append_observation): penalty for out-of-bounds proposals, append to the
    history X [B, ld_n, D] / f [B, ld_n] at row n, running best and its float32 rounding, packed record rec [B, 2D+3]."""
    lib = load_library()
    B, ld_n, D = X.shape
    check(lib.pcabo_append_observation(B, D, ld_n, int(n), _d(x_new, "x_new"), _d(z_new, "z_new"), _d(f_new, "f_new"),
                                       _i(oob, "oob"), _d(acq, "acq"), float(penalty), int(bool(maximization)),
                                       _d(X, "X"), _d(f, "f"), _d(best, "best"), _d(best_f, "best_f"), _d(rec, "rec"),
                                       _stream()), "pcabo_append_observation")


LBFGSB_STATUS = {0: "CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL", 1: "CONVERGENCE: RELATIVE REDUCTION OF F <= FACTR*EPSMCH",
                 2: "STOP: TOTAL NO. OF ITERATIONS REACHED LIMIT", 3: "ABNORMAL"}


def lbfgsb_workspace_doubles(n, kmax, k_bound, num_restarts=10, batch_limit=5, nsel=1):
    """Doubles of caller-owned device workspace pcabo_lbfgsb_logei_runs_ws needs for `nsel` runs: 0 when the whole
    L-BFGS-B state fits shared memory, -1 when the problem is not supported at all (the library answers
    PCABO_E_UNSUPPORTED)."""
    return int(load_library().pcabo_lbfgsb_workspace_doubles(int(n), int(kmax), int(k_bound), int(num_restarts),
                                                             int(batch_limit), int(nsel)))


def lbfgsb_plan(n, kmax, k_bound, batch_limit=5, have_workspace=True):
    """The launcher's shared-memory plan as a dict (panel_rows, ring_doubles, resident, pairs_in_workspace, smem_bytes)."""
    out = (ctypes.c_int * 6)()
    check(load_library().pcabo_lbfgsb_plan(int(n), int(kmax), int(k_bound), int(batch_limit), int(bool(have_workspace)), out),
          "pcabo_lbfgsb_plan")
    d = dict(zip(("panel_rows", "ring_doubles", "resident", "reserved", "pairs_in_workspace", "smem_bytes"), list(out)))
    d.pop("reserved")
    return d


def lbfgsb_fits(n, kmax, k_bound, batch_limit=5):
    """Whether the device optimiser accepts this size (with a workspace if it needs one)."""
    return kmax <= 104 and batch_limit <= 5 and lbfgsb_workspace_doubles(n, kmax, k_bound, 10, batch_limit) >= 0


def lbfgsb_logei(buf: GPBuffers, ics, box_lo, box_hi, batch_limit=5, maxiter=200, factr=1e7, pgtol=1e-5, out=None,
                 prof=None, run_idx=None, k_bound=0, workspace=None):
    """ics [B, num_restarts, kmax] -> dict(cand [B, R, kmax], fval/status/nit/nfev [B, nchunks]); one kernel launch.

    run_idx (device int32 [nsel]): optimise only these runs, the rows of the other runs in `out` stay as they are.
    k_bound: an upper bound of k over the runs optimised (sizes the shared memory; 0 = kmax).
    workspace: device float64 tensor for the BFGS pairs of problems that do not fit shared memory (config C5); allocated
    here when it is needed and not given (pass one to keep the hot loop allocation-free)."""
    lib = load_library()
    B, R, kmax = ics.shape
    if B != buf.B or kmax != buf.kmax:
        raise PcaboError("lbfgsb_logei: shape mismatch with the GP buffers")
    nch = (R + batch_limit - 1) // batch_limit
    dev = ics.device
    if out is None:
        out = dict(cand=torch.empty(B, R, kmax, dtype=torch.float64, device=dev),
                   fval=torch.empty(B, nch, dtype=torch.float64, device=dev),
                   status=torch.empty(B, nch, dtype=torch.int32, device=dev),
                   nit=torch.empty(B, nch, dtype=torch.int32, device=dev),
                   nfev=torch.empty(B, nch, dtype=torch.int32, device=dev))
    m = model_struct(buf)
    nsel = B if run_idx is None else int(run_idx.numel())
    need = lbfgsb_workspace_doubles(buf.n, kmax, k_bound, R, batch_limit, nsel)
    if need > 0 and (workspace is None or workspace.numel() < need):
        workspace = torch.empty(need, dtype=torch.float64, device=dev)
    check(lib.pcabo_lbfgsb_logei_runs_ws(ctypes.byref(m), nsel, _i(run_idx, "run_idx"), int(k_bound), R, int(batch_limit),
                                         int(maxiter), float(factr), float(pgtol),
                                         _d(ics, "ics"), _d(box_lo, "box_lo"), _d(box_hi, "box_hi"),
                                         _d(out["cand"], "cand"), _d(out["fval"], "fval"), _i(out["status"], "status"),
                                         _i(out["nit"], "nit"), _i(out["nfev"], "nfev"),
                                         _dev_ptr(prof, torch.int64, "prof"),
                                         _d(workspace, "workspace") if need > 0 else None,
                                         int(workspace.numel()) if need > 0 else 0, _stream()),
          "pcabo_lbfgsb_logei_runs_ws")
    return out


def sobol_points(npts, k, shift, ltm_dots, state0, box_lo, box_hi, out=None):
    """shift [B, kmax] int32, ltm_dots [B, kmax, 30] int32, state0 [kmax, 30] int32 -> X [B, npts, kmax] (device)."""
    lib = load_library()
    B, kmax = shift.shape
    if out is None:
        out = torch.empty(B, npts, kmax, dtype=torch.float64, device=shift.device)
    check(lib.pcabo_sobol_points(B, int(npts), kmax, _i(k, "k"), _i(shift, "shift"), _i(ltm_dots, "ltm_dots"),
                                 _i(state0, "state0"), _d(box_lo, "box_lo"), _d(box_hi, "box_hi"), _d(out, "X"),
                                 _stream()), "pcabo_sobol_points")
    return out


def gather_rows(X, idx, out=None):
    """X [B, npts, kmax], idx [B, nsel] int32 -> out [B, nsel, kmax]."""
    lib = load_library()
    B, npts, kmax = X.shape
    nsel = idx.shape[1]
    if out is None:
        out = torch.empty(B, nsel, kmax, dtype=torch.float64, device=X.device)
    check(lib.pcabo_gather_rows(B, npts, nsel, kmax, _d(X, "X"), _i(idx, "idx"), _d(out, "out"), _stream()),
          "pcabo_gather_rows")
    return out


def sobol_scramble_bits(k, mt_state, mt_pos, kmax, out=None):
    """mt_state [B, 624] int32 (uint32 words), mt_pos [B] int32, advanced in place -> (shift [B, kmax], dots [B, kmax, 30])."""
    lib = load_library()
    B = mt_state.shape[0]
    dev = mt_state.device
    if out is None:
        out = (torch.empty(B, kmax, dtype=torch.int32, device=dev), torch.empty(B, kmax, 30, dtype=torch.int32, device=dev))
    check(lib.pcabo_sobol_scramble_bits(B, kmax, _i(k, "k"), _i(mt_state, "mt_state"), _i(mt_pos, "mt_pos"),
                                        _i(out[0], "shift"), _i(out[1], "ltm_dots"), _stream()),
          "pcabo_sobol_scramble_bits")
    return out


class SweepWorkspace:
    """Scratch and result buffers of pcabo_ei_sweep (allocate once, reuse)."""

    def __init__(self, device, nparts=None):
        if nparts is None:
            nparts = torch.cuda.get_device_properties(device).multi_processor_count
        self.nparts = int(nparts)
        self.part_val = torch.empty(self.nparts, dtype=torch.float64, device=device)
        self.part_idx = torch.empty(self.nparts, dtype=torch.int64, device=device)
        self.best_val = torch.empty(1, dtype=torch.float64, device=device)
        self.best_idx = torch.empty(1, dtype=torch.int64, device=device)


def ei_sweep(buf: GPBuffers, ws: SweepWorkspace, ncand, idx0=0, run=0, X=None, sobol=None, vals=None):
    """log-EI of `ncand` candidates against run `run` with a fused arg-max.

    X: explicit candidates [ncand, kmax] (device) or None with sobol = (shift [kmax] int32, sstate [kmax, 30] int32,
    lo [kmax], hi [kmax]) device tensors. Returns (best_val [1], best_idx [1]) device tensors (GLOBAL index)."""
    lib = load_library()
    m = model_struct(buf)
    sh = st = lo = hi = None
    if X is None:
        sh, st, lo, hi = sobol
    check(lib.pcabo_ei_sweep(ctypes.byref(m), int(run), int(ncand), int(idx0), _d(X, "X"), _i(sh, "shift"),
                             _i(st, "sstate"), _d(lo, "lo"), _d(hi, "hi"), _d(vals, "vals"),
                             _d(ws.part_val, "part_val"), _dev_ptr(ws.part_idx, torch.int64, "part_idx"), ws.nparts,
                             _d(ws.best_val, "best_val"), _dev_ptr(ws.best_idx, torch.int64, "best_idx"), _stream()),
          "pcabo_ei_sweep")
    return ws.best_val, ws.best_idx


def bbob_eval(fid, xopt, R, M, sign, fopt, lam, f0, X, out=None, aux=None):
    """Batched BBOB f15-f17 / f20-f24: X [B, m, D] (device) -> out [B, m] (f_opt included). aux [B, stride]: the
    Gallagher peak data of f21 / f22 (include/pcabo_b200.h)."""
    lib = load_library()
    B, m, D = X.shape
    if out is None:
        out = torch.empty(B, m, dtype=torch.float64, device=X.device)
    check(lib.pcabo_bbob_eval_aux(B, m, D, _i(fid, "fid"), _d(xopt, "xopt"), _d(R, "R"), _d(M, "M"), _d(sign, "sign"),
                                  _d(fopt, "fopt"), _d(lam, "lam"), float(f0), _d(aux, "aux"),
                                  int(aux.shape[1]) if aux is not None else 0, _d(X, "X"), _d(out, "out"), _stream()),
          "pcabo_bbob_eval_aux")
    return out


def select_initial_conditions(vals, mt_state, mt_pos, X, nsel, eta=1.0, out=None):
    """vals [B, nraw], X [B, nraw, kmax], mt_state [B, 624] int32 / mt_pos [B] int32 (advanced in place) ->
    dict(ics [B, nsel, kmax], idx [B, nsel] int32, flag [B] int32)."""
    lib = load_library()
    B, nraw, kmax = X.shape
    dev = X.device
    if out is None:
        out = dict(ics=torch.zeros(B, nsel, kmax, dtype=torch.float64, device=dev),
                   idx=torch.zeros(B, nsel, dtype=torch.int32, device=dev),
                   flag=torch.zeros(B, dtype=torch.int32, device=dev))
    check(lib.pcabo_select_initial_conditions(B, nraw, int(nsel), kmax, float(eta), _d(vals, "vals"),
                                              _i(mt_state, "mt_state"), _i(mt_pos, "mt_pos"), _d(X, "X"),
                                              _d(out["ics"], "ics"), _i(out["idx"], "idx"), _i(out["flag"], "flag"),
                                              _stream()), "pcabo_select_initial_conditions")
    return out
