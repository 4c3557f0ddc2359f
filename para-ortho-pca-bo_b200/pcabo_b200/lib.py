"""ctypes loader for libpcabo_b200.so and thin typed wrappers around its C ABI (include/pcabo_b200.h)."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EXPORTED_SYMBOLS = [
    "pcabo_abi_version", "pcabo_error_string", "pcabo_last_cuda_error",
    "pcabo_weights_pca_fit", "pcabo_inverse_map", "pcabo_gp_fit", "pcabo_gp_fit_jitter", "pcabo_gp_fit_hyper", "pcabo_gp_mll_grad_hyper", "pcabo_gp_pack_w",
    "pcabo_gp_packed_doubles", "pcabo_gp_mll", "pcabo_gp_mll_grad",
    "pcabo_logei", "pcabo_argmax", "pcabo_lbfgsb_logei", "pcabo_lbfgsb_logei_prof", "pcabo_lbfgsb_logei_runs", "pcabo_lbfgsb_logei_runs_ws",
    "pcabo_lbfgsb_workspace_doubles", "pcabo_lbfgsb_plan", "pcabo_ei_sweep", "pcabo_logei_raw_runs", "pcabo_logei_raw_supported", "pcabo_pick_best", "pcabo_append_observation",
    "pcabo_sobol_points", "pcabo_sobol_scramble_bits", "pcabo_gather_rows", "pcabo_bbob_eval", "pcabo_bbob_eval_aux", "pcabo_select_initial_conditions",
]


class PcaboError(RuntimeError):
    pass


class GpModelStruct(ctypes.Structure):
    """struct pcabo_gp_model (include/pcabo_b200.h)."""
    _fields_ = [
        ("B", ctypes.c_int), ("n", ctypes.c_int), ("kmax", ctypes.c_int), ("ld_n", ctypes.c_int),
        ("maximize", ctypes.c_int), ("acq", ctypes.c_int),
        ("lengthscale", ctypes.c_double),
        ("k", ctypes.c_void_p), ("norm_off", ctypes.c_void_p), ("norm_coef", ctypes.c_void_p),
        ("ZnT", ctypes.c_void_p), ("alpha", ctypes.c_void_p), ("W", ctypes.c_void_p),
        ("ymean", ctypes.c_void_p), ("ystd", ctypes.c_void_p), ("best_f", ctypes.c_void_p),
        ("WT", ctypes.c_void_p), ("Zn", ctypes.c_void_p),
        ("Wp", ctypes.c_void_p), ("wp_stride", ctypes.c_longlong),
        ("lengthscales", ctypes.c_void_p),
    ]


def library_path() -> str:
    return os.path.join(HERE, "libpcabo_b200.so")


def load_library(build_if_missing: bool = False):
    """Load the shared library; raises PcaboError (never falls back) when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        if build_if_missing:
            from .build import build
            build()
        else:
            raise PcaboError(f"{path} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                             "There is no CPU fallback for the PCA-BO hot path.")
    lib = ctypes.CDLL(path)
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    lib.pcabo_abi_version.restype = ci
    lib.pcabo_error_string.restype = ctypes.c_char_p
    lib.pcabo_error_string.argtypes = [ci]
    lib.pcabo_last_cuda_error.restype = ctypes.c_char_p
    lib.pcabo_weights_pca_fit.restype = ci
    lib.pcabo_weights_pca_fit.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp, cd, ci, ci] + [vp] * 9 + [vp]
    lib.pcabo_inverse_map.restype = ci
    lib.pcabo_inverse_map.argtypes = [ci, ci] + [vp] * 9 + [vp]
    lib.pcabo_gp_fit.restype = ci
    lib.pcabo_gp_fit.argtypes = [ci, ci, ci, ci, ci, vp, vp, vp, vp, vp, cd, cd] + [vp] * 13 + [vp]
    lib.pcabo_gp_fit_jitter.restype = ci
    lib.pcabo_gp_fit_jitter.argtypes = [ci, ci, ci, ci, ci, vp, vp, vp, vp, vp, cd, cd, vp] + [vp] * 13 + [vp]
    lib.pcabo_gp_fit_hyper.restype = ci
    lib.pcabo_gp_fit_hyper.argtypes = [ci, ci, ci, ci, ci, vp, vp, vp, vp, vp, cd, cd, vp, vp, vp] + [vp] * 13 + [vp]
    lib.pcabo_gp_mll_grad_hyper.restype = ci
    lib.pcabo_gp_mll_grad_hyper.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp, cd, vp, vp, vp]
    lib.pcabo_gp_pack_w.restype = ci
    lib.pcabo_gp_pack_w.argtypes = [ci, ci, ci, vp, vp, ctypes.c_longlong, vp]
    lib.pcabo_gp_packed_doubles.restype = ctypes.c_longlong
    lib.pcabo_gp_packed_doubles.argtypes = [ci]
    lib.pcabo_gp_mll.restype = ci
    lib.pcabo_gp_mll.argtypes = [ci, ci, ci] + [vp] * 6 + [vp]
    lib.pcabo_gp_mll_grad.restype = ci
    lib.pcabo_gp_mll_grad.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp, cd, vp, vp]
    lib.pcabo_logei.restype = ci
    lib.pcabo_logei.argtypes = [ctypes.POINTER(GpModelStruct), ci, vp, vp, vp, vp, vp, vp]
    lib.pcabo_lbfgsb_logei.restype = ci
    lib.pcabo_lbfgsb_logei.argtypes = [ctypes.POINTER(GpModelStruct), ci, ci, ci, cd, cd] + [vp] * 8 + [vp]
    lib.pcabo_lbfgsb_logei_prof.restype = ci
    lib.pcabo_lbfgsb_logei_prof.argtypes = [ctypes.POINTER(GpModelStruct), ci, ci, ci, cd, cd] + [vp] * 9 + [vp]
    lib.pcabo_lbfgsb_logei_runs.restype = ci
    lib.pcabo_lbfgsb_logei_runs.argtypes = [ctypes.POINTER(GpModelStruct), ci, vp, ci, ci, ci, ci, cd, cd] + [vp] * 9 + [vp]
    lib.pcabo_lbfgsb_logei_runs_ws.restype = ci
    lib.pcabo_lbfgsb_logei_runs_ws.argtypes = [ctypes.POINTER(GpModelStruct), ci, vp, ci, ci, ci, ci, cd, cd] + [vp] * 9 + \
                                              [vp, ctypes.c_longlong, vp]
    lib.pcabo_lbfgsb_workspace_doubles.restype = ctypes.c_longlong
    lib.pcabo_lbfgsb_workspace_doubles.argtypes = [ci] * 6
    lib.pcabo_lbfgsb_plan.restype = ci
    lib.pcabo_lbfgsb_plan.argtypes = [ci] * 5 + [ctypes.POINTER(ci)]
    lib.pcabo_ei_sweep.restype = ci
    lib.pcabo_ei_sweep.argtypes = [ctypes.POINTER(GpModelStruct), ci, ctypes.c_longlong, ctypes.c_longlong] + [vp] * 8 + \
                                  [ci, vp, vp, vp]
    lib.pcabo_logei_raw_supported.restype = ci
    lib.pcabo_logei_raw_supported.argtypes = [ci, ci, ci, ci]
    lib.pcabo_logei_raw_runs.restype = ci
    lib.pcabo_logei_raw_runs.argtypes = [ctypes.POINTER(GpModelStruct), ci, vp, vp, ci, vp]
    lib.pcabo_pick_best.restype = ci
    lib.pcabo_pick_best.argtypes = [ci, ci, ci, vp, vp, vp, vp, vp, vp]
    lib.pcabo_append_observation.restype = ci
    lib.pcabo_append_observation.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp, vp, cd, ci, vp, vp, vp, vp, vp, vp]
    lib.pcabo_sobol_points.restype = ci
    lib.pcabo_sobol_points.argtypes = [ci, ci, ci] + [vp] * 7 + [vp]
    lib.pcabo_sobol_scramble_bits.restype = ci
    lib.pcabo_sobol_scramble_bits.argtypes = [ci, ci, vp, vp, vp, vp, vp, vp]
    lib.pcabo_gather_rows.restype = ci
    lib.pcabo_gather_rows.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp]
    lib.pcabo_select_initial_conditions.restype = ci
    lib.pcabo_select_initial_conditions.argtypes = [ci, ci, ci, ci, cd] + [vp] * 7 + [vp]
    lib.pcabo_bbob_eval.restype = ci
    lib.pcabo_bbob_eval.argtypes = [ci, ci, ci] + [vp] * 7 + [cd, vp, vp, vp]
    lib.pcabo_bbob_eval_aux.restype = ci
    lib.pcabo_bbob_eval_aux.argtypes = [ci, ci, ci] + [vp] * 7 + [cd, vp, ci, vp, vp, vp]
    lib.pcabo_argmax.restype = ci
    lib.pcabo_argmax.argtypes = [ci, ci, vp, vp, vp, vp]
    _LIB = lib
    return lib


def check(code: int, what: str):
    if code != 0:
        lib = load_library()
        msg = lib.pcabo_error_string(code).decode()
        if code == -3:
            msg += ": " + lib.pcabo_last_cuda_error().decode()
        raise PcaboError(f"{what} failed: {msg} (code {code})")
