"""ctypes binding of the C ABI in include/du_b200.h (libdu_b200.so, hand-written sm_100a CUDA).

There is NO fallback: if the shared library is missing the import of any op raises, loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdu_b200.so")

DU_MAX_M = 64
F32, F16, BF16 = 0, 1, 2

# du_status -> exception type (SURVEY.md §8b "Errors")
_EXC = {-1: ValueError, -2: RuntimeError, -3: RuntimeError, -4: RuntimeError, -5: RuntimeError, -6: RuntimeError}

# du_moments_mode
MOM_VAR_UNBIASED, MOM_CENTERED, MOM_VAR_WITH_CENTER, MOM_RAW, MOM_STD_UNBIASED, MOM_PARTIAL_M2 = range(6)
# du_znorm_mode
ZN_BELOW, ZN_ABOVE, ZN_MULTISCALE = range(3)
# du_prediction_type
PRED_EPSILON, PRED_SAMPLE, PRED_V = range(3)
# du_guidance
GUIDE_NONE, GUIDE_POSTERIOR, GUIDE_GRAD_BLEND, GUIDE_GRAD_ADD, GUIDE_WEIGHTS, GUIDE_LINCOMB, GUIDE_SIGN_ADD, GUIDE_MUL_BLEND = range(8)

i64, i32, f32, vp, sz = C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_size_t


class DdimCoeffs(C.Structure):
    _fields_ = [("sqrt_alpha_t", f32), ("sqrt_beta_t", f32), ("sqrt_alpha_prev", f32), ("dir_coef", f32),
                ("sigma", f32), ("clip_range", f32), ("prediction_type", i32), ("clip_sample", i32),
                ("use_clipped_model_output", i32), ("add_noise", i32)]


class GuidedParams(C.Structure):
    _fields_ = [
        ("eps", vp), ("eps_stride", i64), ("eps_dtype", i32), ("guidance", i32),
        ("sample", vp), ("sample_stride", i64), ("sample_dtype", i32), ("higher", i32),
        ("u", vp), ("u_stride", i64),
        ("thr", vp),
        ("mask", vp), ("mask_stride", i64),
        ("aux", vp), ("aux_stride", i64), ("aux_dtype", i32), ("aux_broadcast", i32),
        ("lam", f32), ("post_M", f32), ("inv_alpha_hat", f32), ("skip_ddim", i32),
        ("ddim", DdimCoeffs),
        ("B", i64), ("n", i64),
        ("prev_out", vp), ("prev_stride", i64), ("prev_dtype", i32), ("_pad0", i32),
        ("x0_out", vp), ("x0_stride", i64), ("x0_dtype", i32), ("x0_unguided", i32),
        ("eps_out", vp), ("eps_out_stride", i64), ("eps_out_dtype", i32), ("_pad2", i32),
        ("mask_out", vp), ("mask_out_stride", i64),
        ("mask_period", i64),
    ]


class FusedParams(C.Structure):
    _fields_ = [
        ("scores", vp * DU_MAX_M), ("M", i32), ("score_dtype", i32), ("score_stride", i64),
        ("eps", vp), ("eps_stride", i64),
        ("sample", vp), ("sample_stride", i64), ("sample_dtype", i32), ("moments_mode", i32),
        ("S", vp), ("S_stride", i64), ("S_broadcast", i32), ("higher", i32),
        ("q", f32), ("lerp_fma", i32), ("post_M", f32), ("inv_alpha_hat", f32),
        ("ddim", DdimCoeffs),
        ("B", i64), ("n", i64),
        ("unc_out", vp), ("unc_stride", i64),
        ("thr_out", vp),
        ("prev_out", vp), ("prev_stride", i64), ("prev_dtype", i32), ("S_overlap", i32),
        ("x0_out", vp), ("x0_stride", i64),
        ("eps_out", vp), ("eps_out_stride", i64),
        ("mask_out", vp), ("mask_out_stride", i64),
        ("skip_ddim", i32), ("_reserved0", i32),
    ]


# name -> (restype, argtypes); must list EVERY symbol include/du_b200.h declares (tests check this)
PROTOTYPES = {
    "du_last_error": (C.c_char_p, []),
    "du_version": (C.c_int, []),
    "du_num_sms": (C.c_int, [C.c_int]),
    "du_set_device": (C.c_int, [C.c_int]),
    "du_moments": (C.c_int, [C.POINTER(vp), C.c_int, i64, C.c_int, vp, i64, C.c_int, C.c_int, i64, i64,
                             vp, i64, C.c_int, vp, i64, vp]),
    "du_moments_merge": (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.c_int, i64, vp, vp, vp]),
    "du_quantile_scratch_bytes": (sz, [i64, i64]),
    "du_quantile_threshold": (C.c_int, [vp, i64, i64, i64, f32, C.c_int, vp, vp, vp, vp, sz, vp]),
    "du_threshold_mask": (C.c_int, [vp, i64, C.c_int, vp, C.c_int, i64, i64, vp, i64, vp]),
    "du_tensor_threshold_mask": (C.c_int, [vp, i64, C.c_int, vp, C.c_int, C.c_int, i64, i64, vp, i64, vp]),
    "du_znorm_scratch_bytes": (sz, [i64, i64]),
    "du_znorm_stats": (C.c_int, [vp, i64, C.c_int, i64, i64, vp, vp, sz, vp]),
    "du_znorm_stats_combine": (C.c_int, [vp, C.c_int, vp, vp]),
    "du_znorm_weights": (C.c_int, [vp, i64, C.c_int, vp, C.c_int, C.c_int, f32, i64, i64, vp, i64, vp, i64, vp]),
    "du_ddim_step": (C.c_int, [vp, i64, C.c_int, vp, i64, C.c_int, vp, i64, C.c_int, C.POINTER(DdimCoeffs), i64, i64,
                               vp, i64, C.c_int, vp, i64, C.c_int, vp, i64, C.c_int, vp]),
    "du_guided_step": (C.c_int, [C.POINTER(GuidedParams), vp]),
    "du_batch_sum": (C.c_int, [vp, i64, C.c_int, i64, i64, vp, vp]),
    "du_perturb": (C.c_int, [vp, i64, C.c_int, vp, i64, C.c_int, f32, f32, i64, i64, vp, i64, C.c_int, vp]),
    "du_perturb_rows": (C.c_int, [vp, i64, C.c_int, vp, i64, C.c_int, vp, vp, i64, i64, vp, i64, C.c_int, vp]),
    "du_ema_update": (C.c_int, [vp, vp, C.c_int, f32, f32, f32, i64, vp, vp, vp, vp]),
    "du_accumulate_slot": (C.c_int, [vp, i64, C.c_int, i64, i64, vp, i64, C.c_int, vp]),
    "du_dpm_solver_update": (C.c_int, [vp, i64, C.c_int, vp, i64, C.c_int, vp, i64, C.c_int, f32, f32, f32, f32, i64, i64, vp, i64,
                                       C.c_int, vp]),
    "du_perturb_randn": (C.c_int, [vp, C.c_int, i64, C.c_uint64, C.c_uint64, vp, f32, f32, vp, C.c_int, vp, C.c_int, vp]),
    "du_randn_offset_increment": (C.c_int, [i64, C.POINTER(C.c_uint64)]),
    "du_rng_advance": (C.c_int, [vp, C.c_uint64, vp]),
    "du_image_uint8": (C.c_int, [vp, i64, C.c_int, i64, i64, vp, i64, vp]),
    "du_fused_uncertainty_step": (C.c_int, [C.POINTER(FusedParams), vp]),
    "du_fused_supported": (C.c_int, [i64, C.c_int]),
    "du_fused_last_kernel": (C.c_int, []),
    "du_flip_h": (C.c_int, [vp, i64, C.c_int, i64, i64, i64, i64, vp, i64, C.c_int, vp]),
    "du_flip_sqdiff": (C.c_int, [vp, i64, C.c_int, vp, i64, C.c_int, i64, i64, i64, i64, C.c_int, vp, i64, vp]),
    "du_moments_backward": (C.c_int, [C.POINTER(vp), C.c_int, i64, C.c_int, vp, i64, C.c_int, C.c_int, vp, i64, C.c_int, i64, i64,
                                      C.POINTER(vp), i64, C.c_int, vp, i64, vp]),
    "du_column_kth": (C.c_int, [vp, C.c_int, i64, i64, i64, i64, vp, vp]),
    "du_row_sum": (C.c_int, [vp, i64, C.c_int, i64, i64, vp, vp]),
    "du_slot_sum": (C.c_int, [vp, i64, i64, C.c_int, i64, C.c_int, i64, vp, i64, vp]),
}

_lib = None
_lock = threading.Lock()


class DuError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libdu_b200.so (once).  Raises ImportError with build instructions if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension was not built.  Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or diffusion-uncertainty_b200/csrc/build.sh). "
                "There is no CPU or PyTorch fallback for the uncertainty path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int):
    if rc == 0:
        return
    msg = load().du_last_error().decode("utf-8", "replace")
    raise _EXC.get(rc, DuError)(f"[du_b200 {rc}] {msg}")
