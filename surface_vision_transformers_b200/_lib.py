"""ctypes binding of libsvit_b200.so (the C ABI declared in include/svit_b200.h).

There is no CPU fallback: if the library is missing (and cannot be built) importing this module raises.
"""
import ctypes
import os

from . import build as _build

vp = ctypes.c_void_p
ci = ctypes.c_int
cf = ctypes.c_float
cll = ctypes.c_longlong
csz = ctypes.c_size_t


class SvitConfig(ctypes.Structure):
    _fields_ = [(n, ci) for n in ("dim", "depth", "heads", "dim_head", "mlp_dim", "num_patches", "num_vertices",
                                  "num_channels", "num_classes", "pool_mean")]


class AdamSegment(ctypes.Structure):
    _fields_ = [("offset", cll), ("numel", cll), ("bias_corr1", cf), ("bias_corr2", cf), ("active", ci), ("step", ci)]


ADAM_BLOCK_ELEMS = 4096
PROGRESS_FN = ctypes.CFUNCTYPE(None, ctypes.c_int, ctypes.c_void_p)

# name -> (restype, argtypes); mirrors include/svit_b200.h one to one
SIGNATURES = {
    "svit_last_error": (ctypes.c_char_p, []),
    "svit_version": (ci, []),
    "svit_launch_count": (ctypes.c_ulonglong, []),
    "svit_create": (vp, [ctypes.POINTER(SvitConfig)]),
    "svit_destroy": (None, [vp]),
    "svit_num_params": (ci, [vp]),
    "svit_param_offset": (cll, [vp, ci]),
    "svit_param_numel": (cll, [vp, ci]),
    "svit_flat_numel": (cll, [vp]),
    "svit_shadow_bytes": (csz, [vp]),
    "svit_mpp_shadow_bytes": (csz, [vp]),
    "svit_workspace_bytes": (csz, [vp, ci, ci, ci]),
    "svit_set_check_mode": (ci, [vp, ci]),
    "svit_get_check_mode": (ci, [vp]),
    "svit_set_dropout": (ci, [vp, cf, cf, ctypes.c_ulonglong, ctypes.c_ulonglong]),
    "svit_dropout_mask": (ci, [vp, csz, cf, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_uint, vp]),
    "svit_prepare_weights": (ci, [vp, vp, vp, vp]),
    "svit_mpp_prepare_weights": (ci, [vp, vp, vp, vp]),
    "svit_forward": (ci, [vp, vp, vp, vp, csz, vp, ci, vp, ci, vp, vp, vp, ci, vp]),
    "svit_forward_ex": (ci, [vp, vp, vp, vp, csz, vp, ci, ci, vp, ci, vp, vp, vp, ci, vp]),
    "svit_backward": (ci, [vp, vp, vp, vp, ci, vp, vp, vp, vp, vp]),
    "svit_encoder_forward": (ci, [vp, vp, vp, vp, csz, vp, ci, vp, ci, vp]),
    "svit_encoder_backward": (ci, [vp, vp, vp, vp, ci, vp, vp, vp, vp, vp]),
    "svit_mpp_forward": (ci, [vp, vp, vp, vp, vp, vp, vp, csz, vp, ci, vp, vp, vp, vp, vp, vp, ci, vp]),
    "svit_mpp_backward": (ci, [vp, vp, vp, vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "svit_gather_patches": (ci, [vp, vp, vp, ci, ci, ci, ci, ci, vp]),
    "svit_adamw_step": (ci, [vp, vp, vp, vp, vp, ci, vp, ci, cf, cf, cf, cf, cf, ci, cf, vp]),
    "svit_adamw_advance": (ci, [vp, ci, cf, cf, vp]),
    "svit_sgd_step": (ci, [vp, vp, vp, cll, cf, cf, cf, cf, ci, ci, cf, vp]),
    "svit_regression_loss": (ci, [vp, vp, ci, ci, vp, vp, vp]),
    "svit_gemm_tn": (ci, [vp] * 7 + [ci] * 10 + [vp]),
    "svit_gemm_ln": (ci, [vp] * 10 + [ci] * 5 + [cf, ci, vp]),
    "svit_gemm_wgrad": (ci, [vp, vp, vp] + [ci] * 7 + [vp]),
    "svit_gemm_wgrad_bias": (ci, [vp, vp, vp, vp] + [ci] * 7 + [vp]),
    "svit_attn_fwd": (ci, [vp, vp, vp, ci, ci, ci, cf, vp]),
    "svit_attn_cls_fwd": (ci, [vp, vp, vp, ci, ci, ci, cf, vp]),
    "svit_attn_cls_bwd": (ci, [vp, vp, vp, vp, ci, ci, ci, cf, vp]),
    "svit_attn_bwd": (ci, [vp, vp, vp, vp, vp, ci, ci, ci, cf, vp]),
    "svit_layernorm_fwd": (ci, [vp, vp, vp, vp, vp, vp, ci, ci, cf, vp]),
    "svit_layernorm_bwd": (ci, [vp] * 11 + [ci, ci, vp]),
}

_lib = None


def load():
    """Loads (building first if needed) the shared library; raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    # a library that is stale relative to csrc/ or include/ is rebuilt, never loaded silently
    # (SVIT_LIB=/path/to/other.so: developer override for A/B-ing a library built with other compile-time options)
    path = os.environ.get("SVIT_LIB") or _build.build(force=os.environ.get("SVIT_REBUILD") == "1")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SvitError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = load().svit_last_error()
        raise SvitError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr(t):
    return vp(t.data_ptr()) if t is not None else vp(0)
