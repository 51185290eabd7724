"""ctypes binding of ``libwft_b200.so`` (the C ABI declared in ``include/wft.h``).

This is the only place the package touches native code.  There is no fallback: if the shared library has
not been built (``python -c "import __graft_entry__ as g; g.build()"``) or no CUDA device is present, every
compute call raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libwft_b200.so")

WFT_PCM_F32 = 0
WFT_PCM_I16 = 1
WFT_WS_MEMSET, WFT_WS_PHASE_A, WFT_WS_PHASE_B, WFT_WS_RING = 0, 1, 2, 16
WFT_WS_PHASES = 16
WFT_LAUNCH_PDL, WFT_LAUNCH_OVERLAP = 1, 2
WFT_ERR_INVALID = -1
WFT_ERR_CUDA = -2
ABI_VERSION = 9


class FrontendArgs(Structure):
    """Mirror of ``struct wft_frontend_args`` (include/wft.h)."""

    _fields_ = [
        ("pcm", c_void_p),
        ("pcm_dtype", c_int32),
        ("batch", c_int32),
        ("clip_stride", c_int64),
        ("n_samples", c_int32),
        ("padding", c_int32),
        ("lengths", c_void_p),
        ("n_mels", c_int32),
        ("n_frames_out", c_int32),
        ("n_valid_frames", c_void_p),
        ("mask_params", c_void_p),
        ("mask_value", c_float),
        ("out", c_void_p),
        ("workspace", c_void_p),
        ("workspace_bytes", c_size_t),
        ("workspace_mode", c_int32),
        ("launch_flags", c_int32),
        ("draw_masks", c_int32),
        ("draw_time_mask_param", c_int32),
        ("draw_freq_mask_param", c_int32),
        ("draw_p", c_float),
        ("draw_seed", c_uint64),
        ("draw_clip_offset", c_uint64),
    ]


class AugmentArgs(Structure):
    """Mirror of ``struct wft_augment_args`` (include/wft.h)."""

    _fields_ = [
        ("out", c_void_p),
        ("warp_params", c_void_p),
        ("mask_params", c_void_p),
        ("extremes", c_void_p),
        ("mask_value", c_float),
        ("spline_f32", c_int32),
        ("draw", c_int32),
        ("draw_time_mask_param", c_int32),
        ("draw_freq_mask_param", c_int32),
        ("draw_time_warp_w", c_int32),
        ("draw_p", c_float),
        ("draw_seed", c_uint64),
        ("draw_clip_offset", c_uint64),
    ]


# name -> (restype, argtypes); tests check that the library exports exactly these (and the header declares them)
SIGNATURES = {
    "wft_abi_version": (c_int, []),
    "wft_last_error": (c_char_p, []),
    "wft_frontend_workspace_bytes": (c_int, [c_int32, c_int32, c_int32, POINTER(c_size_t)]),
    "wft_frontend_forward": (c_int, [POINTER(FrontendArgs), c_void_p]),
    "wft_frontend_augment_forward": (c_int, [POINTER(FrontendArgs), POINTER(AugmentArgs), c_void_p]),
    "wft_pad_or_trim_f32": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "wft_specaug_apply_f32": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_float, c_void_p]),
    "wft_specaug_draw": (c_int, [c_uint64, c_uint64, c_int32, c_int32, c_int32, c_int32, c_int32, c_float,
                                 c_void_p, c_void_p]),
    "wft_time_warp_f32": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "wft_time_warp_draw": (c_int, [c_uint64, c_uint64, c_int32, c_int32, c_int32, c_float, c_void_p, c_void_p]),
    "wft_augment_f32": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_float, c_int32,
                                c_void_p]),
    "wft_augment_drawn_f32": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_uint64, c_uint64, c_int32, c_int32, c_int32,
                                      c_float, c_void_p, c_float, c_int32, c_void_p]),
    "wft_mask_bsd": (c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                             c_uint32, c_void_p]),
    "wft_launch_count": (c_int64, [c_int]),
    "wft_frontend_grid": (c_int, [c_int32, c_int32, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "wft_debug_set_max_ctas": (c_int, [c_int32]),
    "wft_debug_set_augment_generic": (c_int, [c_int32]),
    "wft_debug_set_extra_smem": (c_int, [c_int32]),
}

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raise loudly if the CUDA library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: the sm_100a CUDA library is not built. "
            "Run `python -c \"import __graft_entry__ as g; g.build()\"` at the repo root. "
            "There is no CPU fallback for this front end."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.wft_abi_version() != ABI_VERSION:
        raise NativeLibraryMissing(f"{LIB_PATH}: ABI version {lib.wft_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int) -> None:
    """0 -> ok; WFT_ERR_INVALID -> ValueError; anything else -> RuntimeError (reference raises ValueError for bad
    parameters at data_loader.py:111-114; CUDA failures surface as RuntimeError like torch)."""
    if rc == 0:
        return
    msg = load().wft_last_error().decode("utf-8", "replace")
    if rc == WFT_ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)
