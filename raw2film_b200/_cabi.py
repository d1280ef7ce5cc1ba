"""ctypes binding of libr2f_b200.so (include/r2f_b200.h).  No CPU fallback: if the CUDA
library is missing this module raises ImportError and every product entry point fails."""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libr2f_b200.so")

from .flags import BURN, GRAIN, GRAIN_BW, HALATION, MTF, SPATIAL, TAPS  # noqa: F401  (include/r2f_b200.h)

ABI_VERSION = 2

EXPORTS = [
    "r2f_abi_version", "r2f_last_error", "r2f_create", "r2f_destroy", "r2f_select_slot", "r2f_clear_slot",
    "r2f_set_lut2d", "r2f_set_curve1d",
    "r2f_set_lut3d", "r2f_set_halation_kernel", "r2f_set_mtf_kernel", "r2f_set_grain", "r2f_set_grain_seed", "r2f_set_option",
    "r2f_set_burn",
    "r2f_workspace_bytes", "r2f_render", "r2f_render_ex", "r2f_band_row", "r2f_render_banded", "r2f_render_tap", "r2f_render_tap_ex", "r2f_render_host", "r2f_convolve2d",
    "r2f_generate_noise", "r2f_chroma_nr", "r2f_histogram", "r2f_histogram_image", "r2f_calc_exposure", "r2f_canvas_paste", "r2f_resize", "r2f_present", "r2f_launch_count", "r2f_stream_mark", "r2f_fast_chain_stats", "r2f_profile_enable", "r2f_profile_read",
]
OPT_CONV_PATH = 1
OPT_CONV_SYM = 2
OPT_FUSE_MTF = 3
OPT_FAST_CHAIN = 4
MAX_SLOTS = 16
PIX_F32, PIX_U8 = 0, 2
INTER_AREA, INTER_LANCZOS4 = 0, 1
IN_F32, IN_U16 = 0, 1
PROF_NAMES = ["pointwise", "expose", "halation", "density", "mtf", "noise", "grain", "burn", "finish",
              "fft_rows_fwd", "fft_cols", "fft_rows_inv"]


class R2FError(RuntimeError):
    """Non-zero status from the C ABI (the reference raises plain exceptions)."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m raw2film_b200.build` "
            "(nvcc, sm_100a).  raw2film_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, fp, u8p = ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.c_void_p
    ci, cu, cf, cd, sz, u64 = (ctypes.c_int, ctypes.c_uint, ctypes.c_float, ctypes.c_double, ctypes.c_size_t,
                               ctypes.c_uint64)
    sig = {
        "r2f_abi_version": (ci, []),
        "r2f_last_error": (ctypes.c_char_p, []),
        "r2f_create": (ci, [ci, ctypes.POINTER(vp)]),
        "r2f_destroy": (ci, [vp]),
        "r2f_select_slot": (ci, [vp, ci]),
        "r2f_clear_slot": (ci, [vp, ci]),
        "r2f_set_lut2d": (ci, [vp, fp, ci]),
        "r2f_set_curve1d": (ci, [vp, fp, ci, cf]),
        "r2f_set_lut3d": (ci, [vp, fp, ci, cd]),
        "r2f_set_halation_kernel": (ci, [vp, fp, ci]),
        "r2f_set_mtf_kernel": (ci, [vp, fp, ci]),
        "r2f_set_grain": (ci, [vp, fp, ci, fp, ci, u64]),
        "r2f_set_grain_seed": (ci, [vp, u64]),
        "r2f_set_option": (ci, [vp, ci, ci]),
        "r2f_set_burn": (ci, [vp, cf, cf, cf]),
        "r2f_workspace_bytes": (sz, [ci, ci, cu]),
        "r2f_render": (ci, [vp, vp, ci, ci, ci, u8p, cu, vp, ci, vp, sz, vp]),
        "r2f_render_ex": (ci, [vp, vp, ci, cf, ci, ci, ci, u8p, cu, vp, ci, vp, sz, vp]),
        "r2f_band_row": (ci, [ci, ci, ci]),
        "r2f_render_banded": (ci, [vp, vp, ci, cf, ci, ci, ci, u8p, cu, vp, ci, vp, sz, ci, ctypes.POINTER(vp),
                                  ctypes.POINTER(vp), vp]),
        "r2f_render_tap": (ci, [vp, vp, ci, ci, ci, cu, vp, ci, vp, sz, ci, vp, vp]),
        "r2f_render_tap_ex": (ci, [vp, vp, ci, cf, ci, ci, ci, cu, vp, ci, vp, sz, ci, vp, vp]),
        "r2f_render_host": (ci, [vp, vp, ci, ci, ci, vp, cu, vp, ci]),
        "r2f_convolve2d": (ci, [vp, vp, vp, ci, ci, fp, ci, vp, sz, vp]),
        "r2f_generate_noise": (ci, [vp, vp, ci, ci, ci, u64, vp]),
        "r2f_chroma_nr": (ci, [vp, vp, ci, vp, ci, ci, fp, ci, vp, sz, vp]),
        "r2f_histogram": (ci, [vp, vp, ci, ci, vp, vp]),
        "r2f_histogram_image": (ci, [vp, vp, ci, ci, vp, ci, vp, vp]),
        "r2f_calc_exposure": (ci, [vp, vp, ci, ci, ci, ci, cd, ctypes.POINTER(cd), vp]),
        "r2f_canvas_paste": (ci, [vp, vp, ci, ci, vp, ci, ci, ci, ci, ci, ci, ci, vp]),
        "r2f_resize": (ci, [vp, vp, ci, ci, ci, ci, vp, ci, ci, ci, vp]),
        "r2f_present": (ci, [vp, vp, ci, ci, vp, ci, ci, fp, ci, ci, ci, vp]),
        "r2f_launch_count": (u64, [vp]),
        "r2f_stream_mark": (ci, [vp, vp]),
        "r2f_fast_chain_stats": (ci, [vp, ctypes.POINTER(u64), ctypes.POINTER(cf)]),
        "r2f_profile_enable": (ci, [vp, ci]),
        "r2f_profile_read": (ci, [vp, ctypes.POINTER(cd), ctypes.POINTER(u64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.r2f_abi_version() != ABI_VERSION:
        raise ImportError("libr2f_b200.so ABI version mismatch; rebuild with `python -m raw2film_b200.build`")
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        raise R2FError(f"libr2f_b200 error {rc}: {lib.r2f_last_error().decode()}")


def f32_ptr(arr):
    return arr.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
