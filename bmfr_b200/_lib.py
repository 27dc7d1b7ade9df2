"""ctypes binding of libbmfr_b200.so (include/bmfr_b200.h).  There is no fallback: if the CUDA
library is missing this raises, and without a CUDA device bmfr_create reports BMFR_ERR_NO_DEVICE."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

PKG = Path(__file__).resolve().parent
# BMFR_B200_LIB selects another build of the same library (kernel-tuning experiments); it is still
# this package's CUDA library, never a fallback.
LIB_PATH = Path(os.environ.get("BMFR_B200_LIB") or PKG / "libbmfr_b200.so")

STATUS = {0: "BMFR_OK", -1: "BMFR_ERR_INVALID_ARGUMENT", -2: "BMFR_ERR_NO_DEVICE", -3: "BMFR_ERR_CUDA",
          -4: "BMFR_ERR_OUT_OF_MEMORY", -5: "BMFR_ERR_UNSUPPORTED", -6: "BMFR_ERR_HALO_TOO_SMALL",
          -7: "BMFR_ERR_SEQUENCE"}
MODE_STAGED, MODE_FUSED = 0, 1
BUF = dict(noisy_acc=0, spp=1, prev_pixels=2, accept=3, tmp_data=4, weights=5, mins_maxs=6, filtered=7,
           accum=8, tone_mapped=9, result=10, noise_tile=11)
STAGES = ("accum_noisy", "fitter", "weighted_sum", "accum_filtered", "taa", "total")


class BmfrError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status


class Params(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("device", C.c_int), ("mode", C.c_int),
                ("noise_amount", C.c_double), ("blend_alpha", C.c_float), ("second_blend_alpha", C.c_float),
                ("taa_blend_alpha", C.c_float), ("position_limit_squared", C.c_float),
                ("normal_limit_squared", C.c_float), ("tmp_half", C.c_int), ("profile", C.c_int),
                ("strip_y0", C.c_int), ("strip_y1", C.c_int), ("halo_rows", C.c_int), ("stream", C.c_void_p),
                ("reference_order", C.c_int), ("overlap_frames", C.c_int), ("fit_method", C.c_int),
                ("halo_timeout_ms", C.c_int), ("feature_set", C.c_int)]


class Geometry(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("width", "height", "workset_width", "workset_height", "margin_width",
                                        "margin_height", "blocks_x", "blocks_y", "row0", "row1", "own_y0", "own_y1",
                                        "block_row0", "block_row1")]


class HaloPlan(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("send_y0", "send_y1", "recv_y0", "recv_y1")]


# every symbol include/bmfr_b200.h declares: name -> (restype, argtypes)
_P, _I, _F = C.c_void_p, C.c_int, C.POINTER(C.c_float)
SYMBOLS = {
    "bmfr_last_error": (C.c_char_p, []),
    "bmfr_abi_version": (_I, []),
    "bmfr_default_params": (None, [C.POINTER(Params), _I, _I]),
    "bmfr_feature_counts": (_I, [_I, C.POINTER(_I), C.POINTER(_I)]),
    "bmfr_block_offset": (None, [_I, C.POINTER(_I), C.POINTER(_I)]),
    "bmfr_create": (_I, [C.POINTER(Params), C.POINTER(_P)]),
    "bmfr_destroy": (None, [_P]),
    "bmfr_get_geometry": (_I, [_P, C.POINTER(Geometry)]),
    "bmfr_denoise_frame": (_I, [_P, _I, _P, _P, _P, _P, _F, _F, _P]),
    "bmfr_denoise_frame_host": (_I, [_P, _I, _P, _P, _P, _P, _F, _F, _P]),
    "bmfr_sync": (_I, [_P]),
    "bmfr_join": (_I, [_P]),
    "bmfr_get_buffer": (_I, [_P, _I, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "bmfr_read_buffer": (_I, [_P, _I, _P, C.c_size_t]),
    "bmfr_get_stage_ms": (_I, [_P, _I, _F]),
    "bmfr_get_fused_kernel_busy_ms": (_I, [_P, _I, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "bmfr_get_fused_kernel_stamps": (_I, [_P, _I, C.POINTER(C.c_ulonglong)]),
    "bmfr_get_fused_kernel_ms": (_I, [_P, _I, _F]),
    "bmfr_kernel_launches": (C.c_longlong, [_P]),
    "bmfr_get_halo_plan": (_I, [_P, _I, C.POINTER(HaloPlan)]),
    "bmfr_halo_export": (_I, [_P, _P, C.c_size_t]),
    "bmfr_halo_connect": (_I, [_P, _I, _P, C.c_size_t]),
    "bmfr_halo_connect_local": (_I, [_P, _I, _P]),
    "bmfr_synth_camera": (None, [_I, _I, _I, _I, _F, _F]),
    "bmfr_synth_limits": (None, [_F, _F]),
    "bmfr_synth_frame_host": (_I, [_I, _I, _I, _I, _I, C.c_uint, _P, _P, _P, _P]),
    "bmfr_synth_frame_device": (_I, [_I, _I, _I, _I, _I, C.c_uint, _P, _P, _P, _P, _P]),
}

_lib = None


def load(build_if_missing: bool = True):
    """Loads the CUDA library (building it in-tree with nvcc when absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists() and build_if_missing:
        from . import build
        build.build_library()
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m bmfr_b200.build` "
                          "(nvcc, sm_100a). bmfr_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        raise BmfrError(status, load().bmfr_last_error().decode(errors="replace"))
