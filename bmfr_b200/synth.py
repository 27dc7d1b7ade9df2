"""synth-v1: the deterministic synthetic sequence that stands in for the BMFR dataset
(/root/reference/opencl/bmfr.cpp:44-53 + camera_matrices.h).  Host and CUDA twins are bit-identical
(csrc/synth_core.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

SEED = 0x424D4652


def camera(frame: int, width: int, height: int, jitter: bool = False):
    """(camera_matrices[frame] as 16 floats, pixel_offsets[frame] as 2 floats)."""
    lib = _lib.load()
    m, o = (C.c_float * 16)(), (C.c_float * 2)()
    lib.bmfr_synth_camera(frame, width, height, int(jitter), m, o)
    return np.array(m, dtype=np.float32), np.array(o, dtype=np.float32)


def limits():
    """(position_limit_squared, normal_limit_squared) of the synthetic scene."""
    lib = _lib.load()
    a, b = C.c_float(), C.c_float()
    lib.bmfr_synth_limits(C.byref(a), C.byref(b))
    return float(a.value), float(b.value)


def frame_host(width: int, height: int, frame: int, y0: int = 0, y1: int | None = None, seed: int = SEED):
    """albedo, normal, position, noisy as float32 [rows, W, 3] for image rows [y0, y1)."""
    lib = _lib.load()
    y1 = height if y1 is None else y1
    out = [np.empty((y1 - y0, width, 3), dtype=np.float32) for _ in range(4)]
    _lib.check(lib.bmfr_synth_frame_host(width, height, y0, y1, frame, seed, *[a.ctypes.data_as(C.c_void_p) for a in out]))
    return out


def frame_device(width: int, height: int, frame: int, d_ptrs, y0: int = 0, y1: int | None = None, seed: int = SEED,
                 stream: int = 0):
    """Fills four device buffers (raw pointers) with rows [y0, y1) of the frame."""
    lib = _lib.load()
    y1 = height if y1 is None else y1
    _lib.check(lib.bmfr_synth_frame_device(width, height, y0, y1, frame, seed, *[C.c_void_p(p) for p in d_ptrs],
                                           C.c_void_p(stream)))
