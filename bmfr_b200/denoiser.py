"""Host-side mirror of the reference's per-frame contract (the loop body of
/root/reference/opencl/bmfr.cpp:417-485) on top of the C ABI in include/bmfr_b200.h.

`Denoiser` plays the role of `tasks()`'s buffer set + frame loop: create it once, hand it the four
input images, the previous frame's camera matrix, this frame's pixel offset and the frame number,
get `result_buffer.current()` back.  All arithmetic happens in libbmfr_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BUF, MODE_FUSED, MODE_STAGED, STAGES, BmfrError, Geometry, HaloPlan, Params

FUSED_KERNELS = ("reproject_kernel", "fit_gram_kernel", "post_tma_kernel")  # the default FUSED path, in launch order


def fused_kernel_names(width, fit="gram"):
    """The three kernels a FUSED frame launches: the fit by bmfr_params.fit_method, the post pass by whether its tensor
    maps can be built for this width (csrc/bmfr_post.cu: W % 16 == 0)."""
    return ("reproject_kernel", "fit_gram_kernel" if fit == "gram" else "fit_qr_kernel",
            "post_tma_kernel" if width % 16 == 0 else "post_kernel")

_BUF_DTYPE = dict(noisy_acc=np.float32, spp=np.uint8, prev_pixels=np.float32, accept=np.uint8, tmp_data=np.float32,
                  weights=np.float32, mins_maxs=np.float32, filtered=np.float32, accum=np.float32,
                  tone_mapped=np.float32, result=np.float32, noise_tile=np.float64)


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def block_offset(frame: int):
    """BLOCK_OFFSETS[frame % 16] (bmfr.cl:267-285)."""
    x, y = C.c_int(), C.c_int()
    _lib.load().bmfr_block_offset(frame, C.byref(x), C.byref(y))
    return x.value, y.value


class Denoiser:
    def __init__(self, width, height, *, mode="fused", device=0, profile=False, stream=0, strip=None, halo_rows=0,
                 position_limit_squared=None, normal_limit_squared=None, noise_amount=None, blend_alpha=None,
                 second_blend_alpha=None, taa_blend_alpha=None, tmp_half=0, reference_order=0, overlap_frames=0, fit="gram",
                 halo_timeout_ms=0, feature_set=0):
        self.lib = _lib.load()
        p = Params()
        self.lib.bmfr_default_params(C.byref(p), width, height)
        p.device = device
        p.mode = {"staged": MODE_STAGED, "fused": MODE_FUSED}[mode]
        p.profile = int(profile)
        p.tmp_half, p.reference_order = int(tmp_half), int(reference_order)
        p.overlap_frames = int(overlap_frames)
        p.fit_method = {"gram": 0, "tsqr": 1}[fit]
        p.halo_timeout_ms = int(halo_timeout_ms)
        p.feature_set = int(feature_set)
        nf, ns = C.c_int(), C.c_int()
        _lib.check(self.lib.bmfr_feature_counts(p.feature_set, C.byref(nf), C.byref(ns)))
        self.features, self.features_scaled = nf.value, ns.value
        p.stream = C.c_void_p(stream or None)
        if strip is not None:
            p.strip_y0, p.strip_y1, p.halo_rows = int(strip[0]), int(strip[1]), int(halo_rows)
        for k, v in dict(position_limit_squared=position_limit_squared, normal_limit_squared=normal_limit_squared,
                         noise_amount=noise_amount, blend_alpha=blend_alpha, second_blend_alpha=second_blend_alpha,
                         taa_blend_alpha=taa_blend_alpha).items():
            if v is not None:
                setattr(p, k, v)
        self.params, self.mode = p, mode
        self.fused_kernels = fused_kernel_names(width, fit)
        self._h = C.c_void_p()
        _lib.check(self.lib.bmfr_create(C.byref(p), C.byref(self._h)))
        g = Geometry()
        _lib.check(self.lib.bmfr_get_geometry(self._h, C.byref(g)))
        self.geometry = g
        self.W, self.H = width, height
        self.rows = g.row1 - g.row0
        self._keep = []

    # -- lifetime -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.bmfr_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- the frame loop body --------------------------------------------------------------------
    def denoise_frame(self, frame, d_albedo, d_normal, d_position, d_noisy, cam_prev, pixel_offset, d_out=0):
        """Device-pointer entry (ints).  Asynchronous on the context's stream."""
        cam = np.ascontiguousarray(cam_prev if cam_prev is not None else np.zeros(16), dtype=np.float32)
        off = np.ascontiguousarray(pixel_offset, dtype=np.float32)
        _lib.check(self.lib.bmfr_denoise_frame(self._h, frame, C.c_void_p(d_albedo), C.c_void_p(d_normal),
                                               C.c_void_p(d_position), C.c_void_p(d_noisy), _fptr(cam), _fptr(off),
                                               C.c_void_p(d_out or None)))

    def denoise_frame_host(self, frame, albedo, normal, position, noisy, cam_prev, pixel_offset, out=None):
        """Host-array entry: uploads, five kernels' worth of work, read-back (bmfr.cpp:420-480).
        `out` (float32 [rows, W, 3]) is valid after sync()."""
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (albedo, normal, position, noisy)]
        n = self.rows * self.W * 3
        for a in arrs:
            if a.size != n:
                raise ValueError(f"input has {a.size} floats, context holds {self.rows} rows x {self.W} x 3 = {n}")
        cam = np.ascontiguousarray(cam_prev if cam_prev is not None else np.zeros(16), dtype=np.float32)
        off = np.ascontiguousarray(pixel_offset, dtype=np.float32)
        self._keep = (self._keep + [arrs, out])[-8:]  # keep sources alive while copies are in flight
        _lib.check(self.lib.bmfr_denoise_frame_host(self._h, frame, *[a.ctypes.data_as(C.c_void_p) for a in arrs],
                                                    _fptr(cam), _fptr(off),
                                                    out.ctypes.data_as(C.c_void_p) if out is not None else None))

    def sync(self):
        _lib.check(self.lib.bmfr_sync(self._h))

    def join(self):
        """Orders the context's stream after all submitted frames (contexts with overlap_frames; else a no-op)."""
        _lib.check(self.lib.bmfr_join(self._h))

    # -- inspection -----------------------------------------------------------------------------
    def buffer_ptr(self, name):
        p, n = C.c_void_p(), C.c_size_t()
        _lib.check(self.lib.bmfr_get_buffer(self._h, BUF[name], C.byref(p), C.byref(n)))
        return p.value, n.value

    def read(self, name):
        """One of the loop's buffers as of the last frame, in the reference's layout."""
        _, nbytes = self.buffer_ptr(name)
        a = np.empty(nbytes // np.dtype(_BUF_DTYPE[name]).itemsize, dtype=_BUF_DTYPE[name])
        _lib.check(self.lib.bmfr_read_buffer(self._h, BUF[name], a.ctypes.data_as(C.c_void_p), nbytes))
        if name in ("noisy_acc", "filtered", "accum", "tone_mapped", "result"):
            return a.reshape(self.rows, self.W, 3)
        if name == "prev_pixels":
            return a.reshape(self.rows, self.W, 2)
        if name in ("spp", "accept"):
            return a.reshape(self.rows, self.W)
        if name == "weights":
            return a.reshape(-1, self.features, 3)
        if name == "mins_maxs":
            return a.reshape(-1, self.features_scaled, 2)
        if name == "tmp_data":
            return a.reshape(-1, 13, 32, 32)
        if name == "noise_tile":
            return a.reshape(self.features - 1, 1024)
        return a

    def stage_ms(self, frame):
        ms = (C.c_float * 6)()
        _lib.check(self.lib.bmfr_get_stage_ms(self._h, frame, ms))
        return dict(zip(STAGES, ms))

    def fused_kernel_ms(self, frame):
        """Device time of reproject / fit_qr / post for one frame (FUSED, profile=True)."""
        ms = (C.c_float * 3)()
        _lib.check(self.lib.bmfr_get_fused_kernel_ms(self._h, frame, ms))
        return dict(zip(self.fused_kernels, ms))

    def fused_kernel_busy_ms(self, frame):
        """(per-kernel first-CTA-start .. last-CTA-end in ms, the same for the whole frame); FUSED, profile=2."""
        ms, total = (C.c_float * 3)(), C.c_float()
        _lib.check(self.lib.bmfr_get_fused_kernel_busy_ms(self._h, frame, ms, C.byref(total)))
        return dict(zip(self.fused_kernels, ms)), float(total.value)

    def fused_kernel_stamps(self, frame):
        """[(first CTA start, last CTA end)] of reproject / fit / post in ns of the device's globaltimer; FUSED, profile=2."""
        ns = (C.c_ulonglong * 6)()
        _lib.check(self.lib.bmfr_get_fused_kernel_stamps(self._h, frame, ns))
        return [(int(ns[2 * k]), int(ns[2 * k + 1])) for k in range(3)]

    @property
    def kernel_launches(self):
        return int(self.lib.bmfr_kernel_launches(self._h))

    # -- peer-to-peer halo exchange of sharded contexts -------------------------------------------
    def halo_export(self) -> bytes:
        """Opaque blob (CUDA IPC handles + geometry) a neighbouring rank passes to halo_connect()."""
        buf = C.create_string_buffer(1024)
        _lib.check(self.lib.bmfr_halo_export(self._h, buf, len(buf)))
        return buf.raw

    def halo_connect(self, side, blob: bytes):
        """side 0: the strip above, 1: below.  From then on every frame pushes / waits for halo rows itself."""
        buf = C.create_string_buffer(blob, len(blob))
        _lib.check(self.lib.bmfr_halo_connect(self._h, side, buf, len(blob)))

    def halo_connect_local(self, side, other: "Denoiser"):
        _lib.check(self.lib.bmfr_halo_connect_local(self._h, side, other._h))

    def halo_plan(self, side):
        hp = HaloPlan()
        _lib.check(self.lib.bmfr_get_halo_plan(self._h, side, C.byref(hp)))
        return hp
