"""TEST INFRASTRUCTURE — ctypes front end of the checkers (oracle/bmfr_oracle.h).

    Oracle("port")       oracle/libbmfr_oracle.so      plain-C restatement of bmfr.cl
    Oracle("reference")  oracle/_ref/libbmfr_clref.so  the reference's own bmfr.cl through the CL shim (CPU)
    Oracle("opencl")     oracle/_ref/libbmfr_clgpu.so  the reference's unmodified bmfr.cl on the box's OpenCL
                                                       device (the B200 through NVIDIA's ICD); needs a GPU box
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent

BUF = dict(noisy_acc=0, spp=1, prev_pixels=2, accept=3, tmp_data=4, weights=5, mins_maxs=6, filtered=7,
           accum=8, tone_mapped=9, result=10, noise_tile=11)
_DTYPE = dict(noisy_acc=np.float32, spp=np.uint8, prev_pixels=np.float32, accept=np.uint8, tmp_data=np.float32,
              weights=np.float32, mins_maxs=np.float32, filtered=np.float32, accum=np.float32,
              tone_mapped=np.float32, result=np.float32, noise_tile=np.float64)


class OracleParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("noise_amount", C.c_double), ("blend_alpha", C.c_float),
                ("second_blend_alpha", C.c_float), ("taa_blend_alpha", C.c_float),
                ("position_limit_squared", C.c_float), ("normal_limit_squared", C.c_float), ("tmp_half", C.c_int),
                ("keep_tmp", C.c_int), ("k1_schedule", C.c_int), ("threads", C.c_int)]


# feature lists the checkers are built for (include/bmfr_b200.h, bmfr_feature_set): (features, scaled features)
FEATURE_SETS = {0: (10, 6), 1: (7, 3), 2: (7, 6)}


def lib_path(kind: str, feature_set: int = 0, variant: str = "") -> Path:
    sfx = "" if feature_set == 0 else f"_fs{feature_set}"
    var = f"_{variant}" if variant else ""  # reference only: build_oracle.TOGGLE_VARIANTS
    return {"port": HERE / f"libbmfr_oracle{sfx}.so", "reference": HERE / "_ref" / f"libbmfr_clref{sfx}{var}.so",
            "opencl": HERE / "_ref" / f"libbmfr_clgpu{sfx}.so"}[kind]


def available(kind: str, feature_set: int = 0, variant: str = "") -> bool:
    return lib_path(kind, feature_set, variant).exists()


def _load(kind: str, feature_set: int = 0, variant: str = ""):
    path = lib_path(kind, feature_set, variant)
    if not path.exists():
        from . import build_oracle
        if variant:
            build_oracle.build_reference(False, feature_set, variant)
        else:
            dict(port=build_oracle.build_port, reference=build_oracle.build_reference, opencl=build_oracle.build_opencl_host)[kind](False, feature_set)
    if not path.exists():
        raise FileNotFoundError(f"{path} is missing (kind={kind})")
    lib = C.CDLL(str(path))
    lib.oracle_kind.restype = C.c_char_p
    lib.oracle_create.restype = C.c_void_p
    lib.oracle_create.argtypes = [C.POINTER(OracleParams)]
    lib.oracle_destroy.argtypes = [C.c_void_p]
    lib.oracle_frame.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
    lib.oracle_frame.restype = C.c_int
    lib.oracle_buffer.restype = C.c_void_p
    lib.oracle_buffer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
    lib.oracle_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.oracle_random.restype = C.c_float
    lib.oracle_random.argtypes = [C.c_uint]
    assert lib.oracle_kind().decode() == kind
    if kind == "opencl":
        lib.oracle_last_error.restype = C.c_char_p
        lib.oracle_device_name.restype = C.c_char_p
        lib.oracle_device_name.argtypes = [C.c_void_p]
    return lib


_LIBS = {}


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.c_void_p)


class Oracle:
    """The reference's host loop (bmfr.cpp:315-347,417-485) on CPU, one frame per call."""

    def __init__(self, kind, width, height, *, noise_amount=1e-2, blend_alpha=0.2, second_blend_alpha=0.1,
                 taa_blend_alpha=0.2, position_limit_squared, normal_limit_squared, tmp_half=0, keep_tmp=0,
                 k1_schedule=0, threads=0, feature_set=0, variant=""):
        if variant and kind != "reference":
            raise ValueError("variant: the reference's tuning toggles only exist in the shim build of its kernels")
        if (kind, feature_set, variant) not in _LIBS:
            _LIBS[(kind, feature_set, variant)] = _load(kind, feature_set, variant)
        self.lib, self.kind, self.W, self.H = _LIBS[(kind, feature_set, variant)], kind, width, height
        self.features, self.features_scaled = FEATURE_SETS[feature_set]
        self.params = OracleParams(width, height, noise_amount, blend_alpha, second_blend_alpha, taa_blend_alpha,
                                   position_limit_squared, normal_limit_squared, tmp_half, keep_tmp, k1_schedule,
                                   threads)
        self.h = self.lib.oracle_create(C.byref(self.params))
        if not self.h:
            why = self.lib.oracle_last_error().decode(errors="replace") if kind == "opencl" else ""
            raise RuntimeError(f"oracle_create({kind}) failed{': ' + why if why else ''}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.oracle_destroy(self.h)
            self.h = None

    __del__ = close

    def frame(self, frame, albedo, normals, positions, noisy, cam_prev, pixel_offset):
        keep = [_f32(a) for a in (albedo, normals, positions, noisy, cam_prev if cam_prev is not None else np.zeros(16), pixel_offset)]
        rc = self.lib.oracle_frame(self.h, frame, *[k[1] for k in keep])
        if rc != 0:
            raise RuntimeError(f"oracle_frame({frame}) failed: {rc}")

    def buffer(self, name):
        n = C.c_size_t()
        p = self.lib.oracle_buffer(self.h, BUF[name], C.byref(n))
        if not p:
            raise KeyError(f"oracle buffer {name} not available")
        raw = (C.c_char * n.value).from_address(p)
        a = np.frombuffer(raw, dtype=_DTYPE[name]).copy()
        if name in ("noisy_acc", "filtered", "accum", "tone_mapped", "result"):
            a = a.reshape(self.H, self.W, 3)
        elif name == "prev_pixels":
            a = a.reshape(self.H, self.W, 2)
        elif name in ("spp", "accept"):
            a = a.reshape(self.H, self.W)
        elif name == "weights":
            a = a.reshape(-1, self.features, 3)
        elif name == "mins_maxs":
            a = a.reshape(-1, self.features_scaled, 2)
        elif name == "tmp_data":
            a = a.reshape(-1, self.features + 3, 32, 32)
        elif name == "noise_tile":
            a = a.reshape(self.features - 1, 1024)
        return a

    @property
    def device_name(self):
        return self.lib.oracle_device_name(self.h).decode(errors="replace") if self.kind == "opencl" else "host CPU"

    def stage_ms(self):
        ms = (C.c_double * 6)()
        self.lib.oracle_stage_ms(self.h, ms)
        return list(ms)


def random_hash(kind, a: int) -> float:
    if (kind, 0) not in _LIBS:
        _LIBS[(kind, 0)] = _load(kind)
    return float(_LIBS[(kind, 0)].oracle_random(C.c_uint(a & 0xFFFFFFFF)))
