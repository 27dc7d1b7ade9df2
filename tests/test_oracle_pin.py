"""Pins the CPU restatement (oracle/bmfr_oracle.c) against the reference ITSELF: the reference's own
kernel file /root/reference/opencl/bmfr.cl, compiled as C++ through oracle/cl_shim and driven with
bmfr.cpp's buffer set / argument binding / launch geometry.  Every buffer of every frame must be
bit-identical.  (The reference repository holds no tests, golden vectors or fixtures of its own.)"""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import util

ALL = ("noisy_acc", "spp", "prev_pixels", "accept", "tmp_data", "weights", "mins_maxs", "filtered", "accum",
       "tone_mapped", "result", "noise_tile")


@pytest.fixture(scope="module", autouse=True)
def _need_reference():
    if not orc.available("reference"):
        from oracle import build_oracle
        if build_oracle.build_reference() is None:
            pytest.skip("oracle/_ref/libbmfr_clref.so absent and /root/reference not mounted")


def _compare(w, h, frames, **kw):
    a = util.run_oracle("port", w, h, frames, keep=ALL, **kw)
    b = util.run_oracle("reference", w, h, frames, keep=ALL, **kw)
    for f, (x, y) in enumerate(zip(a, b)):
        for k in ALL:
            assert util.bits_equal(x[k], y[k]), f"frame {f}: {k} differs between the port and the reference kernels"


@pytest.mark.parametrize("size", [(64, 32), (160, 96), (200, 120), (97, 75)])
def test_port_equals_reference_kernels(size):
    _compare(*size, 18)  # 18 frames: all 16 block offsets + wrap-around


def test_port_equals_reference_kernels_jitter():
    _compare(200, 120, 6, jitter=True)


@pytest.mark.parametrize("feature_set", [1, 2])
@pytest.mark.parametrize("half", [0, 1])
def test_port_equals_reference_kernels_other_feature_lists(feature_set, half):
    """The reference's feature list is a compile-time string (bmfr.cpp:63-77).  For the two other lists the CUDA library
    instantiates, the reference's bmfr.cl is compiled through the shim with those strings (oracle/build_oracle.py
    FEATURE_SETS) and the port with -DBMFR_FEATURE_SET: every buffer of every frame bit for bit again."""
    if not orc.available("reference", feature_set):
        from oracle import build_oracle
        if build_oracle.build_reference(False, feature_set) is None:
            pytest.skip("the shim build for this feature list is absent and /root/reference not mounted")
    _compare(200, 120, 17, feature_set=feature_set, tmp_half=half)


def test_port_equals_reference_kernels_half_tmp_data():
    """USE_HALF_PRECISION_IN_TMP_DATA=1, the reference's shipped default (bmfr.cpp:88)."""
    _compare(160, 96, 5, tmp_half=1)


def test_port_equals_reference_kernels_plain_schedule():
    """k1_schedule=1: plain row-major work-item order, where mirrored work-items may read the
    in-place store of their twin (bmfr.cl:322 vs :481)."""
    _compare(160, 96, 5, k1_schedule=1)


def test_k1_race_is_confined_to_margin_blocks():
    """The two legal K1 schedules differ only through mirrored work-items, i.e. in blocks that touch
    the image border; interior blocks' weights are identical."""
    w, h, frames = 200, 120, 4
    a = util.run_oracle("port", w, h, frames, keep=("weights", "accept", "spp", "noisy_acc"), every_frame=False)[0]
    b = util.run_oracle("port", w, h, frames, keep=("weights", "accept", "spp", "noisy_acc"), every_frame=False, k1_schedule=1)[0]
    for k in ("accept", "spp", "noisy_acc"):
        assert util.bits_equal(a[k], b[k])       # per-pixel outputs never depend on the schedule
    bx = (32 * ((w + 31) // 32) + 32) // 32
    by = (32 * ((h + 31) // 32) + 32) // 32
    wa, wb = a["weights"].reshape(by, bx, 30), b["weights"].reshape(by, bx, 30)
    assert np.array_equal(wa[2:-2, 2:-2], wb[2:-2, 2:-2])
