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


def _toggle_runs(variant, half):
    from oracle import build_oracle
    if not orc.available("reference", 0, variant) and build_oracle.build_reference(False, 0, variant) is None:
        pytest.skip("the shim build for this toggle is absent and /root/reference not mounted")
    a = util.run_oracle("reference", 160, 96, 6, keep=ALL, tmp_half=half)
    b = util.run_oracle("reference", 160, 96, 6, keep=ALL, tmp_half=half, variant=variant)
    return a, b


@pytest.mark.parametrize("half", [0, 1])
def test_compressed_r_toggle_does_not_change_any_buffer(half):
    """COMPRESSED_R 0 (bmfr.cpp:82): the reference's fitter keeps R as a full R_EDGE x R_EDGE array in local memory instead
    of the packed triangle (bmfr.cl:100-119,664-688; local size bmfr.cpp:355-361).  The reference's kernels compiled that
    way produce every buffer of every frame bit for bit: a scratch layout, not a behaviour — the compatibility mode
    (csrc/bmfr_reforder.cu) needs no second code path for it."""
    a, b = _toggle_runs("r0", half)
    for f, (x, y) in enumerate(zip(a, b)):
        for k in ALL:
            assert util.bits_equal(x[k], y[k]), f"frame {f}, tmp_half={half}: {k} differs with COMPRESSED_R=0"


def test_cache_tmp_data_toggle_is_a_different_filter():
    """CACHE_TMP_DATA 0 (bmfr.cpp:84) is NOT a tuning switch: without the private cache the transform loop re-reads the column
    and calls add_random() on it again — unconditionally, in every reflector pass and on the colour columns too
    (bmfr.cl:643-649), while the dot product in front of it saw the noise only in the first pass on the feature columns
    (bmfr.cl:623-627).  The reference's own kernels compiled that way give other weights from frame 0 on (everything in
    front of the fit is untouched); the reference ships with 1 and the compatibility mode reproduces that."""
    a, b = _toggle_runs("c0", 0)
    for k in ("noisy_acc", "spp", "prev_pixels", "accept", "mins_maxs", "noise_tile"):
        assert util.bits_equal(a[0][k], b[0][k]), f"{k} is computed in front of the fit and cannot depend on the toggle"
    assert not util.bits_equal(a[0]["weights"], b[0]["weights"])
    rel = np.abs(a[0]["result"] - b[0]["result"]).max() / max(np.abs(a[0]["result"]).max(), 1e-6)
    assert rel > 1e-4, rel  # visibly another result, not a rounding difference
