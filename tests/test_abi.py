"""The C-ABI library loads on a CPU-only box, exports every symbol include/bmfr_b200.h declares, and
fails loudly (no CPU fallback) when asked to compute without a CUDA device."""
import ctypes as C
import re
from pathlib import Path

import pytest

from bmfr_b200 import BmfrError, Denoiser, _lib, block_offset

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "bmfr_b200.h").read_text()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(bmfr_[a-z0-9_]+)\s*\(", body)))


def test_header_symbols_are_exported(lib):
    names = declared_functions()
    assert len(names) >= 26
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bmfr_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert lib.bmfr_abi_version() == 5


def test_block_offsets_table():
    """BLOCK_OFFSETS, bmfr.cl:268-285 (integer, must be exact)."""
    expect = [(-14, -14), (4, -6), (-8, 14), (8, 0), (-10, -8), (2, 12), (12, -12), (-10, 0), (12, 14), (-8, -16),
              (6, 6), (-2, -2), (6, -14), (-16, 12), (14, -4), (-6, 4)]
    assert [block_offset(f) for f in range(16)] == expect
    assert [block_offset(f + 16) for f in range(16)] == expect


def test_default_params(lib):
    p = _lib.Params()
    lib.bmfr_default_params(C.byref(p), 1920, 1080)
    assert (p.width, p.height, p.tmp_half) == (1920, 1080, 0)
    assert p.noise_amount == 1e-2 and abs(p.blend_alpha - 0.2) < 1e-7          # bmfr.cpp:58,60
    assert abs(p.second_blend_alpha - 0.1) < 1e-7 and abs(p.taa_blend_alpha - 0.2) < 1e-7   # bmfr.cpp:61-62


def _has_gpu(lib):
    p = _lib.Params()
    lib.bmfr_default_params(C.byref(p), 64, 64)
    h = C.c_void_p()
    st = lib.bmfr_create(C.byref(p), C.byref(h))
    if st == 0:
        lib.bmfr_destroy(h)
    return st == 0


def test_no_cpu_fallback(lib):
    if _has_gpu(lib):
        pytest.skip("a CUDA device is present")
    with pytest.raises(BmfrError) as e:
        Denoiser(128, 64)
    assert e.value.status == -2 and "no CPU path" in str(e.value)


def test_invalid_arguments(lib):
    h = C.c_void_p()
    assert lib.bmfr_create(None, C.byref(h)) == -1
    p = _lib.Params()
    lib.bmfr_default_params(C.byref(p), 16, 16)      # smaller than one block: mirror() precondition, bmfr.cl:312-313
    assert lib.bmfr_create(C.byref(p), C.byref(h)) == -1
    assert b"32x32" in lib.bmfr_last_error()
    lib.bmfr_default_params(C.byref(p), 64, 64)
    p.tmp_half = 1
    assert lib.bmfr_create(C.byref(p), C.byref(h)) in (-5, -2)
    assert lib.bmfr_sync(None) == -1
    assert lib.bmfr_join(None) == -1
    # the defaults keep the reference's in-order queue semantics
    lib.bmfr_default_params(C.byref(p), 64, 64)
    assert p.overlap_frames == 0 and p.reference_order == 0 and p.tmp_half == 0


def test_unmirrorable_geometry_is_refused(lib):
    """mirror() only folds indices less than one image size out of bounds (bmfr.cl:207-208); sizes whose
    32-pixel block margin reaches further (e.g. 33x40: workset 64x64) make the reference read out of bounds
    and are refused here before any device is touched."""
    for w, h in ((33, 40), (40, 100), (100, 40)):
        p = _lib.Params()
        lib.bmfr_default_params(C.byref(p), w, h)
        hnd = C.c_void_p()
        assert lib.bmfr_create(C.byref(p), C.byref(hnd)) == -5, (w, h)
        assert b"mirrored" in lib.bmfr_last_error()
