"""Dataset ingestion (include/bmfr_io.h): the EXR reader, the camera_matrices.h parser and the
PNG writer that stand in for OpenImageIO and the compiled-in dataset header of the reference
(bmfr.cpp:46-47, 145-165, 520-539).  The EXR files are produced here by an independent writer (numpy +
zlib, following the OpenEXR file layout), so the test does not depend on the code under test."""
import ctypes as C

import numpy as np
import pytest

from bmfr_b200 import build

from .exr_util import FLOAT, HALF, NONE, RLE, ZIP, ZIPS, fixture_image, write_exr

OK, ERR_OPEN, ERR_FORMAT, ERR_UNSUPPORTED, ERR_MISMATCH = 0, -2, -3, -4, -5


@pytest.fixture(scope="module")
def io():
    lib = C.CDLL(str(build.build_io()))
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib.bmfr_io_exr_info.argtypes = [C.c_char_p, ip, ip, ip]
    lib.bmfr_io_read_exr_rgb.argtypes = [C.c_char_p, C.c_int, C.c_int, fp]
    lib.bmfr_io_parse_camera_header.argtypes = [C.c_char_p, C.c_int, fp, fp, ip, ip, fp, fp]
    lib.bmfr_io_write_png_rgb.argtypes = [C.c_char_p, C.c_int, C.c_int, fp, C.c_size_t]
    lib.bmfr_io_last_error.restype = C.c_char_p
    lib.bmfr_io_psnr.argtypes = [fp, fp, C.c_size_t, C.c_float, C.POINTER(C.c_double)]
    lib.bmfr_io_ssim_rgb.argtypes = [fp, fp, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_double)]
    lib.bmfr_io_tone_map.argtypes = [fp, C.c_size_t]
    lib.bmfr_io_tone_map.restype = None
    return lib


def _image(h, w, seed):
    rng = np.random.default_rng(seed)
    img = rng.normal(0, 3, (h, w, 3)).astype(np.float32)
    img[: h // 2] = np.linspace(0, 4, w, dtype=np.float32)[None, :, None]  # compressible rows; the random ones end up stored raw
    img[0, 0] = [0.0, -0.0, 65504.0]
    img[1, 1] = [np.inf, 6e-8, 1e-3]  # a subnormal half, a small normal one
    return img


def _read(io, path, w, h):
    out = np.full((h, w, 3), np.nan, dtype=np.float32)
    st = io.bmfr_io_read_exr_rgb(str(path).encode(), w, h, out.ctypes.data_as(C.POINTER(C.c_float)))
    return st, out


@pytest.mark.parametrize("compression", [NONE, RLE, ZIPS, ZIP])
@pytest.mark.parametrize("pixel_type", [HALF, FLOAT])
def test_exr_round_trip(io, tmp_path, compression, pixel_type):
    w, h = 37, 41  # odd sizes: a ragged last ZIP chunk (41 = 2 * 16 + 9)
    img = _image(h, w, 7 * compression + pixel_type)
    path = tmp_path / "color0.exr"
    write_exr(path, {"R": img[..., 0], "G": img[..., 1], "B": img[..., 2]}, dict(R=pixel_type, G=pixel_type, B=pixel_type), compression)
    wi, hi, ci = C.c_int(), C.c_int(), C.c_int()
    assert io.bmfr_io_exr_info(str(path).encode(), C.byref(wi), C.byref(hi), C.byref(ci)) == OK
    assert (wi.value, hi.value, ci.value) == (w, h, 3)
    st, out = _read(io, path, w, h)
    assert st == OK, io.bmfr_io_last_error()
    with np.errstate(over="ignore"):
        want = img.astype(np.float16).astype(np.float32) if pixel_type == HALF else img
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))  # bit for bit, -0 and inf included


@pytest.mark.parametrize("name", ["none_f32", "rle_f16", "zips_f16", "zip_f32", "piz_f32", "piz_f16", "pxr24_f32",
                                  "piz_f32_tall", "zip_f16_tall"])
def test_exr_files_written_by_the_openexr_library(io, name):
    """tests/golden/exr/*.exr were written by OpenCV's bundled OpenEXR (tests/golden/make_exr_fixtures.py)."""
    from pathlib import Path

    golden = Path(__file__).resolve().parent / "golden"
    tall = name.endswith("_tall")
    img = fixture_image(tall=tall)
    h, w = img.shape[:2]
    st, out = _read(io, golden / "exr" / f"{name}.exr", w, h)
    if name.startswith("pxr24"):
        assert st == ERR_UNSUPPORTED and b"not covered" in io.bmfr_io_last_error()
        return
    assert st == OK, io.bmfr_io_last_error()
    with np.errstate(over="ignore"):
        want = img.astype(np.float16).astype(np.float32) if "f16" in name else img
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


def test_exr_channel_order_window_and_line_order(io, tmp_path):
    w, h = 20, 35
    img = _image(h, w, 3)
    # alphabetical storage is B, G, R: the reader must hand back R, G, B (as OpenImageIO does);
    # mixed channel types, a data window that does not start at 0, chunks stored bottom-up
    path = tmp_path / "a.exr"
    write_exr(path, {"R": img[..., 0], "G": img[..., 1], "B": img[..., 2]}, dict(R=FLOAT, G=HALF, B=FLOAT), ZIP, xmin=-5, ymin=12,
              decreasing=True)
    st, out = _read(io, path, w, h)
    assert st == OK, io.bmfr_io_last_error()
    want = img.copy()
    with np.errstate(over="ignore"):
        want[..., 1] = img[..., 1].astype(np.float16).astype(np.float32)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))
    # layer-prefixed names and X / Y / Z (world positions)
    path = tmp_path / "b.exr"
    write_exr(path, {"pos.Z": img[..., 2], "pos.X": img[..., 0], "pos.Y": img[..., 1]}, {"pos.X": FLOAT, "pos.Y": FLOAT, "pos.Z": FLOAT}, ZIPS)
    st, out = _read(io, path, w, h)
    assert st == OK and np.array_equal(out.view(np.uint32), img.view(np.uint32))
    # unknown names: file (alphabetical) order
    path = tmp_path / "c.exr"
    write_exr(path, {"u": img[..., 0], "v": img[..., 1], "w": img[..., 2]}, dict(u=FLOAT, v=FLOAT, w=FLOAT), NONE)
    st, out = _read(io, path, w, h)
    assert st == OK and np.array_equal(out.view(np.uint32), img.view(np.uint32))


def test_exr_errors_follow_the_reference(io, tmp_path):
    w, h = 16, 16
    img = _image(h, w, 5)
    rgb = {"R": img[..., 0], "G": img[..., 1], "B": img[..., 2]}
    f3 = dict(R=FLOAT, G=FLOAT, B=FLOAT)
    assert _read(io, tmp_path / "missing.exr", w, h)[0] == ERR_OPEN
    good = tmp_path / "good.exr"
    write_exr(good, rgb, f3, ZIP)
    # bmfr.cpp:150-155: wrong width / height / channel count is "wrong type"
    assert _read(io, good, w + 1, h)[0] == ERR_MISMATCH
    assert b"wrong type" in io.bmfr_io_last_error()
    rgba = tmp_path / "rgba.exr"
    write_exr(rgba, dict(rgb, A=img[..., 0]), dict(f3, A=FLOAT), ZIP)
    assert _read(io, rgba, w, h)[0] == ERR_MISMATCH
    piz = tmp_path / "piz.exr"
    write_exr(piz, rgb, f3, NONE)
    data = bytearray(piz.read_bytes())
    at = data.index(b"compression\0compression\0") + len(b"compression\0compression\0") + 4
    data[at] = 6  # B44
    piz.write_bytes(bytes(data))
    assert _read(io, piz, w, h)[0] == ERR_UNSUPPORTED
    tiled = tmp_path / "tiled.exr"
    write_exr(tiled, rgb, f3, NONE, flags=0x200)
    assert _read(io, tiled, w, h)[0] == ERR_UNSUPPORTED
    blob = good.read_bytes()
    for cut in (3, 40, len(blob) // 2, len(blob) - 5):
        t = tmp_path / f"cut{cut}.exr"
        t.write_bytes(blob[:cut])
        assert _read(io, t, w, h)[0] == ERR_FORMAT, cut
    # a data window whose corners overflow 32-bit arithmetic (found by scripts/fuzz_io.cpp)
    blob_w = bytearray(blob)
    at = blob_w.index(b"dataWindow\0box2i\0") + len(b"dataWindow\0box2i\0") + 4
    blob_w[at:at + 4] = (-2**31).to_bytes(4, "little", signed=True)
    t = tmp_path / "window.exr"
    t.write_bytes(bytes(blob_w))
    assert _read(io, t, w, h)[0] == ERR_FORMAT
    junk = tmp_path / "junk.exr"
    junk.write_bytes(b"not an exr file at all")
    assert _read(io, junk, w, h)[0] == ERR_FORMAT
    # a damaged zlib stream (a smooth image, so that the chunk really is stored compressed)
    smooth = np.tile(np.linspace(0, 1, w, dtype=np.float32), (h, 1))
    z = tmp_path / "smooth.exr"
    write_exr(z, dict(R=smooth, G=smooth, B=smooth), f3, ZIP)
    assert _read(io, z, w, h)[0] == OK
    corrupt = bytearray(z.read_bytes())
    assert len(corrupt) < len(blob)
    corrupt[-20] ^= 0xFF  # inside the only ZIP chunk
    t = tmp_path / "corrupt.exr"
    t.write_bytes(bytes(corrupt))
    assert _read(io, t, w, h)[0] == ERR_FORMAT


HEADER = """\
// camera_matrices.h -- written by the dataset exporter
#pragma once
/* limits used by the reprojection test
   (squared distances) */
const float position_limit_squared = 0.001600f;
const float normal_limit_squared = 2.5e-1;
#define FRAMES 3
static const float camera_matrices[FRAMES][4][4] = {
    { // frame 0
        {1.0f, 0.f, -0.0f, 0},
        {0, 1.5, 0, 0},
        {0, 0, -1.0002f, -1.f},
        {0.25, -.5, 3e+2f, 1e-3}
    },
    { {2,0,0,0}, {0,2,0,0}, {0,0,2,0}, {0,0,0,2} },
    { {+3.5,0,0,0}, {0,3,0,0}, {0,0,3,0}, {1,2,3,4} },
};
const float pixel_offsets[FRAMES][2] = { {0.5f, 0.5f}, {0.25, 0.75}, {-0.125f, .875f} };
"""


def test_camera_header(io, tmp_path):
    path = tmp_path / "camera_matrices.h"
    path.write_text(HEADER)
    m = np.zeros((4, 16), dtype=np.float32)
    o = np.zeros((4, 2), dtype=np.float32)
    nm, no = C.c_int(), C.c_int()
    pl, nl = C.c_float(-1), C.c_float(-1)
    fp = C.POINTER(C.c_float)
    st = io.bmfr_io_parse_camera_header(str(path).encode(), 4, m.ctypes.data_as(fp), o.ctypes.data_as(fp), C.byref(nm), C.byref(no),
                                        C.byref(pl), C.byref(nl))
    assert st == OK, io.bmfr_io_last_error()
    assert (nm.value, no.value) == (3, 3)
    assert pl.value == np.float32(0.0016) and nl.value == np.float32(0.25)
    assert m[0].tolist() == [np.float32(v) for v in (1, 0, -0.0, 0, 0, 1.5, 0, 0, 0, 0, -1.0002, -1, 0.25, -0.5, 300, 1e-3)]
    assert m[1].reshape(4, 4).tolist() == (2 * np.eye(4)).tolist()
    assert m[2, 12:].tolist() == [1, 2, 3, 4] and m[2, 0] == 3.5
    assert not m[3].any()
    assert o[:3].tolist() == [[0.5, 0.5], [0.25, 0.75], [-0.125, 0.875]]
    # fewer slots than frames: only the first ones are stored, the counts still say what the file holds
    m2 = np.zeros((1, 16), dtype=np.float32)
    o2 = np.zeros((1, 2), dtype=np.float32)
    assert io.bmfr_io_parse_camera_header(str(path).encode(), 1, m2.ctypes.data_as(fp), o2.ctypes.data_as(fp), C.byref(nm), C.byref(no),
                                          None, None) == OK
    assert nm.value == 3 and m2[0, 5] == 1.5 and o2[0].tolist() == [0.5, 0.5]
    # a header without the arrays
    bad = tmp_path / "bad.h"
    bad.write_text("const float position_limit_squared = 1.0f;\nextern const float camera_matrices[60][4][4];\n")
    assert io.bmfr_io_parse_camera_header(str(bad).encode(), 4, m.ctypes.data_as(fp), o.ctypes.data_as(fp), None, None, None,
                                          None) == ERR_FORMAT
    assert io.bmfr_io_parse_camera_header(str(tmp_path / "nope.h").encode(), 4, m.ctypes.data_as(fp), o.ctypes.data_as(fp), None, None,
                                          None, None) == ERR_OPEN


def test_png_writer(io, tmp_path):
    from PIL import Image

    w, h, stride = 19, 11, 24 * 3  # rows of a wider buffer (the reference passes WORKSET_WIDTH * 3, bmfr.cpp:536-537)
    rng = np.random.default_rng(11)
    buf = rng.uniform(-0.2, 1.2, (h, stride)).astype(np.float32)
    buf[0, 0], buf[0, 1], buf[0, 2] = np.nan, np.inf, -np.inf
    buf[1, 0], buf[1, 1], buf[1, 2] = 0.5, 1.0, 0.0
    path = tmp_path / "output0.png"
    st = io.bmfr_io_write_png_rgb(str(path).encode(), w, h, buf.ctypes.data_as(C.POINTER(C.c_float)), stride)
    assert st == OK, io.bmfr_io_last_error()
    got = np.asarray(Image.open(path))
    assert got.shape == (h, w, 3) and got.dtype == np.uint8
    crop = buf[:, : w * 3].reshape(h, w, 3)
    want = np.floor(np.clip(np.nan_to_num(crop, nan=0.0, posinf=1.0, neginf=0.0), 0, 1) * np.float32(255) + np.float32(0.5)).astype(np.uint8)
    assert np.array_equal(got, want)
    assert got[0, 0].tolist() == [0, 255, 0] and got[1, 0].tolist() == [128, 255, 0]
    assert io.bmfr_io_write_png_rgb(str(tmp_path / "no_such_dir" / "x.png").encode(), w, h, buf.ctypes.data_as(C.POINTER(C.c_float)),
                                    stride) == ERR_OPEN


def test_io_library_exports_every_declared_symbol(io):
    import re
    from pathlib import Path

    header = (Path(__file__).resolve().parent.parent / "include" / "bmfr_io.h").read_text()
    names = sorted(set(re.findall(r"\b(bmfr_io_\w+)\s*\(", header)))
    assert len(names) == 8, names
    for n in names:
        assert hasattr(io, n), n


def _ssim_reference(a, b, peak):
    """Wang et al. 2004 as usually implemented (11x11 Gaussian, sigma 1.5, valid windows), in float64 with scipy."""
    from scipy.signal import convolve2d

    g = np.exp(-((np.arange(11) - 5.0) ** 2) / (2 * 1.5 ** 2))
    g /= g.sum()
    win = np.outer(g, g)
    c1, c2 = (0.01 * peak) ** 2, (0.03 * peak) ** 2
    vals = []
    for ch in range(3):
        x, y = a[..., ch].astype(np.float64), b[..., ch].astype(np.float64)
        f = lambda im: convolve2d(im, win, mode="valid")  # noqa: E731 (the window is symmetric)
        mx, my = f(x), f(y)
        vx, vy, cxy = f(x * x) - mx * mx, f(y * y) - my * my, f(x * y) - mx * my
        vals.append((((2 * mx * my + c1) * (2 * cxy + c2)) / ((mx * mx + my * my + c1) * (vx + vy + c2))).mean())
    return float(np.mean(vals))


def test_quality_metrics(io):
    fp = C.POINTER(C.c_float)
    rng = np.random.default_rng(4)
    h, w = 37, 52
    a = rng.uniform(0, 1, (h, w, 3)).astype(np.float32)
    a[:, : w // 2] = np.linspace(0, 1, w // 2, dtype=np.float32)[None, :, None]  # structure, not only noise
    b = np.clip(a + rng.normal(0, 0.05, a.shape), 0, 1).astype(np.float32)
    v = C.c_double()
    assert io.bmfr_io_psnr(a.ctypes.data_as(fp), b.ctypes.data_as(fp), a.size, 1.0, C.byref(v)) == OK
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    assert abs(v.value - 10 * np.log10(1.0 / mse)) < 1e-9
    assert io.bmfr_io_psnr(a.ctypes.data_as(fp), a.ctypes.data_as(fp), a.size, 1.0, C.byref(v)) == OK and v.value == np.inf
    for x, y, peak in ((a, b, 1.0), (a, a, 1.0), (a, 1 - a, 1.0), (2 * a, 2 * b, 2.0)):
        assert io.bmfr_io_ssim_rgb(x.ctypes.data_as(fp), y.ctypes.data_as(fp), w, h, peak, C.byref(v)) == OK
        assert abs(v.value - _ssim_reference(x, y, peak)) < 1e-9, (v.value, _ssim_reference(x, y, peak))
    assert io.bmfr_io_ssim_rgb(a.ctypes.data_as(fp), a.ctypes.data_as(fp), w, h, 1.0, C.byref(v)) == OK and abs(v.value - 1) < 1e-12
    bad = a.copy()
    bad[3, 4, 1] = np.nan
    assert io.bmfr_io_psnr(a.ctypes.data_as(fp), bad.ctypes.data_as(fp), a.size, 1.0, C.byref(v)) == ERR_FORMAT
    assert io.bmfr_io_ssim_rgb(a.ctypes.data_as(fp), bad.ctypes.data_as(fp), w, h, 1.0, C.byref(v)) == ERR_FORMAT
    assert io.bmfr_io_ssim_rgb(a.ctypes.data_as(fp), b.ctypes.data_as(fp), 10, 10, 1.0, C.byref(v)) == -1
    # the display transform of bmfr.cl:852-856
    lin = np.array([-1.0, 0.0, 0.18, 1.0, 7.5, np.nan, np.inf], dtype=np.float32)
    t = lin.copy()
    io.bmfr_io_tone_map(t.ctypes.data_as(fp), t.size)
    want = np.clip(np.power(np.maximum(np.nan_to_num(lin, nan=0.0), 0), np.float32(0.454545)), 0, 1).astype(np.float32)
    assert np.allclose(t, want, rtol=1e-6, atol=0) and t[0] == 0 and t[3] == 1 and t[5] == 0 and t[6] == 1
