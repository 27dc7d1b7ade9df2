"""bmfr_run, the reference's driver program on the C ABI: a dataset in the reference's
format (EXR frames + camera_matrices.h, bmfr.cpp:43-52) goes in, the reference's profiling tables and PNG
frames come out; the frames must be the ones the library produces for the same inputs."""
import re
import subprocess

import numpy as np
import pytest

from bmfr_b200 import Denoiser, build, synth

from .exr_util import FLOAT, ZIP, write_exr

W, H, FRAMES = 128, 72, 4
# not the defaults, so that the test notices if the header's limits do not reach the kernels
LIMITS = tuple(float(np.float32(v)) for v in (0.75 * synth.limits()[0], 0.5 * synth.limits()[1]))


def _write_dataset(root):
    pl, nl = LIMITS
    mats, offs, frames = [], [], []
    for f in range(FRAMES):
        a, n, p, c = synth.frame_host(W, H, f)
        frames.append((a, n, p, c))
        for stem, img in (("albedo", a), ("shading_normal", n), ("world_position", p), ("color", c)):
            img = np.asarray(img, dtype=np.float32).reshape(H, W, 3)
            write_exr(root / f"{stem}{f}.exr", {"R": img[..., 0], "G": img[..., 1], "B": img[..., 2]}, dict(R=FLOAT, G=FLOAT, B=FLOAT), ZIP)
        m, o = synth.camera(f, W, H)
        mats.append(np.asarray(m, dtype=np.float32).reshape(4, 4))
        offs.append(np.asarray(o, dtype=np.float32))

    def lit(v):  # the shortest text that reads back as the same float
        return np.format_float_scientific(np.float32(v), unique=True) + "f"

    h = "// camera_matrices.h of a synthetic test sequence\n"
    h += f"const float position_limit_squared = {lit(pl)};\nconst float normal_limit_squared = {lit(nl)};\n"
    h += f"const float camera_matrices[{FRAMES}][4][4] = {{\n"
    for m in mats:
        h += "    { " + ", ".join("{" + ", ".join(lit(v) for v in row) + "}" for row in m) + " },\n"
    h += "};\n" + f"const float pixel_offsets[{FRAMES}][2] = {{ " + ", ".join("{" + lit(o[0]) + ", " + lit(o[1]) + "}" for o in offs) + " };\n"
    (root / "camera_matrices.h").write_text(h)
    return frames


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fused", "staged"])
def test_driver_runs_a_dataset_in_the_reference_format(tmp_path, mode):
    from PIL import Image

    exe = build.build_driver()
    data, out = tmp_path / "frames", tmp_path / "outputs"
    data.mkdir()
    out.mkdir()
    frames = _write_dataset(data)
    # --truth: the noisy input stands in for a ground-truth sequence (same file format: linear radiance)
    r = subprocess.run([str(exe), "--data", str(data), "--frames", str(FRAMES), "--out", str(out), "--truth", str(data / "color"), mode],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    reported = [tuple(float(x) for x in m) for m in re.findall(r"frame +\d+ : PSNR +([0-9.]+) dB +SSIM ([0-9.]+)", r.stdout)]
    assert len(reported) == FRAMES, r.stdout
    for line in ("Initialize.", "Loading input data.", "Run and profile kernels.", "Total time in all kernels"):  # bmfr.cpp:181,252,387
        assert line in r.stdout
    # the same frames through the library
    with Denoiser(W, H, mode=mode, position_limit_squared=LIMITS[0], normal_limit_squared=LIMITS[1]) as d:
        for f, (a, n, p, c) in enumerate(frames):
            cam_prev, _ = synth.camera(max(f - 1, 0), W, H)
            _, off = synth.camera(f, W, H)
            res = np.empty((H, W, 3), dtype=np.float32)
            d.denoise_frame_host(f, a, n, p, c, cam_prev, off, res)
            d.sync()
            want = np.floor(np.clip(np.nan_to_num(res, nan=0.0), 0, 1) * np.float32(255) + np.float32(0.5)).astype(np.uint8)
            truth = np.clip(np.power(np.maximum(np.asarray(c, dtype=np.float32).reshape(H, W, 3), 0), np.float32(0.454545)), 0, 1)
            mse = np.mean((res.astype(np.float64) - truth.astype(np.float64)) ** 2)
            assert abs(reported[f][0] - 10 * np.log10(1.0 / mse)) < 2e-3, (f, reported[f], 10 * np.log10(1.0 / mse))
            assert 0.0 < reported[f][1] <= 1.0
            got = np.asarray(Image.open(out / f"output{f}.png"))
            assert got.shape == (H, W, 3)
            assert np.array_equal(got, want), f"frame {f}: {np.abs(got.astype(int) - want.astype(int)).max()} levels off"


@pytest.mark.gpu
def test_driver_reports_a_broken_dataset_like_the_reference(tmp_path):
    exe = build.build_driver()
    data = tmp_path / "frames"
    data.mkdir()
    _write_dataset(data)
    (data / "world_position2.exr").unlink()
    r = subprocess.run([str(exe), "--data", str(data), "--frames", str(FRAMES)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1
    assert "Position buffer loading failed, reason:" in r.stdout  # bmfr.cpp:289-291
    assert "One or more errors occurred during buffer loading" in r.stdout  # bmfr.cpp:310
