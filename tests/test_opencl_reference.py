"""The reference's UNMODIFIED bmfr.cl, compiled by the box's own OpenCL driver and run on the B200
(oracle/_ref/libbmfr_clgpu.so, built from /root/reference by oracle/build_oracle.py), as a second, independent pin:

  * against the CPU port oracle with FP_CONTRACT OFF and correctly rounded division: frame 0 and the block min/max bit
    for bit; from frame 1 on within last-bit differences of prev_pixels — the vendor's dot() built-in (the camera
    projection, bmfr.cl:343-347) is fused whatever the pragma says, which OpenCL C allows — hence integer buffers equal
    up to 1e-4 of the pixels and colour within the north star's tolerance;
  * against the CUDA path in the vendor compiler's default mode (FMA contraction allowed by OpenCL C): colour within
    the north star's tolerance, integer buffers within a handful of pixels (last-bit differences of prev_pixels).
Skipped where no OpenCL platform can be opened (this container; a box without the NVIDIA ICD)."""
import os

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _opencl(w, h, strict, feature_set=0):
    from bmfr_b200 import synth
    from oracle import oracle as orc
    if not orc.available("opencl", feature_set):
        pytest.skip("oracle/_ref/libbmfr_clgpu*.so was not built (needs /root/reference at build time)")
    os.environ["BMFR_OPENCL_STRICT_FP"] = "1" if strict else "0"
    pl, nl = synth.limits()
    try:
        return orc.Oracle("opencl", w, h, position_limit_squared=pl, normal_limit_squared=nl, feature_set=feature_set)
    except RuntimeError as e:
        pytest.skip(f"no usable OpenCL platform here: {e}")


def test_vendor_compiled_reference_kernels_match_the_port():
    from bmfr_b200 import synth
    from oracle.oracle import Oracle
    w, h, frames = 416, 250, 10
    cl = _opencl(w, h, strict=True)
    pl, nl = synth.limits()
    port = Oracle("port", w, h, position_limit_squared=pl, normal_limit_squared=nl)
    for fr in util.sequence(w, h, frames):
        cl.frame(*fr)
        port.frame(*fr)
        assert util.bits_equal(cl.buffer("mins_maxs"), port.buffer("mins_maxs")), f"frame {fr[0]}: mins_maxs"
        if fr[0] == 0:  # no reprojection yet: everything up to the fit's inputs is exact arithmetic
            for k in ("spp", "accept", "prev_pixels", "noisy_acc"):
                assert util.bits_equal(cl.buffer(k), port.buffer(k)), f"frame 0: {k} of the OpenCL run differs from the port"
        for k in ("spp", "accept"):
            bad = int((cl.buffer(k) != port.buffer(k)).sum())
            assert bad <= 1e-4 * w * h, f"frame {fr[0]}: {k} differs in {bad} pixels"
        dp = np.abs(cl.buffer("prev_pixels") - port.buffer("prev_pixels"))
        assert float(dp.max()) <= 1e-3, f"frame {fr[0]}: prev_pixels differ by {dp.max()} pixels"
        for k in ("noisy_acc", "filtered", "accum", "result"):
            util.assert_colour_close(cl.buffer(k), port.buffer(k), f"frame {fr[0]} {k}")
    cl.close(); port.close()


@pytest.mark.parametrize("feature_set,size", [(0, (1280, 720)), (1, (416, 250)), (2, (416, 250))])
def test_cuda_path_matches_the_reference_kernels_on_the_same_gpu(feature_set, size):
    """... for the reference's shipped feature list at 720p and for the two other lists (the -D FEATURE_BUFFERS strings of
    oracle/build_oracle.py FEATURE_SETS, compiled by the OpenCL driver)."""
    from bmfr_b200 import Denoiser
    (w, h), frames = size, 12
    cl = _opencl(w, h, strict=False, feature_set=feature_set)
    with Denoiser(w, h, mode="fused", feature_set=feature_set) as d:
        for fr in util.sequence(w, h, frames):
            cl.frame(*fr)
            d.denoise_frame_host(*fr)
            for k in ("spp", "accept"):
                bad = int((cl.buffer(k) != d.read(k)).sum())
                assert bad <= 1e-4 * w * h, f"frame {fr[0]}: {k} differs in {bad} pixels"
            assert util.floats_equal_mod_zero_sign(cl.buffer("mins_maxs"), d.read("mins_maxs"))
            dp = np.abs(cl.buffer("prev_pixels") - d.read("prev_pixels"))
            assert float(dp.max()) <= 1e-3, f"frame {fr[0]}: prev_pixels differ by {dp.max()} pixels"
            # Two legal evaluations of the same source (contracted on the OpenCL side): a pixel whose accept mask or sample
            # count flips takes a different temporal path in the two runs, so the per-pixel bound is asked of all but 1e-4 of
            # the pixels and the PSNR bound of the whole frame.
            for k in ("accum", "result"):
                a, b = d.read(k).astype(np.float64), cl.buffer(k).astype(np.float64)
                bad = np.abs(a - b) > util.REL * np.maximum(np.abs(b), util.EPS)
                assert bad.mean() <= 1e-4, f"frame {fr[0]} {k}: {bad.sum()} elements outside the tolerance"
                _, psnr = util.colour_error(a, b)
                assert psnr >= util.PSNR_DB, f"frame {fr[0]} {k}: PSNR {psnr:.1f} dB"
    cl.close()
