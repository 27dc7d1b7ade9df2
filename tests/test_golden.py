"""Committed golden vectors produced by the reference's own kernels (tests/golden/make_golden.py).
CPU: the plain-C port reproduces every buffer of every frame bit for bit, and the synthetic inputs
have not drifted.  GPU: the CUDA path matches the stored reference buffers."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from bmfr_b200 import synth
from tests import util
from tests.util import Runner

GOLD = np.load(Path(__file__).parent / "golden" / "bmfr_ref_128x72.npz")
META = json.loads(bytes(GOLD["meta"]).decode())
BUFFERS = ("noisy_acc", "spp", "prev_pixels", "accept", "tmp_data", "weights", "mins_maxs", "filtered", "accum",
           "tone_mapped", "result")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def frames():
    w, h = META["width"], META["height"]
    for f in range(META["frames"]):
        a, n, p, c = synth.frame_host(w, h, f, seed=META["seed"])
        cam, _ = synth.camera(max(f - 1, 0), w, h, True)
        _, off = synth.camera(f, w, h, True)
        yield f, a, n, p, c, cam, off


def test_fixture_came_from_the_reference_kernels():
    assert META["kind"] == "reference" and META["tmp_half"] == 0


def test_synthetic_inputs_have_not_drifted():
    for f, *arrs in frames():
        assert [sha(x) for x in arrs] == META["input_sha256"][f], f"frame {f}: synth-v1 output changed; regenerate the fixture"


def test_port_reproduces_golden_bits():
    r = Runner("port", META["width"], META["height"], position_limit_squared=META["position_limit_squared"],
               normal_limit_squared=META["normal_limit_squared"])
    for f, *arrs in frames():
        r.frame(f, *arrs)
        for k in BUFFERS:
            assert sha(r.get(k)) == META["buffer_sha256"][f][k], f"frame {f}: {k}"
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("backend", ["cuda-staged", "cuda-fused"])
def test_cuda_matches_golden(backend):
    r = Runner(backend, META["width"], META["height"], position_limit_squared=META["position_limit_squared"],
               normal_limit_squared=META["normal_limit_squared"])
    for f, *arrs in frames():
        r.frame(f, *arrs)
        for k in ("spp", "accept", "noisy_acc", "prev_pixels"):      # bit-exact obligations
            assert sha(r.get(k)) == META["buffer_sha256"][f][k], f"frame {f}: {k}"
        if backend == "cuda-staged":
            assert sha(r.get("tmp_data")) == META["buffer_sha256"][f]["tmp_data"]
        if f in (1, 7):
            assert util.floats_equal_mod_zero_sign(r.get("mins_maxs"), GOLD[f"f{f}_mins_maxs"])
            for k in ("accum", "result"):
                util.assert_colour_close(r.get(k), GOLD[f"f{f}_{k}"], f"frame {f} {k}")
    r.close()


@pytest.mark.gpu
def test_reference_order_mode_reproduces_the_reference_kernels_bits():
    """STAGED + reference_order against the SHA-256 of the buffers the reference's own kernels produced
    (tests/golden/make_golden.py): the fit coefficients, the weighted sum and the accumulated colour are
    the reference's bits, not just the reprojection outputs."""
    from bmfr_b200 import Denoiser
    d = Denoiser(META["width"], META["height"], mode="staged", reference_order=1,
                 position_limit_squared=META["position_limit_squared"], normal_limit_squared=META["normal_limit_squared"])
    for f, *arrs in frames():
        d.denoise_frame_host(f, *arrs)
        for k in ("spp", "accept", "noisy_acc", "prev_pixels", "mins_maxs", "weights", "filtered", "accum"):
            assert sha(d.read(k)) == META["buffer_sha256"][f][k], f"frame {f}: {k}"
    d.close()
