"""Parity proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded synthetic inputs.  Integer buffers and the per-pixel fp32 stages that contain no reduction
must be bit-exact; fitted / accumulated colour must be within the north star's tolerance
(per-pixel relative error <= 1e-3, PSNR >= 60 dB)."""
import numpy as np
import pytest

from bmfr_b200 import Denoiser
from tests import util

pytestmark = pytest.mark.gpu

KEEP_FUSED = ("noisy_acc", "spp", "prev_pixels", "accept", "weights", "mins_maxs", "accum", "result", "noise_tile")
KEEP_STAGED = KEEP_FUSED + ("tmp_data", "filtered", "tone_mapped")


def _check_frames(cuda, ref, staged):
    worst = dict(rel=0.0, psnr=1e9)
    for f, (c, r) in enumerate(zip(cuda, ref)):
        for k in util.INTEGER_BUFFERS:
            assert util.bits_equal(c[k], r[k]), f"frame {f}: {k} differs in {(c[k] != r[k]).sum()} pixels"
        for k in ("noisy_acc", "prev_pixels", "noise_tile") + (("tmp_data",) if staged else ()):
            assert util.bits_equal(c[k], r[k]), f"frame {f}: {k} not bit-identical"
        # floor(prev pixel) is what the north star names; it follows from the float equality above
        assert np.array_equal(np.floor(c["prev_pixels"]), np.floor(r["prev_pixels"]))
        assert util.floats_equal_mod_zero_sign(c["mins_maxs"], r["mins_maxs"]), f"frame {f}: mins_maxs differ"
        for k in util.COLOUR_BUFFERS + (("filtered", "tone_mapped") if staged else ()):
            rel, psnr = util.assert_colour_close(c[k], r[k], f"frame {f} {k}")
            worst["rel"], worst["psnr"] = max(worst["rel"], rel), min(worst["psnr"], psnr)
    return worst


@pytest.mark.parametrize("mode", ["staged", "fused"])
@pytest.mark.parametrize("size", [(160, 96), (200, 120), (416, 250)])
def test_sequence_matches_oracle(mode, size):
    """20 frames: every BLOCK_OFFSETS entry, ragged sizes (not multiples of 32), frame-0 path."""
    w, h = size
    keep = KEEP_STAGED if mode == "staged" else KEEP_FUSED
    ref = util.run_oracle("port", w, h, 20, keep=keep)
    cuda = util.run_cuda(w, h, 20, mode=mode, keep=keep)
    worst = _check_frames(cuda, ref, mode == "staged")
    print(f"{mode} {w}x{h}: worst rel {worst['rel']:.2e}, worst PSNR {worst['psnr']:.1f} dB")


@pytest.mark.parametrize("size", [(32, 32), (48, 47), (64, 64), (97, 75), (98, 70), (132, 100)])
def test_fused_edge_geometries(size):
    """One-block images, odd widths, widths that are not multiples of 4 (the TMA tile path of the fit needs
    W % 4 == 0 and falls back to per-pixel loads otherwise) — every block offset of the 16-entry table."""
    w, h = size
    ref = util.run_oracle("port", w, h, 17, keep=KEEP_FUSED)
    cuda = util.run_cuda(w, h, 17, mode="fused", keep=KEEP_FUSED)
    _check_frames(cuda, ref, False)


def test_nan_and_far_inputs_follow_the_oracle():
    """NaN normals / positions are scrubbed before the fit (bmfr.cl:468-469) but not in the weighted sum
    (bmfr.cl:725-750): the NaN pattern of every colour buffer must be the oracle's."""
    w, h, frames = 160, 96, 4
    seq = list(util.sequence(w, h, frames))
    rng = np.random.default_rng(7)
    ys, xs = rng.integers(0, h, 40), rng.integers(0, w, 40)
    for k, (f, a, n, p, c, cam, off) in enumerate(seq):
        n, p = n.copy(), p.copy()
        n[ys[:20], xs[:20], 0] = np.nan
        p[ys[20:], xs[20:], 1] = np.nan
        seq[k] = (f, a, n, p, c, cam, off)
    from bmfr_b200 import Denoiser, synth
    from oracle.oracle import Oracle
    pl, nl = synth.limits()
    o = Oracle("port", w, h, position_limit_squared=pl, normal_limit_squared=nl, keep_tmp=0)
    with Denoiser(w, h, mode="fused") as d:
        for fr in seq:
            o.frame(*fr)
            d.denoise_frame_host(*fr)
            for k in util.INTEGER_BUFFERS:
                assert util.bits_equal(d.read(k), o.buffer(k)), k
            r_c, r_o = d.read("result"), o.buffer("result")
            assert np.array_equal(np.isnan(r_c), np.isnan(r_o)), f"frame {fr[0]}: NaN pattern of result differs"
            m = ~np.isnan(r_o)
            util.assert_colour_close(np.where(m, r_c, 0), np.where(m, r_o, 0), f"frame {fr[0]} result")
    o.close()


@pytest.mark.parametrize("half", [0, 1], ids=["fp32_tmp", "fp16_tmp"])
@pytest.mark.parametrize("size", [(160, 96), (200, 120)])
def test_reference_order_mode_is_bit_identical(size, half):
    """STAGED with reference_order=1 evaluates the fitter and the weighted sum in the reference's own
    operation order: every buffer up to the accumulated colour — the fit coefficients included — is
    bit-identical to the oracle, for the fp32 tmp_data of the north star and for the reference's shipped
    default USE_HALF_PRECISION_IN_TMP_DATA=1 (bmfr.cpp:88).  Only the tone map (powr: 16 ulp in OpenCL,
    different libm implementations) keeps tone_mapped / result within tolerance."""
    w, h = size
    keep = ("noisy_acc", "spp", "prev_pixels", "accept", "weights", "mins_maxs", "filtered", "accum", "tone_mapped", "result")
    ref = util.run_oracle("port", w, h, 18, keep=keep, tmp_half=half)
    cuda = util.run_cuda(w, h, 18, mode="staged", keep=keep, tmp_half=half, reference_order=1)
    for f, (c, r) in enumerate(zip(cuda, ref)):
        for k in ("noisy_acc", "spp", "prev_pixels", "accept", "mins_maxs", "weights", "filtered", "accum"):
            assert util.bits_equal(c[k], r[k]), f"frame {f}: {k} differs in {(c[k] != r[k]).sum()} elements"
        for k in ("tone_mapped", "result"):
            util.assert_colour_close(c[k], r[k], f"frame {f} {k}")


def test_jittered_offsets_match_oracle():
    ref = util.run_oracle("port", 200, 120, 8, keep=KEEP_FUSED, jitter=True)
    cuda = util.run_cuda(200, 120, 8, mode="fused", keep=KEEP_FUSED, jitter=True)
    _check_frames(cuda, ref, False)


def test_720p_matches_oracle():
    """BASELINE.json configs[1] geometry, first frames."""
    ref = util.run_oracle("port", 1280, 720, 4, keep=KEEP_FUSED)
    cuda = util.run_cuda(1280, 720, 4, mode="fused", keep=KEEP_FUSED)
    _check_frames(cuda, ref, False)


def test_tsqr_fit_still_matches_oracle():
    """bmfr_params.fit_method = BMFR_FIT_TSQR: the round-1 two-level QR fit stays selectable and in tolerance."""
    ref = util.run_oracle("port", 416, 250, 8, keep=KEEP_FUSED)
    cuda = util.run_cuda(416, 250, 8, mode="fused", keep=KEEP_FUSED, fit="tsqr")
    _check_frames(cuda, ref, False)


def test_busy_stamps_do_not_change_results_and_are_plausible():
    """profile = 2: every kernel stamps the start of its first and the end of its last CTA (device globaltimer) instead of
    event records between the launches.  Same bits as an unprofiled run; times positive, ordered, below a frame's span."""
    w, h, frames = 416, 250, 6
    seq = list(util.sequence(w, h, frames))
    outs = {}
    for prof in (0, 2):
        with Denoiser(w, h, mode="fused", profile=prof) as d:
            for fr in seq:
                d.denoise_frame_host(*fr)
            outs[prof] = {k: d.read(k) for k in ("result", "accum", "spp", "accept")}
            if prof == 2:
                for f in range(1, frames):
                    busy, span = d.fused_kernel_busy_ms(f)
                    assert set(busy) == set(d.fused_kernels)
                    assert all(0.0 < v < 5.0 for v in busy.values()), busy
                    assert max(busy.values()) <= span <= sum(busy.values()) + 1.0, (busy, span)
                    # the stamps behind them: start <= end per kernel, the chain R -> F -> P starts in that order, and
                    # their differences are the busy times
                    st = d.fused_kernel_stamps(f)
                    assert all(a <= b for a, b in st) and st[0][0] <= st[1][0] <= st[2][0], st
                    for (a, b), v in zip(st, busy.values()):
                        assert abs((b - a) * 1e-6 - v) < 1e-3
                    if f > 1:  # one clock for all frames of a device: frames follow each other
                        assert st[0][0] > prev_start
                    prev_start = st[0][0]
    for k in outs[0]:
        assert util.bits_equal(outs[0][k], outs[2][k]), k


@pytest.mark.parametrize("size", [(416, 250), (200, 120)])
@pytest.mark.parametrize("feature_set", [1, 2])
def test_other_feature_lists_match_their_oracle(feature_set, size):
    """bmfr_params.feature_set: the FUSED kernels instantiated for 1, n | p (7 features, 3 scaled) and 1 | p, p^2 (7 features,
    6 scaled), against the port built for the same list (which equals the reference's kernels compiled with that list,
    tests/test_oracle_pin.py).  Same bars as the default list; 416 wide runs the TMA post pass, 200 the per-thread one."""
    w, h = size
    ref = util.run_oracle("port", w, h, 18, keep=KEEP_FUSED, feature_set=feature_set)
    cuda = util.run_cuda(w, h, 18, mode="fused", keep=KEEP_FUSED, feature_set=feature_set)
    assert cuda[0]["weights"].shape == ref[0]["weights"].shape and cuda[0]["mins_maxs"].shape == ref[0]["mins_maxs"].shape
    _check_frames(cuda, ref, False)


def test_feature_lists_need_the_fused_gram_path():
    from bmfr_b200 import BmfrError
    for kw in (dict(mode="staged"), dict(mode="fused", fit="tsqr")):
        with pytest.raises(BmfrError) as e:
            Denoiser(160, 96, feature_set=1, **kw)
        assert e.value.status == -5


def test_staged_and_fused_agree():
    """The two kernel structures share the reprojection code (bit-identical K1 outputs, block min/max
    and noise tile); the fit and the post-fit passes are separate implementations held to the colour
    tolerance."""
    a = util.run_cuda(416, 250, 6, mode="staged", keep=KEEP_FUSED, every_frame=False)[0]
    b = util.run_cuda(416, 250, 6, mode="fused", keep=KEEP_FUSED, every_frame=False)[0]
    for k in ("noisy_acc", "spp", "prev_pixels", "accept", "noise_tile"):
        assert util.bits_equal(a[k], b[k]), k
    assert util.floats_equal_mod_zero_sign(a["mins_maxs"], b["mins_maxs"])
    for k in util.COLOUR_BUFFERS:
        util.assert_colour_close(b[k], a[k], k)


@pytest.mark.parametrize("size", [(256, 160), (1920, 1080)])
def test_overlapped_frames_are_bit_identical(size):
    """params.overlap_frames: the three kernels on three event-linked streams, per-frame temporaries double-
    buffered.  Same kernels on the same inputs: every output frame and every buffer must equal the in-order
    run bit for bit, through the host entry (upload ring, read-back) and through the device-pointer entry with
    all frames resident (nothing between the frames but the events)."""
    import torch

    w, h = size
    frames = 9
    seq = list(util.sequence(w, h, frames))

    def run_host(overlap):
        outs = [np.empty((h, w, 3), dtype=np.float32) for _ in range(frames)]
        with Denoiser(w, h, mode="fused", overlap_frames=overlap) as d:
            for (f, a, n, p, c, cam, off), out in zip(seq, outs):
                d.denoise_frame_host(f, a, n, p, c, cam, off, out)
            d.sync()
            return outs, {k: d.read(k) for k in KEEP_FUSED}

    def run_device(overlap):
        dev = [[torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (a, n, p, c)] for (_, a, n, p, c, _, _) in seq]
        outs = [torch.empty((h, w, 3), dtype=torch.float32, device="cuda") for _ in range(frames)]
        torch.cuda.synchronize()
        with Denoiser(w, h, mode="fused", overlap_frames=overlap) as d:
            for (f, _, _, _, _, cam, off), t, out in zip(seq, dev, outs):
                d.denoise_frame(f, *[x.data_ptr() for x in t], cam, off, out.data_ptr())
            d.sync()
            return [o.cpu().numpy() for o in outs], {k: d.read(k) for k in KEEP_FUSED}

    def run_joined():
        """bmfr_join: work on the context's stream after it sees the frames complete, with no host synchronisation."""
        stream = torch.cuda.Stream()
        dev = [[torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (a, n, p, c)] for (_, a, n, p, c, _, _) in seq]
        outs = [torch.empty((h, w, 3), dtype=torch.float32, device="cuda") for _ in range(frames)]
        torch.cuda.synchronize()
        with Denoiser(w, h, mode="fused", overlap_frames=1, stream=stream.cuda_stream) as d:
            for (f, _, _, _, _, cam, off), t, out in zip(seq, dev, outs):
                d.denoise_frame(f, *[x.data_ptr() for x in t], cam, off, out.data_ptr())
            d.join()
            with torch.cuda.stream(stream):
                snap = [o.clone() for o in outs]  # stream-ordered after the join
            stream.synchronize()  # the context's stream only: the internal streams are not waited for by the host
            got = [x.cpu().numpy() for x in snap]
            d.sync()
        return got

    joined = run_joined()
    ref_outs, _ = run_device(0)
    for f in range(frames):
        assert util.bits_equal(joined[f], ref_outs[f]), f"bmfr_join: frame {f} was copied before it was complete"

    for run in (run_host, run_device):
        ref_outs, ref_bufs = run(0)
        for rep in range(2):
            outs, bufs = run(1)
            for f in range(frames):
                assert util.bits_equal(outs[f], ref_outs[f]), f"{run.__name__}: output of frame {f} differs (pass {rep})"
            for k in KEEP_FUSED:
                assert util.bits_equal(bufs[k], ref_bufs[k]), f"{run.__name__}: {k} differs (pass {rep})"
