"""Parity on the configurations BASELINE.json names and bench.py times (SURVEY.md 8c-8, 8e "Correctness bar"):

  * 1920x1080 x 60 frames (the reference's FRAME_COUNT, bmfr.cpp:42), FUSED, against the CPU oracle: the integer
    and exact buffers bit for bit on EVERY frame, the colour buffers within the north star's tolerance on every
    frame — once through the in-order stream and once with overlap_frames = 1 (what bench.py's `value` times),
    all 60 frames submitted back to back;
  * 3840x2160 over 4 strips (BASELINE.json configs[3]) against the oracle;
  * 7680x4320 over 8 strips (configs[4]) against the whole-image CUDA run.
The strips of the last two live in one process on one GPU (LocalStripSet); the cross-process CUDA-IPC transport
has its own test (tests/test_ipc_halo.py) and bench.py's sharded line carries a parity field.
"""
import numpy as np
import pytest

from bmfr_b200 import Denoiser, sharding, synth
from tests import util

pytestmark = pytest.mark.gpu

EXACT = ("spp", "accept", "prev_pixels", "noisy_acc")


def _oracle(w, h):
    from oracle.oracle import Oracle
    pl, nl = synth.limits()
    return Oracle("port", w, h, position_limit_squared=pl, normal_limit_squared=nl)


def test_1080p_60_frames_match_oracle_every_frame_in_order_and_overlapped():
    import torch
    w, h, frames = 1920, 1080, 60
    o = _oracle(w, h)
    ref_result = []
    worst = dict(rel=0.0, psnr=1e9)
    dev_in = torch.empty((frames, 4, h, w, 3), dtype=torch.float32, device="cuda")  # 5.97 GB, like bench.py
    cams, offs = [], []
    with Denoiser(w, h, mode="fused") as d:
        for f, a, n, p, c, cam, off in util.sequence(w, h, frames):
            o.frame(f, a, n, p, c, cam, off)
            d.denoise_frame_host(f, a, n, p, c, cam, off)
            for k, x in enumerate((a, n, p, c)):
                dev_in[f, k].copy_(torch.from_numpy(x))
            cams.append(cam); offs.append(off)
            for k in EXACT:
                got, want = d.read(k), o.buffer(k)
                assert util.bits_equal(got, want), f"in order, frame {f}: {k} differs in {(got != want).sum()} elements"
            assert util.floats_equal_mod_zero_sign(d.read("mins_maxs"), o.buffer("mins_maxs")), f"frame {f}: mins_maxs"
            for k in util.COLOUR_BUFFERS:
                rel, psnr = util.assert_colour_close(d.read(k), o.buffer(k), f"in order, frame {f} {k}")
                worst["rel"], worst["psnr"] = max(worst["rel"], rel), min(worst["psnr"], psnr)
            ref_result.append(o.buffer("result"))
    last = {k: o.buffer(k) for k in EXACT + util.COLOUR_BUFFERS}
    o.close()
    print(f"1080p x60 in order: worst rel {worst['rel']:.2e}, worst PSNR {worst['psnr']:.1f} dB")

    # the benchmarked mode: overlap_frames = 1, device-pointer entry, nothing between the frames
    outs = torch.empty((frames, h, w, 3), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    with Denoiser(w, h, mode="fused", overlap_frames=1) as d:
        for f in range(frames):
            d.denoise_frame(f, *[dev_in[f, k].data_ptr() for k in range(4)], cams[f], offs[f], outs[f].data_ptr())
        d.sync()
        for f in range(frames):
            util.assert_colour_close(outs[f].cpu().numpy(), ref_result[f], f"overlapped, frame {f} result")
        for k in EXACT:
            assert util.bits_equal(d.read(k), last[k]), f"overlapped, last frame: {k} not bit-identical to the oracle"
        for k in util.COLOUR_BUFFERS:
            util.assert_colour_close(d.read(k), last[k], f"overlapped, last frame {k}")


def _device_sequence(w, h, frames):
    import torch
    seq = []
    for f in range(frames):
        t = torch.empty((4, h, w, 3), dtype=torch.float32, device="cuda")
        synth.frame_device(w, h, f, [t[k].data_ptr() for k in range(4)])
        seq.append(t)
    torch.cuda.synchronize()
    cams = [synth.camera(max(f - 1, 0), w, h)[0] for f in range(frames)]
    offs = [synth.camera(f, w, h)[1] for f in range(frames)]
    return seq, cams, offs


def _halo(h):
    return sharding.default_halo(h)  # what bench.py's sharded arm uses


@pytest.mark.parametrize("overlap", [0, 1])
def test_4k_over_4_strips_matches_oracle(overlap):
    """BASELINE.json configs[3]: 3840x2160, four strips.  Owned rows of every strip against the CPU oracle."""
    import torch
    w, h, frames, n = 3840, 2160, 5, 4
    seq, cams, offs = _device_sequence(w, h, frames)
    o = _oracle(w, h)
    ss = sharding.LocalStripSet(w, h, n, halo=_halo(h), exchange="p2p", overlap_frames=overlap)
    outs = [torch.zeros((h, w, 3), dtype=torch.float32, device="cuda") for _ in range(frames)]
    refs = []
    for f in range(frames):
        ss.denoise_frame(f, [seq[f][k] for k in range(4)], cams[f], offs[f], outs[f])
        if not overlap:
            ss.sync()
        host = seq[f].cpu().numpy()
        o.frame(f, host[0], host[1], host[2], host[3], cams[f], offs[f])
        refs.append(o.buffer("result"))
        if not overlap:  # every buffer, every frame
            for c in ss.ctx:
                y0, y1 = c.strip
                for k in EXACT:
                    got = c.d.read(k)[y0 - c.row0:y1 - c.row0]
                    assert util.bits_equal(got, o.buffer(k)[y0:y1]), f"frame {f}, strip {c.strip}: {k} differs"
                for k in util.COLOUR_BUFFERS:
                    util.assert_colour_close(c.d.read(k)[y0 - c.row0:y1 - c.row0], o.buffer(k)[y0:y1], f"frame {f} strip {c.strip} {k}")
    ss.sync()
    torch.cuda.synchronize()
    for f in range(frames):
        util.assert_colour_close(outs[f].cpu().numpy(), refs[f], f"4K over {n} strips (overlap {overlap}), frame {f}")
    for c in ss.ctx:  # last frame: exact buffers of the owned rows
        y0, y1 = c.strip
        for k in EXACT:
            assert util.bits_equal(c.d.read(k)[y0 - c.row0:y1 - c.row0], o.buffer(k)[y0:y1]), f"last frame, strip {c.strip}: {k}"
    ss.close()
    o.close()


def test_8k_over_8_strips_equals_whole_image_run_bitwise():
    """BASELINE.json configs[4]: 7680x4320, eight strips, overlapped frames, against the whole-image run of the same
    kernels: every output frame and the owned rows of every state buffer bit for bit."""
    import torch
    w, h, frames, n = 7680, 4320, 3, 8
    seq, cams, offs = _device_sequence(w, h, frames)
    out_w = [torch.zeros((h, w, 3), dtype=torch.float32, device="cuda") for _ in range(frames)]
    with Denoiser(w, h, mode="fused") as whole:
        for f in range(frames):
            whole.denoise_frame(f, *[seq[f][k].data_ptr() for k in range(4)], cams[f], offs[f], out_w[f].data_ptr())
        whole.sync()
        ref = {k: whole.read(k) for k in ("noisy_acc", "spp", "accum", "result", "accept")}
    ss = sharding.LocalStripSet(w, h, n, halo=_halo(h), exchange="p2p", overlap_frames=1)
    out_s = [torch.zeros((h, w, 3), dtype=torch.float32, device="cuda") for _ in range(frames)]
    for f in range(frames):
        ss.denoise_frame(f, [seq[f][k] for k in range(4)], cams[f], offs[f], out_s[f])
    ss.sync()
    torch.cuda.synchronize()
    for f in range(frames):
        assert torch.equal(out_s[f], out_w[f]), f"8K over 8 strips: output of frame {f} differs from the whole-image run"
    for k, r in ref.items():
        for c in ss.ctx:
            y0, y1 = c.strip
            assert np.array_equal(c.d.read(k)[y0 - c.row0:y1 - c.row0].view(np.uint8), r[y0:y1].view(np.uint8)), (k, c.strip)
    ss.close()
