"""Known-answer tests derived from the reference kernels' own invariants (the reference has no
tests; SURVEY.md 8c lists these).  Each runs on the CPU checkers (port, reference kernels) and —
marked gpu — on both CUDA kernel structures through the C ABI."""
import numpy as np
import pytest

from bmfr_b200 import block_offset, synth
from oracle import oracle as orc
from tests import util
from tests.util import Runner, backend_params

BACKENDS = backend_params()


def _skip_if_missing(backend):
    if backend == "reference" and not orc.available("reference"):
        pytest.skip("oracle/_ref not built")


# ---------------------------------------------------------------- integer helpers
def py_random(a):
    """random() of bmfr.cl:162-171 in Python integers."""
    m = 0xFFFFFFFF
    a = ((a + 0x7ed55d16) + (a << 12)) & m
    a = ((a ^ 0xc761c23c) ^ (a >> 19)) & m
    a = ((a + 0x165667b1) + (a << 5)) & m
    a = ((a + 0xd3a2646c) ^ (a << 9)) & m
    a = ((a + 0xfd7046c5) + (a << 3)) & m
    a = ((a ^ 0xb55a4f09) ^ (a >> 16)) & m
    return np.float32(a) / np.float32(4294967295)


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_hash_kat(kind):
    _skip_if_missing(kind)
    for seed in [0, 1, 2, 255, 256, 1023, 1024, 13311, 13312, 0x7FFFFFFF, 0xFFFFFFFF, 123456789]:
        assert orc.random_hash(kind, seed) == float(py_random(seed)), seed
    vals = np.array([orc.random_hash(kind, s) for s in range(4096)])
    assert 0 <= vals.min() and vals.max() <= 1 and abs(vals.mean() - 0.5) < 0.02


@pytest.mark.parametrize("backend", BACKENDS)
def test_noise_tile_kat(backend):
    """add_random() increments (bmfr.cl:173-182): fp64 value of NOISE_AMOUNT*2.f*(random(seed)-0.5f),
    seed = id + 256*s + 1024*feature + 13*1024*frame, identical for every block."""
    _skip_if_missing(backend)
    r = Runner(backend, 64, 64)
    a, n, p, c = synth.frame_host(64, 64, 0)
    for f in (0, 3):
        r.frame(f, a, n, p, c, np.eye(4, dtype=np.float32).ravel(), [0.5, 0.5])
        tile = r.get("noise_tile")
        for fb, idx in [(1, 0), (1, 1023), (5, 300), (9, 777)]:
            seed = idx + 1024 * fb + 13 * 1024 * f
            expect = np.float64(1e-2) * np.float64(2.0) * np.float64(py_random(seed) - np.float32(0.5))
            assert tile[fb - 1, idx] == expect
    r.close()


def test_block_grid_mapping_identity():
    """K1's gid = pixel + 16 - offset (bmfr.cl:314-315) and K3's group index (bmfr.cl:719-722) name the
    same block for every offset and every image corner; mirror() KAT (bmfr.cl:209-216)."""
    def mirror(i, size):
        return -i - 1 if i < 0 else (2 * size - i - 1 if i >= size else i)
    assert [mirror(i, 10) for i in (-3, -1, 0, 9, 10, 12)] == [2, 0, 0, 9, 9, 7]
    for w, h in [(1280, 720), (1920, 1080), (200, 120)]:
        wm = 32 * ((w + 31) // 32) + 32
        hm = 32 * ((h + 31) // 32) + 32
        for f in range(16):
            ox, oy = block_offset(f)
            assert -16 <= ox <= 14 and -16 <= oy <= 14 and ox % 2 == 0 and oy % 2 == 0
            for x, y in [(0, 0), (w - 1, 0), (0, h - 1), (w - 1, h - 1), (w // 2, h // 2)]:
                gx, gy = x + 16 - ox, y + 16 - oy
                assert 0 <= gx < wm and 0 <= gy < hm
                group_k1 = (gy // 32) * (wm // 32) + gx // 32
                group_k3 = ((x + 16 - ox) // 32) + ((y + 16 - oy) // 32) * (wm // 32)
                assert group_k1 == group_k3
            # every margin work-item mirrors to an in-image pixel
            for g, size, o in [(0, w, ox), (wm - 1, w, ox), (0, h, oy), (hm - 1, h, oy)]:
                assert 0 <= mirror(g - 16 + o, size) < size


# ---------------------------------------------------------------- K1
@pytest.mark.parametrize("backend", BACKENDS)
def test_k1_frame0(backend):
    """Frame 0 (bmfr.cl:336): accept 0, spp 1, prev pixel == own pixel, colour copied through."""
    _skip_if_missing(backend)
    w, h = 96, 64
    r = Runner(backend, w, h)
    a, n, p, c = synth.frame_host(w, h, 0)
    r.frame(0, a, n, p, c, None, [0.5, 0.5])
    assert (r.get("accept") == 0).all() and (r.get("spp") == 1).all()
    pp = r.get("prev_pixels")
    ys, xs = np.mgrid[0:h, 0:w]
    assert np.array_equal(pp[..., 0], xs.astype(np.float32)) and np.array_equal(pp[..., 1], ys.astype(np.float32))
    assert util.bits_equal(r.get("noisy_acc"), c)
    if backend != "cuda-fused":
        t = r.get("tmp_data").reshape((h + 32 + 31) // 32 * 0 + (32 * ((h + 31) // 32) + 32) // 32, -1, 13, 32, 32)
        # colour planes of the block that starts at margin (16-ox,16-oy) hold the input (offset frame 0 = -14,-14)
        ox, oy = block_offset(0)
        by, bx, yi, xi = (40 + 16 - oy) // 32, (50 + 16 - ox) // 32, (40 + 16 - oy) % 32, (50 + 16 - ox) % 32
        assert np.array_equal(t[by, bx, 10:13, yi, xi], c[40, 50])
        assert t[by, bx, 0, yi, xi] == 1.0 and np.array_equal(t[by, bx, 4:7, yi, xi], p[40, 50])
        assert np.array_equal(t[by, bx, 7:10, yi, xi], p[40, 50] * p[40, 50])
    # taa and accumulation are pass-through on frame 0: result == tone mapped == clamp(pow(albedo*filtered))
    res = r.get("result")
    assert np.isfinite(res).all() and res.min() >= 0 and res.max() <= 1
    r.close()


def _static_plane(w, h):
    """World position == pixel coordinates, constant normal, and a matrix that maps it back to
    pixel + (0.25, 0.25): every interior reprojection hits all four taps."""
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
    pos = np.stack([xs, ys, np.zeros_like(xs)], axis=2)
    nrm = np.zeros((h, w, 3), np.float32)
    nrm[..., 2] = 1
    M = np.zeros((4, 4), np.float32)
    M[0, 0], M[3, 0] = 2.0 / w, 1.0 / w - 1.0     # clip.x = 2(x+0.5)/W - 1
    M[1, 1], M[3, 1] = 2.0 / h, 1.0 / h - 1.0
    M[3, 3] = 1.0
    return pos, nrm, M.ravel(), np.array([0.25, 0.75], np.float32)   # subtracts (0.25, 1-0.75)


@pytest.mark.parametrize("backend", BACKENDS)
def test_k1_static_camera_accumulates_to_saturation(backend):
    """Identity reprojection: accept == 0x0F in the interior, spp = frame+1 up to the 255 saturation
    (bmfr.cl:433-441), accumulated colour = running mean until alpha hits BLEND_ALPHA (bmfr.cl:427-428)."""
    _skip_if_missing(backend)
    w, h = 64, 64
    pos, nrm, M, off = _static_plane(w, h)
    alb = np.full((h, w, 3), 0.5, np.float32)
    r = Runner(backend, w, h, position_limit_squared=4.0, normal_limit_squared=0.1)
    rng = np.random.default_rng(1)
    for f in range(262):
        col = rng.random((h, w, 3), dtype=np.float32)
        r.frame(f, alb, nrm, pos, col, M, off)
        if f in (1, 2, 5, 200, 254, 255, 261):
            acc, spp = r.get("accept"), r.get("spp")
            assert (acc[:-1, :-1] == 0x0F).all()
            assert (acc[:-1, -1] == 0x05).all() and (acc[-1, :-1] == 0x03).all() and acc[-1, -1] == 0x01
            assert (spp == min(f + 1, 255)).all(), (f, np.unique(spp))
            pp = r.get("prev_pixels")
            assert np.array_equal(np.floor(pp[..., 0]), pos[..., 0]) and np.array_equal(np.floor(pp[..., 1]), pos[..., 1])
    r.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_k1_rejections(backend):
    """Random world positions -> every tap fails the distance test -> behaves like frame 0;
    reprojection far outside the image -> in-image guard (bmfr.cl:380-381)."""
    _skip_if_missing(backend)
    w, h = 96, 64
    _, nrm, M, off = _static_plane(w, h)
    rng = np.random.default_rng(2)
    alb = np.full((h, w, 3), 0.5, np.float32)
    r = Runner(backend, w, h, position_limit_squared=1e-3, normal_limit_squared=0.1)
    for f in range(3):
        pos = (rng.random((h, w, 3), dtype=np.float32) * 50 + 100 * f).astype(np.float32)   # never near the previous frame
        pos[..., 0] = np.mgrid[0:h, 0:w][1]
        pos[..., 1] = np.mgrid[0:h, 0:w][0]
        col = rng.random((h, w, 3), dtype=np.float32)
        r.frame(f, alb, nrm, pos, col, M, off)
        assert (r.get("accept") == 0).all() and (r.get("spp") == 1).all()
        assert util.bits_equal(r.get("noisy_acc"), col)
    r.close()
    pos, nrm, M, off = _static_plane(w, h)
    M = M.copy().reshape(4, 4)
    M[3, 0] += 5.0                                  # shifts every reprojection 2.5 image widths to the right
    r = Runner(backend, w, h, position_limit_squared=1e9, normal_limit_squared=1e9)
    for f in range(2):
        r.frame(f, alb, nrm, pos, alb, M.ravel(), off)
    assert (r.get("accept") == 0).all() and (r.get("spp") == 1).all()
    assert util.bits_equal(r.get("result"), r.get("result"))   # taa copy-through path (bmfr.cl:884-890) ran without NaN
    assert np.isfinite(r.get("result")).all()
    r.close()


# ---------------------------------------------------------------- K2 / K3
def _random_frame(w, h, rng):
    nrm = rng.normal(size=(h, w, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
    pos = (rng.random((h, w, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
    col = rng.random((h, w, 3), dtype=np.float32)
    alb = np.full((h, w, 3), 0.5, np.float32)
    return alb, nrm, pos, col


@pytest.mark.parametrize("backend", BACKENDS)
def test_k2_matches_fp64_least_squares(backend):
    """Well-conditioned random blocks: weights == fp64 lstsq of the scaled + noised feature matrix
    built from K1's tmp_data, mins_maxs and the noise tile (bmfr.cl:511-542,623-627)."""
    _skip_if_missing(backend)
    w, h = 96, 64
    rng = np.random.default_rng(3)
    alb, nrm, pos, col = _random_frame(w, h, rng)
    ref = Runner("port", w, h)                       # tmp_data of K1 (bit-identical on every backend)
    ref.frame(0, alb, nrm, pos, col, None, [0.5, 0.5])
    tmp = ref.get("tmp_data").reshape(-1, 13, 1024).astype(np.float64)
    ref.close()
    r = Runner(backend, w, h)
    r.frame(0, alb, nrm, pos, col, None, [0.5, 0.5])
    weights, mm, noise = r.get("weights"), r.get("mins_maxs"), r.get("noise_tile")
    filt = r.get("filtered") if backend != "cuda-fused" else None
    r.close()
    for g in range(tmp.shape[0]):
        A = tmp[g, :10].copy()
        for k in range(6):
            lo, hi = tmp[g, 4 + k].min(), tmp[g, 4 + k].max()
            assert mm[g, k, 0] == np.float32(lo) and mm[g, k, 1] == np.float32(hi)
            f32 = tmp[g, 4 + k].astype(np.float32)
            d = np.float32(hi) - np.float32(lo)
            A[4 + k] = ((f32 - np.float32(lo)) / d if abs(d) > 1 else f32 - np.float32(lo)).astype(np.float64)
        A[1:10] = (A[1:10] + noise).astype(np.float32)
        x, *_ = np.linalg.lstsq(A.T, tmp[g, 10:13].T, rcond=None)
        err = np.abs(weights[g] - x).max() / max(np.abs(x).max(), 1e-6)
        assert err < 1e-4, (g, err)
    if filt is not None:
        assert np.isfinite(filt).all() and filt.min() >= 0     # negative clamp, bmfr.cl:750


@pytest.mark.parametrize("backend", BACKENDS)
def test_k2_constant_colour_block(backend):
    """A constant colour is reproduced by feature 0 (the constant 1) alone: w[0] = colour, rest ~ 0,
    R_00 = 32."""
    _skip_if_missing(backend)
    w, h = 96, 64
    rng = np.random.default_rng(4)
    alb, nrm, pos, _ = _random_frame(w, h, rng)
    col = np.empty((h, w, 3), np.float32)
    col[...] = (0.25, 0.5, 0.75)
    r = Runner(backend, w, h)
    r.frame(0, alb, nrm, pos, col, None, [0.5, 0.5])
    wts = r.get("weights")
    assert np.abs(wts[:, 0] - np.array([0.25, 0.5, 0.75])).max() < 1e-4
    assert np.abs(wts[:, 1:]).max() < 1e-4
    acc = r.get("accum")
    assert np.abs(acc - col).max() < 1e-4
    r.close()


# ---------------------------------------------------------------- K4 / K5
@pytest.mark.parametrize("backend", BACKENDS)
def test_k4_k5_constant_image_is_a_fixed_point(backend):
    """Constant colour + static camera: accumulated colour stays the constant (alpha floors at
    SECOND_BLEND_ALPHA, bmfr.cl:838-839), tone map = clamp(pow(albedo*c, 0.454545)) with end points
    0 and 1 (bmfr.cl:852-856), and taa leaves a constant image unchanged (bmfr.cl:967-973)."""
    _skip_if_missing(backend)
    w, h = 64, 64
    pos, nrm, M, off = _static_plane(w, h)
    col = np.empty((h, w, 3), np.float32)
    col[...] = (0.0, 0.5, 4.0)
    alb = np.ones((h, w, 3), np.float32)
    r = Runner(backend, w, h, position_limit_squared=4.0, normal_limit_squared=0.1)
    for f in range(14):
        r.frame(f, alb, nrm, pos, col, M, off)
    acc, res = r.get("accum"), r.get("result")
    assert np.abs(acc - col).max() < 2e-4
    expect = np.clip(np.power(np.maximum(col.astype(np.float64), 0), 0.454545), 0, 1)
    assert np.abs(res - expect).max() < 2e-4
    assert res[..., 0].max() < 1e-3 and res[..., 2].min() > 0.9999 and res[..., 2].max() <= 1.0
    r.close()
