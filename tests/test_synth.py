"""synth-v1 generator: determinism, the camera convention of bmfr.cl:343-355, host/device twins."""
import numpy as np
import pytest

from bmfr_b200 import synth


def test_deterministic_and_sane():
    a = synth.frame_host(160, 96, 3)
    b = synth.frame_host(160, 96, 3)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    albedo, normal, position, noisy = a
    assert np.isfinite(position).all() and np.abs(position).max() < 256      # bmfr.cpp:86-88
    assert albedo.min() >= 0.2 and albedo.max() <= 0.9
    nn = np.linalg.norm(normal, axis=2)
    assert np.all((np.abs(nn - 1) < 1e-4) | (nn == 0))                        # unit normals, 0 for sky
    assert noisy.min() >= 0
    c = synth.frame_host(160, 96, 4)
    assert not np.array_equal(c[3], noisy)                                     # noise depends on the frame


def test_rows_subset_matches_full_frame():
    full = synth.frame_host(96, 80, 2)
    part = synth.frame_host(96, 80, 2, y0=17, y1=49)
    for f, p in zip(full, part):
        assert np.array_equal(f[17:49], p)


@pytest.mark.parametrize("size", [(160, 96), (1920, 1080)])
def test_camera_matrix_reprojects_to_own_pixel(size):
    """clip_j = sum_i p_i M[i][j]; uv = (ndc+1)/2; pixel = uv*(W,H) - (off.x, 1-off.y)  (bmfr.cl:343-355).
    A point seen at pixel (x,y) of frame f must reproject to (x,y) with frame f's own matrix."""
    w, h = size
    f = 5
    _, _, pos, _ = synth.frame_host(w, h, f, y0=h // 2, y1=h // 2 + 2)
    M, off = synth.camera(f, w, h)
    M = M.reshape(4, 4).astype(np.float64)
    p = np.concatenate([pos.astype(np.float64), np.ones(pos.shape[:2] + (1,))], axis=2)
    clip = p @ M
    uv = (clip[..., :2] / clip[..., 3:4] + 1) / 2
    px = uv[..., 0] * w - off[0]
    py = uv[..., 1] * h - (1 - off[1])
    xs = np.arange(w)[None, :]
    ys = np.arange(h // 2, h // 2 + 2)[:, None]
    assert np.abs(px - xs).max() < 2e-3 * w / 160 and np.abs(py - ys).max() < 2e-3 * w / 160
    assert clip[..., 3].min() > 0.5                                           # w stays away from 0


def test_camera_motion_is_a_few_pixels():
    w, h = 1920, 1080
    _, _, pos, _ = synth.frame_host(w, h, 10, y0=500, y1=501)
    M, off = synth.camera(9, w, h)
    p = np.concatenate([pos[0].astype(np.float64), np.ones((w, 1))], axis=1) @ M.reshape(4, 4).astype(np.float64)
    px = (p[:, 0] / p[:, 3] + 1) / 2 * w - off[0]
    motion = np.abs(px - np.arange(w))
    assert 0.5 < np.median(motion) < 12


@pytest.mark.gpu
def test_device_twin_is_bit_identical():
    import torch
    w, h, f = 416, 250, 7
    host = synth.frame_host(w, h, f)
    dev = torch.empty((4, h, w, 3), dtype=torch.float32, device="cuda")
    synth.frame_device(w, h, f, [dev[k].data_ptr() for k in range(4)], stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    d = dev.cpu().numpy()
    for k in range(4):
        assert np.array_equal(host[k].view(np.uint32), d[k].view(np.uint32)), f"image {k}"
