"""Shared helpers of the parity tests: the synthetic sequence, the oracle run and the CUDA run."""
from __future__ import annotations

import numpy as np

from bmfr_b200 import Denoiser, synth
from oracle.oracle import Oracle

ALL_BUFFERS = ("noisy_acc", "spp", "prev_pixels", "accept", "weights", "mins_maxs", "accum", "result")
STAGED_ONLY = ("tmp_data", "filtered", "tone_mapped")
INTEGER_BUFFERS = ("spp", "accept")          # bit-exact by the north star
EXACT_FLOAT_BUFFERS = ("noisy_acc", "prev_pixels", "mins_maxs")   # bit-exact by construction
COLOUR_BUFFERS = ("accum", "result")         # within tolerance
# per-pixel relative tolerance of the north star, with the floor it leaves open: |a-b| <= REL * max(|b|, EPS)
REL, EPS, PSNR_DB = 1e-3, 1e-2, 60.0


def sequence(width, height, frames, jitter=False, seed=synth.SEED):
    """Yields (frame, albedo, normal, position, noisy, cam_prev, pixel_offset) like bmfr.cpp:417-445."""
    for f in range(frames):
        a, n, p, c = synth.frame_host(width, height, f, seed=seed)
        cam_prev, _ = synth.camera(max(f - 1, 0), width, height, jitter)
        _, off = synth.camera(f, width, height, jitter)   # matrix f-1 is paired with offset f (bmfr.cpp:440-444)
        yield f, a, n, p, c, cam_prev, off


def run_oracle(kind, width, height, frames, *, keep=("result",), jitter=False, every_frame=True, **kw):
    pl, nl = synth.limits()
    o = Oracle(kind, width, height, position_limit_squared=kw.pop("position_limit_squared", pl),
               normal_limit_squared=kw.pop("normal_limit_squared", nl), keep_tmp=int("tmp_data" in keep), **kw)
    out = []
    for f, a, n, p, c, cam, off in sequence(width, height, frames, jitter):
        o.frame(f, a, n, p, c, cam, off)
        if every_frame or f == frames - 1:
            out.append({k: o.buffer(k) for k in keep})
    o.close()
    return out


def run_cuda(width, height, frames, *, mode="fused", keep=("result",), jitter=False, every_frame=True, **kw):
    """Runs the CUDA path through the C ABI's host-pointer entry."""
    out = []
    with Denoiser(width, height, mode=mode, **kw) as d:
        for f, a, n, p, c, cam, off in sequence(width, height, frames, jitter):
            d.denoise_frame_host(f, a, n, p, c, cam, off)
            if every_frame or f == frames - 1:
                out.append({k: d.read(k) for k in keep})
        d.sync()
    return out


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def floats_equal_mod_zero_sign(a, b):
    """Equal as IEEE values (+0 == -0), NaNs in the same places."""
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def colour_error(a, b):
    """(max per-pixel relative error with floor EPS, PSNR in dB against a peak of max|b|)."""
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    diff = np.abs(a64 - b64)
    rel = float(np.max(diff / np.maximum(np.abs(b64), EPS)))
    mse = float(np.mean(diff ** 2))
    peak = max(float(np.max(np.abs(b64))), 1e-12)
    psnr = float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)
    return rel, psnr


def assert_colour_close(a, b, what):
    assert not np.isnan(a).any() or np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN pattern differs"
    m = ~np.isnan(b)
    rel, psnr = colour_error(np.where(m, a, 0), np.where(m, b, 0))
    assert rel <= REL, f"{what}: max relative error {rel:.3e} > {REL} (eps floor {EPS})"
    assert psnr >= PSNR_DB, f"{what}: PSNR {psnr:.1f} dB < {PSNR_DB}"
    return rel, psnr


# ------------------------------------------------------------------------------------------------
# one interface over the four implementations, for tests that feed hand-made inputs
# ------------------------------------------------------------------------------------------------
BACKENDS_CPU = ("port", "reference")
BACKENDS_GPU = ("cuda-staged", "cuda-fused")


class Runner:
    """backend: "port" | "reference" (CPU checkers) | "cuda-staged" | "cuda-fused" (the product)."""

    def __init__(self, backend, width, height, **params):
        self.backend, self.W, self.H = backend, width, height
        pl, nl = synth.limits()
        params.setdefault("position_limit_squared", pl)
        params.setdefault("normal_limit_squared", nl)
        if backend.startswith("cuda"):
            self.impl = Denoiser(width, height, mode=backend.split("-")[1], **params)
        else:
            self.impl = Oracle(backend, width, height, keep_tmp=1, **params)

    def frame(self, f, albedo, normal, position, noisy, cam_prev, pixel_offset):
        if self.backend.startswith("cuda"):
            self.impl.denoise_frame_host(f, albedo, normal, position, noisy, cam_prev, pixel_offset)
        else:
            self.impl.frame(f, albedo, normal, position, noisy, cam_prev, pixel_offset)

    def get(self, name):
        return self.impl.read(name) if self.backend.startswith("cuda") else self.impl.buffer(name)

    def close(self):
        self.impl.close()


def backend_params(gpu_too=True):
    import pytest
    ps = [pytest.param(b, id=b) for b in BACKENDS_CPU]
    if gpu_too:
        ps += [pytest.param(b, id=b, marks=pytest.mark.gpu) for b in BACKENDS_GPU]
    return ps
