#!/usr/bin/env python
"""Per-CTA timeline of fit_gram_kernel (tuning build -DBMFR_QR_TIMING, BMFR_B200_LIB=...): where a CTA's time goes."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bmfr_b200 import Denoiser, _lib, synth  # noqa: E402

w, h, frames = 1920, 1080, 8
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
inputs = torch.empty((frames, 4, h, w, 3), dtype=torch.float32, device="cuda")
for f in range(frames):
    synth.frame_device(w, h, f, [inputs[f, k].data_ptr() for k in range(4)], stream=stream.cuda_stream)
d = Denoiser(w, h, mode="fused", stream=stream.cuda_stream, profile=True)
for f in range(frames):
    d.denoise_frame(f, *[inputs[f, k].data_ptr() for k in range(4)], synth.camera(max(f - 1, 0), w, h)[0], synth.camera(f, w, h)[1], 0)
d.sync()
print("kernel ms (events):", d.fused_kernel_ms(frames - 1))
lib = _lib.load()
n = 1024 * 16
buf = (C.c_longlong * n)()
fn = lib.bmfr_debug_gram_cta
fn.restype, fn.argtypes = C.c_int, [C.POINTER(C.c_longlong), C.c_int]
assert fn(buf, n) == 0
t = np.array(buf[:], dtype=np.int64).reshape(1024, 16)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
us = lambda a: (a - t0) / 1e3
print(f"CTAs {len(t)}; start spread {us(t[:,0]).max():.1f} us; level-1 end median {np.median(us(t[:,14])):.1f} max {us(t[:,14]).max():.1f}; solves end median {np.median(us(t[:,15])):.1f} max {us(t[:,15]).max():.1f} us")
first_wait = (t[:, 2] - t[:, 0]) / 1e3
print(f"first block: data ready after median {np.median(first_wait):.2f} us (p90 {np.quantile(first_wait, 0.9):.2f})")
for it in range(6):
    ok = t[:, 3 + 2 * it] > 0
    if not ok.any():
        break
    comp = (t[ok, 3 + 2 * it] - t[ok, 2 + 2 * it]) / 1e3
    line = f"block {it}: {ok.sum():4d} CTAs, compute median {np.median(comp):.2f} us (p90 {np.quantile(comp, 0.9):.2f})"
    if it > 0:
        ok2 = ok & (t[:, 2 + 2 * it] > 0)
        wait = (t[ok2, 2 + 2 * it] - t[ok2, 1 + 2 * it]) / 1e3
        line += f", wait for its tiles median {np.median(wait):.2f} (p90 {np.quantile(wait, 0.9):.2f})"
    print(line)
sol = (t[:, 15] - t[:, 14]) / 1e3
print(f"solves: median {np.median(sol):.2f} us (p90 {np.quantile(sol, 0.9):.2f})")
nb = ((t[:, 3::2][:, :6] > 0).sum(axis=1))
print("blocks per CTA (capped at 6):", np.bincount(nb))
