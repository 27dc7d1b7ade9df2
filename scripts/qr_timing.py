#!/usr/bin/env python
"""Phase timers of fit_qr_kernel (tuning build with -DBMFR_QR_TIMING): runs a few 1080p frames and prints
the clock64 deltas of CTA 0's first thread per block iteration, and per CTA the end of level 1 (its last
block) and of level 2 (the solves of its blocks)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bmfr_b200 import Denoiser, _lib, synth  # noqa: E402

w, h, frames = 1920, 1080, 6
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
inputs = torch.empty((frames, 4, h, w, 3), dtype=torch.float32, device="cuda")
for f in range(frames):
    synth.frame_device(w, h, f, [inputs[f, k].data_ptr() for k in range(4)], stream=stream.cuda_stream)
out = torch.empty((h, w, 3), dtype=torch.float32, device="cuda")
d = Denoiser(w, h, mode="fused", stream=stream.cuda_stream)
for f in range(frames):
    cam = synth.camera(max(f - 1, 0), w, h)[0]
    off = synth.camera(f, w, h)[1]
    d.denoise_frame(f, *[inputs[f, k].data_ptr() for k in range(4)], cam, off, out.data_ptr())
d.sync()
lib = _lib.load()
buf = (C.c_longlong * 512)()
fn = lib.bmfr_debug_qr_timing
fn.restype, fn.argtypes = C.c_int, [C.POINTER(C.c_longlong), C.c_int]
assert fn(buf, 512) == 0
t = np.array(buf[:], dtype=np.int64)
comp = t[:64].reshape(8, 8)
t0 = comp[0, 0]
names = ["lds", "minmax", "barrier+draw+finish", "scale+noise", "-", "level-1 QR"]
print("compute warp 0 of CTA 0 (cycles):")
for it in range(8):
    if comp[it, 0] == 0:
        break
    deltas = np.diff(comp[it, :7])
    print(f"  block {it}: start +{comp[it, 0] - t0:7d}  " + "  ".join(f"{n} {v}" for n, v in zip(names, deltas)) + f"  total {comp[it, 6] - comp[it, 0]}")
cta = (C.c_longlong * 4096)()
fn2 = lib.bmfr_debug_qr_cta
fn2.restype, fn2.argtypes = C.c_int, [C.POINTER(C.c_longlong), C.c_int]
assert fn2(cta, 4096) == 0
c = np.array(cta[:], dtype=np.int64).reshape(1024, 4)
ncta = int((c[:, 0] != 0).sum())
c = c[:ncta]
g0 = c[:, 0].min()
start, cend, send, sm = c[:, 0] - g0, c[:, 1] - g0, c[:, 2] - g0, c[:, 3]
print(f"{ncta} CTAs; per CTA (ns from first start): start  min {start.min()} max {start.max()};  level-1 end  min {cend.min()} median {int(np.median(cend))} max {cend.max()};"
      f"  level-2 end  min {send.min()} median {int(np.median(send))} max {send.max()}")
print("  level-1 end histogram (2 us bins from 28):", np.histogram(cend, bins=np.arange(28000, 60000, 2000))[0].tolist())
print("  level-2 end histogram (2 us bins from 28):", np.histogram(send, bins=np.arange(28000, 60000, 2000))[0].tolist())
order = np.argsort(send)[-5:]
print("  slowest CTAs:", [(int(i), int(sm[i]), int(start[i]), int(cend[i]), int(send[i])) for i in order], "(cta, sm, start, level-1 end, level-2 end)")
per_sm = {}
for i in range(ncta):
    per_sm.setdefault(int(sm[i]), []).append(int(send[i]))
ends = np.array([max(v) for v in per_sm.values()])
print(f"  SMs used {len(per_sm)}, CTAs per SM {sorted(set(len(v) for v in per_sm.values()))}, per-SM last end: min {ends.min()} median {int(np.median(ends))} max {ends.max()}")
d.close()
