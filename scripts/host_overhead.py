#!/usr/bin/env python
"""Host cost of bmfr_denoise_frame (device pointers): wall time of submitting frames without waiting for the GPU,
against the GPU time of the same frames.  The GPU must be the bound (host per frame well below device per frame)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bmfr_b200 import Denoiser, synth  # noqa: E402

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
frames = 60
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
inputs = torch.empty((4, 4, h, w, 3), dtype=torch.float32, device="cuda")
for f in range(4):
    synth.frame_device(w, h, f, [inputs[f, k].data_ptr() for k in range(4)], stream=stream.cuda_stream)
cams = [(synth.camera(max(f - 1, 0), w, h)[0], synth.camera(f, w, h)[1]) for f in range(frames)]
d = Denoiser(w, h, mode="fused", stream=stream.cuda_stream)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f in range(frames):
        d.denoise_frame(f, *[inputs[f % 4, k].data_ptr() for k in range(4)], cams[f][0], cams[f][1], 0)
    t1 = time.perf_counter()
    d.sync()
    t2 = time.perf_counter()
    print(f"pass {rep}: submit {1e6 * (t1 - t0) / frames:.1f} us/frame (host, through ctypes), until done {1e6 * (t2 - t0) / frames:.1f} us/frame")
d.close()
