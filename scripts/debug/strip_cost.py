"""GPU box (1 GPU): what a strip costs per kernel against a whole image of the same pixel count."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from bmfr_b200 import Denoiser, sharding, synth

W, H, frames, halo = 3840, 2160, 12, 48
if len(sys.argv) > 2 and sys.argv[2] == "8k":
    W, H, halo = 7680, 4320, 60
def inputs(w, h, y0, y1):
    t = torch.empty((frames, 4, y1 - y0, w, 3), dtype=torch.float32, device="cuda")
    for f in range(frames):
        synth.frame_device(w, h, f, [t[f, k].data_ptr() for k in range(4)], y0=y0, y1=y1)
    torch.cuda.synchronize()
    return t
def cams(w, h):
    return [synth.camera(max(f - 1, 0), w, h)[0] for f in range(frames)], [synth.camera(f, w, h)[1] for f in range(frames)]
def report(tag, d):
    ms = np.array([[d.fused_kernel_ms(f)[k] for k in d.fused_kernels] for f in range(2, frames)]).mean(axis=0)
    print(f"{tag:52s} " + "  ".join(f"{k} {1e3 * v:7.1f} us" for k, v in zip(d.fused_kernels, ms)), flush=True)

# (a) whole image with the pixel count of one strip
w, h = (3840, 1088) if W == 3840 else (7680, 544)
t = inputs(w, h, 0, h); cm, of = cams(w, h)
with Denoiser(w, h, mode="fused", profile=True) as d:
    for f in range(frames):
        d.denoise_frame(f, *[t[f, k].data_ptr() for k in range(4)], cm[f], of[f], 0)
    d.sync(); report(f"whole {w}x{h}", d)
del t
# (b) the upper strip of 3840x2160, not connected (no halo duties; halo rows go stale, timing only)
cm, of = cams(W, H)
for strip in (((0, 1088), (1088, 2160)) if W == 3840 else ((0, 544), (1632, 2176))):
    d = Denoiser(W, H, mode="fused", profile=True, strip=strip, halo_rows=halo)
    g = d.geometry
    t = inputs(W, H, g.row0, g.row1)
    for f in range(frames):
        d.denoise_frame(f, *[t[f, k].data_ptr() for k in range(4)], cm[f], of[f], 0)
    try:
        d.sync()
    except Exception as e:
        print("  (sync:", str(e)[:60], ")")
    report(f"strip {strip} unconnected, rows {g.row1 - g.row0}", d); d.close(); del t
if len(sys.argv) > 1 and sys.argv[1] == "unconnected":
    sys.exit(0)
# (c) two connected strips on this one device (same-device wait kernels + in-kernel pushes)
full = inputs(W, H, 0, H)
ss = sharding.LocalStripSet(W, H, 2, halo=halo, exchange="p2p", profile=True)
out = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
for f in range(frames):
    ss.denoise_frame(f, [full[f, k] for k in range(4)], cm[f], of[f], out)
ss.sync()
for c in ss.ctx:
    report(f"strip {c.strip} connected locally", c.d)
