"""GPU box (1 GPU): frames/s of a whole image against an unconnected strip of the same pixel count (overlap_frames = 1),
i.e. what the strip geometry alone costs (halo rows reprojected twice, partial tiles at the strip edges) — what is left
of the N-GPU figure is the exchange.  Timing only: the strip's halo rows go stale."""
import sys
import time
import torch
sys.path.insert(0, ".")
from bmfr_b200 import Denoiser, sharding, synth

W, H = 7680, 4320
n_strips = int(sys.argv[1]) if len(sys.argv) > 1 else 8
frames, reps = 8, 12
halo = sharding.default_halo(H)
rows = H // n_strips

def inputs(w, h, y0, y1):
    t = torch.empty((frames, 4, y1 - y0, w, 3), dtype=torch.float32, device="cuda")
    for f in range(frames):
        synth.frame_device(w, h, f, [t[f, k].data_ptr() for k in range(4)], y0=y0, y1=y1)
    torch.cuda.synchronize()
    return t

def run(tag, d, t, cm, of):
    for overlap_note in (0,):
        n = frames * reps
        for f in range(frames):  # warm-up
            d.denoise_frame(f, *[t[f % frames, k].data_ptr() for k in range(4)], cm[f % frames], of[f % frames], 0)
        try:
            d.sync()
        except Exception as e:
            print("  (sync:", str(e)[:60], ")")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        t0 = time.perf_counter()
        for f in range(frames, frames + n):
            d.denoise_frame(f, *[t[f % frames, k].data_ptr() for k in range(4)], cm[f % frames], of[f % frames], 0)
        try:
            d.sync()
        except Exception as e:
            print("  (sync:", str(e)[:60], ")")
        dt = time.perf_counter() - t0
        print(f"{tag:48s} {1e6 * dt / n:7.1f} us/frame", flush=True)

for overlap in (1, 0):
    w, h = W, (rows + 31) // 32 * 32
    t = inputs(w, h, 0, h)
    cm = [synth.camera(max(f - 1, 0), w, h)[0] for f in range(frames)]; of = [synth.camera(f, w, h)[1] for f in range(frames)]
    with Denoiser(w, h, mode="fused", overlap_frames=overlap) as d:
        run(f"whole {w}x{h} overlap={overlap}", d, t, cm, of)
    del t
    cm = [synth.camera(max(f - 1, 0), W, H)[0] for f in range(frames)]; of = [synth.camera(f, W, H)[1] for f in range(frames)]
    for strip in ((0, rows), (3 * rows, 4 * rows)):
        d = Denoiser(W, H, mode="fused", overlap_frames=overlap, strip=strip, halo_rows=halo)
        g = d.geometry
        t = inputs(W, H, g.row0, g.row1)
        run(f"strip {strip} unconnected, rows held {g.row1 - g.row0} overlap={overlap}", d, t, cm, of)
        d.close(); del t
