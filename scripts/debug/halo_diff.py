"""GPU box: which rows of which state buffer of a strip context differ from the whole-image run, frame by frame."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from bmfr_b200 import Denoiser, sharding, synth

w, h, frames, halo, n = 320, 384, 3, 40, 2
overlap = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev = torch.device("cuda:0")
seq = [[torch.from_numpy(x).to(dev) for x in synth.frame_host(w, h, f)] for f in range(frames)]
cams = [synth.camera(max(f - 1, 0), w, h)[0] for f in range(frames)]
offs = [synth.camera(f, w, h)[1] for f in range(frames)]
whole = Denoiser(w, h, mode="fused")
ss = sharding.LocalStripSet(w, h, n, halo=halo, exchange="p2p", overlap_frames=overlap)
out_s = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
out_w = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
for f in range(frames):
    whole.denoise_frame(f, *[t.data_ptr() for t in seq[f]], cams[f], offs[f], out_w.data_ptr())
    ss.denoise_frame(f, seq[f], cams[f], offs[f], out_s)
    whole.sync(); ss.sync(); torch.cuda.synchronize()
    print(f"frame {f}: output equal {bool(torch.equal(out_s, out_w))}")
    for name in ("noisy_acc", "spp", "accept", "prev_pixels", "weights", "mins_maxs", "accum", "result"):
        ref = whole.read(name)
        for c in ss.ctx:
            got = c.d.read(name)
            if name in ("weights", "mins_maxs"):
                bad = np.argwhere((got != ref).reshape(got.shape[0], -1).any(axis=1)).ravel()
                bx = c.d.geometry.blocks_x
                print(f"  {name:12s} strip {c.strip}: blocks differing {len(bad)} rows {sorted(set((bad // bx).tolist()))[:12]}")
                continue
            r = ref[c.row0:c.row1]
            diff = (got.reshape(got.shape[0], -1) != r.reshape(r.shape[0], -1)).any(axis=1)
            rows = np.argwhere(diff).ravel() + c.row0
            own = [y for y in rows if c.strip[0] <= y < c.strip[1]]
            print(f"  {name:12s} strip {c.strip} rows [{c.row0},{c.row1}): differing rows {len(rows)} (own {len(own)}) first {rows[:6].tolist()} last {rows[-6:].tolist()}")
