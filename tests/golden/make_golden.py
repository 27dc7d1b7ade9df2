"""Generates tests/golden/bmfr_ref_128x72.npz from the REFERENCE KERNELS: the reference's own
/root/reference/opencl/bmfr.cl executed through oracle/cl_shim (oracle/_ref/libbmfr_clref.so) on the
synth-v1 sequence.  Run in the build container, where /root/reference is mounted:

    python tests/golden/make_golden.py

The fixture holds, for an 8-frame 128x72 jittered sequence: sha256 of every input image and of every
buffer of the host loop after every frame, plus the full buffers of frames 1 and 7.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from bmfr_b200 import synth  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

W, H, FRAMES, FULL = 128, 72, 8, (1, 7)
BUFFERS = ("noisy_acc", "spp", "prev_pixels", "accept", "tmp_data", "weights", "mins_maxs", "filtered", "accum",
           "tone_mapped", "result")
OUT = Path(__file__).with_name("bmfr_ref_128x72.npz")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main(kind="reference"):
    pl, nl = synth.limits()
    o = Oracle(kind, W, H, position_limit_squared=pl, normal_limit_squared=nl, keep_tmp=1)
    meta = dict(kind=kind, width=W, height=H, frames=FRAMES, jitter=True, seed=synth.SEED, position_limit_squared=pl,
                normal_limit_squared=nl, noise_amount=1e-2, blend_alpha=0.2, second_blend_alpha=0.1, taa_blend_alpha=0.2,
                tmp_half=0, k1_schedule=0, input_sha256=[], buffer_sha256=[])
    arrays = {}
    for f in range(FRAMES):
        a, n, p, c = synth.frame_host(W, H, f)
        cam, _ = synth.camera(max(f - 1, 0), W, H, True)
        _, off = synth.camera(f, W, H, True)
        meta["input_sha256"].append([sha(x) for x in (a, n, p, c, cam, off)])
        o.frame(f, a, n, p, c, cam, off)
        bufs = {k: o.buffer(k) for k in BUFFERS}
        meta["buffer_sha256"].append({k: sha(v) for k, v in bufs.items()})
        if f in FULL:
            for k, v in bufs.items():
                if k not in ("tmp_data", "filtered", "tone_mapped"):
                    arrays[f"f{f}_{k}"] = v
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print(f"wrote {OUT} ({OUT.stat().st_size / 1024:.0f} KiB) from kind={kind}")


if __name__ == "__main__":
    main()
