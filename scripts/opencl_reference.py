#!/usr/bin/env python
"""GPU box: the reference's unmodified bmfr.cl on the B200 through NVIDIA's OpenCL ICD
(oracle/_ref/libbmfr_clgpu.so), timed with OpenCL events in the reference's six-table format
(CLUtils.hpp:313-332, bmfr.cpp:488-517), beside this repository's CUDA path on the same frames, and
compared with it buffer by buffer.  Writes a JSON summary; a measured baseline, not a bench value.

    python scripts/opencl_reference.py [--width 1920 --height 1080 --frames 60 --half 0|1] --out FILE
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

LABELS = ["Accumulation of noisy data", "Fitting feature buffers to noisy data", "Weighted sum",
          "Accumulation of filtered data", "TAA", "Total time in all kernels (including intermediate launch overheads)"]


def table(label, t):
    t = np.asarray(t, dtype=np.float64)
    w = 9
    return (f"\n {label}\n {'-' * len(label)}\n   Mean   : {t.mean():{w}.3f} ms\n   Min    : {t.min():{w}.3f} ms\n"
            f"   Max    : {t.max():{w}.3f} ms\n   Total  : {t.sum():{w}.3f} ms\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--half", type=int, default=0, help="USE_HALF_PRECISION_IN_TMP_DATA (the reference ships 1; the north star asks for 0)")
    ap.add_argument("--compare", type=int, default=1, help="also run the CUDA path and compare every frame")
    ap.add_argument("--port-frames", type=int, default=0, help="also compare the first N frames with the CPU port oracle")
    ap.add_argument("--out", default="")
    a = ap.parse_args()

    from bmfr_b200 import Denoiser, synth
    from oracle.oracle import Oracle
    from tests import util

    w, h = a.width, a.height
    pl, nl = synth.limits()
    try:
        o = Oracle("opencl", w, h, position_limit_squared=pl, normal_limit_squared=nl, tmp_half=a.half)
    except Exception as e:  # a logged, specific reason
        msg = {"opencl": "unavailable", "why": str(e)}
        print(json.dumps(msg))
        if a.out:
            Path(a.out).write_text(json.dumps(msg, indent=1))
        return 0
    dev = o.device_name
    d = Denoiser(w, h, mode="fused") if a.compare else None
    port = Oracle("port", w, h, position_limit_squared=pl, normal_limit_squared=nl, tmp_half=a.half) if a.port_frames else None
    stage = []
    cmp_rows = []
    keep = ("spp", "accept", "prev_pixels", "noisy_acc", "mins_maxs", "weights", "accum", "result")
    for f, alb, nrm, pos, col, cam, off in util.sequence(w, h, a.frames):
        o.frame(f, alb, nrm, pos, col, cam, off)
        stage.append(o.stage_ms())
        if d is None:
            continue
        d.denoise_frame_host(f, alb, nrm, pos, col, cam, off)
        row = {"frame": f}
        want = f in (0, 1, 2, 5, 17, a.frames // 2, a.frames - 1)
        if want:
            for k in keep:
                r, c = o.buffer(k), d.read(k)
                if k in ("spp", "accept"):
                    row[k + "_mismatch_pixels"] = int((r != c).sum())
                elif k in ("prev_pixels", "noisy_acc", "mins_maxs"):
                    row[k + "_bit_identical"] = bool(util.bits_equal(r, c))
                    row[k + "_mismatch_elems"] = int((r.view(np.uint32) != c.view(np.uint32)).sum())
                elif k == "weights":
                    row["weights_max_abs_diff"] = float(np.nanmax(np.abs(r - c)))
                else:
                    m = ~np.isnan(r)
                    rel, psnr = util.colour_error(np.where(m, c, 0), np.where(m, r, 0))
                    row[k + "_max_rel"], row[k + "_psnr_db"] = rel, psnr
            cmp_rows.append(row)
        if port is not None and f < a.port_frames:
            port.frame(f, alb, nrm, pos, col, cam, off)
            prow = {"frame": f, "vs": "port"}
            for k in keep:
                r, c = o.buffer(k), port.buffer(k)
                prow[k + "_mismatch_elems"] = int((np.ascontiguousarray(r).view(np.uint8) != np.ascontiguousarray(c).view(np.uint8)).sum())
                if k in ("accum", "result"):
                    prow[k + "_max_rel"], prow[k + "_psnr_db"] = util.colour_error(r, c)
            cmp_rows.append(prow)
    st = np.array(stage)  # [frames, 6]
    # frame-0 exclusion of bmfr.cpp:392-397,488-506: K1, K4, K5 and the total are profiled for frames 1..N-1
    txt = f"Using device named: {dev}\n"
    sel = {0: slice(1, None), 1: slice(0, None), 2: slice(0, None), 3: slice(1, None), 4: slice(1, None), 5: slice(1, None)}
    means = {}
    for i, label in enumerate(LABELS):
        t = st[sel[i], i]
        txt += table(label, t)
        means[label] = {"mean_ms": float(t.mean()), "min_ms": float(t.min()), "max_ms": float(t.max())}
    print(txt)
    total = float(st[1:, 5].mean())
    summary = {
        "opencl": "ok", "device": dev, "workload": f"{w}x{h} x{a.frames} frames synth-v1", "tmp_half": a.half,
        "stages": means, "frames_per_s_from_total": 1e3 / total,
        "kernel_sum_ms": float(st[1:, :5].sum(axis=1).mean()),
        "compare_with_cuda_fused": cmp_rows,
        "note": "OpenCL profiling events of the five reference kernels; uploads and read-back outside, as bmfr.cpp:415-416,478",
    }
    print(json.dumps(summary))
    if a.out:
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        Path(a.out).write_text(json.dumps(summary, indent=1))
        Path(a.out).with_suffix(".txt").write_text(txt)
    return 0


if __name__ == "__main__":
    sys.exit(main())
