#!/usr/bin/env python
"""bench.py — frames/s of the BMFR denoise path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path over one batch = the 60-frame synthetic sequence (synth-v1) of
the workload.  N=1: 1920x1080 (the headline single-GPU config).  N>1 (torchrun, one rank per GPU): the
multi-GPU configs of BASELINE.json, strip-sharded along block rows with peer-to-peer halo pushes over
NVLink: N=2 and N=4 -> 3840x2160, N=8 -> 7680x4320.

  value  : frames/s, inputs resident in HBM, device-timed (CUDA events on the launching stream)
  e2e    : frames/s through the C ABI's host-pointer entry: pinned-host uploads + read-back inside
  roofline / kernels : per-kernel algorithmic bytes / measured launch duration
  cpu_baseline : the CPU oracle on the box's host cores, bounded sample (reported, not a target)

`--impl reference` times the reference's CPU implementation of the path on the host cores (the plain-C port
of bmfr.cl, pinned bit for bit against the reference's own kernels; --ref-shim also times those through the
OpenCL-C shim of oracle/_ref), all host threads, every step a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FRAMES = 60
FUSED_KERNELS = ("reproject_kernel", "fit_gram_kernel", "post_tma_kernel")  # replaced by the context's own list at run time
SINGLE_GPU_WORKLOAD = (1920, 1080)


def sharded_workload(n_gpus):
    """BASELINE.json's multi-GPU configs: 3840x2160 over 2 and 4 GPUs, 7680x4320 over 8 (other N: a 3840x540 strip
    per rank).  `value` stays in 1080p-equivalent frames/s so that the series is comparable across N; other
    pairings (4K over 8, the old 3840x(540 N) weak series) run with --width/--height."""
    return {2: (3840, 2160), 4: (3840, 2160), 8: (7680, 4320)}.get(n_gpus, (3840, 540 * n_gpus))


def workload_config(w, h, n_gpus):
    """The `config` object of the JSON line — identical for the B200 arm and the reference arm."""
    return {"workload": f"{w}x{h} x{FRAMES} frames synth-v1, 32x32 blocks, 10 features, fp32 tmp_data",
            "value_unit": "1080p-equivalent frames/s = native frames/s x (W*H)/(1920*1080) (= frames/s at 1920x1080)",
            "l2": "inputs of a sequence (5.97 GB at 1080p) are larger than L2; no explicit flush",
            "parallelism": "single GPU" if n_gpus == 1 else f"strips{n_gpus} (block rows)"}


def geometry(w, h):
    ww, hw = 32 * ((w + 31) // 32), 32 * ((h + 31) // 32)
    wm, hm = ww + 32, hw + 32
    return dict(P=w * h, M=wm * hm, NB=(wm // 32) * (hm // 32))


def algorithmic_bytes(w, h):
    """Bytes per frame each kernel must move (fp32 tmp_data, frames >= 1)."""
    g = geometry(w, h)
    P, M, NB = g["P"], g["M"], g["NB"]
    return {
        "accumulate_noisy_data": 36 * P + 37 * P + 22 * P + 13 * 4 * M,
        "fitter": 13 * 4 * M + 168 * NB,
        "weighted_sum": 24 * P + 168 * NB + 12 * P,
        "accumulate_filtered_data": 46 * P + 24 * P,
        "taa": 32 * P + 12 * P,
        # FUSED: tmp_data / filtered / tone_mapped never reach HBM
        "reproject_kernel": 95 * P,                        # K1 per image pixel: 73 B in, 22 B out
        "reproject_tma_kernel": 95 * P,
        "fit_qr_kernel": 36 * P + 216 * NB,                # normals, positions, accumulated colour once; weights + min/max out
        "fit_gram_kernel": 36 * P + 216 * NB,
        "post_kernel": 94 * P + 168 * NB,                  # K3+K4+K5
        "post_tma_kernel": 94 * P + 168 * NB,
    }


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons while the timed region runs (B200_PROFILING.md), polled through NVML
    every 2 ms from a helper thread (nvidia-smi -lms 100 as the fallback when NVML is not importable)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.stop, self.thread, self.source = index, [], None, False, None, None

    def _nvml_loop(self, nv, handle):
        bits = [(getattr(nv, n, 0), name) for n, name in (
            ("nvmlClocksEventReasonHwSlowdown", "hw_slowdown"), ("nvmlClocksEventReasonHwThermalSlowdown", "hw_thermal_slowdown"),
            ("nvmlClocksEventReasonSwThermalSlowdown", "sw_thermal_slowdown"), ("nvmlClocksEventReasonSwPowerCap", "sw_power_cap"))]
        if not any(b for b, _ in bits):  # older binding names
            bits = [(getattr(nv, n, 0), name) for n, name in (
                ("nvmlClocksThrottleReasonHwSlowdown", "hw_slowdown"), ("nvmlClocksThrottleReasonHwThermalSlowdown", "hw_thermal_slowdown"),
                ("nvmlClocksThrottleReasonSwThermalSlowdown", "sw_thermal_slowdown"), ("nvmlClocksThrottleReasonSwPowerCap", "sw_power_cap"))]
        mx = nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                r = get_reasons(handle)
                self.rows.append([str(sm), str(mx)] + ["Active" if (b and (r & b)) else "Not Active" for b, _ in bits])
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            idx = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:
                    pass
            handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            self.source = "nvml, 2 ms"
            return self
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            self.source = "nvidia-smi -lms 100"
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        self.stop = True
        if self.thread:
            self.thread.join(timeout=1)
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(self.NAMES, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def reference_gpu(w, h):
    """The reference's unmodified bmfr.cl on the B200 through NVIDIA's OpenCL ICD: the committed measurement of
    scripts/opencl_reference.py for this workload (profiles/), or None.  A reported baseline, never part of `value`."""
    for name in sorted((ROOT / "profiles").glob("r*_reference_opencl_b200*.json"), reverse=True):
        try:
            m = json.loads(name.read_text())
            if m.get("opencl") == "ok" and m.get("workload", "").startswith(f"{w}x{h} "):
                return {"frames_per_s": m["frames_per_s_from_total"], "kernel_sum_ms": m["kernel_sum_ms"], "device": m["device"],
                        "tmp_half": m["tmp_half"], "stages_ms": {k: v["mean_ms"] for k, v in m["stages"].items()},
                        "source": f"profiles/{name.name}"}
        except Exception:
            continue
    return None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
# CPU arms (the only places bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def time_oracle(kind, w, h, nframes, inputs=None, min_seconds=0.0, threads=0):
    """frames/s of `kind` ("port" | "reference") over the first `nframes` frames of the workload; the pass is
    repeated (restarting at frame 0) until `min_seconds` of CPU work have been timed.  -> (fps, seconds, stage ms, passes)"""
    from bmfr_b200 import synth
    from oracle.oracle import Oracle
    pl, nl = synth.limits()
    o = Oracle(kind, w, h, position_limit_squared=pl, normal_limit_squared=nl, threads=threads or host_cores())
    frames = []
    for f in range(nframes):
        a, n, p, c = inputs(f) if inputs else synth.frame_host(w, h, f)
        cam, _ = synth.camera(max(f - 1, 0), w, h)
        _, off = synth.camera(f, w, h)
        frames.append((f, a, n, p, c, cam, off))
    stage = np.zeros(6)
    passes = 0
    t0 = time.perf_counter()
    while True:
        for fr in frames:
            o.frame(*fr)
            stage += np.array(o.stage_ms())
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    o.close()
    return nframes * passes / dt, dt, (stage / (nframes * passes)).tolist(), passes


class OracleSequence:
    """The reference's frame loop on the host cores as one continuing run: frame numbers walk through the 60-frame
    sequence and wrap (frame 0, the only frame without a temporal path, comes up once per 60 frames as in the full
    workload).  Inputs are generated outside the timed calls."""

    def __init__(self, kind, w, h, threads):
        from bmfr_b200 import synth
        from oracle.oracle import Oracle
        self.synth, self.w, self.h, self.f = synth, w, h, 0
        pl, nl = synth.limits()
        self.o = Oracle(kind, w, h, position_limit_squared=pl, normal_limit_squared=nl, threads=threads)

    def run(self, nframes):
        """Processes the next `nframes` frames; returns the seconds spent inside the oracle (inputs excluded)."""
        spent = 0.0
        for _ in range(nframes):
            f = self.f % FRAMES
            a, n, p, c = self.synth.frame_host(self.w, self.h, f)
            cam, _ = self.synth.camera(max(f - 1, 0), self.w, self.h)
            _, off = self.synth.camera(f, self.w, self.h)
            t0 = time.perf_counter()
            self.o.frame(f, a, n, p, c, cam, off)
            spent += time.perf_counter() - t0
            self.f += 1
        return spent


def run_reference_arm(args):
    """Rank 0 only (the other ranks exit 0 without work).  Each of the `steps` steps is a bounded sample of the
    workload — F consecutive frames of the continuing sequence — with F calibrated on one untimed frame so that
    warm-up + steps take about --ref-budget seconds (default 60) at any size; `steps`, `ms_per_step` and
    `frames_per_step` are what was actually run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    w, h = SINGLE_GPU_WORKLOAD if args.gpus == 1 else sharded_workload(args.gpus)
    w, h = args.width or w, args.height or h
    norm = (w * h) / float(SINGLE_GPU_WORKLOAD[0] * SINGLE_GPU_WORKLOAD[1])  # 1080p-equivalent frames per frame
    cores = host_cores()
    kinds = ["port"] + (["reference"] if args.ref_shim and orc.available("reference") else [])
    results = {}
    for kind in kinds:
        seq = OracleSequence(kind, w, h, cores)
        t_cal = seq.run(1) + seq.run(1)        # calibration (untimed for the result): frames 0 and 1
        per_frame = max(t_cal / 2.0, 1e-4)
        nsteps = max(1, args.steps)
        if args.ref_frames > 0:
            fps_step = args.ref_frames
        else:
            fps_step = int(args.ref_budget / per_frame / (nsteps + max(0, args.warmup)))
            fps_step = max(1, min(FRAMES, fps_step))
        for _ in range(max(0, args.warmup)):
            seq.run(fps_step)
        step_s = [seq.run(fps_step) for _ in range(nsteps)]
        total = float(sum(step_s))
        results[kind] = dict(fps=fps_step * nsteps / total, seconds=total, frames_per_step=fps_step,
                             ms_per_step=1e3 * total / nsteps)
        seq.o.close()
    kind = max(results, key=lambda k: results[k]["fps"])
    r = results[kind]
    native = r["fps"]
    v = native * norm  # same unit as the B200 arm: 1080p-equivalent frames/s (identical to frames/s at N = 1)
    line = {
        "impl": "reference", "metric": "frames/sec", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": max(1, args.steps), "warmup": max(0, args.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "frames_per_s_native": native, "frames_per_step": r["frames_per_step"],
        "config": workload_config(w, h, args.gpus),
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"every step = {r['frames_per_step']} consecutive frames of the continuing 60-frame sequence "
                                   f"({r['seconds']:.1f} s of CPU work in the {max(1, args.steps)} timed steps), oracle/bmfr_oracle.c, "
                                   f"OpenMP on {cores} host threads (set explicitly: torchrun exports OMP_NUM_THREADS=1)",
                         "all": {k: {"frames_per_s": x["fps"], "frames_per_step": x["frames_per_step"]} for k, x in results.items()}},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm, one GPU
# ------------------------------------------------------------------------------------------------
def run_single_gpu(args):
    import torch
    from bmfr_b200 import Denoiser, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the bmfr_b200 path has no CPU fallback")
    w, h = args.width or SINGLE_GPU_WORKLOAD[0], args.height or SINGLE_GPU_WORKLOAD[1]
    torch.cuda.set_device(0)
    # an explicit stream: torch's default stream has handle 0, which the C ABI reads as "create one"
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0

    # inputs resident in HBM: 60 frames x 4 images (5.97 GB at 1080p, far larger than the 126 MB L2)
    inputs = torch.empty((FRAMES, 4, h, w, 3), dtype=torch.float32, device="cuda")
    for f in range(FRAMES):
        synth.frame_device(w, h, f, [inputs[f, k].data_ptr() for k in range(4)], stream=sp)
    cams = [synth.camera(max(f - 1, 0), w, h)[0] for f in range(FRAMES)]
    offs = [synth.camera(f, w, h)[1] for f in range(FRAMES)]
    torch.cuda.synchronize()

    # Like the reference's timed region (bmfr.cpp:446-476; the read-back at :479-480 is outside its timers)
    # the device-timed sequence leaves every frame's result in the context's result buffer
    # (BMFR_BUF_RESULT); the e2e arm below is the one that copies results out.
    def run_sequence(d):
        for f in range(FRAMES):
            d.denoise_frame(f, inputs[f, 0].data_ptr(), inputs[f, 1].data_ptr(), inputs[f, 2].data_ptr(),
                            inputs[f, 3].data_ptr(), cams[f], offs[f], None)

    overlap = int(args.overlap and args.mode == "fused")

    def timed(overlap_frames, steps, warmup, sample_clocks=False):
        """-> (total ms of `steps` steps, kernel launches, clock summary): CUDA events on the context's stream."""
        d = Denoiser(w, h, mode=args.mode, stream=sp, overlap_frames=overlap_frames, fit=args.fit)
        for _ in range(warmup):
            run_sequence(d)
        torch.cuda.synchronize()
        l0 = d.kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(0) if sample_clocks else None
        if sampler:
            sampler.__enter__()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            run_sequence(d)
        d.join()  # overlapped frames run on the context's internal streams: order the closing event after them
        e1.record(stream)
        torch.cuda.synchronize()
        if sampler:
            sampler.__exit__()
        d.sync()
        ms, n = e0.elapsed_time(e1), d.kernel_launches - l0
        d.close()
        return ms, n, (sampler.summary() if sampler else None)

    total_ms, launches, clock_summary = timed(overlap, args.steps, args.warmup, sample_clocks=True)
    ms_per_step = total_ms / args.steps
    fps = FRAMES * args.steps / (total_ms * 1e-3)
    # the same K steps on the drop-in in-order stream (the reference's queue semantics, bmfr.cpp:191)
    in_order = None
    if overlap:
        ms_io, _, _ = timed(0, args.steps, args.warmup)
        in_order = {"value": FRAMES * args.steps / (ms_io * 1e-3), "unit": "frames/s", "ms_per_frame": ms_io / args.steps / FRAMES,
                    "what": "overlap_frames = 0: one in-order stream, three kernels per frame chained with programmatic dependent launch"}
    # a longer timed region of the same mode, so that one DVFS wobble cannot move the figure
    sustained = None
    if args.sustain_seconds > 0:
        k = max(args.steps, int(np.ceil(args.sustain_seconds * 1e3 / ms_per_step)))
        ms_s, _, clk_s = timed(overlap, k, 1, sample_clocks=True)
        sustained = {"value": FRAMES * k / (ms_s * 1e-3), "unit": "frames/s", "steps": k, "seconds": ms_s * 1e-3, "clocks": clk_s}

    # per-kernel durations: same sequence with event pairs around every launch (separate pass, so the
    # events do not perturb the headline number)
    peak, peak_src = measured_peak()
    alg = algorithmic_bytes(w, h)
    kernels = {}
    for mode in ([args.mode] if not args.all_modes else ["fused", "staged"]):
        dp = Denoiser(w, h, mode=mode, stream=sp, profile=True, fit=args.fit)
        run_sequence(dp)
        run_sequence(dp)
        dp.sync()
        if mode == "staged":
            names = ["accumulate_noisy_data", "fitter", "weighted_sum", "accumulate_filtered_data", "taa"]
            ms = np.array([[dp.stage_ms(f)[k] for k in ("accum_noisy", "fitter", "weighted_sum", "accum_filtered", "taa", "total")]
                           for f in range(1, FRAMES)])  # frame 0 excluded like bmfr.cpp:392-397
        else:
            names = list(dp.fused_kernels)
            ms = np.array([[dp.fused_kernel_ms(f)[k] for k in names] + [0.0, 0.0, dp.stage_ms(f)["total"]]
                           for f in range(1, FRAMES)])
        mean = ms.mean(axis=0)
        for i, name in enumerate(names):
            if name:
                gbs = alg[name] / (mean[i] * 1e-3) / 1e9
                kernels[name] = {"ms": float(mean[i]), "min_ms": float(ms[:, i].min()), "max_ms": float(ms[:, i].max()),
                                 "algorithmic_bytes": alg[name], "achieved_gbs": gbs, "frac": gbs / peak}
        kernels[f"total_{mode}"] = {"ms": float(mean[5])}
        dp.close()
    # the same kernels timed INSIDE the chained in-order run (no event records between them, so their programmatic
    # launches overlap as in production): first CTA start .. last CTA end per kernel from device-side globaltimer stamps
    if args.mode == "fused":
        dq = Denoiser(w, h, mode="fused", stream=sp, profile=2, fit=args.fit)
        run_sequence(dq)
        run_sequence(dq)
        dq.sync()
        busy = [dq.fused_kernel_busy_ms(f) for f in range(1, FRAMES)]
        for name in dq.fused_kernels:
            b = float(np.mean([x[0][name] for x in busy]))
            kernels[name]["busy_ms"] = b
            kernels[name]["frac_busy"] = alg[name] / (b * 1e-3) / 1e9 / peak
        kernels["total_fused"]["busy_ms"] = float(np.mean([x[1] for x in busy]))
        dq.close()
    staged_names = ("accumulate_noisy_data", "fitter", "weighted_sum", "accumulate_filtered_data", "taa")
    own = [k for k in kernels if not k.startswith("total_") and (k not in staged_names) == (args.mode == "fused")]
    dom = max(own, key=lambda k: kernels[k]["ms"])
    # DRAM traffic of the same kernel from the committed ncu --set full capture (profiles/), per launch
    traffic, traffic_src = None, None
    try:
        cap = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        if args.mode == "fused" and (w, h) == SINGLE_GPU_WORKLOAD and dom in cap["kernels"]:
            k = cap["kernels"][dom]
            traffic, traffic_src = k["dram_bytes_read"] + k["dram_bytes_write"], cap["capture"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"], "ms_per_launch": kernels[dom]["ms"],
                "frac_busy": kernels[dom].get("frac_busy"),
                "how": "ms_per_launch: CUDA events around every launch of an in-order pass (includes the launch gap an isolated kernel "
                       "pays); frac_busy: the same bytes over first-CTA-start .. last-CTA-end inside the chained run (profile = 2)"}

    # end to end through the host-pointer entry of the C ABI: pinned-host inputs, uploads and the
    # read-back of every frame's result inside the timed region
    e2e = None
    if not args.no_e2e:
        host_in = torch.empty((FRAMES, 4, h, w, 3), dtype=torch.float32, pin_memory=True)
        host_in.copy_(inputs)
        host_out = torch.empty((2, h, w, 3), dtype=torch.float32, pin_memory=True)
        torch.cuda.synchronize()
        hin, hout = host_in.numpy(), host_out.numpy()
        dh = Denoiser(w, h, mode=args.mode, overlap_frames=overlap, fit=args.fit)

        def run_host_sequence():
            for f in range(FRAMES):
                dh.denoise_frame_host(f, hin[f, 0], hin[f, 1], hin[f, 2], hin[f, 3], cams[f], offs[f], hout[f & 1])
        run_host_sequence()
        dh.sync()
        steps = max(1, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(steps):
            run_host_sequence()
        dh.sync()
        dt = time.perf_counter() - t0
        dh.close()
        e2e = {"value": FRAMES * steps / dt, "unit": "frames/s", "h2d_bytes_per_step": FRAMES * 4 * w * h * 12,
               "d2h_bytes_per_step": FRAMES * w * h * 12, "steps": steps,
               "api": "bmfr_denoise_frame_host (C ABI), pinned host buffers, async upload ring + read-back"}
        del host_in, host_out

    cpu = None
    if not args.no_cpu:
        nfr = args.cpu_frames
        hin_small = inputs[:nfr].cpu().numpy()
        fps_cpu, dt_cpu, stage, passes = time_oracle("port", w, h, nfr, inputs=lambda f: [hin_small[f, k] for k in range(4)],
                                                     min_seconds=args.cpu_seconds, threads=host_cores())
        cpu = {"value": fps_cpu, "unit": "frames/s", "cores": host_cores(), "kind": "port",
               "sample": f"first {nfr} frames of the workload, {passes} passes ({dt_cpu:.1f} s of CPU work), "
                         "oracle/bmfr_oracle.c with OpenMP",
               "stage_ms": stage}

    line = {
        "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "ms_per_frame": ms_per_step / FRAMES, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(w, h, 1),
        "run": {"mode": args.mode, "overlap_frames": overlap, "fit_method": args.fit},
        "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "in_order": in_order, "sustained": sustained, "reference_gpu": reference_gpu(w, h),
        "clocks": clock_summary,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=150, help="timed steps (150 x 60 frames at 1080p is a 1.2 s timed region)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="fused", choices=["fused", "staged"])
    ap.add_argument("--all-modes", action="store_true", help="also time the five staged kernels one by one")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--cpu-frames", type=int, default=8, help="frames of the workload the cpu_baseline runs")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="the cpu_baseline repeats its frames until this much CPU work is timed")
    ap.add_argument("--ref-frames", type=int, default=0, help="--impl reference: frames per step (0 = calibrated to --ref-budget)")
    ap.add_argument("--ref-budget", type=float, default=60.0, help="--impl reference: seconds of CPU work for warm-up + steps")
    ap.add_argument("--ref-shim", action="store_true", help="--impl reference: also time the reference's own kernels through the CL shim")
    ap.add_argument("--parity-frames", type=int, default=3, help="N > 1: frames compared bit for bit with a whole-image run (outside the timed region)")
    ap.add_argument("--sustain-seconds", type=float, default=1.0, help="length of the extra long timed region reported as `sustained`")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="halo transport of the sharded run (N > 1)")
    ap.add_argument("--halo-rows", type=int, default=0, help="N > 1: halo_rows of the strip contexts (0 = sharding.default_halo(height))")
    ap.add_argument("--overlap", type=int, default=1, choices=[0, 1],
                    help="bmfr_params.overlap_frames of the timed contexts: 1 = consecutive frames overlap on the device "
                         "(three event-linked streams), 0 = one in-order stream")
    ap.add_argument("--fit", default="gram", choices=["gram", "tsqr"], help="bmfr_params.fit_method of the FUSED contexts")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        from bmfr_b200 import sharding
        return sharding.bench_sharded(args, sharded_workload(max(args.gpus, world)), FRAMES)
    return run_single_gpu(args)


if __name__ == "__main__":
    main()
