"""Strip sharding of one frame over several GPUs along block rows.

The reference is single-device; this is the B200-side answer to the 4K / 8K configs of BASELINE.json.
Rank r owns a contiguous band of image rows (a whole number of 32-row block rows) and stores
`halo_rows` extra rows on each interior side.  Per frame:

  1. every rank runs the two fused kernels on its strip: blocks that straddle a strip boundary are
     fitted redundantly by both neighbours from identical data, so their coefficients are identical
     bits and no coefficient exchange is needed;
  2. the four temporal state buffers (accumulated noisy colour, spp, accumulated filtered colour,
     TAA result) have their halo rows refreshed from the owning neighbour — the only communication,
     neighbour-to-neighbour, no collective reduction.  Two transports:
       "p2p"  (default): the contexts are connected once (CUDA IPC handles exchanged through
              torch.distributed); from then on every bmfr_denoise_frame call pushes its boundary rows
              straight into the neighbours' halo rows over NVLink and waits for theirs on the device
              (csrc/bmfr_pipeline.cu, "Peer-to-peer halo exchange") — no host work per frame;
       "nccl" : batched torch.distributed isend/irecv of the same rows after every frame.

Ownership is fixed, so the sharded output equals the single-GPU output bit for bit as long as
`halo_rows >= 34 + max vertical reprojection distance`; a gather that leaves strip + halo raises
BMFR_ERR_HALO_TOO_SMALL instead of silently diverging.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

from . import synth
from .denoiser import Denoiser

STATE_BUFFERS = ("noisy_acc", "spp", "accum", "result")   # what frame f+1 gathers from frame f
_BYTES_PER_PIXEL = dict(noisy_acc=12, spp=1, accum=12, result=12)
DEFAULT_HALO = 48


def default_halo(height: int) -> int:
    """halo_rows for the synthetic sequence at this image height: 34 rows (one straddling block row + the TAA ring) + the
    largest vertical reprojection distance (5.6 pixels per frame at 1080 rows in synth-v1, measured with the oracle over the
    60 frames; it scales with the height) + the second bilinear tap row + one row of slack.  A gather that leaves strip +
    halo is reported (BMFR_ERR_HALO_TOO_SMALL), never silently wrong."""
    return 36 + (6 * height + 1079) // 1080


def partition(height: int, n: int):
    """Contiguous bands of block rows, sizes differing by at most one block row (e.g. 69 -> 35/34)."""
    nb = (height + 31) // 32
    if n < 1 or n > nb:
        raise ValueError(f"cannot split {nb} block rows over {n} ranks")
    base, extra = divmod(nb, n)
    out, b = [], 0
    for r in range(n):
        rows = base + (1 if r < extra else 0)
        out.append((min(b * 32, height), min((b + rows) * 32, height)))
        b += rows
    return out


def storage_rows(strip, height, halo):
    return max(0, strip[0] - halo), min(height, strip[1] + halo)


def halo_messages(strips, height, halo):
    """[(src_rank, dst_rank, y0, y1)]: image rows [y0,y1) owned by src that dst stores as halo."""
    msgs = []
    for r, (y0, y1) in enumerate(strips):
        lo, hi = storage_rows((y0, y1), height, halo)
        for s, (sy0, sy1) in enumerate(strips):
            if s == r:
                continue
            a, b = max(lo, sy0), min(hi, sy1)
            if a < b:
                msgs.append((s, r, a, b))
    return msgs


def check_partition(strips, height, halo):
    if halo < 34:
        raise ValueError("halo_rows must cover one block row plus the TAA ring (>= 34) plus the camera motion")
    for (y0, y1) in strips:
        if y1 - y0 < 32 and y1 != height:
            raise ValueError("strips must hold at least one block row")


class _DevBuf:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can address it."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def _as_tensor(ptr, nbytes, device):
    import torch
    return torch.as_tensor(_DevBuf(ptr, nbytes), device=device)


class StripContext:
    """One rank's strip: a Denoiser created with strip_y0/strip_y1/halo_rows + views of its state."""

    def __init__(self, width, height, strip, halo, device=0, stream=0, mode="fused", **kw):
        self.W, self.H, self.strip, self.halo, self.device = width, height, strip, halo, device
        whole = strip == (0, height)
        self.d = Denoiser(width, height, mode=mode, device=device, stream=stream,
                          strip=None if whole else strip, halo_rows=0 if whole else halo, **kw)
        g = self.d.geometry
        self.row0, self.row1 = g.row0, g.row1
        self._views = {}

    def rows_view(self, name, y0, y1):
        """uint8 tensor over image rows [y0,y1) of the state buffer as of the last frame."""
        ptr, nbytes = self.d.buffer_ptr(name)
        key = (name, ptr)
        if key not in self._views:
            self._views[key] = _as_tensor(ptr, nbytes, f"cuda:{self.device}")
        pitch = self.W * _BYTES_PER_PIXEL[name]
        return self._views[key][(y0 - self.row0) * pitch:(y1 - self.row0) * pitch]

    def close(self):
        self.d.close()


class LocalStripSet:
    """All ranks' strips inside ONE process on ONE GPU, exchanging halos with device copies.  This is
    how the strip logic is verified bit for bit on a single-GPU box (ranks emulated over all ranks'
    data, as B200_PROFILING.md prescribes when there are fewer GPUs than ranks)."""

    def __init__(self, width, height, n, halo=DEFAULT_HALO, device=0, mode="fused", exchange="p2p", **kw):
        import torch
        self.W, self.H, self.n, self.exchange = width, height, n, exchange
        self.strips = partition(height, n)
        check_partition(self.strips, height, halo)
        # one explicit stream for all strips and for the halo copies (torch's default stream has
        # handle 0, which the C ABI reads as "create a private stream")
        self._stream = torch.cuda.Stream()
        torch.cuda.set_stream(self._stream)
        self.stream = self._stream.cuda_stream
        self.ctx = [StripContext(width, height, s, halo, device, self.stream, mode, **kw) for s in self.strips]
        self.msgs = halo_messages(self.strips, height, halo)
        if exchange == "p2p":  # the library pushes / waits for halo rows itself from now on
            for r in range(n - 1):
                self.ctx[r].d.halo_connect_local(1, self.ctx[r + 1].d)
                self.ctx[r + 1].d.halo_connect_local(0, self.ctx[r].d)

    def denoise_frame(self, frame, full_inputs, cam_prev, pixel_offset, out_full):
        """full_inputs: four torch CUDA tensors [H, W, 3]; out_full: [H, W, 3] receives owned rows."""
        for c in self.ctx:
            ptrs = [t[c.row0:c.row1].data_ptr() for t in full_inputs]
            c.d.denoise_frame(frame, *ptrs, cam_prev, pixel_offset, out_full[c.row0:c.row1].data_ptr())
        if self.exchange != "p2p":
            for name in STATE_BUFFERS:
                for src, dst, y0, y1 in self.msgs:
                    self.ctx[dst].rows_view(name, y0, y1).copy_(self.ctx[src].rows_view(name, y0, y1))

    def sync(self):
        for c in self.ctx:
            c.d.sync()

    def close(self):
        for c in self.ctx:
            c.close()


def exchange_distributed(views_by_name, msgs, rank):
    """Neighbour halo refresh with torch.distributed point-to-point ops (NCCL on GPUs, gloo on CPU).
    views_by_name(name, y0, y1) -> contiguous uint8 tensor over those image rows."""
    import torch.distributed as dist
    ops = []
    for name in STATE_BUFFERS:
        for src, dst, y0, y1 in msgs:
            if src == rank:
                ops.append(dist.P2POp(dist.isend, views_by_name(name, y0, y1), dst))
            elif dst == rank:
                ops.append(dist.P2POp(dist.irecv, views_by_name(name, y0, y1), src))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


# ------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): BASELINE.json's multi-GPU configs strip-sharded, one rank per GPU
# ------------------------------------------------------------------------------------------------
def _connect(ctx, rank, world, dist):
    """One-time exchange of IPC handles between neighbouring ranks; afterwards no host-side communication per frame."""
    blobs = [None] * world
    dist.all_gather_object(blobs, ctx.d.halo_export())
    if rank > 0:
        ctx.d.halo_connect(0, blobs[rank - 1])
    if rank < world - 1:
        ctx.d.halo_connect(1, blobs[rank + 1])
    dist.barrier()


def strip_algorithmic_bytes(w, strip):
    """Algorithmic bytes per frame of the three FUSED kernels for the rows a rank OWNS (SURVEY.md 8d's per-pixel figures;
    the redundant rows of straddling blocks and the halo traffic are not credited)."""
    P = w * (strip[1] - strip[0])
    NB = ((w + 31) // 32 + 1) * ((strip[1] - strip[0] + 31) // 32)
    return {"reproject": 95 * P, "fit": 36 * P + 216 * NB, "post": 94 * P + 168 * NB}


def bench_sharded(args, workload, frames):
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, measured_peak, workload_config

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    w, h = args.width or workload[0], args.height or workload[1]
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    halo = int(getattr(args, "halo_rows", 0)) or default_halo(h)
    strips = partition(h, world)
    check_partition(strips, h, halo)
    msgs = halo_messages(strips, h, halo)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    exchange = getattr(args, "exchange", "p2p")
    fit = getattr(args, "fit", "gram")
    # the host-driven NCCL exchange works on the context's stream between frames and needs the in-order mode
    overlap = int(bool(getattr(args, "overlap", 1)) and exchange == "p2p" and args.mode == "fused")

    def make_ctx(**kw):
        c = StripContext(w, h, strips[rank], halo, local, sp, args.mode, fit=fit, **kw)
        if exchange == "p2p":
            _connect(c, rank, world, dist)
        return c

    def replace_ctx(old, **kw):
        """A neighbour may still be storing into this context's memory until every rank is done with the frames."""
        old.d.sync()
        torch.cuda.synchronize()
        dist.barrier()
        old.close()
        return make_ctx(**kw)

    ctx = make_ctx(overlap_frames=overlap)
    rows = ctx.row1 - ctx.row0
    inputs = torch.empty((frames, 4, rows, w, 3), dtype=torch.float32, device="cuda")
    for f in range(frames):
        synth.frame_device(w, h, f, [inputs[f, k].data_ptr() for k in range(4)], y0=ctx.row0, y1=ctx.row1, stream=sp)
    cams = [synth.camera(max(f - 1, 0), w, h)[0] for f in range(frames)]
    offs = [synth.camera(f, w, h)[1] for f in range(frames)]

    def run_sequence(c, out_ptr=None, nframes=frames):
        for f in range(nframes):
            # like the single-GPU arm, the device-timed run leaves the result in the context's buffer
            c.d.denoise_frame(f, inputs[f, 0].data_ptr(), inputs[f, 1].data_ptr(), inputs[f, 2].data_ptr(),
                              inputs[f, 3].data_ptr(), cams[f], offs[f], out_ptr(f) if out_ptr else None)
            if exchange != "p2p":
                exchange_distributed(c.rows_view, msgs, rank)

    def timed(c, steps, warmup, clocks=False):
        """max over ranks of the device time of `steps` steps (CUDA events on the context's stream), launches"""
        for _ in range(warmup):
            run_sequence(c)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        l0 = c.d.kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local) if clocks else None
        if sampler:
            sampler.__enter__()
        e0.record(stream)
        for _ in range(steps):
            run_sequence(c)
        c.d.join()  # overlapped frames: order the closing event after the frames on the internal streams
        e1.record(stream)
        torch.cuda.synchronize()
        if sampler:
            sampler.__exit__()
        dist.barrier()
        torch.cuda.synchronize()
        c.d.sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), c.d.kernel_launches - l0, (sampler.summary() if sampler else None)

    # ---- parity, outside the timed region: the first K frames of THIS transport (IPC peer stores, device flags) against a
    # whole-image context on the same GPU, owned rows of every output frame bit for bit
    K = max(0, min(int(getattr(args, "parity_frames", 3)), frames))
    parity = None
    if K > 0:
        y0, y1 = strips[rank]
        out_s = torch.zeros((K, rows, w, 3), dtype=torch.float32, device="cuda")
        run_sequence(ctx, out_ptr=lambda f: out_s[f].data_ptr(), nframes=K)
        ctx.d.sync()
        torch.cuda.synchronize()
        ok = True
        full = torch.empty((2, 4, h, w, 3), dtype=torch.float32, device="cuda")   # this frame's and the previous frame's inputs
        out_w = torch.empty((h, w, 3), dtype=torch.float32, device="cuda")
        with Denoiser(w, h, mode=args.mode, device=local, stream=sp, fit=fit) as whole:
            for f in range(K):
                synth.frame_device(w, h, f, [full[f & 1, k].data_ptr() for k in range(4)], stream=sp)
                whole.denoise_frame(f, *[full[f & 1, k].data_ptr() for k in range(4)], cams[f], offs[f], out_w.data_ptr())
                whole.sync()
                ok = ok and bool(torch.equal(out_w[y0:y1], out_s[f, y0 - ctx.row0:y1 - ctx.row0]))
        del full, out_w, out_s
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = {"frames": K, "bit_identical": bool(flag.item() == 1),
                  "what": "owned rows of every rank's output frames vs a whole-image context of the same library on the same GPU"}
        # a fresh context for the timed run: frame numbers restart at 0, and so do the halo flags
        ctx = replace_ctx(ctx, overlap_frames=overlap)

    total_ms, launches, clock_summary = timed(ctx, args.steps, args.warmup, clocks=True)
    halo_bytes = sum((y1 - y0) * w * sum(_BYTES_PER_PIXEL.values()) for s, d_, y0, y1 in msgs if d_ == rank)
    norm = (w * h) / float(1920 * 1080)
    native = frames * args.steps / (total_ms * 1e-3)

    # the same K steps on in-order streams (one stream per context, the reference's queue semantics)
    in_order = None
    if overlap:
        ctx = replace_ctx(ctx, overlap_frames=0)
        ms_io, _, _ = timed(ctx, args.steps, args.warmup)
        in_order = {"value": frames * args.steps / (ms_io * 1e-3) * norm, "unit": "frames/s", "ms_per_frame": ms_io / args.steps / frames}

    # per-kernel durations on every rank (event pairs around the launches, in-order profile contexts connected like the
    # timed ones), roofline of rank 0's dominant kernel against the bytes of the rows it owns
    kernels = roofline = None
    if args.mode == "fused" and exchange == "p2p":
        ctx = replace_ctx(ctx, profile=True)
        run_sequence(ctx)
        run_sequence(ctx)
        ctx.d.sync()
        names = list(ctx.d.fused_kernels)
        ms = np.array([[ctx.d.fused_kernel_ms(f)[k] for k in names] for f in range(1, frames)]).mean(axis=0)
        per_rank = torch.tensor(ms, device="cuda", dtype=torch.float64)
        gathered = [torch.zeros_like(per_rank) for _ in range(world)]
        dist.all_gather(gathered, per_rank)
        peak, peak_src = measured_peak()
        alg = strip_algorithmic_bytes(w, strips[0])
        kernels = {}
        for i, (name, key) in enumerate(zip(names, ("reproject", "fit", "post"))):
            gbs = alg[key] / (float(gathered[0][i]) * 1e-3) / 1e9
            kernels[name] = {"ms": float(gathered[0][i]), "ms_per_rank": [float(g[i]) for g in gathered], "algorithmic_bytes": alg[key],
                             "achieved_gbs": gbs, "frac": gbs / peak}
        dom = max(kernels, key=lambda k: kernels[k]["ms"])
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["frac"], "traffic": None, "peak_source": peak_src, "rank": 0,
                    "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"], "ms_per_launch": kernels[dom]["ms"],
                    "what": "rank 0's strip: bytes of the rows it owns / the kernel's mean launch duration on rank 0"}
        ctx = replace_ctx(ctx, overlap_frames=overlap)

    # timeline of the timed configuration (overlapped, connected) from the kernels' own globaltimer stamps (profile = 2),
    # every rank, and next to it the same strip without neighbours (what the strip geometry alone costs: halo rows
    # reprojected twice, partial tiles at the strip edges); the difference is the exchange
    timeline = None
    if args.mode == "fused" and exchange == "p2p":
        def stamps_summary(c):
            run_sequence(c)
            run_sequence(c)
            c.d.sync()
            t = np.array([[x for pair in c.d.fused_kernel_stamps(f) for x in pair] for f in range(8, frames)], dtype=np.float64) * 1e-3
            r0 = t[:, 0:1]
            rel = (t - r0).mean(axis=0)
            return [float(np.diff(t[:, 0]).mean())] + [float(v) for v in rel[1:]] + [float((t[:, 2 * k + 1] - t[:, 2 * k]).mean()) for k in range(3)]
        ctx = replace_ctx(ctx, overlap_frames=overlap, profile=2)
        mine = stamps_summary(ctx)
        ctx.d.sync()
        torch.cuda.synchronize()
        dist.barrier()
        lone = StripContext(w, h, strips[rank], halo, local, sp, args.mode, fit=fit, overlap_frames=overlap, profile=2)
        try:
            alone = stamps_summary(lone)
        except Exception:  # (stale halo rows can trip the out-of-rows flag; timing only)
            alone = [float("nan")] * 9
        lone.close()
        both = torch.tensor([mine, alone], device="cuda", dtype=torch.float64)
        gathered = [torch.zeros_like(both) for _ in range(world)]
        dist.all_gather(gathered, both)
        keys = ("period", "reproject_end", "fit_start", "fit_end", "post_start", "post_end", "reproject_busy", "fit_busy", "post_busy")
        timeline = {"unit": "us", "what": "per rank, mean over frames 8.. of one pass: frame period (start of reprojection to the next one) "
                    "and first-CTA-start / last-CTA-end of the three kernels relative to the start of the frame's reprojection; "
                    "'connected' = the timed configuration, 'alone' = the same strip without neighbours",
                    "connected": {k: [round(float(g[0][i]), 1) for g in gathered] for i, k in enumerate(keys)},
                    "alone": {k: [round(float(g[1][i]), 1) for g in gathered] for i, k in enumerate(keys)}}
        ctx = replace_ctx(ctx, overlap_frames=overlap)

    # end to end: every rank uploads its strip of each frame from pinned host memory and reads its rows of the
    # result back, through the C ABI's host entry (bounded to the first frames of the sequence to bound pinned memory)
    e2e = None
    if not getattr(args, "no_e2e", False):
        nf = min(frames, 20)
        host_in = torch.empty((nf, 4, rows, w, 3), dtype=torch.float32, pin_memory=True)
        host_in.copy_(inputs[:nf])
        host_out = torch.empty((2, rows, w, 3), dtype=torch.float32, pin_memory=True)
        torch.cuda.synchronize()
        hin, hout = host_in.numpy(), host_out.numpy()

        def run_host_sequence():
            for f in range(nf):
                ctx.d.denoise_frame_host(f, hin[f, 0], hin[f, 1], hin[f, 2], hin[f, 3], cams[f], offs[f], hout[f & 1])
                if exchange != "p2p":
                    exchange_distributed(ctx.rows_view, msgs, rank)
        run_host_sequence()
        ctx.d.sync()
        dist.barrier()
        steps_h = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(steps_h):
            run_host_sequence()
        ctx.d.sync()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        own_rows = strips[rank][1] - strips[rank][0]
        e2e = {"value": nf * steps_h / float(dt.item()) * norm, "unit": "frames/s",
               "frames_per_s_native": nf * steps_h / float(dt.item()),
               "h2d_bytes_per_step": nf * 4 * rows * w * 12, "d2h_bytes_per_step": nf * own_rows * w * 12, "steps": steps_h,
               "frames_per_step": nf, "bytes_are": "per rank (rank 0)",
               "api": "bmfr_denoise_frame_host on every rank: pinned host buffers, async upload ring + read-back, p2p halos"}
        del host_in, host_out
    if rank == 0:
        line = {
            "metric": "frames/sec", "value": native * norm, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "ms_per_frame": total_ms / args.steps / frames,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "frames_per_s_native": native,
            "config": workload_config(w, h, world),
            "run": {"mode": args.mode, "overlap_frames": overlap, "fit_method": fit, "strips": strips, "halo_rows": halo,
                    "exchange": ("in-kernel peer stores of the state halo rows over NVLink (CUDA IPC mappings), device-side flags"
                                 if exchange == "p2p" else "NCCL send/recv of state halo rows, neighbours only")},
            "parity": parity, "in_order": in_order, "kernels": kernels, "roofline": roofline, "timeline": timeline,
            "halo_bytes_per_frame_rank0": halo_bytes, "gpu_launches": int(launches),
            "cpu_baseline": None, "e2e": e2e, "clocks": clock_summary,
        }
        print(json.dumps(line))
    ctx.d.sync()
    torch.cuda.synchronize()
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
