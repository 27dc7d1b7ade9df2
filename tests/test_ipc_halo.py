"""The cross-process halo transport that bench.py --gpus N times: bmfr_halo_export -> bmfr_halo_connect (CUDA IPC
handles, peer stores over NVLink, device-side flags written by the neighbour's GPU), one process per GPU.  Owned rows
of every state buffer and every output frame must equal the single-GPU whole-image run bit for bit, with
overlap_frames 0 and 1 (SURVEY.md 8e "Correctness bar").

Needs two GPUs: ranks whose kernels wait on each other's flags must not share a device across processes (the
time-sliced contexts of two processes never run concurrently, B200_PROFILING.md), so the test is skipped on a
one-GPU box; there tests/test_sharding.py covers the same protocol with all strips in one process.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H, FRAMES, HALO = 320, 384, 7, 40
STATE = ("noisy_acc", "spp", "accum", "result", "accept")


def _worker(rank, world, overlap, inbox, outbox):
    try:
        import torch
        from bmfr_b200 import sharding, synth
        torch.cuda.set_device(rank)
        strips = sharding.partition(H, world)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        ctx = sharding.StripContext(W, H, strips[rank], HALO, rank, stream.cuda_stream, "fused", overlap_frames=overlap)
        outbox.put((rank, "blob", ctx.d.halo_export()))
        blobs = inbox.get(timeout=120)
        if rank > 0:
            ctx.d.halo_connect(0, blobs[rank - 1])
        if rank < world - 1:
            ctx.d.halo_connect(1, blobs[rank + 1])
        outbox.put((rank, "connected", None))
        assert inbox.get(timeout=120) == "go"
        rows = ctx.row1 - ctx.row0
        outs = torch.zeros((FRAMES, rows, W, 3), dtype=torch.float32, device="cuda")
        ins = []
        for f in range(FRAMES):
            a, n, p, c = synth.frame_host(W, H, f, y0=ctx.row0, y1=ctx.row1)
            ins.append([torch.from_numpy(x).cuda() for x in (a, n, p, c)])
        torch.cuda.synchronize()
        for f in range(FRAMES):
            ctx.d.denoise_frame(f, *[t.data_ptr() for t in ins[f]], synth.camera(max(f - 1, 0), W, H)[0], synth.camera(f, W, H)[1],
                                outs[f].data_ptr())
        ctx.d.sync()
        torch.cuda.synchronize()
        y0, y1 = strips[rank]
        res = {"out": outs[:, y0 - ctx.row0:y1 - ctx.row0].cpu().numpy()}
        for k in STATE:
            res[k] = ctx.d.read(k)[y0 - ctx.row0:y1 - ctx.row0].copy()
        outbox.put((rank, "done", res))
        inbox.get(timeout=120)  # keep the exported memory alive until the neighbour is finished too
        ctx.close()
    except Exception as e:  # noqa: BLE001 — reported to the parent, which fails the test
        import traceback
        outbox.put((rank, "error", traceback.format_exc() + repr(e)))


@pytest.mark.parametrize("overlap", [0, 1])
def test_two_processes_two_gpus_equal_single_gpu_bitwise(overlap):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (one process per GPU); single-GPU boxes run tests/test_sharding.py instead")
    import torch.multiprocessing as mp
    from bmfr_b200 import Denoiser, sharding, synth
    world = 2
    mpc = mp.get_context("spawn")
    inboxes, outbox = [mpc.Queue() for _ in range(world)], mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, overlap, inboxes[r], outbox)) for r in range(world)]
    for p in procs:
        p.start()

    def collect(tag):
        got = {}
        while len(got) < world:
            rank, what, payload = outbox.get(timeout=300)
            assert what != "error", f"rank {rank}: {payload}"
            assert what == tag, (what, tag)
            got[rank] = payload
        return got

    try:
        blobs = collect("blob")
        for q in inboxes:
            q.put([blobs[r] for r in range(world)])
        collect("connected")
        for q in inboxes:
            q.put("go")
        res = collect("done")
    finally:
        for q in inboxes:
            q.put("bye")
        for p in procs:
            p.join(timeout=60)

    outs = np.zeros((FRAMES, H, W, 3), dtype=np.float32)
    with Denoiser(W, H, mode="fused") as whole:
        for f in range(FRAMES):
            a, n, p, c = synth.frame_host(W, H, f)
            whole.denoise_frame_host(f, a, n, p, c, synth.camera(max(f - 1, 0), W, H)[0], synth.camera(f, W, H)[1], outs[f])
        whole.sync()
        ref = {k: whole.read(k) for k in STATE}
    strips = sharding.partition(H, world)
    for r, (y0, y1) in enumerate(strips):
        for f in range(FRAMES):
            assert np.array_equal(res[r]["out"][f].view(np.uint8), outs[f, y0:y1].view(np.uint8)), f"rank {r}: output of frame {f} differs"
        for k in STATE:
            assert np.array_equal(res[r][k].view(np.uint8), ref[k][y0:y1].view(np.uint8)), f"rank {r}: {k} differs"
