"""Strip sharding.  CPU: partition / halo-plan logic and the point-to-point halo
refresh over gloo with world_size 2.  GPU: all ranks emulated inside one process on one GPU must
reproduce the single-GPU run bit for bit."""
import os
import socket

import numpy as np
import pytest

from bmfr_b200 import sharding


def test_partition():
    # 2160 rows = 68 image block rows (the last one ragged) -> 34 / 34
    assert sharding.partition(2160, 2) == [(0, 1088), (1088, 2160)]
    assert sharding.partition(2160, 4) == [(0, 544), (544, 1088), (1088, 1632), (1632, 2160)]
    s8 = sharding.partition(4320, 8)
    assert all(b - a == 17 * 32 for a, b in s8[:-1]) and s8[-1][1] == 4320
    for n in (1, 2, 3, 5, 8):
        st = sharding.partition(1080, n)
        assert st[0][0] == 0 and st[-1][1] == 1080
        assert all(st[i][1] == st[i + 1][0] for i in range(n - 1))
        assert all(a % 32 == 0 for a, _ in st)
    with pytest.raises(ValueError):
        sharding.partition(64, 3)


def test_halo_messages_cover_every_halo_row_once():
    h, halo = 2160, 48
    for n in (2, 4, 8):
        strips = sharding.partition(h, n)
        sharding.check_partition(strips, h, halo)
        msgs = sharding.halo_messages(strips, h, halo)
        for r, (y0, y1) in enumerate(strips):
            lo, hi = sharding.storage_rows((y0, y1), h, halo)
            need = set(range(lo, y0)) | set(range(y1, hi))
            got = []
            for src, dst, a, b in msgs:
                if dst == r:
                    assert strips[src][0] <= a and b <= strips[src][1]
                    got += list(range(a, b))
            assert sorted(got) == sorted(need)
    with pytest.raises(ValueError):
        sharding.check_partition(sharding.partition(h, 2), h, 16)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, h, w, halo, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    strips = sharding.partition(h, world)
    msgs = sharding.halo_messages(strips, h, halo)
    lo, hi = sharding.storage_rows(strips[rank], h, halo)
    bpp = sharding._BYTES_PER_PIXEL
    # every rank fills ONLY its owned rows with a function of (buffer, row, byte); halo rows start as 0xFF
    state = {}
    for bi, name in enumerate(sharding.STATE_BUFFERS):
        pitch = w * bpp[name]
        t = torch.full(((hi - lo) * pitch,), 255, dtype=torch.uint8)
        for y in range(*strips[rank]):
            t[(y - lo) * pitch:(y - lo + 1) * pitch] = (torch.arange(pitch) * 7 + y * 13 + bi * 31) % 251
        state[name] = t

    def view(name, y0, y1):
        pitch = w * bpp[name]
        return state[name][(y0 - lo) * pitch:(y1 - lo) * pitch]

    sharding.exchange_distributed(view, msgs, rank)
    ok = True
    for bi, name in enumerate(sharding.STATE_BUFFERS):
        pitch = w * bpp[name]
        for y in range(lo, hi):
            expect = (torch.arange(pitch) * 7 + y * 13 + bi * 31) % 251
            ok = ok and bool((state[name][(y - lo) * pitch:(y - lo + 1) * pitch] == expect.to(torch.uint8)).all())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_halo_exchange_gloo_world2():
    """The N>1 host path on CPU: after the exchange every stored row (owned + halo) holds the owning
    rank's data."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 160, 24, 40, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["p2p", "copy"])
@pytest.mark.parametrize("n", [2, 3])
def test_sharded_equals_single_gpu_bitwise(n, exchange):
    """All ranks' strips in one process on one GPU: owned rows of every state buffer and the output are
    bit-identical to the whole-image run, frame after frame.  "p2p": the library's own halo pushes and
    device-side flags between locally connected contexts; "copy": halo rows refreshed by the test."""
    import torch
    from bmfr_b200 import Denoiser, synth
    w, h, frames, halo = 320, 384, 7, 40
    dev = torch.device("cuda:0")
    seq = []
    for f in range(frames):
        a, nrm, p, c = synth.frame_host(w, h, f)
        seq.append([torch.from_numpy(x).to(dev) for x in (a, nrm, p, c)])
    cams = [synth.camera(max(f - 1, 0), w, h)[0] for f in range(frames)]
    offs = [synth.camera(f, w, h)[1] for f in range(frames)]
    torch.cuda.synchronize()

    whole = Denoiser(w, h, mode="fused")
    ss = sharding.LocalStripSet(w, h, n, halo=halo, exchange=exchange)
    out_s = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
    out_w = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
    for f in range(frames):
        whole.denoise_frame(f, *[t.data_ptr() for t in seq[f]], cams[f], offs[f], out_w.data_ptr())
        ss.denoise_frame(f, seq[f], cams[f], offs[f], out_s)
        whole.sync(); ss.sync(); torch.cuda.synchronize()
        assert torch.equal(out_s, out_w), f"frame {f}: sharded output differs"
        for name in ("noisy_acc", "spp", "accum", "result", "accept"):
            ref = whole.read(name)
            for c in ss.ctx:
                got = c.d.read(name)
                y0, y1 = c.strip
                assert np.array_equal(got[y0 - c.row0:y1 - c.row0].view(np.uint8), ref[y0:y1].view(np.uint8)), (f, name, c.strip)
    whole.close(); ss.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,size", [(2, (320, 384)), (3, (320, 384)), (4, (1920, 1080))])
def test_overlapped_strips_equal_single_gpu_bitwise(n, size):
    """overlap_frames on strip contexts: three streams per context and the split (early / late) halo flags.  All
    frames are submitted back to back, nothing waits on the host in between; every output frame and the owned rows
    of every state buffer must equal the whole-image in-order run bit for bit."""
    import torch
    from bmfr_b200 import Denoiser, synth
    w, h = size
    frames, halo = 8, 48
    dev = torch.device("cuda:0")
    seq = []
    for f in range(frames):
        a, nrm, p, c = synth.frame_host(w, h, f)
        seq.append([torch.from_numpy(x).to(dev) for x in (a, nrm, p, c)])
    cams = [synth.camera(max(f - 1, 0), w, h)[0] for f in range(frames)]
    offs = [synth.camera(f, w, h)[1] for f in range(frames)]
    torch.cuda.synchronize()

    whole = Denoiser(w, h, mode="fused")
    out_w = [torch.zeros((h, w, 3), dtype=torch.float32, device=dev) for _ in range(frames)]
    for f in range(frames):
        whole.denoise_frame(f, *[t.data_ptr() for t in seq[f]], cams[f], offs[f], out_w[f].data_ptr())
    whole.sync()
    ref = {name: whole.read(name) for name in ("noisy_acc", "spp", "accum", "result", "accept")}
    whole.close()

    for rep in range(2):
        ss = sharding.LocalStripSet(w, h, n, halo=halo, exchange="p2p", overlap_frames=1)
        out_s = [torch.zeros((h, w, 3), dtype=torch.float32, device=dev) for _ in range(frames)]
        for f in range(frames):
            ss.denoise_frame(f, seq[f], cams[f], offs[f], out_s[f])
        ss.sync()
        torch.cuda.synchronize()
        for f in range(frames):
            assert torch.equal(out_s[f], out_w[f]), f"frame {f}: sharded output differs (pass {rep})"
        for name, r in ref.items():
            for c in ss.ctx:
                got = c.d.read(name)
                y0, y1 = c.strip
                assert np.array_equal(got[y0 - c.row0:y1 - c.row0].view(np.uint8), r[y0:y1].view(np.uint8)), (name, c.strip, rep)
        ss.close()


@pytest.mark.gpu
def test_halo_too_small_is_reported():
    import torch
    from bmfr_b200 import BmfrError, Denoiser, synth
    w, h = 320, 384
    d = Denoiser(w, h, mode="fused", strip=(128, 256), halo_rows=8)      # cannot even hold a straddling block
    g = d.geometry
    rows = slice(g.row0, g.row1)
    for f in range(2):
        a, nrm, p, c = [np.ascontiguousarray(x[rows]) for x in synth.frame_host(w, h, f)]
        d.denoise_frame_host(f, a, nrm, p, c, synth.camera(max(f - 1, 0), w, h)[0], synth.camera(f, w, h)[1])
    with pytest.raises(BmfrError) as e:
        d.sync()
    assert e.value.status == -6
    d.close()
