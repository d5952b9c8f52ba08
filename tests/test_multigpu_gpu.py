"""N >= 2 on real GPUs (skipped on a single-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py`, log kept under
profiles/): both exchange modes of TiledFrame — NCCL all-gather + un-tile, and the fused resolve with peer-memory stores and
device-side completion flags (every rank receives / root only, RGBA32F / 8-bit, four frames in flight, every frame different) —
must reproduce the single-GPU frames bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import importlib
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pkg = importlib.import_module("hardware-ray-tracer_b200")
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))  # a real stream: frames are captured into CUDA graphs
    scene = pkg.scenes.make_scene("terrain", small=True)
    w, h = 500, 300
    flags = 3
    n_frames = 9  # more than two rounds over the four gather images: every image is overwritten twice

    def uniform(ctx, k):
        return scene.uniform(ctx, w, h, k, 3)  # the frame number seeds the bounce sampling: every frame is different

    # ---- NCCL all-gather + un-tile: every rank gets every frame
    ctx = pkg.Context(device=rank, tile_rank=rank, tile_world=world)
    scene.upload(ctx)
    frame = pkg.TiledFrame(ctx, w, h, rank, world, dev, mode="nccl")
    imgs = []
    for k in range(3):
        imgs.append(frame.render(uniform(ctx, k), ctx.opts(w, h, 2, flags)).clone())
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"nccl{rank}.npy"), torch.stack(imgs).cpu().numpy())
    ctx.close()

    # ---- fused exchange, every rank receives, synchronous calls: frames 0-2 with the default form (local resolve + small-grid push kernel),
    # frames 3-5 with the one-kernel form (BRT_PEER_PUSH=0, read by brt_create)
    imgs = []
    host = torch.empty(h * w * 4, dtype=torch.float32).pin_memory()
    for form in (None, "0"):
        if form is not None:
            os.environ["BRT_PEER_PUSH"] = form
        ctx = pkg.Context(device=rank, tile_rank=rank, tile_world=world)
        os.environ.pop("BRT_PEER_PUSH", None)
        scene.upload(ctx)
        frame = pkg.TiledFrame(ctx, w, h, rank, world, dev, mode="p2p")
        for k in range(len(imgs), len(imgs) + 3):
            frame.render(uniform(ctx, k), ctx.opts(w, h, 2, flags))
            frame.to_host(host)
            imgs.append(host.numpy().reshape(h, w, 4).copy())
        frame.check()
        dist.barrier()
        ctx.close()
    np.save(os.path.join(out_dir, f"p2p_all{rank}.npy"), np.stack(imgs))

    # ---- fused exchange, root only, FOUR frames in flight, no host barrier anywhere: RGBA32F, then B8G8R8A8_UNORM
    for name, fmt, dtype, per_px in (("f32", 0, torch.float32, 4), ("bgra8", pkg.render_format(pkg.FORMAT_BGRA8_UNORM), torch.uint8, 4)):
        ctx = pkg.Context(device=rank, tile_rank=rank, tile_world=world)
        scene.upload(ctx)
        frame = pkg.TiledFrame(ctx, w, h, rank, world, dev, mode="p2p", root_only=True, root=0)
        hosts = [torch.empty(h * w * per_px, dtype=dtype).pin_memory() for _ in range(4)]
        got = [None] * n_frames
        opts = ctx.opts(w, h, 2, flags | fmt)
        for k in range(n_frames):
            slot = k % 4
            if rank == 0 and k >= 4:  # the host buffer of this slot still holds frame k - 4
                frame.wait_fetch(slot)
                got[k - 4] = hosts[slot].numpy().reshape(h, w, per_px).copy()
            frame.submit(uniform(ctx, 20 + k), opts, slot)
            if rank == 0:
                frame.fetch_async(slot, hosts[slot])
        if rank == 0:
            for k in range(max(0, n_frames - 4), n_frames):
                frame.wait_fetch(k % 4)
                got[k] = hosts[k % 4].numpy().reshape(h, w, per_px).copy()
            np.save(os.path.join(out_dir, f"root_{name}.npy"), np.stack(got))
        for k in range(4):
            ctx.frame_wait(k)
        frame.check()
        dist.barrier()
        ctx.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_exchange_modes(pkg, tmp_path):
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    single = pkg.Context(device=0)
    scene = pkg.scenes.make_scene("terrain", small=True)
    scene.upload(single)
    w, h = 500, 300

    def ref(k, fmt=0):
        return single.render_frame(scene.uniform(single, w, h, k, 3), single.opts(w, h, 2, 3 | fmt))

    for r in range(world):
        nccl = np.load(tmp_path / f"nccl{r}.npy")
        for k in range(3):
            assert np.array_equal(nccl[k].reshape(h, w, 4).view(np.uint32), ref(k).view(np.uint32)), ("nccl", r, k)
        p2p = np.load(tmp_path / f"p2p_all{r}.npy")
        for k in range(6):
            assert np.array_equal(p2p[k].view(np.uint32), ref(k).view(np.uint32)), ("p2p, every rank receives", r, k)
    root = np.load(tmp_path / "root_f32.npy")
    assert len(root) == 9
    distinct = 0
    for k in range(9):
        want = ref(20 + k)
        assert np.array_equal(root[k].view(np.uint32), want.view(np.uint32)), ("root only, four frames in flight", k)
        distinct += k > 0 and not np.array_equal(root[k], root[k - 1])
    assert distinct == 8  # the frames really differ, so a stale or half-overwritten gather image would have been seen
    root8 = np.load(tmp_path / "root_bgra8.npy")
    for k in range(9):
        want = ref(20 + k, pkg.render_format(pkg.FORMAT_BGRA8_UNORM))
        assert np.array_equal(root8[k], want), ("root only, B8G8R8A8_UNORM", k)


def test_flag_protocol_on_one_gpu(pkg):
    """The fused exchange with a world of ONE rank (its own and only receiver): the same kernels as on N GPUs — resolve with peer stores into
    gather image `slot`, arrival flag published by the last block, device-side wait before the copy-out, release, and the producer's
    device-side wait for that release four frames later — run on the single-GPU test box. Nine different frames over the four slots, never
    more than the copy stream between host and device: every frame equals a plain brt_render_frame, in RGBA32F and in B8G8R8A8."""
    a = pkg.Context(device=0)
    scene = pkg.scenes.make_scene("terrain", small=True)
    scene.upload(a)
    w, h = 300, 200
    a.gather_image_open([a.gather_image_export(w, h)])
    a.gather_configure(True, 0)
    copy_stream = torch.cuda.Stream()
    for fmt, dtype in ((0, torch.float32), (pkg.render_format(pkg.FORMAT_BGRA8_UNORM), torch.uint8)):
        hosts = [torch.empty(h * w * 4, dtype=dtype).pin_memory() for _ in range(4)]
        got = []
        n = 9
        for k in range(n):
            slot = k % 4
            if k >= 4:
                copy_stream.synchronize()  # (host buffer reuse; the device side needs no help)
                got.append(hosts[slot].numpy().reshape(h, w, 4).copy())
            a.render_frame_peers_async(scene.uniform(a, w, h, 30 + k, 3), a.opts(w, h, 2, 3 | fmt), slot)
            a.gather_copy_to_host(slot, hosts[slot].data_ptr(), copy_stream.cuda_stream)
        copy_stream.synchronize()
        for k in range(max(0, n - 4), n):
            got.append(hosts[k % 4].numpy().reshape(h, w, 4).copy())
        for k in range(4):
            a.frame_wait(k)
        assert a.gather_timed_out() == 0
        for k in range(n):
            want = a.render_frame(scene.uniform(a, w, h, 30 + k, 3), a.opts(w, h, 2, 3 | fmt))
            assert np.array_equal(got[k], want), (fmt, k)
    # misuse is an error, not a crash
    with pytest.raises(pkg.BrtError):
        a.render_frame_peers_async(scene.uniform(a, w, h, 0, 3), a.opts(w, h, 2, 3), 4)       # slot >= BRT_GATHER_IMAGES
    with pytest.raises(pkg.BrtError):
        a.render_frame_peers_async(scene.uniform(a, w, h, 0, 3), a.opts(w + 32, h, 2, 3), 0)  # another frame size than exported
    with pytest.raises(pkg.BrtError):
        a.gather_configure(True, 0)                                                            # after the first frame
    a.close()
