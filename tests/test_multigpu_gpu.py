"""N = 2 on real GPUs (skipped on a single-GPU box): both exchange modes of TiledFrame — NCCL all-gather + un-tile, and
the fused resolve with peer-memory stores — must reproduce the single-GPU frame bit for bit."""
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
    ctx = pkg.Context(device=rank, tile_rank=rank, tile_world=world)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))  # a real stream: frames are captured into CUDA graphs (two per slot, one per gather image)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    scene = pkg.scenes.make_scene("terrain", small=True)
    scene.upload(ctx)
    w, h = 500, 300
    u = scene.uniform(ctx, w, h, 0, 3)
    opts = ctx.opts(w, h, 2, 3)
    for mode in ("nccl", "p2p"):
        frame = pkg.TiledFrame(ctx, w, h, rank, world, dev, mode=mode)
        for it in range(2):  # twice: the second frame overwrites the first in the same buffers
            frame.render(u, opts)
            img = torch.empty(h * w * 4, dtype=torch.float32, device=dev)
            # copy the frame out through a raw device pointer
            import ctypes
            cudart = ctypes.CDLL("libcudart.so")
            cudart.cudaMemcpy(ctypes.c_void_p(img.data_ptr()), ctypes.c_void_p(frame.frame_ptr()), ctypes.c_size_t(h * w * 16), 3)
            torch.cuda.synchronize()
            dist.barrier()  # nobody may start the next frame before everybody has read this one
        np.save(os.path.join(out_dir, f"{mode}{rank}.npy"), img.cpu().numpy().reshape(h, w, 4))
        if mode == "p2p":
            # pipelined copy-out: rank 0 copies frame k to the host on a side stream while frame k+1 is traced (the library
            # alternates between two gather images; render() completes the previous copy before its barrier)
            hosts = [torch.empty(h * w * 4, dtype=torch.float32).pin_memory() for _ in range(2)]
            got = []
            for k in range(3):
                frame.render(scene.uniform(ctx, w, h, 10 + k, 3), opts)
                if rank == 0:
                    if k > 0:
                        got.append(hosts[(k - 1) % 2].numpy().reshape(h, w, 4).copy())
                    frame.to_host_async(hosts[k % 2])
            if rank == 0:
                frame.wait_host()
                got.append(hosts[2 % 2].numpy().reshape(h, w, 4).copy())
                np.save(os.path.join(out_dir, "pipelined.npy"), np.stack(got))
            # two frames in flight: frame k+1 is submitted before frame k is completed
            got = []
            n = 5
            for k in range(n):
                frame.submit(scene.uniform(ctx, w, h, 20 + k, 3), opts, k % 2)
                if k > 0:
                    frame.complete((k - 1) % 2)
                    if rank == 0:
                        frame.to_host(hosts[0])
                        got.append(hosts[0].numpy().reshape(h, w, 4).copy())
            frame.complete((n - 1) % 2)
            if rank == 0:
                frame.to_host(hosts[0])
                got.append(hosts[0].numpy().reshape(h, w, 4).copy())
                np.save(os.path.join(out_dir, "inflight.npy"), np.stack(got))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_exchange_modes(pkg, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    single = pkg.Context(device=0)
    scene = pkg.scenes.make_scene("terrain", small=True)
    scene.upload(single)
    w, h = 500, 300
    u = scene.uniform(single, w, h, 0, 3)
    ref = single.render_frame(u, single.opts(w, h, 2, 3))
    for mode in ("nccl", "p2p"):
        for r in range(world):
            img = np.load(tmp_path / f"{mode}{r}.npy")
            assert np.array_equal(img.view(np.uint32), ref.view(np.uint32)), (mode, r)
    inflight = np.load(tmp_path / "inflight.npy")
    for k in range(5):
        ref = single.render_frame(scene.uniform(single, w, h, 20 + k, 3), single.opts(w, h, 2, 3))
        assert np.array_equal(inflight[k].view(np.uint32), ref.view(np.uint32)), ("in flight", k)
    piped = np.load(tmp_path / "pipelined.npy")
    for k in range(3):
        ref = single.render_frame(scene.uniform(single, w, h, 10 + k, 3), single.opts(w, h, 2, 3))
        assert np.array_equal(piped[k].view(np.uint32), ref.view(np.uint32)), k
