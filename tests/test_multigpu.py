"""N > 1 path on the CPU: two gloo ranks, each with a tile_rank/tile_world context of the host-emulated
kernels (tests/emu), all-gather of the packed tile buffers, un-tile; the result must equal the single-rank
frame bit for bit. (On the GPU box bench.py --gpus N runs the same code over NCCL.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import ctypes
    import importlib
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("hardware-ray-tracer_b200")
    emu = ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libbrt_emu.so"))
    ctx = pkg.binding.SceneApi(emu, "brt_", 0, rank, world, 0)
    scene = pkg.scenes.make_scene("cornell", small=True)
    scene.upload(ctx)
    w, h = 100, 70
    u = scene.uniform(ctx, w, h, 0, 3)
    frame = pkg.TiledFrame(ctx, w, h, rank, world, torch.device("cpu"))
    img = frame.render(u, ctx.opts(w, h, 2, 3)).numpy().copy()
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), img)
    # every rank also checks that it only traced its own share
    st = ctx.get_stats()
    np.save(os.path.join(out_dir, f"rays{rank}.npy"), np.array([st.rays_closest]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tile_parallel_gather_gloo(pkg, emu_lib, tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    single = pkg.binding.SceneApi(emu_lib, "brt_", 0, 0, 1, 0)
    scene = pkg.scenes.make_scene("cornell", small=True)
    scene.upload(single)
    w, h = 100, 70
    u = scene.uniform(single, w, h, 0, 3)
    ref = single.render_frame(u, single.opts(w, h, 2, 3))
    total = 0
    for r in range(world):
        img = np.load(tmp_path / f"rank{r}.npy")
        assert np.array_equal(img.view(np.uint32), ref.view(np.uint32)), r
        total += int(np.load(tmp_path / f"rays{r}.npy")[0])
    assert total == single.get_stats().rays_closest  # the ranks partition the work exactly
