#!/usr/bin/env python
"""The fused exchange at N GPUs under different settings, one torchrun job: for every variant (environment read by brt_create) a fresh
context per rank, the C5 frame tiled over the ranks, fused resolve + peer stores to rank 0, 4 frames in flight, in-stream L2 flush per
frame like bench.py; ms per frame = max over ranks of one timed region.

  torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/ab_peers.py --variants "" BRT_PEER_GRID=148 BRT_PEER_GRID=32
"""
import argparse, importlib, json, os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before CUDA initialises (set it to 8 on the command line for the old behaviour)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5")
    ap.add_argument("--frames", type=int, default=80)
    ap.add_argument("--slots", type=int, default=4)
    ap.add_argument("--bgra8", action="store_true", help="gather images in the 8-bit present format (4 bytes per pixel over NVLink)")
    ap.add_argument("--variants", nargs="+", default=[""])
    ap.add_argument("--no-exchange", action="store_true", help="also: every rank renders its tiles with no exchange at all")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("hardware-ray-tracer_b200")
    cfg = dict(pkg.scenes.CONFIGS[a.config])
    scene = pkg.scenes.make_scene(cfg.pop("scene"))
    w, h = cfg["width"], cfg["height"]
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
    S = a.slots
    per_rank = []

    def timed(ctx, body):
        streams = [torch.cuda.ExternalStream(ctx.frame_stream(k), device=dev) for k in range(S)]

        def run(n):
            for i in range(n):
                k = i % S
                with torch.cuda.stream(streams[k]):
                    flush.fill_(i & 0xff)
                body(i, k)
            for k in range(S):
                ctx.frame_wait(k)
            torch.cuda.synchronize()

        run(2 * S)
        dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        run(a.frames)
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(streams[0])
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank[:] = [round(float(x.item()) / a.frames, 4) for x in every]
        return max(per_rank)

    for v in a.variants:
        for kv in v.split(","):
            if "=" in kv:
                k_, v_ = kv.split("=")
                os.environ[k_] = v_
        ctx = pkg.Context(device=local, tile_rank=rank, tile_world=world)
        scene.upload(ctx)
        u = scene.uniform(ctx, w, h, 0, cfg["depth_max"])
        opts = ctx.opts(w, h, cfg["spp"], cfg["flags"] | (pkg.render_format(pkg.FORMAT_BGRA8_UNORM) if a.bgra8 else 0))
        frame = pkg.TiledFrame(ctx, w, h, rank, world, dev, mode="p2p", root_only=True, root=0)

        def body(i, k):
            frame.submit(u, opts, k)
            if rank == 0:
                frame.release(k)

        ms = timed(ctx, body)
        ranks_ms = list(per_rank)
        frame.check()
        ms_alone = None
        if a.no_exchange:
            ms_alone = timed(ctx, lambda i, k: ctx.render_frame_async(u, opts, k, None))
        if rank == 0:
            print(json.dumps({"variant": v, "world": world, "config": a.config, "slots": S, "bgra8": a.bgra8, "ms_per_frame": round(ms, 4), "per_rank": ranks_ms, "ms_no_exchange": ms_alone and round(ms_alone, 4)}), flush=True)
        dist.barrier()
        ctx.close()
        for kv in v.split(","):
            if "=" in kv:
                os.environ.pop(kv.split("=")[0], None)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
