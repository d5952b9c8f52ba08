#!/usr/bin/env python
"""ONE rank's share of a tiled multi-GPU frame, measured on one GPU: the context is created as rank 0 of `--world` ranks, so it traces
every `world`-th 32x32 tile. K frames rotate over `--slots` frame slots (frames in flight), one timed region (CUDA events on the slots'
streams). Shows what bounds a rank of an N-GPU job once its wavefronts are small, without needing N GPUs.

  python tools/bench_rank.py --config c5 --world 8 --slots 1 2 3 4
"""
import argparse
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5")
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0, help="which rank's tiles")
    ap.add_argument("--slots", type=int, nargs="+", default=[1, 2, 3, 4])
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--depth", type=int, default=0, help="override the config's depthMax (rounds per frame)")
    ap.add_argument("--spp", type=int, default=0, help="override the config's samples per pixel (fixed cost per frame vs cost per path)")
    ap.add_argument("--flush-mb", type=int, default=0, help="write this many MiB on the frame's stream before every frame (bench.py: 160)")
    ap.add_argument("--serial", action="store_true", help="also one frame at a time on the serial schedule: per-kernel-class CUDA-event times")
    args = ap.parse_args()
    pkg = importlib.import_module("hardware-ray-tracer_b200")
    cfg = dict(pkg.scenes.CONFIGS[args.config])
    scene = pkg.scenes.make_scene(cfg.pop("scene"))
    dev = torch.device("cuda", 0)
    ctx = pkg.Context(device=0, tile_rank=args.rank, tile_world=args.world, flags=pkg.CFG_NO_GRAPH if args.no_graph else 0)
    scene.upload(ctx)
    w, h = cfg["width"], cfg["height"]
    if args.depth:
        cfg["depth_max"] = args.depth
    u = scene.uniform(ctx, w, h, 0, cfg["depth_max"])
    if args.spp:
        cfg["spp"] = args.spp
    opts = ctx.opts(w, h, cfg["spp"], cfg["flags"])
    flush = torch.empty(args.flush_mb << 20, dtype=torch.uint8, device=dev) if args.flush_mb else None
    out = {"config": args.config, "world": args.world, "rank": args.rank, "flush_mb": args.flush_mb, "graph": not args.no_graph, "runs": []}
    for n_slots in args.slots:
        streams = [torch.cuda.ExternalStream(ctx.frame_stream(k), device=dev) for k in range(n_slots)]

        def submit(i):
            k = i % n_slots
            ctx.frame_wait(k)
            if flush is not None:
                with torch.cuda.stream(streams[k]):
                    flush.fill_(i & 0xff)
            ctx.render_frame_async(u, opts, k, None)

        for i in range(2 * n_slots):
            submit(i)
        for k in range(n_slots):
            ctx.frame_wait(k)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for i in range(args.frames):
            submit(i)
        ends = []
        for k in range(n_slots):
            e = torch.cuda.Event(enable_timing=True)
            e.record(streams[k])
            ends.append(e)
        for k in range(n_slots):
            ctx.frame_wait(k)
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e) for e in ends) / args.frames
        st = ctx.get_stats()
        rays = st.rays_closest + st.rays_occlusion
        out["runs"].append({"slots": n_slots, "ms_per_frame": ms, "rays": int(rays), "mrays_rank": rays / ms / 1e3, "launches": st.launches_total})
    if args.serial:
        sctx = pkg.Context(device=0, tile_rank=args.rank, tile_world=args.world, flags=pkg.CFG_NO_OVERLAP)
        scene.upload(sctx)
        rows = []
        for _ in range(5):
            sctx.render_frame(u, opts, want_image=False)
            st = sctx.get_stats()
            rows.append((st.ms_total, st.ms_trace_closest, st.ms_trace_occlusion, st.ms_shade, st.ms_raygen, st.ms_accumulate, st.ms_resolve))
        rows.sort()
        m = rows[len(rows) // 2]
        out["serial_frame_ms"] = dict(zip(("total", "closest", "occlusion", "shade", "raygen", "accumulate", "resolve"), [round(x, 4) for x in m]))
        out["serial_frame_ms"]["sum_of_kernels"] = round(sum(m[1:]), 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
