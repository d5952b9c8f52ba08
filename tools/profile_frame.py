#!/usr/bin/env python
"""Renders a few frames of one BASELINE config through the C ABI — the short command the ncu captures wrap.

  python tools/profile_frame.py --config c2 --frames 3 [--counters] [--no-treelet]
"""
import argparse
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--counters", action="store_true")
    ap.add_argument("--no-treelet", action="store_true")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--greedy-collapse", action="store_true")
    ap.add_argument("--hit-sort", action="store_true", help="bounce rounds shaded in the order of their hit positions (BRT_CFG_HIT_SORT)")
    ap.add_argument("--graph", action="store_true", help="product schedule: two streams + replayed frame graph (total time only)")
    ap.add_argument("--fast-shading", action="store_true", help="BRT_RENDER_FAST_SHADING")
    ap.add_argument("--flush", action="store_true", help="write 256 MiB (larger than the L2) before every frame, like bench.py")
    ap.add_argument("--spp", type=int, default=0, help="override the config's samples per pixel")
    args = ap.parse_args()
    pkg = importlib.import_module("hardware-ray-tracer_b200")
    cfg = dict(pkg.scenes.CONFIGS[args.config])
    scene = pkg.scenes.make_scene(cfg.pop("scene"), small=args.small)
    # per-kernel event times need individually launched kernels
    flags = ((0 if args.graph else pkg.CFG_NO_GRAPH) | (pkg.CFG_HIT_SORT if args.hit_sort else 0) | (pkg.CFG_COUNTERS if args.counters else 0) | (pkg.CFG_NO_TREELET if args.no_treelet else 0)
             | (pkg.CFG_NO_OVERLAP if args.no_overlap else 0) | (pkg.CFG_GREEDY_COLLAPSE if args.greedy_collapse else 0))
    ctx = pkg.Context(device=0, flags=flags)
    t0 = time.perf_counter()
    scene.upload(ctx)
    build_s = time.perf_counter() - t0
    st = ctx.get_stats()
    out = {"config": args.config, "build_s": build_s, "ms_blas": st.ms_blas_build, "ms_tlas": st.ms_tlas_build, "bvh_nodes": st.bvh_nodes,
           "sah": st.sah_cost, "sah_lbvh": st.sah_cost_lbvh, "frames": []}
    w, h = cfg["width"], cfg["height"]
    u = scene.uniform(ctx, w, h, 0, cfg["depth_max"])
    if args.spp:
        cfg["spp"] = args.spp
    if args.fast_shading:
        cfg["flags"] |= pkg.FAST_SHADING
    rt, flush_buf = None, None
    if args.flush:
        import ctypes
        rt = ctypes.CDLL("libcudart.so")
        flush_buf = ctypes.c_void_p()
        assert rt.cudaMalloc(ctypes.byref(flush_buf), ctypes.c_size_t(256 << 20)) == 0
    for f in range(args.frames):
        if rt is not None:
            assert rt.cudaMemset(flush_buf, f & 0xff, ctypes.c_size_t(256 << 20)) == 0 and rt.cudaDeviceSynchronize() == 0
        ctx.render_frame(u, ctx.opts(w, h, cfg["spp"], cfg["flags"]), want_image=False)
        s = ctx.get_stats()
        rays = s.rays_closest + s.rays_occlusion
        out["frames"].append({"ms_total": s.ms_total, "closest": s.ms_trace_closest, "occl": s.ms_trace_occlusion, "shade": s.ms_shade,
                              "raygen": s.ms_raygen, "accum": s.ms_accumulate, "resolve": s.ms_resolve, "rays": rays,
                              "mrays": rays / max(s.ms_total, 1e-9) / 1e3, "nodes_c": s.nodes_visited_closest, "prims_c": s.prims_tested_closest,
                              "nodes_o": s.nodes_visited_occlusion, "prims_o": s.prims_tested_occlusion,
                              "rays_c": s.rays_closest, "rays_o": s.rays_occlusion})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
