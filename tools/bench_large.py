#!/usr/bin/env python
"""Traversal outside the L2: the C2 frame (1920x1080, primary + 2 bounces, 3 lights) on terrains of 8.9 M and 34 M unique
triangles, whose node and triangle records (0.5 / 2.0 GB) do not fit the 126 MB L2 — the regime where SURVEY.md §8(d) expects the
traversal to be bound by HBM. (Measured, profiles/r1c_large_scenes.md: it is not — the rays of a frame share enough of the tree that
DRAM runs at 9-12 % of its peak and the kernels stay bound by instruction issue; the frame costs 25-40 % more than on the 1M scene.)

  python tools/bench_large.py [--n 2048 4096] [--frames 5]

Per scene: build time, nodes, frame and per-kernel times (serial schedule, CUDA events), visit counters of an instrumented run of
the same kernels, algorithmic bytes (80 B per node, 48 B per triangle record, 96 B per instance entered, 48 B per ray) and the
logical fetch rate against MEASURED_PEAKS.json. Wrap the same command in `ncu --metrics dram__bytes_read.sum,...
-k regex:k_trace` (with --frames 2 --no-counters) for the real DRAM traffic next to it.
"""
import argparse
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
NODE_BYTES, TRI_BYTES, INST_BYTES, RAY_IO_BYTES = 80, 48, 96, 48  # as bench.py / DESIGN.md §5


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[2048, 4096], help="heightfield resolution: 2 n^2 terrain triangles")
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--no-counters", action="store_true")
    args = ap.parse_args()
    pkg = importlib.import_module("hardware-ray-tracer_b200")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = None
    try:
        peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = None
    if isinstance(peak, dict):
        for k in ("hbm_gbs_burst", "hbm_gbs", "hbm_copy_gbs"):
            if k in peak:
                hbm = float(peak[k])
                break
        if hbm is None:
            for k, v in peak.items():
                if "hbm" in k.lower() and isinstance(v, (int, float)):
                    hbm = float(v)
                    break
    w, h, depth, rflags = 1920, 1080, 3, pkg.BOUNCE_REFLECT | pkg.BOUNCE_REFRACT
    for n in args.n:
        t0 = time.perf_counter()
        scene = pkg.scenes.terrain_icospheres(n=n)
        gen_s = time.perf_counter() - t0
        out = {"terrain_n": n, "triangles": scene.triangles(), "generate_s": round(gen_s, 2)}
        runs = [("timed", pkg.CFG_NO_OVERLAP)] + ([] if args.no_counters else [("counters", pkg.CFG_NO_OVERLAP | pkg.CFG_COUNTERS)])
        for label, flags in runs:
            ctx = pkg.Context(device=0, flags=flags)
            t0 = time.perf_counter()
            scene.upload(ctx)
            up_s = time.perf_counter() - t0
            st = ctx.get_stats()
            u = scene.uniform(ctx, w, h, 0, depth)
            frames = []
            for f in range(args.frames):
                ctx.render_frame(u, ctx.opts(w, h, 1, rflags), want_image=False)
                frames.append(ctx.get_stats())
            s = frames[-1]
            if label == "timed":
                med = sorted(frames[1:], key=lambda x: x.ms_total)[len(frames[1:]) // 2]
                out.update({"upload_build_s": round(up_s, 2), "ms_blas": st.ms_blas_build, "bvh_nodes": st.bvh_nodes,
                            "bvh_bytes": st.bvh_bytes, "sah": st.sah_cost,
                            "ms_frame": med.ms_total, "ms_closest": med.ms_trace_closest, "ms_occlusion": med.ms_trace_occlusion,
                            "ms_shade": med.ms_shade, "rays": med.rays_closest + med.rays_occlusion,
                            "mrays_s": (med.rays_closest + med.rays_occlusion) / max(med.ms_total, 1e-9) / 1e3})
            else:
                bc = (s.nodes_visited_closest * NODE_BYTES + s.prims_tested_closest * TRI_BYTES + s.spheres_tested_closest * INST_BYTES
                      + s.rays_closest * RAY_IO_BYTES)
                bo = (s.nodes_visited_occlusion * NODE_BYTES + s.prims_tested_occlusion * TRI_BYTES
                      + s.spheres_tested_occlusion * INST_BYTES + s.rays_occlusion * RAY_IO_BYTES)
                out.update({"nodes_per_ray_closest": s.nodes_visited_closest / max(1, s.rays_closest),
                            "prims_per_ray_closest": s.prims_tested_closest / max(1, s.rays_closest),
                            "nodes_per_ray_occlusion": s.nodes_visited_occlusion / max(1, s.rays_occlusion),
                            "prims_per_ray_occlusion": s.prims_tested_occlusion / max(1, s.rays_occlusion),
                            "algorithmic_bytes_closest": int(bc), "algorithmic_bytes_occlusion": int(bo)})
                if "ms_closest" in out and out["ms_closest"] > 0:
                    out["logical_gbs_closest"] = bc / out["ms_closest"] / 1e6
                    out["logical_gbs_occlusion"] = bo / out["ms_occlusion"] / 1e6
                    if hbm:
                        out["hbm_peak_gbs"] = hbm
                        out["frac_closest"] = out["logical_gbs_closest"] / hbm
                        out["frac_occlusion"] = out["logical_gbs_occlusion"] / hbm
            ctx.close()
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
