#!/usr/bin/env python
"""C4's per-frame dynamic work, the way Scene::prepareRendering drives it: the animated mesh gets new vertices, then ONE call
(brt_smart_cull) rebuilds its BLAS, computes the screen-space footprints and builds the TLAS once; then the frame is traced.

  python tools/c4_dynamic.py [--frames 8]

Prints per frame: BLAS rebuild / Smart Culling / TLAS build (CUDA events inside the library), the host wall time of the prepare
call (vertex upload not included) and the frame's device time."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    args = ap.parse_args()
    pkg = importlib.import_module("hardware-ray-tracer_b200")
    cfg = dict(pkg.scenes.CONFIGS["c4"])
    scene = pkg.scenes.make_scene(cfg.pop("scene"))
    ctx = pkg.Context(device=0)
    scene.upload(ctx)
    w, h = cfg["width"], cfg["height"]
    u = scene.uniform(ctx, w, h, 0, cfg["depth_max"])
    base = scene.meshes[1][1]
    rows = []
    vis = 0
    for f in range(args.frames):
        ctx.mesh_update_vertices(1, pkg.scenes.animate_icosphere(base, f))
        t0 = time.perf_counter()
        vis = ctx.smart_cull(u, w, h, 4.0, 0.25)
        wall = (time.perf_counter() - t0) * 1e3
        s = ctx.get_stats()
        ctx.render_frame(u, ctx.opts(w, h, cfg["spp"], cfg["flags"]), want_image=False)
        s3 = ctx.get_stats()
        rows.append((s.ms_blas_build, s.ms_cull, s.ms_tlas_build, wall, s3.ms_total))
    a = np.array(rows[2:])
    med = np.median(a, axis=0)
    print(json.dumps({"blas_rebuild_ms(20480 tris)": round(float(med[0]), 3), "cull_ms(513 inst)": round(float(med[1]), 3),
                      "tlas_ms": round(float(med[2]), 3), "prepare_wall_ms": round(float(med[3]), 3), "frame_ms": round(float(med[4]), 3),
                      "visible": int(vis), "frames": [[round(float(x), 3) for x in r] for r in rows]}))


if __name__ == "__main__":
    main()
