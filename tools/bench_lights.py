#!/usr/bin/env python
"""Many-light scaling: the reference's O(numLights) loop of calculateColor (SH/raytracing.slang:72-88) against the light BVH
(LightBVHNode, RT/Scene.h:123-130: one importance-sampled light and one shadow ray per hit) on the C2 scene.

  python tools/bench_lights.py [--config c2]"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    args = ap.parse_args()
    pkg = importlib.import_module("hardware-ray-tracer_b200")
    cfg = dict(pkg.scenes.CONFIGS[args.config])
    w, h, spp, flags, depth = cfg["width"], cfg["height"], cfg["spp"], cfg["flags"], cfg["depth_max"]
    rows = []
    for n_lights in (3, 16, 256, 4096):
        scene = pkg.scenes.make_scene(cfg["scene"])
        ctx = pkg.Context(device=0)
        scene.upload(ctx, build=False)
        rng = np.random.default_rng(1)
        for _ in range(n_lights - len(scene.lights)):
            pos = (rng.random(3) * np.array([14.0, 2.0, 14.0]) - np.array([7.0, 5.0, 7.0])).tolist()
            ctx.light_create(pos, (0.3 + 0.7 * rng.random(3)).tolist(), float(0.5 + 2.0 * rng.random()))
        ctx.scene_build()
        u = scene.uniform(ctx, w, h, 0, depth)
        for mode, mflag in (("loop", 0), ("light_bvh", pkg.LIGHT_BVH)):
            if mode == "loop" and n_lights > 16:
                continue
            ms = []
            for _ in range(4):
                ctx.render_frame(u, ctx.opts(w, h, spp, flags | mflag), want_image=False)
                st = ctx.get_stats()
                ms.append(st.ms_total)
            rows.append({"lights": n_lights, "mode": mode, "ms_frame": round(float(np.median(ms[1:])), 3), "rays_closest": int(st.rays_closest),
                         "rays_occlusion": int(st.rays_occlusion), "mrays_s": round((st.rays_closest + st.rays_occlusion) / np.median(ms[1:]) / 1e3, 1)})
        ctx.close()
    print(json.dumps(rows, indent=1))
    print("\n| lights | mode | ms/frame | closest rays | shadow rays | Mrays/s |\n|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['lights']} | {r['mode']} | {r['ms_frame']} | {r['rays_closest']} | {r['rays_occlusion']} | {r['mrays_s']} |")


if __name__ == "__main__":
    main()
