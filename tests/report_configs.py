#!/usr/bin/env python
"""TEST INFRASTRUCTURE (it uses the oracle as the checker, which only tests/ may do): runs the five BASELINE.json configurations on one
GPU and prints the table BASELINE.md §3 asks for (ms/frame, Mrays/s, build / cull times for C4, the CPU oracle on a bounded crop, parity
on that crop).   python tests/report_configs.py"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("hardware-ray-tracer_b200")
from oracle import binding as ob  # noqa: E402  (checker only)
from util import compare_frames  # noqa: E402

rows = []
for name in ("c1", "c2", "c3", "c4", "c5"):
    cfg = dict(pkg.scenes.CONFIGS[name])
    scene = pkg.scenes.make_scene(cfg.pop("scene"))
    w, h, spp, flags, depth = cfg["width"], cfg["height"], cfg["spp"], cfg["flags"], cfg["depth_max"]
    ctx = pkg.Context(device=0)
    t0 = time.perf_counter()
    scene.upload(ctx)
    build_wall = time.perf_counter() - t0
    bst = ctx.get_stats()
    u = scene.uniform(ctx, w, h, 0, depth)
    extra = {}
    if name == "c4":  # per frame: animate one mesh, rebuild its BLAS (LBVH), Smart Culling, TLAS rebuild
        base = scene.meshes[1][1]
        cull, tlas, blas = [], [], []
        for f in range(4):
            ctx.mesh_update_vertices(1, pkg.scenes.animate_icosphere(base, f))
            vis = ctx.smart_cull(u, w, h, 4.0, 0.25)  # one call: BLAS of the updated mesh, footprints, TLAS (Scene::prepareRendering)
            s2 = ctx.get_stats()
            blas.append(s2.ms_blas_build); cull.append(s2.ms_cull); tlas.append(s2.ms_tlas_build)
        extra = {"blas_rebuild_ms(20480 tris)": round(float(np.median(blas)), 3), "cull_ms(513 inst)": round(float(np.median(cull)), 3),
                 "tlas_ms": round(float(np.median(tlas)), 3), "visible": vis}
    ms, rays = [], 0
    for f in range(4):
        ctx.render_frame(u, ctx.opts(w, h, spp, flags), want_image=False)
        st = ctx.get_stats()
        ms.append(st.ms_total)
        rays = st.rays_closest + st.rays_occlusion
    ms_frame = float(np.median(ms[1:]))
    # oracle on a bounded crop, same uniform: parity + CPU rate
    orc = ob.Oracle(pkg)
    scene.upload(orc)
    if name == "c4":
        orc.mesh_update_vertices(1, pkg.scenes.animate_icosphere(scene.meshes[1][1], 3))
        orc.scene_build()
        orc.smart_cull(u, w, h, 4.0, 0.25)
    cw, ch = min(w, 256), min(h, 144)
    crop = ((w - cw) // 2, (h - ch) // 2, cw, ch)
    ospp = min(spp, 2)
    t0 = time.perf_counter()
    r = compare_frames(pkg, ctx, orc, u, w, h, flags, ospp, crop)
    ost = r["ref_stats"]
    # time the oracle alone
    t0 = time.perf_counter()
    orc.render_frame(u, orc.opts(w, h, ospp, flags, crop))
    odt = time.perf_counter() - t0
    rows.append({"config": name, "scene": scene.name, "triangles": int(bst.total_triangles), "frame": f"{w}x{h} spp{spp} depth{depth}",
                 "ms_frame": round(ms_frame, 3), "rays_frame": int(rays), "mrays_s": round(rays / ms_frame / 1e3, 1),
                 "build_ms(blas+tlas)": round(bst.ms_blas_build + bst.ms_tlas_build, 2), "sah_lbvh->sah": f"{bst.sah_cost_lbvh:.1f}->{bst.sah_cost:.1f}",
                 "oracle_mrays_s": round((ost.rays_closest + ost.rays_occlusion) / odt / 1e6, 2), "oracle_threads": orc._f("get_threads")(orc.ctx),
                 "crop_id_agreement": r["id_agreement"], "crop_rel_rmse": r["rmse"], "crop_bit_exact": r["bit_exact"], **extra})
    ctx.close()
print(json.dumps(rows, indent=1))
keys = ["config", "triangles", "frame", "ms_frame", "rays_frame", "mrays_s", "build_ms(blas+tlas)", "sah_lbvh->sah", "oracle_mrays_s", "oracle_threads",
        "crop_id_agreement", "crop_rel_rmse", "crop_bit_exact"]
print("\n| " + " | ".join(keys) + " |\n|" + "---|" * len(keys))
for r in rows:
    print("| " + " | ".join(str(r[k]) for k in keys) + " |")
for r in rows:
    if r["config"] == "c4":
        print("\nC4 per-frame dynamic work:", {k: v for k, v in r.items() if k not in keys and k not in ("scene",)})
