#!/usr/bin/env python
"""Steady-state BVH build times: rebuilds the BLAS of the 1M-triangle terrain scene and of the C4 animated mesh."""
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("hardware-ray-tracer_b200")

out = {}
for name, flags in (("treelets", pkg.CFG_TREELET_ON_REBUILD), ("lbvh_only", pkg.CFG_NO_TREELET)):
    scene = pkg.scenes.terrain_icospheres()
    ctx = pkg.Context(device=0, flags=flags)
    scene.upload(ctx)
    hv = scene.meshes[0][1]
    ms = []
    for it in range(5):
        ctx.mesh_update_vertices(0, hv)
        t0 = time.perf_counter()
        ctx.scene_build()
        wall = (time.perf_counter() - t0) * 1e3
        st = ctx.get_stats()
        ms.append((round(st.ms_blas_build, 3), round(st.ms_tlas_build, 3), round(wall, 3)))
    out[name + "_524k_tris(ms_blas,ms_tlas,wall_ms)"] = ms
    out[name + "_sah"] = (st.sah_cost_lbvh, st.sah_cost)
    ctx.close()
scene = pkg.scenes.instanced_lattice()
ctx = pkg.Context(device=0)
scene.upload(ctx)
base = scene.meshes[1][1]
ms = []
for it in range(6):
    v = pkg.scenes.animate_icosphere(base, it)
    ctx.mesh_update_vertices(1, v)
    t0 = time.perf_counter()
    ctx.scene_build()
    wall = (time.perf_counter() - t0) * 1e3
    st = ctx.get_stats()
    ms.append((round(st.ms_blas_build, 3), round(st.ms_tlas_build, 3), round(wall, 3)))
out["c4_20k_tri_blas_rebuild+513_inst_tlas(ms_blas,ms_tlas,wall_ms)"] = ms
print(json.dumps(out, indent=1))
