import importlib, sys, numpy as np
sys.path.insert(0,'/root/repo')
pkg = importlib.import_module("hardware-ray-tracer_b200")
cfg = dict(pkg.scenes.CONFIGS["c4"]); scene = pkg.scenes.make_scene(cfg.pop("scene"))
ctx = pkg.Context(device=0); scene.upload(ctx)
w,h = cfg["width"], cfg["height"]
u = scene.uniform(ctx, w, h, 0, cfg["depth_max"])
base = scene.meshes[1][1]
cull, tlas, blas, fr = [], [], [], []
for f in range(6):
    ctx.mesh_update_vertices(1, pkg.scenes.animate_icosphere(base, f)); ctx.scene_build(); s1 = ctx.get_stats()
    vis = ctx.smart_cull(u, w, h, 4.0, 0.25); s2 = ctx.get_stats()
    ctx.render_frame(u, ctx.opts(w, h, cfg["spp"], cfg["flags"]), want_image=False); s3 = ctx.get_stats()
    blas.append(s1.ms_blas_build); cull.append(s2.ms_cull); tlas.append(s2.ms_tlas_build); fr.append(s3.ms_total)
print("blas", np.round(blas,3), "cull", np.round(cull,3), "tlas", np.round(tlas,3), "frame", np.round(fr,3), "visible", vis)
