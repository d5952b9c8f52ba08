"""What makes a C3 frame slower in bench.py than in tools/profile_frame.py --graph? One knob at a time: the asynchronous entry point
(brt_render_frame_async + brt_frame_wait), torch in the process, the in-stream L2 flush. Wall clock around a synchronous wait, per frame."""
import argparse, importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--async", dest="asyn", action="store_true")
ap.add_argument("--torch", action="store_true")
ap.add_argument("--flush", action="store_true")
ap.add_argument("--slots", type=int, default=1)
ap.add_argument("--no-graph", action="store_true", help="two real streams, kernels launched individually (BRT_CFG_NO_GRAPH)")
ap.add_argument("--serial", action="store_true", help="BRT_CFG_NO_OVERLAP: one stream")
a = ap.parse_args()
pkg = importlib.import_module("hardware-ray-tracer_b200")
torch = None
if a.torch or a.flush:
    import torch
    torch.cuda.set_device(0)
    buf = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
cfg = dict(pkg.scenes.CONFIGS[a.config])
scene = pkg.scenes.make_scene(cfg.pop("scene"))
ctx = pkg.Context(device=0, flags=(pkg.CFG_NO_GRAPH if a.no_graph else 0) | (pkg.CFG_NO_OVERLAP if a.serial else 0))
scene.upload(ctx)
w, h = cfg["width"], cfg["height"]
u = scene.uniform(ctx, w, h, 0, cfg["depth_max"])
opts = ctx.opts(w, h, cfg["spp"], cfg["flags"])
ms = []
for f in range(a.frames):
    k = f % a.slots
    if a.flush:
        st = torch.cuda.ExternalStream(ctx.frame_stream(k)) if a.asyn else torch.cuda.current_stream()
        with torch.cuda.stream(st):
            buf.fill_(f & 0xff)
        if not a.asyn:
            torch.cuda.synchronize()
    t0 = time.perf_counter()
    if a.asyn:
        ctx.render_frame_async(u, opts, k)
        ctx.frame_wait(k)
    else:
        ctx.render_frame(u, opts, want_image=False)
    ms.append(round((time.perf_counter() - t0) * 1e3, 2))
print(json.dumps({"lib": os.environ.get("BRT_LIB", "x/lib/libbrt.so").split("/")[-2], "async": a.asyn, "torch": a.torch, "flush": a.flush, "slots": a.slots, "graph": not a.no_graph, "serial": a.serial, "wall_ms": ms}))
