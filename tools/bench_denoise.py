#!/usr/bin/env python
"""Times the denoiser slot (brt_denoise) on a BASELINE config and reports it against the HBM roofline.

  python tools/bench_denoise.py --config c2 [--frames 6]

Algorithmic bytes per pixel: temporal pass 16 colour + 40 G-buffer + 44 history read, 40 written = 140;
every a-trous / bilateral pass 16 + 24 G-buffer read, 16 written = 56 (the 25 taps of a pass are shared between neighbouring
lanes through L1 / L2)."""
import argparse
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--iterations", type=int, default=4)
    args = ap.parse_args()
    pkg = importlib.import_module("hardware-ray-tracer_b200")
    cfg = dict(pkg.scenes.CONFIGS[args.config])
    scene = pkg.scenes.make_scene(cfg.pop("scene"))
    ctx = pkg.Context(device=0)
    scene.upload(ctx)
    w, h = cfg["width"], cfg["height"]
    dop = ctx.denoise_opts(iterations=args.iterations, flags=pkg.DENOISE_BILATERAL)
    ms = []
    for f in range(args.frames):
        u = scene.uniform(ctx, w, h, f, cfg["depth_max"])
        ctx.render_frame(u, ctx.opts(w, h, 1, cfg["flags"] | pkg.GBUFFER | pkg.JITTER), want_image=False)
        ctx.denoise(u, dop, w, h, want_image=False)
        st = ctx.get_stats()
        ms.append(st.ms_denoise)
    passes = args.iterations + 1
    bytes_px = 140 + 56 * passes
    best = min(ms[1:]) if len(ms) > 1 else ms[0]
    peak = 6556.2
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    gbs = w * h * bytes_px / (best * 1e-3) / 1e9
    print(json.dumps({"config": args.config, "width": w, "height": h, "passes": 1 + passes, "ms_denoise": ms, "best_ms": best,
                      "algorithmic_bytes_per_pixel": bytes_px, "achieved_gbs": gbs, "hbm_peak_gbs": peak, "frac": gbs / peak,
                      "frame_ms": st.ms_total}))


if __name__ == "__main__":
    main()
