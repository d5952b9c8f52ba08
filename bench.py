#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native Bloon RT hot path (BASELINE.json metric: Mrays/s).

A *step* is one frame of the workload: every ray query (primary, bounce, shadow) of the frame traced and
shaded by the CUDA path, the scene (BVH, tables) already resident in HBM. `value` = rays of all ranks per
second of device time. `e2e` = the same through the reference-facing C-ABI calls with a HOST framebuffer
(uniform from host memory in, RGBA32F image copied back to pinned host memory inside the timed region).

ONE workload at every N: C3 (BASELINE.json configs[2]: 1M triangles, 4K, 16 spp, 4-bounce GI — the configuration the metric is
quoted on for 1/2/4/8 GPUs; it fits one GPU), so the 1 -> 8 curve is strong scaling of one frame. The other configs ride along as
sub-records under `configs` (N = 1: c1, c2 with its own roofline, c4 with cull / TLAS / rebuild times, c5; N > 1: c5).
N = 1: K frames rotate over the library's frame slots (brt_render_frame_async / brt_frame_wait — the reference keeps
MAX_FRAMES_IN_FLIGHT = 2 frames in flight), every frame a replayed CUDA graph; ONE timed region around all K steps, the L2 flush of
every step enqueued in-stream inside it. `single_frame_latency` is one synchronous brt_render_frame at a time.
N > 1 (torchrun, one process per GPU): the frame is split into 32x32 tiles round-robin over the ranks, scene replicated; the resolve
kernel of every rank stores its pixels into rank 0's gather image through NVLink peer memory, completion by device-side flags, four
frames in flight per rank (--exchange p2p, default), or the packed tiles are all-gathered with NCCL and un-tiled (--exchange nccl).
The gathered frame is compared bit for bit with the same frame rendered by rank 0 alone (`parity_vs_1gpu`, both exchange modes).

  python bench.py                                             # product arm, N = 1
  torchrun --nproc-per-node 8 bench.py --gpus 8               # the same workload on eight GPUs
  python bench.py --impl reference --steps 2 --warmup 1       # CPU arm: the oracle (kind "port") on all host threads

Only the cpu_baseline leg and --impl reference touch oracle/ (the checker); the product arm needs
lib/libbrt.so and a CUDA device and fails loudly without them.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# before anything initialises CUDA: the frames in flight and the fused exchange use more streams than the default 8 hardware work
# queues (hardware-ray-tracer_b200/__init__.py sets the same default)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

NODE_BYTES, TRI_BYTES, INST_BYTES, RAY_IO_BYTES = 80, 48, 96, 48  # algorithmic bytes per visit / per ray


def load_pkg():
    return importlib.import_module("hardware-ray-tracer_b200")


def workload(pkg, name):
    cfg = dict(pkg.scenes.CONFIGS[name])
    scene = pkg.scenes.make_scene(cfg.pop("scene"))
    return scene, cfg


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def traversal_bytes(st, which):
    if which == "closest":
        return (st.nodes_visited_closest * NODE_BYTES + st.prims_tested_closest * TRI_BYTES + st.spheres_tested_closest * INST_BYTES
                + st.rays_closest * RAY_IO_BYTES)
    return (st.nodes_visited_occlusion * NODE_BYTES + st.prims_tested_occlusion * TRI_BYTES + st.spheres_tested_occlusion * INST_BYTES
            + st.rays_occlusion * RAY_IO_BYTES)


def oracle_sample(pkg, scene, cfg, budget_s, threads=0):
    """Times the CPU oracle on a centred crop of the workload sized for about `budget_s` seconds."""
    from oracle import binding as ob
    orc = ob.Oracle(pkg, threads=threads)
    scene.upload(orc)
    w, h = cfg["width"], cfg["height"]
    u = scene.uniform(orc, w, h, 0, cfg["depth_max"])
    cores = orc._f("get_threads")(orc.ctx)

    spp = min(cfg["spp"], 2)  # a bounded sample: the per-ray cost does not depend on the sample count

    def run(cw, ch):
        crop = ((w - cw) // 2, (h - ch) // 2, cw, ch)
        t0 = time.perf_counter()
        orc.render_frame(u, orc.opts(w, h, spp, cfg["flags"], crop))
        dt = time.perf_counter() - t0
        st = orc.get_stats()
        return dt, st.rays_closest + st.rays_occlusion

    dt, rays = run(max(32, w // 16), max(32, h // 16))  # calibration + warm-up
    frac = min(1.0, (budget_s / max(dt, 1e-4)) / 256.0)
    scale = max(1.0 / 16.0, frac ** 0.5)
    cw, ch = max(32, int(w * scale)), max(32, int(h * scale))
    dt, rays = run(cw, ch)
    return {"mrays": rays / dt / 1e6, "cores": int(cores), "seconds": dt, "rays": int(rays), "spp": spp,
            "sample": f"centred {cw}x{ch} crop of the {w}x{h} frame, {spp} of {cfg['spp']} spp, depthMax {cfg['depth_max']}, one pass"}, orc, u


def bench_config(scene, cfg, args):
    """Names the workload. Identical in the product arm and the reference arm (schedules and samples are reported elsewhere)."""
    return {"workload": f"{args.config}: {scene.name} {scene.triangles()} triangles, {cfg['width']}x{cfg['height']}, {cfg['spp']} spp, "
                        f"depthMax {cfg['depth_max']} (primary + {cfg['depth_max'] - 1} bounces), shadow ray per light per hit",
            "render_flags": cfg["flags"], "lights": len(scene.lights), "instances": len(scene.instances),
            "l2": "flushed before every step (write larger than the 126 MB L2)",
            "parallelism": f"image tiles 32x32 round-robin over {args.gpus} GPU(s), scene replicated",
            "workload_choice": "c3 at every N (BASELINE.json configs[2], the configuration quoted for 1/2/4/8 GPUs; it fits one GPU): "
                               "one workload for the whole scaling curve; c1 / c2 / c4 / c5 are sub-records under `configs`"
                               if getattr(args, "config_defaulted", False) else f"--config {args.config}"}


def run_reference(args):
    """--impl reference: the reference cannot run here (Windows/Vulkan/RT cores, no CPU path), so this arm
    times the oracle — the C++ restatement of its shaders — on all host threads, kind 'port'."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = load_pkg()
    scene, cfg = workload(pkg, args.config)
    total = args.steps + args.warmup
    budget = min(20.0, 150.0 / max(total, 1))
    base, orc, u = oracle_sample(pkg, scene, cfg, budget)
    w, h = cfg["width"], cfg["height"]
    cw, ch = [int(x) for x in base["sample"].split()[1].split("x")]
    crop = ((w - cw) // 2, (h - ch) // 2, cw, ch)
    spp = base["spp"]
    times, rays = [], 0
    for i in range(total):
        t0 = time.perf_counter()
        orc.render_frame(u, orc.opts(w, h, spp, cfg["flags"], crop))
        dt = time.perf_counter() - t0
        st = orc.get_stats()
        if i >= args.warmup:
            times.append(dt)
            rays = st.rays_closest + st.rays_occlusion
    ms = 1e3 * float(np.mean(times))
    v = rays / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": bench_config(scene, cfg, args),
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": base["cores"], "kind": "port", "sample": base["sample"] + " per step"},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


class Gpu:
    """What every measurement of the product arm shares: device, torch stream, rank / world, helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product arm has no CPU fallback (use --impl reference for the CPU oracle)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        torch.cuda.set_stream(torch.cuda.Stream(device=self.dev))  # a real (non-default) stream for torch work and frame slot 0
        self.stream = torch.cuda.current_stream()
        self.pkg = load_pkg()
        self.flush = torch.empty(160 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def context(self, flags=0, tiled=True):
        ctx = self.pkg.Context(device=self.local, tile_rank=self.rank if tiled else 0, tile_world=self.world if tiled else 1,
                               flags=flags | (self.pkg.CFG_NO_GRAPH if self.args.no_graph else 0))
        ctx.set_stream(self.stream.cuda_stream)
        return ctx

    def sync_all(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed_sync(self, fn, steps, warmup, collect=None):
        """per-step schedule: flush, barrier + synchronize, events around one synchronous call; max over ranks"""
        torch = self.torch
        for _ in range(warmup):
            fn()
        total_ms = 0.0
        for _ in range(steps):
            self.flush.fill_(1)
            self.sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            fn()
            e1.record(self.stream)
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
            if collect is not None:
                collect()
        return self.max_over_ranks(total_ms) / steps


def pipelined_single(g, ctx, u, opts, n_slots, steps, warmup, hosts=None):
    """N = 1 product schedule: frames in flight (brt_render_frame_async / brt_frame_wait — the reference's MAX_FRAMES_IN_FLIGHT,
    VK/SwapChain.h:8). K frames rotate over the frame slots; ONE timed region brackets all K steps (synchronize on both sides, CUDA
    events on the slots' streams). The L2 flush of every step is enqueued on the frame's stream right before the frame, INSIDE the
    timed region. hosts: pinned buffers, one per slot, for the end-to-end variant (framebuffer copied back inside the region)."""
    torch = g.torch
    streams = [torch.cuda.ExternalStream(ctx.frame_stream(k), device=g.dev) for k in range(n_slots)]

    def submit(i):
        k = i % n_slots
        ctx.frame_wait(k)  # the slot's fence: frame i - n_slots (and its copy to the host) is complete
        with torch.cuda.stream(streams[k]):
            g.flush.fill_(i & 0xff)
        ctx.render_frame_async(u, opts, k, hosts[k].data_ptr() if hosts else None)

    for i in range(max(warmup, n_slots)):  # at least one untimed frame per slot: a slot allocates its wavefront buffers on first use
        submit(i)
    for k in range(n_slots):
        ctx.frame_wait(k)
    g.sync_all()
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for i in range(steps):
        submit(i)
    ends = []
    for k in range(n_slots):
        e = torch.cuda.Event(enable_timing=True)
        e.record(streams[k])
        ends.append(e)
    for k in range(n_slots):
        ctx.frame_wait(k)
    torch.cuda.synchronize()
    return max(e0.elapsed_time(e) for e in ends) / steps


def roofline_record(g, scene, cfg_name, u, opts, steps, tiles_ptr=None):
    """Serial-schedule pass (BRT_CFG_NO_OVERLAP: per-kernel CUDA events, live, same L2 flush between steps) + an instrumented run of
    the same kernels on the same BVH for the visit counters. Returns (kernel_ms dict, roofline dict)."""
    pkg = g.pkg
    k = {"closest": 0.0, "occl": 0.0, "shade": 0.0, "other": 0.0, "n": 0}
    sctx = g.context(pkg.CFG_NO_OVERLAP)
    scene.upload(sctx)

    def step():
        sctx.render_frame(u, opts, want_image=False)

    def collect():
        st = sctx.get_stats()
        k["closest"] += st.ms_trace_closest
        k["occl"] += st.ms_trace_occlusion
        k["shade"] += st.ms_shade
        k["other"] += st.ms_raygen + st.ms_accumulate + st.ms_resolve
        k["n"] += 1

    ms_serial = g.timed_sync(step, steps, 1, collect)
    l2_gbs = sctx.debug_l2_read_gbs(64 << 20, 10) if g.rank == 0 else 0.0
    sctx.close()
    if g.rank != 0:
        return None, None
    cctx = g.context(pkg.CFG_COUNTERS | pkg.CFG_NO_OVERLAP)
    scene.upload(cctx)
    cctx.render_frame(u, opts, want_image=False)
    cst = cctx.get_stats()
    cctx.close()
    n = max(k["n"], 1)
    ms = {"closest": k["closest"] / n, "occlusion": k["occl"] / n}
    peak, peak_src = peaks()
    traffic = issue = None
    tpath, ipath = os.path.join(ROOT, "profiles", "traffic.json"), os.path.join(ROOT, "profiles", "issue.json")
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    ij = json.load(open(ipath)) if os.path.exists(ipath) else {}
    per = {}
    for which in ("closest", "occlusion"):
        kb = traversal_bytes(cst, which)
        kms = ms[which]
        ach = kb / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
        rays = cst.rays_closest if which == "closest" else cst.rays_occlusion
        nodes = cst.nodes_visited_closest if which == "closest" else cst.nodes_visited_occlusion
        prims = cst.prims_tested_closest if which == "closest" else cst.prims_tested_occlusion
        iss = ij.get(cfg_name, {}).get(which)
        lanes = ij.get("active_threads_per_warp_instruction_by_config", {}).get(cfg_name, {}).get(which)
        per[which] = {"kernel": f"k_trace<{'false' if which == 'closest' else 'true'}>", "kernel_ms_per_step": kms, "algorithmic_bytes_per_step": int(kb),
                      "logical_gbs": ach, "frac_of_hbm_peak": ach / peak, "frac_of_l2_read_peak": ach / l2_gbs if l2_gbs > 0 else None,
                      "rays_per_step": int(rays), "gray_per_s": rays / (kms * 1e-3) / 1e9 if kms > 0 else 0.0,
                      "nodes_per_ray": nodes / max(rays, 1), "prims_per_ray": prims / max(rays, 1),
                      "dram_bytes_per_step_ncu": tj.get(cfg_name, {}).get(which),
                      "issue_slot_pct_ncu": iss, "active_lanes_ncu": lanes,
                      "thread_inst_efficiency": (iss / 100.0 * lanes / 32.0) if (iss and lanes) else None}
    which = "closest" if ms["closest"] >= ms["occlusion"] else "occlusion"
    d = per[which]
    roofline = {
        "bound": "issue", "bound_contract": "hbm",
        "kernel": f"{d['kernel']} ({which}-hit traversal)", "achieved": d["logical_gbs"], "peak": peak, "unit": "GB/s", "frac": d["frac_of_hbm_peak"],
        "peak_source": peak_src, "traffic": d["dram_bytes_per_step_ncu"], "algorithmic_bytes_per_step": d["algorithmic_bytes_per_step"],
        "kernel_ms_per_step": d["kernel_ms_per_step"],
        "launches_per_step": int(cst.launches_trace_closest if which == "closest" else cst.launches_trace_occlusion),
        "thread_inst_efficiency": d["thread_inst_efficiency"],
        "l2_read_gbs_measured": l2_gbs, "l2_read_how": "brt_debug_l2_read_gbs: hand-written 16-byte-load read pass over 64 MiB, best of 10, CUDA events",
        "both_traversal_kernels": per, "serial_schedule_ms_per_step": ms_serial,
        "ncu_source": ij.get("source"),
        "note": "achieved / peak / frac follow the contract's formula (algorithmic node + primitive + ray bytes over the kernel's CUDA-event time, "
                "over the measured HBM copy peak): a LOGICAL fetch rate, because the BVH (nodes + triangle records) fits the 126 MB L2 and real DRAM "
                "traffic is a few % of it (`traffic`). What binds is instruction issue (`bound`): thread_inst_efficiency = issue-slot utilisation x "
                "active lanes / 32 from the committed ncu capture (profiles/)."}
    kernel_ms = {"trace_closest": ms["closest"], "trace_occlusion": ms["occlusion"], "shade": k["shade"] / n, "other": k["other"] / n,
                 "schedule": "serial pass (BRT_CFG_NO_OVERLAP); the timed `value` uses the overlapped schedule"}
    return kernel_ms, roofline


def measure_single(g, name, steps, warmup, level):
    """One BASELINE config on ONE GPU (world == 1). level: 'full' (device, e2e, latency, 8-bit, denoised, roofline) or 'lite'."""
    torch, pkg, args = g.torch, g.pkg, g.args
    scene, cfg = workload(pkg, name)
    w, h, spp, flags = cfg["width"], cfg["height"], cfg["spp"], cfg["flags"]
    ctx = g.context()
    t0 = time.perf_counter()
    scene.upload(ctx)
    build_s = time.perf_counter() - t0
    bst = ctx.get_stats()
    u = scene.uniform(ctx, w, h, 0, cfg["depth_max"])
    opts = ctx.opts(w, h, spp, flags)
    n_slots = max(1, args.frames_in_flight)
    hosts = [torch.empty(h * w * 4, dtype=torch.float32).pin_memory() for _ in range(n_slots)]
    ms_dev = pipelined_single(g, ctx, u, opts, n_slots, steps, warmup)
    st = ctx.get_stats()
    rays = st.rays_closest + st.rays_occlusion
    launches_per_step = st.launches_total
    ms_e2e = pipelined_single(g, ctx, u, opts, n_slots, steps, max(2, warmup // 2), hosts)
    rec = {"workload": bench_config(scene, cfg, argparse.Namespace(config=name, gpus=1))["workload"],
           "value": rays / (ms_dev * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_dev, "rays_per_step": int(rays),
           "e2e": {"value": rays / (ms_e2e * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": 140 + 32, "d2h_bytes_per_step": w * h * 16},
           "gpu_launches_per_step": int(launches_per_step), "frames_in_flight": n_slots,
           "scene_build": {"seconds_incl_upload": build_s, "ms_blas": bst.ms_blas_build, "ms_tlas": bst.ms_tlas_build, "bvh_nodes": int(bst.bvh_nodes),
                           "bvh_bytes": int(bst.bvh_bytes), "sah_cost": bst.sah_cost}}
    launches = launches_per_step * (steps + max(warmup, n_slots)) * 2
    if level == "full":
        few = max(2, min(steps, 5))

        def step_device():
            ctx.render_frame(u, opts, want_image=False)

        def step_e2e():
            ctx.render_frame_ptr(u, opts, hosts[0].data_ptr())

        rec["single_frame_latency"] = {"device_ms": g.timed_sync(step_device, few, 1), "e2e_ms": g.timed_sync(step_e2e, few, 1)}
        bgra = pkg.render_format(pkg.FORMAT_BGRA8_UNORM)
        ms8 = pipelined_single(g, ctx, u, ctx.opts(w, h, spp, flags | bgra), n_slots, steps, 2, hosts)
        rec["e2e_bgra8"] = {"value": rays / (ms8 * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms8, "d2h_bytes_per_step": w * h * 4, "format": "B8G8R8A8_UNORM"}
        msd = pipelined_single(g, ctx, u, ctx.opts(w, h, spp, flags | pkg.DENOISE | bgra), n_slots, steps, 3, hosts)
        rec["e2e_denoised_bgra8"] = {"value": rays / (msd * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": msd, "d2h_bytes_per_step": w * h * 4,
                                     "stages": "trace, denoise (temporal + 4 a-trous + bilateral), B8G8R8A8_UNORM, copy"}
        launches += launches_per_step * (2 * few + 2 * steps + 9)
    ctx.close()
    if level == "full":
        kernel_ms, roofline = roofline_record(g, scene, name, u, opts, steps)
        rec["kernel_ms_per_step"], rec["roofline"] = kernel_ms, roofline
        launches += launches_per_step * (steps + 2)
    rec["_launches"] = int(launches)
    rec["_scene"], rec["_cfg"] = scene, cfg
    return rec


def measure_c4_dynamic(g, frames=10):
    """C4's per-frame dynamic work through the one-call entry (Scene::prepareRendering -> brt_smart_cull): new vertices for the animated
    mesh, BLAS rebuild, Smart Culling footprints, TLAS build, then the frame. Median over frames 2.. (CUDA events inside the library)."""
    pkg = g.pkg
    scene, cfg = workload(pkg, "c4")
    ctx = g.context()
    scene.upload(ctx)
    w, h = cfg["width"], cfg["height"]
    u = scene.uniform(ctx, w, h, 0, cfg["depth_max"])
    base = scene.meshes[1][1]
    rows, vis = [], 0
    for f in range(frames):
        ctx.mesh_update_vertices(1, pkg.scenes.animate_icosphere(base, f))
        t0 = time.perf_counter()
        vis = ctx.smart_cull(u, w, h, 4.0, 0.25)
        wall = (time.perf_counter() - t0) * 1e3
        s = ctx.get_stats()
        ctx.render_frame(u, ctx.opts(w, h, cfg["spp"], cfg["flags"]), want_image=False)
        s3 = ctx.get_stats()
        rows.append((s.ms_blas_build, s.ms_cull, s.ms_tlas_build, wall, s3.ms_total, s3.rays_closest + s3.rays_occlusion))
    ctx.close()
    med = np.median(np.array(rows[2:]), axis=0)
    tris = 20480
    return {"blas_rebuild_ms": float(med[0]), "blas_rebuild_mtris_s": tris / (float(med[0]) * 1e-3) / 1e6, "smart_cull_ms": float(med[1]),
            "tlas_build_ms": float(med[2]), "prepare_call_host_wall_ms": float(med[3]), "frame_ms": float(med[4]), "frame_mrays_s": float(med[5]) / float(med[4]) / 1e3,
            "instances": len(scene.instances), "visible": int(vis), "rebuilt_triangles": tris,
            "build_roofline": {"algorithmic_bytes": tris * 320, "ms": float(med[0]), "gbs": tris * 320 / (float(med[0]) * 1e-3) / 1e9,
                               "frac_of_hbm_peak": tris * 320 / (float(med[0]) * 1e-3) / 1e9 / peaks()[0], "bound": "latency (dependent launches), not bytes"}}


def measure_multi(g, name, steps, warmup, exchange):
    """One config tiled over all ranks. Returns the record (rank 0) and the pieces the parity check needs."""
    torch, dist, pkg, args = g.torch, g.dist, g.pkg, g.args
    rank, world, dev = g.rank, g.world, g.dev
    scene, cfg = workload(pkg, name)
    w, h, spp, flags = cfg["width"], cfg["height"], cfg["spp"], cfg["flags"]
    ctx = g.context()
    t0 = time.perf_counter()
    scene.upload(ctx)
    build_s = time.perf_counter() - t0
    bst = ctx.get_stats()
    u = scene.uniform(ctx, w, h, 0, cfg["depth_max"])
    opts = ctx.opts(w, h, spp, flags)
    opts8 = ctx.opts(w, h, spp, flags | pkg.render_format(pkg.FORMAT_BGRA8_UNORM))
    mode = exchange
    try:
        frame = pkg.TiledFrame(ctx, w, h, rank, world, dev, mode=exchange, root_only=True, root=0)
    except Exception as e:  # peer memory not available on this box: NCCL all-gather + un-tile
        if exchange != "p2p":
            raise
        mode = f"nccl (p2p unavailable: {type(e).__name__})"
        frame = pkg.TiledFrame(ctx, w, h, rank, world, dev, mode="nccl")
    S = 4
    rec = {"exchange": mode}
    last_host = None
    if frame.mode == "p2p":
        streams = [torch.cuda.ExternalStream(ctx.frame_stream(k), device=dev) for k in range(S)]
        hosts = [torch.empty(h * w * 4, dtype=torch.float32).pin_memory() for _ in range(S)] if rank == 0 else None

        def run(n, o, to_host):
            for i in range(n):
                k = i % S
                if rank == 0 and to_host:
                    frame.wait_fetch(k)  # the host buffer of this slot still receives frame i - S
                with torch.cuda.stream(streams[k]):
                    g.flush.fill_(i & 0xff)
                frame.submit(u, o, k)
                if rank == 0:
                    if to_host:
                        frame.fetch_async(k, hosts[k])
                    else:
                        frame.release(k)
            for k in range(S):
                ctx.frame_wait(k)
            if rank == 0:
                frame.wait_fetch()
                frame._copy_stream.synchronize()

        def timed_region(n, wu, o, to_host):
            run(max(wu, S), o, to_host)
            g.sync_all()
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record(streams[0])
            run(n, o, to_host)
            ends = []
            for st_ in streams + ([frame._copy_stream] if rank == 0 else []):
                e = torch.cuda.Event(enable_timing=True)
                e.record(st_)
                ends.append(e)
            torch.cuda.synchronize()
            return g.max_over_ranks(max(e0.elapsed_time(e) for e in ends)) / n

        ms_dev = timed_region(steps, warmup, opts, False)
        st = ctx.get_stats()
        ms_e2e = timed_region(steps, max(2, warmup // 2), opts, True)
        if rank == 0:
            last_host = hosts[(steps - 1) % S].clone()  # (the 8-bit pass below reuses the host buffers)
        ms_e2e8 = timed_region(steps, 2, opts8, True)
        frame.check()
        rec["schedule"] = (f"one timed region around all K steps: {S} frames in flight per rank (slot k <-> gather image k on rank 0), the frame's last kernels store its tiles "
                           "straight into rank 0's image over NVLink, completion by device-side flags (no collective, no host barrier); rank 0 copies every "
                           "frame to pinned host memory on a side stream inside the region (e2e); L2 flushed in-stream before every frame")
        rec["frames_in_flight"] = S
    else:
        host = torch.empty(h * w * 4, dtype=torch.float32).pin_memory()

        def step_device():
            frame.render(u, opts)

        def step_e2e():
            frame.render(u, opts)
            if rank == 0:
                frame.to_host(host)
            g.stream.synchronize()

        ms_dev = g.timed_sync(step_device, steps, warmup)
        st = ctx.get_stats()
        ms_e2e = g.timed_sync(step_e2e, steps, max(1, warmup // 2))
        ms_e2e8 = None
        last_host = host
        rec["schedule"] = "per step: trace own tiles, NCCL all-gather of the packed tiles, un-tile; barrier + synchronize around every step"
        rec["frames_in_flight"] = 1
    rays = g.sum_over_ranks(st.rays_closest + st.rays_occlusion)
    launches_per_step = st.launches_total + (0 if frame.mode == "p2p" else 1)
    rec.update({"workload": bench_config(scene, cfg, argparse.Namespace(config=name, gpus=world))["workload"],
                "value": rays / (ms_dev * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_dev, "rays_per_step": int(rays),
                "e2e": {"value": rays / (ms_e2e * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": 140 + 32, "d2h_bytes_per_step": w * h * 16},
                "e2e_bgra8": ({"value": rays / (ms_e2e8 * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_e2e8, "d2h_bytes_per_step": w * h * 4,
                               "format": "B8G8R8A8_UNORM (converted in the resolve kernel, 4 bytes per pixel over NVLink and PCIe)"} if ms_e2e8 else None),
                "gpu_launches_per_step_per_rank": int(launches_per_step),
                "scene_build": {"seconds_incl_upload": build_s, "ms_blas": bst.ms_blas_build, "ms_tlas": bst.ms_tlas_build, "bvh_nodes": int(bst.bvh_nodes)}})
    rec["_launches"] = int(launches_per_step * (2 * steps + 2 * max(warmup, S) + (steps + S if ms_e2e8 else 0)))
    # ---- the same frame by rank 0 alone: strong-scaling reference (>= 10 frames, same frames-in-flight schedule) and bit-exact parity
    dist.barrier()
    parity = {}
    if rank == 0:
        solo = g.context(tiled=False)
        scene.upload(solo)
        ref = solo.render_frame(u, opts)
        got = last_host.numpy().reshape(h, w, 4)
        parity[frame.mode] = bool(np.array_equal(got.view(np.uint32), ref.view(np.uint32)))
        n1 = max(10, min(steps, 20))
        ms1 = pipelined_single(g_single_view(g), solo, u, opts, max(1, args.frames_in_flight), n1, 3)
        st1 = solo.get_stats()
        v1 = (st1.rays_closest + st1.rays_occlusion) / (ms1 * 1e-3) / 1e6
        rec["same_workload_1gpu"] = {"value": v1, "unit": "Mrays/s", "ms_per_step": ms1, "frames": n1, "speedup": rec["value"] / v1,
                                     "efficiency": rec["value"] / v1 / world}
        rec["_launches"] += int(st1.launches_total * (n1 + 4))
        solo.close()
    dist.barrier()
    # the other exchange mode once, outside any timed region, for the parity record
    other = "nccl" if frame.mode == "p2p" else "p2p"
    try:
        ctx2 = g.context()
        scene.upload(ctx2)
        f2 = pkg.TiledFrame(ctx2, w, h, rank, world, dev, mode=other, root_only=True, root=0)
        img = f2.render(u, opts)
        if rank == 0:
            if other == "nccl":
                torch.cuda.synchronize()
                got = img.cpu().numpy()
            else:
                hb = torch.empty(h * w * 4, dtype=torch.float32).pin_memory()
                f2.to_host(hb)
                got = hb.numpy().reshape(h, w, 4)
            parity[other] = bool(np.array_equal(got.view(np.uint32), ref.view(np.uint32)))
        dist.barrier()
        ctx2.close()
    except Exception as e:
        parity[other] = f"not run: {type(e).__name__}: {e}"
    ctx.close()
    rec["parity_vs_1gpu"] = parity
    return rec, scene, cfg


def g_single_view(g):
    """`g` as a one-rank world (rank 0 measuring alone while the others wait at a barrier)."""
    import copy
    v = copy.copy(g)
    v.world = 1
    return v


def run_product(args):
    g = Gpu(args)
    rank, world = g.rank, g.world
    sampler = ClockSampler(g.local) if rank == 0 else None
    line = None
    if world == 1:
        main_rec = measure_single(g, args.config, args.steps, args.warmup, "full")
        clocks = sampler.stop() if sampler else None
        scene, cfg = main_rec.pop("_scene"), main_rec.pop("_cfg")
        launches = main_rec.pop("_launches")
        cpu = None
        if not args.no_cpu_baseline:
            b, _, _ = oracle_sample(g.pkg, scene, cfg, 12.0)
            cpu = {"value": b["mrays"], "unit": "Mrays/s", "cores": b["cores"], "kind": "port", "sample": b["sample"], "seconds": b["seconds"]}
        configs = {}
        if args.config_defaulted and not args.no_sub_records:
            sub_steps = max(5, min(args.steps, 20))
            for name, level in (("c2", "full"), ("c1", "lite"), ("c5", "lite")):
                r = measure_single(g, name, sub_steps, 3, level)
                r.pop("_scene"), r.pop("_cfg")
                launches += r.pop("_launches")
                configs[name] = r
            configs["c4"] = measure_c4_dynamic(g)
            r = measure_single(g, "c4", sub_steps, 3, "lite")
            r.pop("_scene"), r.pop("_cfg")
            launches += r.pop("_launches")
            configs["c4"]["static_frame"] = r
        line = {
            "metric": "Mrays/s", "value": main_rec["value"], "unit": "Mrays/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(scene, cfg, args),
            "schedule": f"K frames rotate over {main_rec['frames_in_flight']} frame slots (brt_render_frame_async / brt_frame_wait; the reference keeps "
                        "MAX_FRAMES_IN_FLIGHT = 2 frames in flight), every frame a replayed CUDA graph; one timed region around all K steps, "
                        "160 MiB L2 flush enqueued on the frame's stream before every frame inside the region",
            "rays_per_step": main_rec["rays_per_step"], "single_frame_latency": main_rec.get("single_frame_latency"),
            "e2e": main_rec["e2e"], "e2e_bgra8": main_rec.get("e2e_bgra8"), "e2e_denoised_bgra8": main_rec.get("e2e_denoised_bgra8"),
            "gpu_launches": int(launches), "gpu_launches_per_step": main_rec["gpu_launches_per_step"],
            "kernel_ms_per_step": main_rec.get("kernel_ms_per_step"), "roofline": main_rec.get("roofline"), "cpu_baseline": cpu, "clocks": clocks,
            "scene_build": main_rec["scene_build"], "configs": configs,
        }
    else:
        main_rec, scene, cfg = measure_multi(g, args.config, args.steps, args.warmup, args.exchange)
        clocks = sampler.stop() if sampler else None
        launches = main_rec.pop("_launches")
        configs = {}
        if args.config_defaulted and not args.no_sub_records:
            r, _, _ = measure_multi(g, "c5", max(10, min(args.steps, 40)), 4, args.exchange)
            launches += r.pop("_launches")
            configs["c5"] = r
        if rank == 0:
            line = {
                "metric": "Mrays/s", "value": main_rec["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": main_rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": bench_config(scene, cfg, args), "exchange": main_rec["exchange"], "schedule": main_rec["schedule"],
                "rays_per_step": main_rec["rays_per_step"], "e2e": main_rec["e2e"], "e2e_bgra8": main_rec["e2e_bgra8"],
                "gpu_launches": int(launches) * world, "gpu_launches_per_step_per_rank": main_rec["gpu_launches_per_step_per_rank"],
                "parity_vs_1gpu": main_rec["parity_vs_1gpu"], "same_workload_1gpu": main_rec.get("same_workload_1gpu"),
                "roofline": None, "cpu_baseline": None, "clocks": clocks, "scene_build": main_rec["scene_build"], "configs": configs,
                "roofline_note": "per-kernel roofline and cpu_baseline are reported by the N = 1 line (same kernels, same workload)",
            }
    ok = True
    if world > 1:
        g.dist.barrier()
        g.dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
        if world > 1:
            par = [line["parity_vs_1gpu"]] + [c.get("parity_vs_1gpu", {}) for c in line["configs"].values()]
            ok = all(v is not False for p in par for v in p.values())
    if not ok:
        raise SystemExit("bench.py: a gathered multi-GPU frame differs from the single-GPU frame (parity_vs_1gpu)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None, choices=["c1", "c2", "c3", "c4", "c5"],
                    help="BASELINE config; default: c3 (4K, 16 spp, 4-bounce GI: the configuration BASELINE.json quotes for 1/2/4/8 GPUs) at every N, "
                         "with the other configs as sub-records")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true", help="skip the c1 / c2 / c4 / c5 sub-records of the default run")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels of every frame individually (default: replay the frame's CUDA graph)")
    ap.add_argument("--frames-in-flight", type=int, default=3, choices=[1, 2, 3, 4],
                    help="N = 1: frames rotate over this many frame slots (the reference keeps 2 frames in flight over a swapchain of "
                         "typically 3 images); 1 = one frame at a time")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: framebuffer exchange — p2p = resolve kernel stores into rank 0's frame through NVLink peer memory, device-side "
                         "completion flags, four frames in flight (fused, default); nccl = all-gather of packed tiles + un-tile kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.config_defaulted = args.config is None
    if args.config is None:
        args.config = "c3"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
