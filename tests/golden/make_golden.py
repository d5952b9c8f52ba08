"""Generates tests/golden/{rng_kat.json, oracle_frames.npz}.

The reference has no tests, golden vectors or CPU path — "parity unpinned". These
files pin what CAN be pinned: the integer KATs of SH/random.slang (checked against
an independent pure-Python evaluation in tests/test_oracle_kat.py) and small frames rendered by the
oracle itself (regression pins, so that an accidental change of the oracle is caught).
Run:  python tests/golden/make_golden.py
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

KAT = {
    "hash": [
        {"in": [0, 0, 0], "hash": "a5e2b579", "pcg1": "19492af8", "state1": "a2fdd892", "pcg2": "aa455eaa", "rand1": 0.0987727046},
        {"in": [1, 0, 0], "hash": "53406116", "pcg1": "1866fb63", "state1": "9c882993", "pcg2": "0f6c0c41", "rand1": 0.0953213796},
        {"in": [0, 1, 0], "hash": "5038d863", "pcg1": "b1305a1c", "state1": "3d934e04", "pcg2": "5d0430a1", "rand1": 0.692144036},
        {"in": [0, 0, 1], "hash": "e5f7e175", "pcg1": "08206188", "state1": "2f2015be", "pcg2": "ec23f4c4", "rand1": 0.0317440927},
        {"in": [511, 511, 0], "hash": "623b51fe", "pcg1": "6c1fea2f", "state1": "1e40559b", "pcg2": "6b432f1e", "rand1": 0.42236197},
        {"in": [1919, 1079, 7], "hash": "f96fd588", "pcg1": "b3b9f09b", "state1": "2c3a7c2d", "pcg2": "f6b14850", "rand1": 0.702055991},
        {"in": [3839, 2159, 15], "hash": "dae69a28", "pcg1": "2c48455c", "state1": "6dd9e14d", "pcg2": "3c666527", "rand1": 0.17297776},
    ],
    "pcg_raw": [
        {"state": "00000000", "out": "07bb2fe2", "next": "ac564b05"},
        {"state": "00000001", "out": "a8beea3c", "next": "d8e8c2ba"},
        {"state": "00000002", "out": "7a7ecc88", "next": "057b3a6f"},
        {"state": "deadbeef", "out": "67299972", "next": "d93d6300"},
    ],
}

FRAMES = {  # name -> (scene kind, w, h, depth, flags, spp)
    "cornell_direct": ("cornell", 64, 64, 1, 0, 1),
    "cornell_paths": ("cornell", 48, 48, 4, 1 | 2 | 4 | 8, 2),
    "rtapp_demo": ("rtapp", 64, 48, 2, 0, 1),
    "terrain_bounce": ("terrain", 64, 36, 3, 1 | 2, 1),
    "lattice_direct": ("lattice", 64, 36, 1, 0, 1),
}


def render_all(pkg, orc_mod):
    out = {}
    for name, (kind, w, h, depth, flags, spp) in FRAMES.items():
        scene = pkg.scenes.make_scene(kind, small=True)
        o = orc_mod.Oracle(pkg)
        scene.upload(o)
        u = scene.uniform(o, w, h, 0, depth)
        img = o.render_frame(u, o.opts(w, h, spp, flags))
        out[name] = (img, o.get_aov(pkg.AOV_PRIM_ID, w, h), o.get_aov(pkg.AOV_INST_ID, w, h))
    out.update(render_widened(pkg, orc_mod))
    return out


def render_widened(pkg, orc_mod):
    """The rows beyond the hot path: a light-BVH frame, a present-format frame (the uint8 image is stored
    widened to float32 so that all pins share one comparison) and the third frame of a denoised sequence."""
    out = {}
    w, h = 48, 48
    scene = pkg.scenes.make_scene("cornell", small=True)
    o = orc_mod.Oracle(pkg)
    scene.upload(o, build=False)
    for k in range(6):
        o.light_create((0.5 - 0.2 * k, -0.7, 0.1 * k - 0.3), (1.0, 0.8 - 0.1 * k, 0.4 + 0.1 * k), 0.1 + 0.05 * k)
    o.scene_build()
    u = scene.uniform(o, w, h, 2, 3)
    img = o.render_frame(u, o.opts(w, h, 2, pkg.LIGHT_BVH | pkg.BOUNCE_DIFFUSE | pkg.JITTER))
    out["cornell_light_bvh"] = (img, o.get_aov(pkg.AOV_PRIM_ID, w, h), o.get_aov(pkg.AOV_INST_ID, w, h))
    img8 = o.render_frame(u, o.opts(w, h, 1, pkg.render_format(pkg.FORMAT_BGRA8_SRGB)))
    out["cornell_bgra8_srgb"] = (img8.astype(np.float32), o.get_aov(pkg.AOV_PRIM_ID, w, h), o.get_aov(pkg.AOV_INST_ID, w, h))
    dop = o.denoise_opts(iterations=3, sigma_n_log2=5, sigma_z=0.05, sigma_l=4.0, clamp_gamma=1.5, max_history=8.0,
                         flags=pkg.DENOISE_BILATERAL | pkg.DENOISE_RESET)
    for k in range(3):
        uk = scene.uniform(o, w, h, 5 + k, 4)
        uk.viewInverse[3] += 0.03 * k
        o.render_frame(uk, o.opts(w, h, 1, pkg.BOUNCE_DIFFUSE | pkg.JITTER | pkg.GBUFFER))
        den = o.denoise(uk, dop, w, h)
        dop.flags = pkg.DENOISE_BILATERAL
    out["cornell_denoised_3"] = (den, o.get_aov(pkg.AOV_PRIM_ID, w, h), o.get_aov(pkg.AOV_INST_ID, w, h))
    return out


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("hardware-ray-tracer_b200")
    from oracle import binding as orc_mod
    json.dump(KAT, open(os.path.join(HERE, "rng_kat.json"), "w"), indent=1)
    arrays = {}
    for name, (img, prim, inst) in render_all(pkg, orc_mod).items():
        arrays[name + "_img"], arrays[name + "_prim"], arrays[name + "_inst"] = img, prim, inst
    np.savez_compressed(os.path.join(HERE, "oracle_frames.npz"), **arrays)
    print("wrote", sorted(arrays))
