"""Pins for the CPU oracle (no GPU): integer KATs of SH/random.slang (re-derived
here in pure Python), analytic ray/geometry cases, BVH == brute force, golden frame regression."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
M = 0xFFFFFFFF


def py_hash(x, y, z):  # SH/random.slang:2-12 in exact u32 arithmetic
    p0, p1, p2, p3 = 2246822519, 3266489917, 668265263, 374761393
    rot = lambda v: ((v << 17) | (v >> 15)) & M
    h = (z + p3 + x * p1) & M
    h = (p2 * rot(h)) & M
    h = (h + y * p1) & M
    h = (p2 * rot(h)) & M
    h = (p0 * (h ^ (h >> 15))) & M
    h = (p1 * (h ^ (h >> 13))) & M
    return h ^ (h >> 16)


def py_pcg(state):  # SH/random.slang:14-19
    prev = (state * 747796405 + 2891336453) & M
    word = (((prev >> ((prev >> 28) + 4)) ^ prev) * 277803737) & M
    return ((word >> 22) ^ word) & M, prev


def test_rng_kat_table(orc_mod):
    lib = orc_mod.load()
    kat = json.load(open(os.path.join(HERE, "golden", "rng_kat.json")))
    for row in kat["hash"]:
        x, y, z = row["in"]
        assert py_hash(x, y, z) == int(row["hash"], 16)
        assert lib.orc_kat_hash(x, y, z) == int(row["hash"], 16)
        s = C.c_uint32(int(row["hash"], 16))
        assert lib.orc_kat_pcg(C.byref(s)) == int(row["pcg1"], 16)
        assert s.value == int(row["state1"], 16)
        assert lib.orc_kat_pcg(C.byref(s)) == int(row["pcg2"], 16)
        s = C.c_uint32(int(row["hash"], 16))
        r1 = lib.orc_kat_rand(C.byref(s))
        assert np.float32(r1) == np.float32(int(row["pcg1"], 16)) * np.float32(2.0 ** -32)
        assert abs(r1 - row["rand1"]) < 1e-7
    for row in kat["pcg_raw"]:
        out, st = py_pcg(int(row["state"], 16))
        assert out == int(row["out"], 16) and st == int(row["next"], 16)
        s = C.c_uint32(int(row["state"], 16))
        assert lib.orc_kat_pcg(C.byref(s)) == out and s.value == st


def test_rng_random_agreement(orc_mod):
    lib = orc_mod.load()
    rng = np.random.default_rng(1)
    for x, y, z in rng.integers(0, 2 ** 32, size=(200, 3), dtype=np.uint64):
        assert lib.orc_kat_hash(int(x), int(y), int(z)) == py_hash(int(x), int(y), int(z))


def test_rand_can_return_one(orc_mod):
    """float(0xffffffff) rounds to 2^32: the largest output maps to exactly 1.0."""
    assert np.float32(0xFFFFFFFF) * np.float32(2.0 ** -32) == np.float32(1.0)


def test_scene_hash_matches_oracle(pkg, orc_mod):
    lib = orc_mod.load()
    xs = np.arange(50, dtype=np.uint32)
    h = pkg.scenes.hash3(xs, xs * 7 + 1, np.uint32(0xB100B200))
    for i, x in enumerate(xs):
        assert int(h[i]) == lib.orc_kat_hash(int(x), int(x * 7 + 1), 0xB100B200)


def test_det_sincos_log2_accuracy(orc_mod):
    lib = orc_mod.load()
    s, c = C.c_float(), C.c_float()
    for x in np.linspace(0.0, 6.2831855, 400, dtype=np.float32):
        lib.orc_kat_sincos(float(x), C.byref(s), C.byref(c))
        assert abs(s.value - np.sin(np.float64(x))) < 2e-6 and abs(c.value - np.cos(np.float64(x))) < 2e-6
    for x in np.geomspace(1e-6, 1e6, 300, dtype=np.float32):
        assert abs(lib.orc_kat_log2(float(x)) - np.log2(np.float64(x))) < 4e-6 * max(1.0, abs(np.log2(np.float64(x))))


def _f3(*v):
    return (C.c_float * 3)(*v)


def test_triangle_analytic(orc_mod):
    lib = orc_mod.load()
    tuv = _f3(0, 0, 0)
    v0, v1, v2 = _f3(0, 0, 1), _f3(1, 0, 1), _f3(0, 1, 1)
    assert lib.orc_kat_intersect_tri(_f3(0.25, 0.25, 0), _f3(0, 0, 1), 0.001, 1e32, v0, v1, v2, tuv) == 1
    assert tuple(tuv) == (1.0, 0.25, 0.25)  # t, u (weight of v1), v (weight of v2)
    # both windings are hit (no face culling, RT/Scene.cpp:188)
    assert lib.orc_kat_intersect_tri(_f3(0.25, 0.25, 0), _f3(0, 0, 1), 0.001, 1e32, v0, v2, v1, tuv) == 1
    # behind the origin / beyond tmax / outside
    assert lib.orc_kat_intersect_tri(_f3(0.25, 0.25, 2), _f3(0, 0, 1), 0.001, 1e32, v0, v1, v2, tuv) == 0
    assert lib.orc_kat_intersect_tri(_f3(0.25, 0.25, 0), _f3(0, 0, 1), 0.001, 1.0, v0, v1, v2, tuv) == 0
    assert lib.orc_kat_intersect_tri(_f3(0.75, 0.75, 0), _f3(0, 0, 1), 0.001, 1e32, v0, v1, v2, tuv) == 0


def test_triangle_watertight_shared_edge(orc_mod):
    """Rays through the shared edge / vertex of two triangles always hit at least one of them."""
    lib = orc_mod.load()
    tuv = _f3(0, 0, 0)
    a, b, c, d = _f3(0, 0, 5), _f3(1, 0, 5.3), _f3(1, 1, 4.9), _f3(0, 1, 5.1)
    rng = np.random.default_rng(3)
    for _ in range(2000):
        s = float(rng.random())
        p = np.array([s, s, 0.0]) * 1.0  # a point on the diagonal a-c projected
        tgt = np.array([a[0] + s * (c[0] - a[0]), a[1] + s * (c[1] - a[1]), a[2] + s * (c[2] - a[2])], np.float32)
        o = np.array([rng.normal(), rng.normal(), -3.0], np.float32)
        dd = (tgt - o).astype(np.float32)
        h1 = lib.orc_kat_intersect_tri(_f3(*o), _f3(*dd), 0.0, 1e32, a, b, c, tuv)
        h2 = lib.orc_kat_intersect_tri(_f3(*o), _f3(*dd), 0.0, 1e32, a, c, d, tuv)
        assert h1 or h2


def test_brdf_early_outs_and_lambert(orc_mod, pkg):
    lib = orc_mod.load()
    m = pkg.Material()
    m.color[:] = (0.5, 0.25, 1.0)
    m.roughness, m.specular = 1.0, 0.0
    out = _f3(0, 0, 0)
    # N.L <= 0 and N.V <= 0 -> 0 (SH/disney.slang:99-100)
    lib.orc_kat_brdf(C.byref(m), _f3(0, 0, 1), _f3(0, 0, 1), _f3(0, 0, -1), out)
    assert tuple(out) == (0.0, 0.0, 0.0)
    lib.orc_kat_brdf(C.byref(m), _f3(0, 0, 1), _f3(0, 0, -1), _f3(0, 0, 1), out)
    assert tuple(out) == (0.0, 0.0, 0.0)
    # normal incidence, rough dielectric with specular 0: diffuse term dominates and is ~ color / pi * FD
    lib.orc_kat_brdf(C.byref(m), _f3(0, 0, 1), _f3(0, 0, 1), _f3(0, 0, 1), out)
    assert out[0] > 0 and out[1] > 0 and out[2] > 0
    assert abs(out[0] / out[2] - 0.5) < 0.2


def test_point_light_inverse_square(pkg, orc_mod):
    """A single quad lit from distance d: radiance scales with 1/d^2 (SH/light.slang:23-39), no cosine term."""
    S = pkg.scenes
    vals = []
    for d in (1.0, 2.0):
        s = S.SceneDesc("quad")
        s.meshes = [("tri",) + S._quad((-50, 0, -50), (50, 0, -50), (50, 0, 50), (-50, 0, 50))]
        s.materials = [S.MaterialDesc((1, 1, 1))]
        s.lights = [((0.0, -d, 0.0), (1.0, 1.0, 1.0), 1.0)]
        s.instances = [(0, 0, S.xform())]
        s.cam_pos, s.cam_rot = (0.0, -1.0, 0.0), (-np.pi / 2, 0.0, 0.0)  # looking straight down (+Y is down)
        o = orc_mod.Oracle(pkg)
        s.upload(o)
        u = s.uniform(o, 33, 33, 0, 1)
        img = o.render_frame(u, o.opts(33, 33))
        vals.append(img[16, 16, 0])
        assert o.get_aov(pkg.AOV_INST_ID, 33, 33)[16, 16] == 0
    assert vals[0] > 0 and abs(vals[0] / vals[1] - 4.0) < 0.05


def test_oracle_bvh_equals_brute_force(pkg, orc_mod):
    from util import random_rays
    for kind in ("cornell", "terrain", "lattice"):
        scene = pkg.scenes.make_scene(kind, small=True)
        a, b = orc_mod.Oracle(pkg), orc_mod.Oracle(pkg, brute_force=True)
        scene.upload(a)
        scene.upload(b)
        rays = random_rays(3000, 7, (-8, -8, -8), (8, 8, 20))
        assert np.array_equal(a.trace_rays(rays, True), b.trace_rays(rays, True))
        assert np.array_equal(a.trace_rays(rays, False)[:, 3], b.trace_rays(rays, False)[:, 3])


def test_oracle_golden_frames(pkg, orc_mod):
    """Regression pin: the committed oracle outputs (tests/golden/make_golden.py) are reproduced bit for bit."""
    g = np.load(os.path.join(HERE, "golden", "oracle_frames.npz"))
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    for name, (img, prim, inst) in make_golden.render_all(pkg, orc_mod).items():
        assert np.array_equal(g[name + "_prim"], prim), name
        assert np.array_equal(g[name + "_inst"], inst), name
        assert np.array_equal(g[name + "_img"].view(np.uint32), img.view(np.uint32)), name


def test_oracle_error_behaviour(pkg, orc_mod):
    o = orc_mod.Oracle(pkg)
    with pytest.raises(pkg.BrtError):
        o.mesh_create(np.zeros((3, 8), np.float32), [0, 1, 5])  # index out of range
    with pytest.raises(pkg.BrtError):
        o.instance_create(0, 0, np.eye(3, 4))  # no such mesh / material
    with pytest.raises(pkg.BrtError):
        u = o.camera_uniform((0, 0, 0), (0, 0, 0), 1.0, 1.0)
        o.render_frame(u, o.opts(4, 4))  # scene not built


def test_det_exp2_srgb_unorm8(orc_mod):
    """The explicit exp2 polynomial (sRGB curve, denoiser weights), the sRGB transfer function and Vulkan's float -> UNORM8 rule."""
    import ctypes as C
    lib = orc_mod.load()
    lib.orc_kat_exp2.restype, lib.orc_kat_exp2.argtypes = C.c_float, [C.c_float]
    lib.orc_kat_srgb.restype, lib.orc_kat_srgb.argtypes = C.c_float, [C.c_float]
    lib.orc_kat_unorm8.restype, lib.orc_kat_unorm8.argtypes = C.c_uint32, [C.c_float]
    xs = np.concatenate([np.linspace(-100, 100, 4001), np.arange(-20, 21), np.array([0.5, -0.5, 1e-3, -1e-3])]).astype(np.float32)
    got = np.array([lib.orc_kat_exp2(float(x)) for x in xs], dtype=np.float64)
    want = np.exp2(xs.astype(np.float64))
    assert np.max(np.abs(got / want - 1.0)) < 4e-7          # a few ulp over the whole range the denoiser / sRGB curve use
    assert all(lib.orc_kat_exp2(float(k)) == 2.0 ** k for k in range(-20, 21))  # exact at integers
    cs = np.linspace(0.0, 1.0, 2001).astype(np.float32)
    s = np.array([lib.orc_kat_srgb(float(c)) for c in cs], dtype=np.float64)
    ref = np.where(cs <= 0.0031308, 12.92 * cs, 1.055 * np.power(cs.astype(np.float64), 1 / 2.4) - 0.055)
    assert np.max(np.abs(s - ref)) < 2e-6 and np.all(np.diff(s) >= 0.0) and s[0] == 0.0 and s[-1] == 1.0
    assert lib.orc_kat_srgb(-1.0) == 0.0 and lib.orc_kat_srgb(7.0) == 1.0 and lib.orc_kat_srgb(float("nan")) == 0.0
    # UNORM8: clamp, NaN -> 0, round to nearest even
    for f, q in ((-0.1, 0), (0.0, 0), (1.0, 255), (3.0, 255), (float("nan"), 0), (0.5 / 255, 0), (1.5 / 255, 2), (2.5 / 255, 2), (0.5, 128), (254.5 / 255, 254)):
        assert lib.orc_kat_unorm8(f) == q, (f, q)


def test_light_bvh_sampler_pdf(pkg, orc_mod):
    """Light BVH descent: for any shading point the leaf probabilities sum to one, every light with flux can be
    reached, the sampler's 1 / pdf is the reciprocal of that probability, and drawing with a stratified random number reproduces it."""
    import ctypes as C
    lib = orc_mod.load()
    f3 = C.POINTER(C.c_float)
    lib.orc_kat_light_pdfs.restype, lib.orc_kat_light_pdfs.argtypes = C.c_int, [C.c_void_p, f3, f3, C.c_uint32]
    lib.orc_kat_light_sample.restype, lib.orc_kat_light_sample.argtypes = C.c_uint32, [C.c_void_p, f3, C.c_float, f3]
    orc = orc_mod.Oracle(pkg)
    pkg.scenes.cornell().upload(orc, build=False)
    rng = np.random.default_rng(3)
    n = 1 + 37
    for _ in range(37):
        orc.light_create((rng.random(3) * 4 - 2).tolist(), (rng.random(3) + 0.05).tolist(), float(0.1 + rng.random() * 3))
    orc.scene_build()
    for P in ([0.0, 0.0, 0.0], [1.9, -1.9, 0.3], [40.0, 5.0, -30.0]):
        p = (C.c_float * 3)(*P)
        pdf = (C.c_float * n)()
        assert lib.orc_kat_light_pdfs(orc.ctx, p, pdf, n) == 0
        pdf = np.array(pdf, dtype=np.float64)
        assert abs(pdf.sum() - 1.0) < 1e-5 and (pdf > 0).all()
        m = 20000
        counts = np.zeros(n)
        inv = C.c_float()
        for k in range(m):
            li = lib.orc_kat_light_sample(orc.ctx, p, (k + 0.5) / m, C.byref(inv))
            counts[li] += 1
            if k % 997 == 0:
                assert abs(inv.value * pdf[li] - 1.0) < 1e-4
        assert np.abs(counts / m - pdf).max() < 2.0 / m * n  # stratified draws: each leaf interval is hit to within its two ends


def test_math_golden_bits(orc_mod):
    """tests/golden/math_kat.json (generator: tests/golden/make_math_kat.py): the deterministic elementary functions keep their bits."""
    import ctypes as C
    import json
    import os
    import struct
    lib = orc_mod.load()
    for name in ("orc_kat_exp2", "orc_kat_srgb", "orc_kat_log2"):
        getattr(lib, name).restype, getattr(lib, name).argtypes = C.c_float, [C.c_float]
    lib.orc_kat_unorm8.restype, lib.orc_kat_unorm8.argtypes = C.c_uint32, [C.c_float]
    kat = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "math_kat.json")))
    f = lambda h: struct.unpack("<f", struct.pack("<I", int(h, 16)))[0]
    b = lambda v: "%08x" % struct.unpack("<I", struct.pack("<f", v))[0]
    for e in kat["exp2"]:
        assert b(lib.orc_kat_exp2(f(e["x"]))) == e["y"]
    for e in kat["log2"]:
        assert b(lib.orc_kat_log2(f(e["x"]))) == e["y"]
    for e in kat["srgb"]:
        assert b(lib.orc_kat_srgb(f(e["x"]))) == e["y"]
    for e in kat["sincos"]:
        s, c = C.c_float(), C.c_float()
        lib.orc_kat_sincos(f(e["x"]), C.byref(s), C.byref(c))
        assert (b(s.value), b(c.value)) == (e["s"], e["c"])
    for e in kat["unorm8"]:
        assert lib.orc_kat_unorm8(f(e["x"])) == e["q"]
