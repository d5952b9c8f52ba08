"""Parity against the reference's SHIPPED SHADER BINARY: tests/golden/spv_kat.json holds frames computed by executing
`shaders/raytracing.slang.spv` (the module the application loads, RT/RTPipeline.cpp:168) instruction by instruction
(oracle/ref/spv_interp.py; generator tests/golden/make_spv_kat.py, run where /root/reference exists), with the Uniform block
and instance transforms produced by the reference's own Camera.cpp / MeshInstance.h / glm (oracle/_ref).

The same scenes go through the C ABI (`brt_*`, CUDA) and through the oracle; both must reproduce the binary's frames:
* primary (instance, primitive) ids identical on every non-fragile pixel (integers: exact),
* radiance within REL = 2e-5 of the pixel's magnitude + 1e-7 absolute + SLACK x the pixel's measured conditioning. SPIR-V leaves
  sqrt / normalize / length / pow / log2 / division and FMA contraction to a few ulps and TraceRay's barycentrics are
  implementation-defined (a binary32 ray/triangle test resolves them to ~4e-6 on these scenes); a clear-coat or GGX lobe of small
  roughness amplifies that — and the binary32 rounding noise of the evaluation itself — a thousandfold. The generator measured the
  amplification per pixel (`slack` = the largest radiance change over sixteen probes: barycentrics moved by up to 4e-6 along the axes and
  at random, the ray direction by 2e-7) and the bar widens by SLACK = 6 times it (observed need: 2.3; the worst-conditioned pixel of the
  seven scenes has slack = 0.18 % of its radiance, the median pixel 1e-6). A wrong constant, a swapped operand or a missing term is an error
  of percent, orders of magnitude above the bar (the median pixel agrees to 1e-7; a third of the pixels are bit-identical).
* `fragile` pixels (a ray within 1e-5 of a triangle edge or interval end: another valid intersector may decide differently) are
  skipped; they must stay below 2 % of any frame.
"""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REL, ABS, SLACK = 2e-5, 1e-7, 6.0


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(ROOT, "tests", "golden", "spv_kat.json")) as f:
        return json.load(f)


def _scene(pkg, case):
    S = pkg.scenes
    s = S.SceneDesc(case["name"])
    for m in case["meshes"]:
        s.meshes.append(("tri", np.array(m["vertices"], np.uint32).view(np.float32).reshape(-1, 8), np.array(m["indices"], np.uint32)))
    for m in case["materials"]:
        extra = {k: m[k] for k in ("subsurface", "specularTint", "anisotropic", "sheen", "sheenTint", "clearCoat", "clearCoatGloss")}
        s.materials.append(S.MaterialDesc(tuple(m["color"]), m["metallic"], m["roughness"], m["specular"], extra))
    s.lights = [(l["pos"], l["color"], l["intensity"], l["type"]) for l in case["lights"]]
    for inst, x in zip(case["instances"], case["xforms"]):
        s.instances.append((inst["mesh"], inst["material"], np.array(x, np.uint32).view(np.float32).reshape(3, 4)))
    s.cam_pos, s.cam_rot, s.fovy = tuple(case["cam_pos"]), tuple(case["cam_rot"]), case["fovy"]
    return s


def _check(pkg, api, case, own_uniform):
    w, h = case["width"], case["height"]
    scene = _scene(pkg, case)
    scene.upload(api)
    if own_uniform:   # the product's / oracle's own camera maths from the pose (whole chain: pose -> pixels)
        u = scene.uniform(api, w, h, frame=0, depth_max=case["depth_max"])
    else:             # the reference's bytes, as RTApp::run would hand them to writeToUniformBuffer
        u = pkg.binding.Uniform.from_buffer_copy(bytes.fromhex(case["uniform_hex"]))
    img = api.render_frame(u, api.opts(w, h))
    ref = np.array(case["rgba_bits"], np.uint32).view(np.float32).reshape(h, w, 4)
    ok = np.array(case["fragile"], np.uint8).reshape(h, w) == 0
    assert (~ok).mean() < 0.02
    prim = api.get_aov(pkg.AOV_PRIM_ID, w, h).astype(np.int64)
    inst = api.get_aov(pkg.AOV_INST_ID, w, h).astype(np.int64)
    rprim, rinst = np.array(case["prim"]).reshape(h, w), np.array(case["inst"]).reshape(h, w)
    hit = rinst >= 0
    # a miss is reported by the product as id 0xFFFFFFFF (or -1)
    miss_ok = ((inst == 0xFFFFFFFF) | (inst < 0))[ok & ~hit].all()
    assert miss_ok
    assert (inst[ok & hit] == rinst[ok & hit]).all() and (prim[ok & hit] == rprim[ok & hit]).all()
    t = api.get_aov(pkg.AOV_HIT_T, w, h)
    rt = np.array(case["hit_t"], np.float64).reshape(h, w)
    assert np.allclose(t[ok & hit], rt[ok & hit], rtol=2e-6)
    err = np.abs(img[..., :3].astype(np.float64) - ref[..., :3])
    slack = np.array(case["slack"], np.float64).reshape(h, w, 1)
    bar = REL * np.abs(ref[..., :3]).max(axis=-1, keepdims=True) + ABS + SLACK * slack
    bad = (err > bar).any(-1) & ok
    assert not bad.any(), (case["name"], int(bad.sum()), float((err / bar)[ok].max()), np.argwhere(bad)[:5].tolist())
    assert np.array_equal(img[..., 3], ref[..., 3])  # alpha = 1 (SH/raytracing.slang:132)
    lit = (ref[..., :3].sum(-1) > 0)[ok].mean()
    return float((err / np.maximum(np.abs(ref[..., :3]).max(axis=-1, keepdims=True), 1e-30))[ok & (ref[..., :3].sum(-1) > 1e-3)].max()), float(lit)


def test_golden_file_is_the_shipped_binary(kat):
    assert len(kat["cases"]) >= 7 and kat["generator"] == "tests/golden/make_spv_kat.py"
    assert len(kat["spv_sha256"]) == 64
    spv = "/root/reference/Hardware Ray Tracer/shaders/raytracing.slang.spv"
    if os.path.exists(spv):  # in the build container the vectors must belong to the binary that is there
        import hashlib
        assert hashlib.sha256(open(spv, "rb").read()).hexdigest() == kat["spv_sha256"]
    for c in kat["cases"]:
        n = c["width"] * c["height"]
        assert len(c["rgba_bits"]) == 4 * n and len(c["prim"]) == n and c["rays"] > n
        assert c["shadow_rays"] > 0


def test_golden_file_regenerates_from_the_binary(kat):
    """Where /root/reference exists (the build container): executing the binary again for a spot check of pixels of every case gives the
    committed bits — the JSON is what tests/golden/make_spv_kat.py writes today, not a hand-edited file."""
    spv = "/root/reference/Hardware Ray Tracer/shaders/raytracing.slang.spv"
    if not os.path.exists(spv):
        pytest.skip("/root/reference is not present (GPU box): the committed vectors are used as they are")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_spv_kat as G
    mod, host = G.Module(spv), G.ref_host()
    by_name = {c["name"]: c for c in kat["cases"]}
    rng = np.random.default_rng(3)
    for sc in G.scenes():
        case = by_name[sc["name"]]
        w, h = sc["width"], sc["height"]
        pix = {(int(rng.integers(0, w)), int(rng.integers(0, h))) for _ in range(6)} | {(w // 2, h // 2)}
        r = G.run_scene(mod, host, sc, pixels=pix)
        want = np.array(case["rgba_bits"], np.uint32).reshape(h, w, 4)
        for x, y in pix:
            assert np.array_equal(r["rgba"][y, x].view(np.uint32), want[y, x]), (sc["name"], x, y)
            assert r["prim"][y, x] == case["prim"][y * w + x] and r["inst"][y, x] == case["inst"][y * w + x]
        assert r["uniform"].hex() == case["uniform_hex"]


@pytest.mark.parametrize("own_uniform", [False, True])
def test_oracle_reproduces_the_shader_binary(pkg, orc_mod, kat, own_uniform):
    worst = 0.0
    for case in kat["cases"]:
        orc = orc_mod.Oracle(pkg)
        rel, lit = _check(pkg, orc, case, own_uniform)
        worst = max(worst, rel)
        assert lit > 0.3
    print(f"oracle vs raytracing.slang.spv: worst relative radiance difference {worst:.2e}")


def test_product_shading_code_reproduces_the_shader_binary_on_the_host_emulation(pkg, emu_lib, kat):
    """The PRODUCT's own kernel bodies (csrc/shading.cuh, render_kernels.cuh, traverse.cuh, the builder) compiled as host loops
    (tests/emu, test infrastructure) against the binary's frames: pins the product's restatement of the shaders in the CPU-only suite too
    (the GPU test below runs the same source as CUDA)."""
    worst = 0.0
    for case in kat["cases"]:
        emu = pkg.binding.SceneApi(emu_lib, "brt_", 0, 0, 1, 0)
        rel, lit = _check(pkg, emu, case, own_uniform=False)
        worst = max(worst, rel)
    print(f"product kernel bodies (host emulation) vs raytracing.slang.spv: worst relative radiance difference {worst:.2e}")


@pytest.mark.gpu
@pytest.mark.parametrize("own_uniform", [False, True])
def test_cuda_reproduces_the_shader_binary(pkg, kat, own_uniform):
    worst = 0.0
    for case in kat["cases"]:
        gpu = pkg.Context(device=0)
        rel, lit = _check(pkg, gpu, case, own_uniform)
        worst = max(worst, rel)
        gpu.close()
    print(f"CUDA vs raytracing.slang.spv: worst relative radiance difference {worst:.2e}")
