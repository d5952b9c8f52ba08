"""BASELINE.json's full-size configurations on the GPU: oracle comparison on crops / bounded samples the CPU
finishes in seconds, plus size-independent properties on the whole frame."""
import numpy as np
import pytest

from util import compare_frames, random_rays

pytestmark = pytest.mark.gpu

R, T, D, J = 1, 2, 4, 8


@pytest.fixture(scope="module")
def terrain(pkg):
    return pkg.scenes.terrain_icospheres()  # 1,015,808 triangles


@pytest.fixture(scope="module")
def gpu_terrain(pkg, terrain):
    a = pkg.Context(device=0)
    terrain.upload(a)
    return a


@pytest.fixture(scope="module")
def orc_terrain(pkg, orc_mod, terrain):
    b = orc_mod.Oracle(pkg)
    terrain.upload(b)
    return b


def test_c2_scene_stats(pkg, terrain, gpu_terrain):
    st = gpu_terrain.get_stats()
    assert st.total_triangles == 1015808 == terrain.triangles()
    assert st.instances_total == 25 and st.instances_visible == 25
    assert 0 < st.bvh_nodes < st.total_triangles


def test_c2_1080p_crops_match_oracle(pkg, terrain, gpu_terrain, orc_terrain):
    """1920x1080, 1 spp, primary + 2 reflection/refraction bounces: three 256x128 windows against the oracle."""
    w, h = 1920, 1080
    u = terrain.uniform(gpu_terrain, w, h, 0, 3)
    for crop in [(832, 476, 256, 128), (0, 900, 256, 128), (1500, 300, 256, 128)]:
        r = compare_frames(pkg, gpu_terrain, orc_terrain, u, w, h, R | T, 1, crop)
        assert r["id_agreement"] >= 0.9999, crop
        assert r["rmse"] <= 1e-3, crop
        assert r["bit_exact"] and r["t_agreement"] == 1.0, crop  # stronger than the bar


def test_c2_full_frame_properties(pkg, terrain, gpu_terrain):
    """Whole 1080p frame: ray accounting, finite non-negative radiance, crop == same window of the full frame."""
    w, h = 1920, 1080
    u = terrain.uniform(gpu_terrain, w, h, 0, 3)
    full = gpu_terrain.render_frame(u, gpu_terrain.opts(w, h, 1, R | T)).copy()
    st = gpu_terrain.get_stats()
    assert w * h <= st.rays_closest <= 3 * w * h
    assert st.rays_occlusion <= 3 * st.rays_closest
    assert np.isfinite(full).all() and (full[..., :3] >= 0).all() and (full[..., 3] == 1).all()
    inst = gpu_terrain.get_aov(pkg.AOV_INST_ID, w, h)
    assert (inst != pkg.AOV_MISS).mean() > 0.3
    again = gpu_terrain.render_frame(u, gpu_terrain.opts(w, h, 1, R | T))
    assert np.array_equal(again.view(np.uint32), full.view(np.uint32))  # run-to-run reproducible at full size
    crop = (700, 400, 320, 200)
    part = gpu_terrain.render_frame(u, gpu_terrain.opts(w, h, 1, R | T, crop))
    assert np.array_equal(part[400:600, 700:1020].view(np.uint32), full[400:600, 700:1020].view(np.uint32))


def test_c2_random_rays_vs_oracle_bvh(pkg, gpu_terrain, orc_terrain):
    """1M-triangle scene: BVH8 (GPU) == independent binary SAH BVH (oracle) for incoherent rays."""
    rays = random_rays(200000, 23, (-8, -4, -8), (8, 1, 8))
    assert np.array_equal(gpu_terrain.trace_rays(rays, True), orc_terrain.trace_rays(rays, True))
    assert np.array_equal(gpu_terrain.trace_rays(rays, False)[:, 3], orc_terrain.trace_rays(rays, False)[:, 3])


def test_c3_4k_gi_crop_matches_oracle(pkg, terrain, gpu_terrain, orc_terrain):
    """3840x2160, 4-bounce GI with jitter; 2 of the 16 samples on a 192x96 window against the oracle."""
    w, h = 3840, 2160
    u = terrain.uniform(gpu_terrain, w, h, 0, 5)
    r = compare_frames(pkg, gpu_terrain, orc_terrain, u, w, h, R | T | D | J, 2, (1800, 1000, 192, 96))
    assert r["id_agreement"] >= 0.9999 and r["rmse"] <= 1e-3


def test_c5_diffuse_stress_crop(pkg, orc_mod):
    """8-bounce all-diffuse GI at 4K: window against the oracle."""
    scene = pkg.scenes.terrain_icospheres(all_diffuse=True)
    a, b = pkg.Context(device=0), orc_mod.Oracle(pkg)
    scene.upload(a)
    scene.upload(b)
    w, h = 3840, 2160
    u = scene.uniform(a, w, h, 0, 9)
    r = compare_frames(pkg, a, b, u, w, h, D, 1, (1700, 1100, 160, 80))
    assert r["id_agreement"] >= 0.9999 and r["rmse"] <= 1e-3


def test_c4_instanced_10m(pkg, orc_mod):
    """10.49 M instanced triangles: Smart Culling + TLAS rebuild + dynamic BLAS rebuild per frame, window vs oracle."""
    scene = pkg.scenes.instanced_lattice()
    a, b = pkg.Context(device=0), orc_mod.Oracle(pkg)
    scene.upload(a)
    scene.upload(b)
    assert a.get_stats().total_triangles == 513 * 20480
    base = scene.meshes[1][1]
    w, h = 1920, 1080
    for frame in range(2):
        v = pkg.scenes.animate_icosphere(base, frame)
        u = scene.uniform(a, w, h, frame, 1)
        for api in (a, b):
            api.mesh_update_vertices(1, v)
            api.scene_build()
        va, vb = a.smart_cull(u, w, h, 4.0, 0.25), b.smart_cull(u, w, h, 4.0, 0.25)
        assert va == vb and np.array_equal(a.get_visibility(513), b.get_visibility(513))
        r = compare_frames(pkg, a, b, u, w, h, 0, 1, (800, 400, 320, 200))
        assert r["id_agreement"] >= 0.9999 and r["rmse"] <= 1e-3


def test_large_scene_outside_l2_matches_oracle(pkg, orc_mod):
    """8.9 M unique triangles (terrain 2048^2 quads + the icospheres): node and triangle records (~0.5 GB) exceed the 126 MB L2.
    A 1080p window and incoherent rays against the oracle's independent binary SAH BVH."""
    scene = pkg.scenes.terrain_icospheres(n=2048)
    a, b = pkg.Context(device=0), orc_mod.Oracle(pkg)
    scene.upload(a)
    scene.upload(b)
    st = a.get_stats()
    assert st.total_triangles == scene.triangles() == 2 * 2048 * 2048 + 24 * 20480
    assert st.bvh_bytes > 3 * 126 * 2 ** 20
    w, h = 1920, 1080
    u = scene.uniform(a, w, h, 0, 3)
    r = compare_frames(pkg, a, b, u, w, h, R | T, 1, (832, 476, 192, 96))
    assert r["id_agreement"] >= 0.9999 and r["rmse"] <= 1e-3
    assert r["bit_exact"] and r["t_agreement"] == 1.0
    rays = random_rays(100000, 29, (-8, -4, -8), (8, 1, 8))
    assert np.array_equal(a.trace_rays(rays, True), b.trace_rays(rays, True))
    assert np.array_equal(a.trace_rays(rays, False)[:, 3], b.trace_rays(rays, False)[:, 3])
