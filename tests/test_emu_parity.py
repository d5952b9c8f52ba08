"""Kernel LOGIC of the CUDA path checked on the CPU: the product's kernel bodies (csrc/*.cuh) compiled by
g++ as sequential loops (tests/emu, test infrastructure only) against the oracle. The GPU tests
(test_gpu_parity.py, -m gpu) run the same cases through libbrt.so."""
import pytest

import parity_cases as pc


@pytest.fixture()
def make(pkg, emu_lib):
    return lambda flags=0: pkg.binding.SceneApi(emu_lib, "brt_", 0, 0, 1, flags)


@pytest.mark.parametrize("case", pc.FRAME_CASES, ids=[c[0] for c in pc.FRAME_CASES])
def test_frames(pkg, orc_mod, make, case):
    pc.frame_case(pkg, orc_mod, make, case)


def test_sky_and_nonpoint_light(pkg, orc_mod, make):
    pc.sky_and_nonpoint_light(pkg, orc_mod, make)


@pytest.mark.parametrize("kind", ["terrain", "lattice", "cornell"])
def test_random_rays_vs_brute_force(pkg, orc_mod, make, kind):
    pc.random_rays_vs_brute_force(pkg, orc_mod, make, kind)


def test_collapse_rules(pkg, make):
    pc.collapse_rules(pkg, make)


def test_blas_structure(pkg, make):
    pc.blas_structure(pkg, make)


def test_grazing_and_axis_aligned_rays(pkg, orc_mod, make):
    pc.grazing_and_axis_aligned_rays(pkg, orc_mod, make)


def test_degenerate_extents(pkg, orc_mod, make):
    pc.degenerate_extents(pkg, orc_mod, make)


def test_nonfinite_vertices(pkg, make):
    pc.nonfinite_vertices(pkg, make)


def test_bvh_quality_guard(pkg, make):
    pc.bvh_quality_guard(pkg, make)


def test_edge_cases(pkg, orc_mod, make):
    pc.edge_cases(pkg, orc_mod, make)


def test_dynamic_rebuild_and_instances(pkg, orc_mod, make):
    pc.dynamic_rebuild_and_instances(pkg, orc_mod, make)


def test_smart_culling(pkg, orc_mod, make):
    pc.smart_culling(pkg, orc_mod, make)


def test_tiles_and_crop(pkg, orc_mod, emu_lib):
    pc.tiles_and_crop(pkg, orc_mod, lambda rank, world: pkg.binding.SceneApi(emu_lib, "brt_", 0, rank, world, 0))


def test_error_behaviour(pkg, make):
    pc.error_behaviour(pkg, make)


def test_general_transforms_and_many_lights(pkg, orc_mod, make):
    pc.general_transforms_and_many_lights(pkg, orc_mod, make)


def test_frames_in_flight(pkg, orc_mod, make):
    pc.frames_in_flight(pkg, orc_mod, make)


def test_present_formats(pkg, orc_mod, make):
    pc.present_formats(pkg, orc_mod, make)


def test_denoiser(pkg, orc_mod, make):
    pc.denoiser(pkg, orc_mod, make)


def test_light_bvh(pkg, orc_mod, make):
    pc.light_bvh(pkg, orc_mod, make)


def test_denoise_in_flight(pkg, orc_mod, make):
    pc.denoise_in_flight(pkg, orc_mod, make)


def test_golden_frames(pkg, make):
    pc.golden_frames(pkg, make)


def test_counters_flag(pkg, orc_mod, make):
    """BRT_CFG_COUNTERS fills the node / primitive visit counters that feed the roofline's algorithmic bytes."""
    scene = pkg.scenes.make_scene("terrain", small=True)
    a = make(pkg.CFG_COUNTERS)
    scene.upload(a)
    u = scene.uniform(a, 64, 36, 0, 2)
    a.render_frame(u, a.opts(64, 36, 1, 1))
    st = a.get_stats()
    assert st.nodes_visited_closest > st.rays_closest > 0 and st.prims_tested_closest > 0
    assert st.nodes_visited_occlusion > 0 and st.total_triangles == scene.triangles()
    assert st.bvh_nodes > 0 and st.bvh_bytes > st.total_triangles


def test_warp_kernels_equal_scalar(pkg, emu_lib, monkeypatch):
    """The device's warp-cooperative collapse (collapse_warp, build_kernels.cuh) on 32 lock-stepped host threads (tests/emu/warp_emu.h)
    against the scalar collapse_body, item by item while real BVHs are built: same node bytes, same queue entries, same primitive
    records, same allocation counters — for both collapse rules, triangle BLASes and an instance TLAS."""
    import ctypes as C
    import numpy as np

    def stats():
        a, b = C.c_ulonglong(), C.c_ulonglong()
        emu_lib.brt_emu_warp_collapse_stats(C.byref(a), C.byref(b))
        return a.value, b.value

    monkeypatch.setenv("BRT_EMU_WARP_CHECK", "1")
    items0, bad0 = stats()
    S = pkg.scenes
    hv, hi = S.heightfield(6)
    sv, si = S.icosphere(1)
    for flags in (0, pkg.CFG_GREEDY_COLLAPSE):
        a = pkg.binding.SceneApi(emu_lib, "brt_", 0, 0, 1, flags)
        mat = a.material_create((0.6, 0.6, 0.6))
        terrain, ball = a.mesh_create(hv, hi), a.mesh_create(sv, si)
        a.instance_create(terrain, mat, S.xform())
        rng = np.random.default_rng(3)
        for k in range(19):  # TLAS of 20 instances + a sphere
            px, py, pz = (rng.random(3) * 2 - 1) * (6.0, 1.0, 6.0)
            a.instance_create(ball, mat, S.xform((0.3, 0.3, 0.3), (px, py - 1.5, pz)))
        a.instance_create(a.sphere_create((0.0, -2.0, 0.0), 0.5), mat, S.xform())
        a.scene_build()
        a.close()
    items, bad = stats()
    assert items - items0 > 60
    assert bad == bad0 == 0
    # ... and the device's warp-cooperative treelet kernel (k_treelet_warp) on a copy of every binary tree that gets treelet passes:
    # a valid tree (parents, boxes, counts, costs consistent, every leaf once) with about the SAH cost of the scalar passes
    checks, invalid, dev = C.c_ulonglong(), C.c_ulonglong(), C.c_float()
    emu_lib.brt_emu_warp_treelet_stats(C.byref(checks), C.byref(invalid), C.byref(dev))
    assert checks.value >= 4 and invalid.value == 0
    assert dev.value < 0.05  # equal on the terrain; the icosphere's many ties let the two versions settle on slightly different trees


def test_shade_hoist_variants_give_the_same_bits(pkg, emu_lib):
    """BRT_SHADE_HOIST_HEAD / BRT_SHADE_HOIST_HIT (opt-in builds: the shade kernels request a path's loads ahead of the branches they sit
    behind, profiles/r2_ncu_summary.md §5) only move loads: same frames, bit for bit, as the default build."""
    import ctypes
    import os
    import subprocess
    import numpy as np
    emu_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
    subprocess.check_call(["make", "-s", "-C", emu_dir, "libbrt_emu_hoist.so"])
    hoist = ctypes.CDLL(os.path.join(emu_dir, "libbrt_emu_hoist.so"))
    scene = pkg.scenes.make_scene("terrain", small=True)
    a = pkg.binding.SceneApi(emu_lib, "brt_", 0, 0, 1, 0)
    b = pkg.binding.SceneApi(hoist, "brt_", 0, 0, 1, 0)
    scene.upload(a)
    scene.upload(b)
    for depth, flags in ((1, 0), (3, 3), (4, 7)):
        u = scene.uniform(a, 96, 54, depth, depth)
        ia = a.render_frame(u, a.opts(96, 54, 2, flags))
        ib = b.render_frame(u, b.opts(96, 54, 2, flags))
        assert np.array_equal(ia.view(np.uint32), ib.view(np.uint32)), (depth, flags)
        assert np.array_equal(a.get_aov(pkg.AOV_PRIM_ID, 96, 54), b.get_aov(pkg.AOV_PRIM_ID, 96, 54))
