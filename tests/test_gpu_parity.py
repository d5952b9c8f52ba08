"""GPU parity tests proper: libbrt.so (hand-written sm_100a CUDA) through the C ABI against the CPU oracle
on the same seeded inputs. Run on the B200 box with `pytest -m gpu`."""
import numpy as np
import pytest

import parity_cases as pc
from util import compare_frames

pytestmark = pytest.mark.gpu


@pytest.fixture()
def make(pkg):
    return lambda flags=0: pkg.Context(device=0, flags=flags)


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


def test_grazing_and_axis_aligned_rays(pkg, orc_mod, make):
    pc.grazing_and_axis_aligned_rays(pkg, orc_mod, make)


def test_edge_cases(pkg, orc_mod, make):
    pc.edge_cases(pkg, orc_mod, make)


def test_dynamic_rebuild_and_instances(pkg, orc_mod, make):
    pc.dynamic_rebuild_and_instances(pkg, orc_mod, make)


def test_smart_culling(pkg, orc_mod, make):
    pc.smart_culling(pkg, orc_mod, make)


def test_tiles_and_crop(pkg, orc_mod):
    pc.tiles_and_crop(pkg, orc_mod, lambda rank, world: pkg.Context(device=0, tile_rank=rank, tile_world=world))


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


def test_counters_match_plain_run(pkg, make):
    """The instrumented kernels (BRT_CFG_COUNTERS) produce the same image as the plain ones."""
    scene = pkg.scenes.make_scene("terrain", small=True)
    a, b = make(pkg.CFG_COUNTERS), make()
    scene.upload(a)
    scene.upload(b)
    u = scene.uniform(a, 160, 90, 0, 3)
    ia = a.render_frame(u, a.opts(160, 90, 1, 3))
    ib = b.render_frame(u, b.opts(160, 90, 1, 3))
    assert np.array_equal(ia.view(np.uint32), ib.view(np.uint32))
    st = a.get_stats()
    assert st.nodes_visited_closest > st.rays_closest > 0 and st.prims_tested_closest > 0
    assert b.get_stats().nodes_visited_closest == 0


def test_c1_cornell_full_size(pkg, orc_mod, make):
    """BASELINE config 1 at its full size: 512x512, 1 spp, primary + shadow."""
    scene = pkg.scenes.cornell()
    a, b = make(), orc_mod.Oracle(pkg)
    scene.upload(a)
    scene.upload(b)
    u = scene.uniform(a, 512, 512, 0, 1)
    r = compare_frames(pkg, a, b, u, 512, 512)
    assert r["id_agreement"] >= 0.9999 and r["rmse"] <= 1e-3
    assert r["bit_exact"]
    assert r["stats"].rays_closest == 512 * 512


def test_repeatable(pkg, make):
    """Same inputs -> same bits, run to run (counter-based per-pixel RNG, ordered accumulation)."""
    scene = pkg.scenes.make_scene("terrain", small=True)
    a = make()
    scene.upload(a)
    u = scene.uniform(a, 192, 108, 5, 5)
    i1 = a.render_frame(u, a.opts(192, 108, 2, 15)).copy()
    i2 = a.render_frame(u, a.opts(192, 108, 2, 15))
    assert np.array_equal(i1.view(np.uint32), i2.view(np.uint32))


def test_fast_shading(pkg, orc_mod, make):
    """BRT_RENDER_FAST_SHADING (opt-in; shade kernels with FMA contraction and approximate division / rsqrt): primary ids and hit distances
    unchanged, ray counts unchanged up to a handful of flipped stochastic branches, radiance within the north star's 1e-3 relative RMSE of the ORACLE."""
    from util import rel_rmse
    scene = pkg.scenes.make_scene("terrain", small=True)
    a, b = make(), orc_mod.Oracle(pkg)
    scene.upload(a)
    scene.upload(b)
    w, h = 256, 144
    for depth, flags, spp in ((3, pkg.BOUNCE_REFLECT | pkg.BOUNCE_REFRACT, 1), (5, pkg.BOUNCE_REFLECT | pkg.BOUNCE_REFRACT | pkg.BOUNCE_DIFFUSE | pkg.JITTER, 4)):
        u = scene.uniform(a, w, h, 0, depth)
        fast = a.render_frame(u, a.opts(w, h, spp, flags | pkg.FAST_SHADING)).copy()
        ids = (a.get_aov(pkg.AOV_PRIM_ID, w, h).copy(), a.get_aov(pkg.AOV_INST_ID, w, h).copy(), a.get_aov(pkg.AOV_HIT_T, w, h).copy())
        rays = a.get_stats().rays_closest + a.get_stats().rays_occlusion
        ref = b.render_frame(u, b.opts(w, h, spp, flags))
        assert np.array_equal(ids[0], b.get_aov(pkg.AOV_PRIM_ID, w, h)) and np.array_equal(ids[1], b.get_aov(pkg.AOV_INST_ID, w, h))
        assert np.array_equal(ids[2].view(np.uint32), b.get_aov(pkg.AOV_HIT_T, w, h).view(np.uint32))
        rb = b.get_stats().rays_closest + b.get_stats().rays_occlusion
        assert abs(int(rays) - int(rb)) <= 1e-4 * rb
        e = rel_rmse(fast, ref)
        print(f"fast shading, depth {depth}: relative RMSE vs the oracle {e:.2e}, rays {rays} vs {rb}")
        assert e <= 1e-3
        assert not np.array_equal(fast.view(np.uint32), ref.view(np.uint32))  # (it really is another arithmetic)


def test_scene_info_buffer_and_tlas_handles(pkg, make):
    """Scene::getSceneInfoBuffer() / getTlas() (RT/Scene.h:150-151, tables of RT/Scene.cpp:313-403): the 80-byte SceneBufferInfo, its device
    copy and every table behind its addresses hold what was uploaded, in the reference's byte layouts (what SH/raytracing.slang:17-32,
    SH/objects.slang:15-59 and SH/light.slang:18-39 read through the buffer addresses)."""
    import ctypes as C
    import torch
    scene = pkg.scenes.make_scene("rtapp")
    a = make()
    scene.upload(a)
    info, d_copy = a.get_scene_info_buffer()
    rt = C.CDLL("libcudart.so")

    def peek(addr, nbytes):
        buf = (C.c_char * nbytes)()
        assert rt.cudaMemcpy(buf, C.c_void_p(addr), C.c_size_t(nbytes), 2) == 0
        return bytes(buf)

    assert (info.mStride, info.lStride, info.vStride, info.sStride, info.skyStride) == (52, 32, 32, 24, 88)
    assert info.lCount == len(scene.lights) == 3
    assert peek(d_copy, 80) == bytes(info)  # descriptor binding 3
    mats = np.frombuffer(peek(info.mBuf, 52 * len(scene.materials)), np.float32).reshape(-1, 13)
    assert np.allclose(mats[0, :3], 1.0) and mats[0, 4] == 1.0 and mats[0, 5] == 1.0 and mats[1, 5] == 0.0 and mats[0, 6] == 0.5  # RT/RTApp.cpp:6-7
    lights = np.frombuffer(peek(info.lBuf, 32 * 3), np.float32).reshape(3, 8)
    assert np.allclose(lights[0, :7], [1, 0, 0, 0, 0, 1, 2]) and np.allclose(lights[2, :7], [0, 0, -1, 1, 0, 0, 2])  # RT/RTApp.cpp:9-11
    inst = np.frombuffer(peek(info.sBuf, 24 * len(scene.instances)), np.uint64).reshape(-1, 3)
    assert [int(x) & 0xFFFFFFFF for x in inst[:, 2]] == [1, 0]  # material ids of the two instances (RT/RTApp.cpp:13-14)
    assert inst[0, 0] == inst[1, 0] and inst[0, 1] == inst[1, 1]  # both instances use mesh 0
    v, idx = scene.meshes[0][1], scene.meshes[0][2]
    assert peek(int(inst[0, 0]), v.nbytes) == v.tobytes() and peek(int(inst[0, 1]), idx.nbytes) == idx.tobytes()
    assert len(peek(info.skyBuf, 88)) == 88
    t = a.get_tlas()
    assert t.n_instances == 2 and t.n_nodes >= 1 and t.address == t.buffer == t.handle != 0 and t.memory != 0
    rec = np.frombuffer(peek(t.memory, 96 * 2), np.uint32).reshape(2, 24)
    assert sorted(int(x) for x in rec[:, 21]) == [0, 1]  # InstRec.inst_id of the two TLAS leaves
    torch.cuda.synchronize()


def test_hit_sort_changes_nothing_but_the_order_of_work(pkg, make):
    """BRT_CFG_HIT_SORT (bounce rounds shaded in the order of their hit positions): same bits, same ray counts, with and without a frame graph."""
    scene = pkg.scenes.make_scene("terrain", small=True)
    a, b, c = make(), make(pkg.CFG_HIT_SORT), make(pkg.CFG_HIT_SORT | pkg.CFG_NO_GRAPH)
    for x in (a, b, c):
        scene.upload(x)
    u = scene.uniform(a, 192, 108, 3, 6)
    flags = pkg.BOUNCE_REFLECT | pkg.BOUNCE_REFRACT | pkg.BOUNCE_DIFFUSE | pkg.JITTER
    ia = a.render_frame(u, a.opts(192, 108, 3, flags))
    for x in (b, c, b):
        ix = x.render_frame(u, x.opts(192, 108, 3, flags))
        assert np.array_equal(ia.view(np.uint32), ix.view(np.uint32))
        sa, sx = a.get_stats(), x.get_stats()
        assert (sa.rays_closest, sa.rays_occlusion) == (sx.rays_closest, sx.rays_occlusion)


@pytest.mark.parametrize("n,bits", [(1, 30), (2, 30), (31, 8), (4096, 30), (4097, 30), (100003, 30), (262144, 30), (262145, 30), (1 << 20, 32), (333333, 16)])
def test_radix_sort_pairs(pkg, make, n, bits):
    """The builder's own radix sort: sorted by the low `bits` bits, stable, a permutation of the input."""
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    if n > 1000:
        keys[: n // 3] = keys[0]  # long runs of equal keys exercise the stability
    vals = np.arange(n, dtype=np.uint32)
    k, v = make().debug_sort_pairs(keys, vals, bits)
    mask = np.uint32(0xFFFFFFFF if bits == 32 else (1 << bits) - 1)
    order = np.argsort(keys & mask, kind="stable")
    assert np.array_equal(v, order.astype(np.uint32))
    assert np.array_equal(k, keys[order])


@pytest.mark.parametrize("n_side", [2, 16, 17])
def test_tlas_sizes_vs_brute_force(pkg, orc_mod, make, n_side):
    """TLAS builds on both sides of the single-block small-build limit (4096 instances; 17^3 = 4913 takes the general path, 2^3 and
    16^3 = 4096 the one-launch path): closest and any-hit queries equal the oracle's brute force over all instances."""
    from util import random_rays
    scene = pkg.scenes.instanced_lattice(n_side, 0, animated=False)
    a, b = make(), orc_mod.Oracle(pkg, brute_force=True)
    scene.upload(a)
    scene.upload(b)
    assert a.get_stats().instances_visible == n_side ** 3
    half = n_side * 1.6 + 1.0
    rays = random_rays(6000, 21 + n_side, (-half, -half, 14.0 - half), (half, half, 14.0 + half))
    rays[::4, 7] = np.random.default_rng(9).random(len(rays[::4])).astype(np.float32) * 10.0
    ha, hb = a.trace_rays(rays, True), b.trace_rays(rays, True)
    assert np.array_equal(ha, hb)
    assert hb[:, 3].sum() > 300
    assert np.array_equal(a.trace_rays(rays, False)[:, 3], b.trace_rays(rays, False)[:, 3])


def test_frame_graph_replay(pkg, orc_mod, make):
    """The captured CUDA graph of a frame shape is replayed with new per-frame values (camera, frame counter, sky travel through
    FrameConsts): a sequence of frames with moving camera, a changing sky and changing frame shapes equals the same sequence
    rendered with individually launched kernels (BRT_CFG_NO_GRAPH) and the oracle's last frame."""
    scene = pkg.scenes.make_scene("terrain", small=True)
    a, b, o = make(), make(pkg.CFG_NO_GRAPH), orc_mod.Oracle(pkg)
    for api in (a, b, o):
        scene.upload(api)
    sky = pkg.Sky()
    sky.horizonColor[:] = (0.8, 0.8, 0.7)
    sky.groundColor[:] = (0.2, 0.15, 0.1)
    sky.upDirection[:] = (0.0, -1.0, 0.0)
    sky.brightness, sky.horizonSize = 1.0, 0.4
    shapes = [(160, 90, 3, 3), (160, 90, 3, 3), (160, 90, 3, 3), (128, 72, 2, 3 | 16), (160, 90, 3, 3), (160, 90, 3, 3 | 16)]
    for k, (w, h, depth, flags) in enumerate(shapes):
        u = scene.uniform(a, w, h, k, depth)
        u.viewInverse[3] += 0.2 * k
        sky.skyColor[:] = (0.1 * k, 0.4, 0.9 - 0.1 * k)
        for api in (a, b, o):
            api.sky_set(sky)
        ia = a.render_frame(u, a.opts(w, h, 2, flags))
        ib = b.render_frame(u, b.opts(w, h, 2, flags))
        assert np.array_equal(ia.view(np.uint32), ib.view(np.uint32)), k
        assert a.get_stats().rays_closest == b.get_stats().rays_closest and a.get_stats().ms_total > 0.0
    io = o.render_frame(u, o.opts(w, h, 2, flags))
    assert np.array_equal(ia.view(np.uint32), io.view(np.uint32))
    # per-class kernel times exist only for individually launched kernels
    assert b.get_stats().ms_trace_closest > 0.0 and a.get_stats().ms_trace_closest == 0.0


@pytest.mark.parametrize("stagger", ["0", "1", "2", "3", "4"])
def test_release_point_of_frames_in_flight(pkg, make, monkeypatch, stagger):
    """BRT_STAGGER / BRT_STAGGER_FIRST decide where in a frame the NEXT frame in flight may start (at once; behind round 0's closest-hit
    trace / shade / occlusion trace of the first sample batch; by depth). Scheduling only: six different multi-batch frames over three
    slots give the bits of the synchronous call under every setting."""
    scene = pkg.scenes.make_scene("terrain", small=True)
    ref_ctx = make()
    monkeypatch.setenv("BRT_STAGGER", stagger)
    monkeypatch.setenv("BRT_WAVEFRONT_PATHS", "20000")  # 160x90 -> one sample per wavefront: four batches per frame
    if stagger == "3":
        monkeypatch.setenv("BRT_STAGGER_FIRST", "0")  # the release point of round 2: the LAST batch
    a = make()
    scene.upload(a)
    scene.upload(ref_ctx)
    w, h, spp, flags = 160, 90, 4, 3
    uni = [scene.uniform(a, w, h, fr, 1 + fr % 4) for fr in range(6)]  # depthMax 1..4: both sides of the by-depth rule
    for k, u in enumerate(uni):
        u.viewInverse[3] += 0.25 * k
    ref = [ref_ctx.render_frame(u, ref_ctx.opts(w, h, spp, flags)).copy() for u in uni]
    out = [np.zeros((h, w, 4), np.float32) for _ in uni]
    for k, u in enumerate(uni):
        a.render_frame_async(u, a.opts(w, h, spp, flags), k % 3, out[k].ctypes.data)
    for s in range(3):
        a.frame_wait(s)
    for k in range(len(uni)):
        assert np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)), (stagger, k)
    assert len({o.tobytes() for o in out}) == len(out)  # the frames really differ
