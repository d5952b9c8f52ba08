"""Parity cases shared by the CPU emulation tests (kernel logic, -m "not gpu") and the GPU tests proper
(libbrt.so through the C ABI, -m gpu). `make` builds a context of the implementation under test,
`orc_mod.Oracle(pkg)` the checker. Bars (BASELINE.json): primary-hit ids identical on >= 99.99 % of the
pixels, linear radiance within 1e-3 relative RMSE; the arithmetic contract of DESIGN.md §3 in fact makes
every case below bit-exact, which is asserted where noted."""
import numpy as np
import pytest

from util import compare_frames, random_rays, render_pair

R, T, D, J, SKY = 1, 2, 4, 8, 16  # BRT_RENDER_* flags

FRAME_CASES = [  # name, scene, w, h, depth, flags, spp
    ("c1_cornell_direct", "cornell", 128, 128, 1, 0, 1),
    ("cornell_reflect_refract", "cornell", 96, 96, 3, R | T, 1),
    ("cornell_gi_jitter_spp", "cornell", 64, 64, 5, R | T | D | J, 3),
    ("cornell_diffuse_only", "cornell", 64, 64, 4, D, 2),
    ("rtapp_demo_depth2", "rtapp", 160, 120, 2, 0, 1),
    ("c2_terrain_bounces", "terrain", 160, 90, 3, R | T, 1),
    ("c3_terrain_gi", "terrain", 96, 54, 5, R | T | D | J, 2),
    ("c4_lattice_direct", "lattice", 160, 90, 1, 0, 1),
    ("c5_terrain_diffuse8", "terrain_diffuse", 96, 54, 9, D, 1),
]


def frame_case(pkg, orc_mod, make, case):
    name, kind, w, h, depth, flags, spp = case
    scene = pkg.scenes.make_scene(kind, small=True)
    r = render_pair(pkg, scene, make(), orc_mod.Oracle(pkg), w, h, depth, flags, spp)
    assert r["id_agreement"] >= 0.9999, name
    assert r["rmse"] <= 1e-3, name
    assert r["stats"].rays_closest == r["ref_stats"].rays_closest, name
    assert r["stats"].rays_occlusion == r["ref_stats"].rays_occlusion, name
    assert r["t_agreement"] == 1.0 and r["bit_exact"], name  # stronger than the bar: identical bits
    return r


def sky_and_nonpoint_light(pkg, orc_mod, make):
    """Sky miss colour (extension) and the constant-direction branch of processLight (SH/light.slang:33-36)."""
    scene = pkg.scenes.make_scene("cornell", small=True)
    a, b = make(), orc_mod.Oracle(pkg)
    sky = pkg.Sky()
    sky.skyColor[:] = (0.2, 0.4, 0.9)
    sky.horizonColor[:] = (0.8, 0.8, 0.7)
    sky.groundColor[:] = (0.2, 0.15, 0.1)
    sky.upDirection[:] = (0.0, -1.0, 0.0)
    sky.brightness, sky.horizonSize = 1.5, 0.4
    for api in (a, b):
        scene.upload(api, build=False)
        api.sky_set(sky)
        api.light_create((0, 0, 0), (1.0, 0.9, 0.8), 0.7, type=2)  # DIRECTIONAL
        api.scene_build()
    u = scene.uniform(a, 96, 96, 0, 3)
    u.viewInverse[11] = -6.0  # step back so that some rays miss the box
    r = compare_frames(pkg, a, b, u, 96, 96, R | T | SKY, 1)
    assert r["id_agreement"] == 1.0 and r["bit_exact"]
    assert (r["img"][..., :3].reshape(-1, 3)[a.get_aov(pkg.AOV_INST_ID, 96, 96).reshape(-1) == pkg.AOV_MISS] > 0).all()


def random_rays_vs_brute_force(pkg, orc_mod, make, kind="terrain"):
    """BVH8 traversal == brute force over all primitives (closest: same t bits, prim, instance; any-hit: same bool)."""
    scene = pkg.scenes.make_scene(kind, small=True)
    a, b = make(), orc_mod.Oracle(pkg, brute_force=True)
    scene.upload(a)
    scene.upload(b)
    box = {"lattice": ((-5, -5, 9), (5, 5, 19)), "cornell": ((-1, -1, -1), (1, 1, 1))}.get(kind, ((-8, -4, -8), (8, 2, 8)))
    rays = random_rays(20000, 11, *box)
    # bounded segments too (shadow-ray style)
    rays[::3, 7] = np.random.default_rng(5).random(len(rays[::3])).astype(np.float32) * 6.0
    ha, hb = a.trace_rays(rays, True), b.trace_rays(rays, True)
    assert np.array_equal(ha, hb)
    assert hb[:, 3].sum() > 1000
    oa, ob_ = a.trace_rays(rays, False), b.trace_rays(rays, False)
    assert np.array_equal(oa[:, 3], ob_[:, 3])


def collapse_rules(pkg, make):
    """The area-optimal choice of 8-wide nodes (wide_cost_body) against the greedy rule (BRT_CFG_GREEDY_COLLAPSE): identical hits and
    frames (results never depend on the BVH's shape), fewer nodes, fewer node visits."""
    scene = pkg.scenes.make_scene("terrain", small=True)
    a, b = make(pkg.CFG_COUNTERS), make(pkg.CFG_COUNTERS | pkg.CFG_GREEDY_COLLAPSE)
    scene.upload(a)
    scene.upload(b)
    rays = random_rays(20000, 23, (-8, -4, -8), (8, 2, 8))
    assert np.array_equal(a.trace_rays(rays, True), b.trace_rays(rays, True))
    assert np.array_equal(a.trace_rays(rays, False)[:, 3], b.trace_rays(rays, False)[:, 3])
    u = scene.uniform(a, 96, 64, 0, 3)
    fl = pkg.BOUNCE_REFLECT | pkg.BOUNCE_REFRACT
    ia, ib = a.render_frame(u, a.opts(96, 64, 1, fl)), b.render_frame(u, b.opts(96, 64, 1, fl))
    assert np.array_equal(ia.view(np.uint32), ib.view(np.uint32))
    sa, sb = a.get_stats(), b.get_stats()
    assert sa.rays_closest == sb.rays_closest and sa.rays_occlusion == sb.rays_occlusion
    assert sa.bvh_nodes < 0.75 * sb.bvh_nodes
    assert sa.nodes_visited_closest <= sb.nodes_visited_closest and sa.nodes_visited_occlusion <= sb.nodes_visited_occlusion


def check_blas_structure(nodes, tris, n_tris_expected):
    """Structural invariants of a compressed 8-wide BLAS: every node reachable exactly once, every primitive in exactly one
    leaf slot, every slot's dequantised box contains everything below it, empty slots can never be hit, permuted masks consistent.
    Returns (mean used slots per node, depth)."""
    n_nodes = len(nodes)
    assert len(tris) == n_tris_expected
    prim_ids = tris[:, 3].copy().view(np.uint32)
    assert np.array_equal(np.sort(prim_ids), np.arange(n_tris_expected, dtype=np.uint32))  # every primitive exactly once
    verts = tris.reshape(-1, 3, 4)[:, :, :3].astype(np.float64)
    tlo, thi = verts.min(axis=1), verts.max(axis=1)
    p = nodes[:, 0:3].copy().view(np.float32).astype(np.float64)
    ew = nodes[:, 3]
    step = np.stack([((ew >> (8 * k)) & 0xFF).astype(np.uint32) << 23 for k in range(3)], axis=1).view(np.float32).astype(np.float64)
    imask, child_base, prim_base, iperm, lperm = ew >> 24, nodes[:, 4], nodes[:, 5], nodes[:, 6], nodes[:, 7]
    leafmask = lperm & 0xFF
    assert np.array_equal(iperm & 0xFF, imask) and not (imask & leafmask).any()
    for x in range(1, 4):  # byte x = the mask with the two low bits of every slot index XORed with x
        for j in range(8):
            assert np.array_equal((iperm >> (8 * x + j)) & 1, (imask >> (j ^ x)) & 1)
            assert np.array_equal((lperm >> (8 * x + j)) & 1, (leafmask >> (j ^ x)) & 1)

    def qbytes(w0, w1):  # two words -> [n, 8] bytes
        return np.stack([(w0 >> (8 * k)) & 0xFF for k in range(4)] + [(w1 >> (8 * k)) & 0xFF for k in range(4)], axis=1).astype(np.float64)

    qlo = np.stack([qbytes(nodes[:, 8], nodes[:, 9]), qbytes(nodes[:, 10], nodes[:, 11]), qbytes(nodes[:, 12], nodes[:, 13])], axis=2)  # [n, 8, 3]
    qhi = np.stack([qbytes(nodes[:, 14], nodes[:, 15]), qbytes(nodes[:, 16], nodes[:, 17]), qbytes(nodes[:, 18], nodes[:, 19])], axis=2)
    blo = p[:, None, :] + qlo * step[:, None, :]
    bhi = p[:, None, :] + qhi * step[:, None, :]
    used = imask | leafmask
    for sl in range(8):
        empty = ((used >> sl) & 1) == 0
        assert (qlo[empty, sl, :] > qhi[empty, sl, :]).all()  # inverted box: no ray can hit it
    # bottom-up over a depth-first order: actual bounds of every node's subtree
    sub_lo, sub_hi = np.full((n_nodes, 3), np.inf), np.full((n_nodes, 3), -np.inf)
    seen = np.zeros(n_nodes, bool)
    seen_prim = np.zeros(n_tris_expected, bool)
    order, stack, depth = [], [(0, 1)], 0
    while stack:
        i, d = stack.pop()
        assert not seen[i]
        seen[i] = True
        depth = max(depth, d)
        order.append(i)
        k = 0
        for sl in range(8):
            if (imask[i] >> sl) & 1:
                c = int(child_base[i]) + k
                assert 0 < c < n_nodes
                stack.append((c, d + 1))
                k += 1
    assert seen.all()
    for i in reversed(order):
        ki = kl = 0
        for sl in range(8):
            if (imask[i] >> sl) & 1:
                c = int(child_base[i]) + ki
                ki += 1
                lo, hi = sub_lo[c], sub_hi[c]
            elif (leafmask[i] >> sl) & 1:
                r = int(prim_base[i]) + kl
                kl += 1
                assert r < n_tris_expected and not seen_prim[r]
                seen_prim[r] = True
                lo, hi = tlo[r], thi[r]
            else:
                continue
            assert (blo[i, sl] <= lo).all() and (bhi[i, sl] >= hi).all(), (i, sl)  # the quantised box contains what is below the slot
            sub_lo[i], sub_hi[i] = np.minimum(sub_lo[i], lo), np.maximum(sub_hi[i], hi)
    assert seen_prim.all()
    fill = float(np.mean([bin(int(u)).count("1") for u in used]))
    return fill, depth


def blas_structure(pkg, make):
    """The builder's output, node by node (brt_debug_get_blas), for both collapse rules, a treelet and a plain LBVH build, a rebuild."""
    scene = pkg.scenes.make_scene("terrain", small=True)
    fills = {}
    for name, flags in (("optimal", 0), ("greedy", pkg.CFG_GREEDY_COLLAPSE), ("lbvh", pkg.CFG_NO_TREELET)):
        a = make(flags)
        scene.upload(a)
        for mesh_id, (_, v, idx) in enumerate(scene.meshes):
            nodes, tris = a.debug_get_blas(mesh_id)
            fill, depth = check_blas_structure(nodes, tris, len(idx) // 3)
            assert depth <= 24
            fills[(name, mesh_id)] = (fill, len(nodes))
        if name == "optimal":  # rebuild after a vertex update ("fast build" path: cost table filled by the refit walk)
            v2 = scene.meshes[1][1].copy()
            v2[:, 0:3] *= 1.25
            a.mesh_update_vertices(1, v2)
            a.scene_build()
            nodes, tris = a.debug_get_blas(1)
            check_blas_structure(nodes, tris, len(scene.meshes[1][2]) // 3)
    for mesh_id in range(len(scene.meshes)):  # the area-optimal choice: fuller nodes, fewer of them
        assert fills[("optimal", mesh_id)][0] > fills[("greedy", mesh_id)][0]
        assert fills[("optimal", mesh_id)][1] < fills[("greedy", mesh_id)][1]
    assert fills[("optimal", 0)][0] > 7.0 and fills[("optimal", 0)][1] < 0.5 * fills[("greedy", 0)][1]  # the 8192-triangle terrain


def grazing_and_axis_aligned_rays(pkg, orc_mod, make):
    """Rays with zero direction components, rays inside box faces and along triangle edges of the Cornell box."""
    scene = pkg.scenes.make_scene("cornell", small=True)
    a, b = make(), orc_mod.Oracle(pkg, brute_force=True)
    scene.upload(a)
    scene.upload(b)
    rays = []
    for ax in range(3):
        for sgn in (1.0, -1.0):
            for off in np.linspace(-1.0, 1.0, 41, dtype=np.float32):
                for off2 in (-1.0, -0.5, 0.0, 0.25, 1.0):
                    o = np.zeros(3, np.float32)
                    o[ax] = -3.0 * sgn
                    o[(ax + 1) % 3] = off
                    o[(ax + 2) % 3] = off2
                    d = np.zeros(3, np.float32)
                    d[ax] = sgn
                    rays.append(np.concatenate([o, [0.001], d, [1e32]]))
    rays = np.array(rays, np.float32)
    assert np.array_equal(a.trace_rays(rays, True), b.trace_rays(rays, True))
    assert np.array_equal(a.trace_rays(rays, False)[:, 3], b.trace_rays(rays, False)[:, 3])


def degenerate_extents(pkg, orc_mod, make):
    """Meshes whose bounds have no extent along one or two axes, or none at all (the extent-adaptive Morton code hands every bit to
    the axes that have some), many coincident triangles (equal codes: the hierarchy falls back to the index), one long strip:
    BVH == brute force for rays from all sides."""
    S = pkg.scenes
    rng = np.random.default_rng(41)

    def grid(nx, nz, sx, sz, y=0.0):  # planar grid in the plane y = const
        gx, gz = np.meshgrid(np.arange(nx + 1), np.arange(nz + 1), indexing="xy")
        v = np.zeros(((nx + 1) * (nz + 1), 8), np.float32)
        v[:, 0], v[:, 1], v[:, 2] = (gx.ravel() / max(nx, 1) - 0.5) * sx, y, (gz.ravel() / max(nz, 1) - 0.5) * sz
        v[:, 4] = -1.0
        i, j = np.meshgrid(np.arange(nx), np.arange(nz), indexing="xy")
        v00 = (j * (nx + 1) + i).ravel()
        idx = np.stack([v00, v00 + 1, v00 + nx + 2, v00, v00 + nx + 2, v00 + nx + 1], axis=1).astype(np.uint32).reshape(-1)
        return v, idx

    plane = grid(48, 48, 4.0, 4.0)                      # no extent in y
    strip = grid(3000, 1, 300.0, 0.01)                  # 6000 triangles, 30000 : 1
    tri = np.zeros((3, 8), np.float32)
    tri[:, 0:3] = [(-0.5, 0.3, -0.5), (0.5, 0.3, -0.5), (0.0, 0.3, 0.5)]
    tri[:, 4] = -1.0
    same = (np.tile(tri, (200, 1)), np.arange(600, dtype=np.uint32))  # 200 coincident triangles: no extent between the centres
    a, b = make(), orc_mod.Oracle(pkg, brute_force=True)
    for api in (a, b):
        mat = api.material_create((0.7, 0.7, 0.7))
        api.light_create((0, -3, 0), (1, 1, 1), 5.0)
        for k, (v, idx) in enumerate((plane, strip, same)):
            api.instance_create(api.mesh_create(v, idx), mat, S.xform(translate=(0.0, 0.5 * k, 0.0)))
        api.scene_build()
    n = 6000
    o = (rng.random((n, 3), dtype=np.float32) * 2 - 1) * np.array([3.0, 2.0, 3.0], np.float32)
    o[: n // 4, 0] *= 50.0  # along the strip
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = o, 0.001, d, 1e32
    ha, hb = a.trace_rays(rays, True), b.trace_rays(rays, True)
    assert np.array_equal(ha, hb)
    assert hb[:, 3].sum() > 500
    assert np.array_equal(a.trace_rays(rays, False)[:, 3], b.trace_rays(rays, False)[:, 3])
    st = a.get_stats()
    assert st.total_triangles == 48 * 48 * 2 + 6000 + 200 and 0 < st.bvh_nodes < st.total_triangles


def nonfinite_vertices(pkg, make):
    """A NaN and an infinite coordinate in a mesh: the build terminates (no endless loop in the Morton code, the cost table, the
    collapse or the quantisation), the frame renders, the finite triangles are still found."""
    v, idx = pkg.scenes.heightfield(16)
    v = v.copy()
    v[7, 0] = np.nan
    v[40, 1] = np.inf
    a = make()
    m = a.mesh_create(v, idx)
    a.instance_create(m, a.material_create((0.5, 0.5, 0.5)), pkg.scenes.xform())
    a.light_create((0, -3, 0), (1, 1, 1), 5.0)
    a.scene_build()
    st = a.get_stats()
    assert st.total_triangles == len(idx) // 3 and 0 < st.bvh_nodes < st.total_triangles
    u = a.camera_uniform((0, -3, -8), (-0.3, 0, 0), 1.0, 1.5, frame=0, depth_max=2)
    a.render_frame(u, a.opts(48, 32, 1, pkg.BOUNCE_REFLECT | pkg.BOUNCE_REFRACT))
    assert (a.get_aov(pkg.AOV_INST_ID, 48, 32) != pkg.AOV_MISS).mean() > 0.2


def bvh_quality_guard(pkg, make):
    """Regression guard for what the frame time follows (profiles/r1c_large_scenes.md: -8 % node visits = -8 % traversal time): node visits
    and primitive tests per ray and the SAH cost of the small terrain scene stay at what the extent-adaptive Morton code + treelets +
    area-optimal collapse deliver (measured 5.85 / 7.73 node visits per closest-hit / occlusion ray, SAH 22.9, 1299 nodes)."""
    scene = pkg.scenes.make_scene("terrain", small=True)
    a = make(pkg.CFG_COUNTERS)
    scene.upload(a)
    u = scene.uniform(a, 192, 108, 0, 3)
    a.render_frame(u, a.opts(192, 108, 1, pkg.BOUNCE_REFLECT | pkg.BOUNCE_REFRACT), want_image=False)
    s = a.get_stats()
    assert s.nodes_visited_closest / s.rays_closest < 6.1
    assert s.nodes_visited_occlusion / s.rays_occlusion < 8.1
    assert s.prims_tested_closest / s.rays_closest < 2.25
    assert s.sah_cost < 24.0 and s.sah_cost <= s.sah_cost_lbvh
    assert s.bvh_nodes < 1400


def edge_cases(pkg, orc_mod, make):
    """Empty scene, empty mesh, single triangle, duplicate triangles (tie-break), tiny and huge coordinates."""
    S = pkg.scenes
    # empty scene: every pixel misses, image is black with alpha 1
    a, b = make(), orc_mod.Oracle(pkg)
    for api in (a, b):
        api.scene_build()
    u = a.camera_uniform((0, 0, -2), (0, 0, 0), 1.0, 1.0, frame=0, depth_max=2)
    r = compare_frames(pkg, a, b, u, 40, 24)
    assert r["bit_exact"] and (a.get_aov(pkg.AOV_INST_ID, 40, 24) == pkg.AOV_MISS).all()
    assert np.array_equal(r["img"][..., 3], np.ones((24, 40), np.float32))
    # empty mesh + single triangle + two coincident triangles in different instances (lowest instance wins)
    tri = np.zeros((3, 8), np.float32)
    tri[:, 0:3] = [(-1, -1, 1), (1, -1, 1), (0, 1, 1)]
    tri[:, 3:6] = (0, 0, -1)
    a, b = make(), orc_mod.Oracle(pkg)
    for api in (a, b):
        e = api.mesh_create(np.zeros((0, 8), np.float32), np.zeros(0, np.uint32))
        m = api.mesh_create(tri, [0, 1, 2])
        dup = api.mesh_create(np.concatenate([tri, tri]), [0, 1, 2, 3, 4, 5])
        mat = api.material_create((0.8, 0.8, 0.8))
        api.light_create((0, 0, -1), (1, 1, 1), 3.0)
        api.instance_create(e, mat, S.xform())
        api.instance_create(m, mat, S.xform())
        api.instance_create(m, mat, S.xform())                       # coincident with the previous one
        api.instance_create(dup, mat, S.xform(translate=(3, 0, 0)))   # coincident primitives 0 and 1
        api.instance_create(m, mat, S.xform(scale=(1e-3, 1e-3, 1e-3), translate=(-0.5, 0.9, -1.9)))
        api.instance_create(m, mat, S.xform(scale=(3e3, 3e3, 3e3), translate=(0, 0, 5e3)))
        api.scene_build()
    u = a.camera_uniform((0.4, 0, -2), (0, 0.2, 0), 1.2, 1.5, frame=3, depth_max=2)
    r = compare_frames(pkg, a, b, u, 96, 64, R | D, 2)
    assert r["id_agreement"] == 1.0 and r["bit_exact"]
    inst = a.get_aov(pkg.AOV_INST_ID, 96, 64)
    prim = a.get_aov(pkg.AOV_PRIM_ID, 96, 64)
    assert (inst == 1).any() and not (inst == 2).any()           # tie between instances 1 and 2 -> 1
    assert ((inst == 3) & (prim == 0)).any() and not ((inst == 3) & (prim == 1)).any()


def dynamic_rebuild_and_instances(pkg, orc_mod, make):
    """Scene::prepareRendering placeholder (RT/Scene.cpp:135-138): vertex update + BLAS rebuild; instance
    set_transform / set_material / destroy (swap-remove, RT/Scene.cpp:122-125)."""
    scene = pkg.scenes.make_scene("lattice", small=True)
    a, b = make(), orc_mod.Oracle(pkg)
    scene.upload(a)
    scene.upload(b)
    base = scene.meshes[1][1]
    w, h = 128, 72
    for frame in range(3):
        v = pkg.scenes.animate_icosphere(base, frame)
        for api in (a, b):
            api.mesh_update_vertices(1, v)
            if frame == 1:
                api.instance_set_transform(0, pkg.scenes.xform((1.5, 1.5, 1.5), (0.5, -0.5, 9.0)))
                api.instance_set_material(2, 1)
            if frame == 2:
                api.instance_destroy(1)
            api.scene_build()
        assert a.get_stats().blas_built == 1
        u = scene.uniform(a, w, h, frame, 1)
        r = compare_frames(pkg, a, b, u, w, h)
        assert r["id_agreement"] == 1.0 and r["bit_exact"], frame


def smart_culling(pkg, orc_mod, make):
    """README.md:15-18: per-instance screen footprint, hysteresis, TLAS over the survivors."""
    scene = pkg.scenes.make_scene("lattice", small=True)
    a, b = make(), orc_mod.Oracle(pkg)
    scene.upload(a)
    scene.upload(b)
    n = len(scene.instances)
    w, h = 128, 72
    seen = set()
    for step, (z, thr) in enumerate([(-12.0, 40.0), (-30.0, 40.0), (-26.0, 40.0), (-12.0, 40.0), (-12.0, 0.0)]):
        u = a.camera_uniform((0.0, 0.0, z), (0, 0, 0), scene.fovy, w / h, frame=step, depth_max=1)
        va, vb = a.smart_cull(u, w, h, thr, 0.25), b.smart_cull(u, w, h, thr, 0.25)
        assert va == vb
        assert np.array_equal(a.get_visibility(n), b.get_visibility(n))
        seen.add(va)
        r = compare_frames(pkg, a, b, u, w, h)
        assert r["id_agreement"] == 1.0 and r["bit_exact"], step
        assert a.get_stats().instances_visible == va
    assert len(seen) > 1 and n in seen  # culling removed something, threshold 0 restored everything
    # the per-frame entry (Scene::prepareRendering): new vertices, then smart_cull alone rebuilds that BLAS and builds the TLAS once
    v = pkg.scenes.animate_icosphere(scene.meshes[1][1], 5)
    a.mesh_update_vertices(1, v)
    b.mesh_update_vertices(1, v)
    b.scene_build()
    u = a.camera_uniform((0.0, 0.0, -26.0), (0, 0, 0), scene.fovy, w / h, frame=9, depth_max=1)
    assert a.smart_cull(u, w, h, 40.0, 0.25) == b.smart_cull(u, w, h, 40.0, 0.25)
    assert a.get_stats().blas_built == 1
    r = compare_frames(pkg, a, b, u, w, h)
    assert r["id_agreement"] == 1.0 and r["bit_exact"]


def tiles_and_crop(pkg, orc_mod, make_ranked):
    """tile_rank / tile_world partition (32x32 tiles round-robin) and the crop window."""
    scene = pkg.scenes.make_scene("cornell", small=True)
    w, h = 100, 70  # not a multiple of the tile size
    full = orc_mod.Oracle(pkg)
    scene.upload(full)
    u = scene.uniform(full, w, h, 0, 2)
    ref = full.render_frame(u, full.opts(w, h, 1, R | T))
    world = 3
    acc = np.zeros_like(ref)
    for rank in range(world):
        api = make_ranked(rank, world)
        scene.upload(api)
        part = api.render_frame(u, api.opts(w, h, 1, R | T))
        ty, tx = np.meshgrid(np.arange(h) // 32, np.arange(w) // 32, indexing="ij")
        tile = ty * ((w + 31) // 32) + tx
        own = (tile % world + world - (tile // world) % world) % world == rank  # tile_of_rank (csrc/render_kernels.cuh)
        assert (part[~own] == 0).all()
        assert np.array_equal(part[own].view(np.uint32), ref[own].view(np.uint32))
        acc += part
    assert np.array_equal(acc.view(np.uint32), ref.view(np.uint32))
    api = make_ranked(0, 1)
    scene.upload(api)
    crop = (17, 9, 50, 40)
    part = api.render_frame(u, api.opts(w, h, 1, R | T, crop))
    mask = np.zeros((h, w), bool)
    mask[9:49, 17:67] = True
    assert (part[~mask] == 0).all() and np.array_equal(part[mask].view(np.uint32), ref[mask].view(np.uint32))


def error_behaviour(pkg, make):
    a = make()
    with pytest.raises(pkg.BrtError):
        a.mesh_create(np.zeros((3, 8), np.float32), [0, 1, 5])
    with pytest.raises(pkg.BrtError):
        a.instance_create(0, 0, np.eye(3, 4))
    with pytest.raises(pkg.BrtError):
        u = a.camera_uniform((0, 0, 0), (0, 0, 0), 1.0, 1.0)
        a.render_frame(u, a.opts(4, 4))
    with pytest.raises(pkg.BrtError):
        a.sphere_create((0, 0, 0), -1.0)


def general_transforms_and_many_lights(pkg, orc_mod, make):
    """Instances with rotation, non-uniform and mirrored (negative) scale — the C ABI takes any affine 3x4, beyond the
    scale + translate of RT/MeshInstance.h:82-85 — plus BRT_MAX_LIGHTS lights, a degenerate (zero-area) triangle and clear-coat /
    sheen / anisotropic / subsurface material fields."""
    S = pkg.scenes
    sv, si = S.icosphere(2)
    quad_v, quad_i = S._quad((-4, 1, -4), (4, 1, -4), (4, 1, 4), (-4, 1, 4))
    deg = np.zeros((3, 8), np.float32)
    deg[:, 0:3] = [(0, 0, 0), (1, 0, 0), (2, 0, 0)]  # collinear
    deg[:, 3:6] = (0, -1, 0)
    a, b = make(), orc_mod.Oracle(pkg)
    rng = np.random.default_rng(42)
    for api in (a, b):
        ms = api.mesh_create(sv, si)
        mq = api.mesh_create(quad_v, quad_i)
        md = api.mesh_create(deg, [0, 1, 2])
        mats = [api.material_create((0.8, 0.7, 0.6), metallic=0.2, roughness=0.6, clearCoat=1.0, clearCoatGloss=0.7, sheenTint=0.5),
                api.material_create((0.3, 0.6, 0.9), metallic=0.9, roughness=0.35, anisotropic=0.8, specularTint=0.6),
                api.material_create((0.9, 0.9, 0.9), metallic=0.0, roughness=0.9, subsurface=0.7, sheen=1.0)]
        for k in range(16):
            ang = 2 * np.pi * k / 16
            api.light_create((3 * np.cos(ang), -2.5, 3 * np.sin(ang)), (0.3 + 0.7 * (k % 3 == 0), 0.3 + 0.7 * (k % 3 == 1), 0.3 + 0.7 * (k % 3 == 2)), 1.5)
        api.instance_create(mq, mats[2], S.xform())
        api.instance_create(md, mats[0], S.xform(translate=(0, 0.5, 0)))
    r2 = np.random.default_rng(7)
    for k in range(12):
        ax = r2.normal(size=3)
        ax /= np.linalg.norm(ax)
        ang = r2.uniform(0, 2 * np.pi)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        Rm = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        sc = np.diag(r2.uniform(0.3, 0.9, size=3) * np.where(r2.random(3) < 0.25, -1.0, 1.0))
        x = np.zeros((3, 4), np.float32)
        x[:, :3] = (Rm @ sc).astype(np.float32)
        x[:, 3] = (r2.uniform(-2.5, 2.5), r2.uniform(-0.8, 0.6), r2.uniform(-2.0, 2.5))
        for api in (a, b):
            api.instance_create(0, k % 3, x)
    for api in (a, b):
        api.scene_build()
    u = a.camera_uniform((0.3, -1.2, -5.0), (-0.2, 0.05, 0.1), 1.0, 1.5, frame=2, depth_max=4)
    r = compare_frames(pkg, a, b, u, 144, 96, R | D | J, 2)
    assert r["id_agreement"] == 1.0 and r["bit_exact"] and r["t_agreement"] == 1.0
    assert r["stats"].rays_occlusion > 16 * 1000
    return r


def frames_in_flight(pkg, orc_mod, make):
    """brt_render_frame_async / brt_frame_wait (the reference's MAX_FRAMES_IN_FLIGHT = 2, VK/SwapChain.h:8): frames
    enqueued on the two slots, different cameras and sizes, give the bits of the synchronous call and of the oracle;
    a scene change drains the slots."""
    import ctypes as C
    scene = pkg.scenes.make_scene("terrain", small=True)
    a, b = make(), orc_mod.Oracle(pkg)
    scene.upload(a)
    scene.upload(b)
    shapes = [(160, 90, 0), (128, 72, 3), (160, 90, 7), (96, 54, 1)]
    uni = [scene.uniform(a, w, h, fr, 3) for (w, h, fr) in shapes]
    for k, u in enumerate(uni):
        u.viewInverse[3] += 0.3 * k  # move the eye sideways a little from frame to frame
    ref = [b.render_frame(u, b.opts(w, h, 2, R | T | J)).copy() for u, (w, h, _) in zip(uni, shapes)]
    out = [np.zeros((h, w, 4), np.float32) for (w, h, _) in shapes]
    for k, (u, (w, h, _)) in enumerate(zip(uni, shapes)):
        a.render_frame_async(u, a.opts(w, h, 2, R | T | J), k % 2, out[k].ctypes.data)  # waits for frame k-2 first
    a.frame_wait(0)
    a.frame_wait(1)
    for k in range(len(shapes)):
        assert np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)), k
    st = a.get_stats()
    assert st.rays_closest > 0 and st.launches_total > 0
    # statistics and AOVs refer to the frame waited for last (slot 1 = frame 3)
    w, h, _ = shapes[3]
    assert np.array_equal(a.get_aov(pkg.AOV_PRIM_ID, w, h), b.get_aov(pkg.AOV_PRIM_ID, w, h))
    # a scene change between asynchronous frames: the build drains the frame in flight, the next frame sees the new scene
    w, h, _ = shapes[0]
    a.render_frame_async(uni[0], a.opts(w, h, 1, 0), 0, out[0].ctypes.data)
    for api in (a, b):
        api.light_create((0.0, -6.0, 0.0), (0.3, 0.9, 0.4), 30.0)
        api.scene_build()
    first = out[0].copy()  # complete: brt_scene_build waited for slot 0
    a.render_frame_async(uni[0], a.opts(w, h, 1, 0), 1, out[2].ctypes.data)
    a.frame_wait(1)
    assert np.array_equal(out[2].view(np.uint32), b.render_frame(uni[0], b.opts(w, h, 1, 0)).view(np.uint32))
    assert not np.array_equal(first, out[2])
    with pytest.raises(pkg.BrtError):
        a.render_frame_async(uni[0], a.opts(w, h, 1, 0), 4, None)  # slot out of range (BRT_FRAMES_IN_FLIGHT = 4)


def present_formats(pkg, orc_mod, make):
    """Present path (Pipeline::rebuildRenderOutput(format, extent) + copyImageToSwapchain, RT/RTPipeline.cpp:49-55,
    RT/RTApp.cpp:87-152): the frame in the 8-bit swapchain formats is byte-identical to the oracle's, agrees with the
    textbook conversion of the float frame to within one code, and BGRA is RGBA with R and B swapped."""
    scene = pkg.scenes.make_scene("cornell", small=True)
    a, b = make(), orc_mod.Oracle(pkg)
    scene.upload(a)
    scene.upload(b)
    w, h = 96, 80
    u = scene.uniform(a, w, h, 0, 3)
    lin = a.render_frame(u, a.opts(w, h, 2, R | T | J)).copy()
    assert lin.dtype == np.float32 and lin[..., :3].max() > 1.0 and (lin[..., :3] == 0).any()  # exercises both clamps
    got = {}
    for fmt in (pkg.FORMAT_RGBA8_UNORM, pkg.FORMAT_BGRA8_UNORM, pkg.FORMAT_RGBA8_SRGB, pkg.FORMAT_BGRA8_SRGB):
        flags = R | T | J | pkg.render_format(fmt)
        img = a.render_frame(u, a.opts(w, h, 2, flags))
        ref = b.render_frame(u, b.opts(w, h, 2, flags))
        assert img.dtype == np.uint8 and img.shape == (h, w, 4)
        assert np.array_equal(img, ref), fmt
        got[fmt] = img
    c = np.clip(lin[..., :3].astype(np.float64), 0.0, 1.0)
    unorm = np.rint(c * 255.0)
    srgb = np.rint(np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(c, 1 / 2.4) - 0.055) * 255.0)
    assert np.abs(got[pkg.FORMAT_RGBA8_UNORM][..., :3].astype(np.float64) - unorm).max() <= 1
    assert (got[pkg.FORMAT_RGBA8_UNORM][..., :3] == unorm).mean() > 0.999
    assert np.abs(got[pkg.FORMAT_RGBA8_SRGB][..., :3].astype(np.float64) - srgb).max() <= 1
    assert (got[pkg.FORMAT_RGBA8_SRGB][..., :3] == srgb).mean() > 0.99
    for rgba, bgra in ((pkg.FORMAT_RGBA8_UNORM, pkg.FORMAT_BGRA8_UNORM), (pkg.FORMAT_RGBA8_SRGB, pkg.FORMAT_BGRA8_SRGB)):
        assert np.array_equal(got[rgba][..., [2, 1, 0, 3]], got[bgra])
        assert (got[rgba][..., 3] == 255).all()
    # the asynchronous entry point honours the format too, and the linear image stays available on the device
    out = np.zeros((h, w, 4), np.uint8)
    a.render_frame_async(u, a.opts(w, h, 2, R | T | J | pkg.render_format(pkg.FORMAT_BGRA8_SRGB)), 1, out.ctypes.data)
    a.frame_wait(1)
    assert np.array_equal(out, got[pkg.FORMAT_BGRA8_SRGB])
    with pytest.raises(pkg.BrtError):
        a.render_frame(u, a.opts(w, h, 1, pkg.render_format(7)))


def denoiser(pkg, orc_mod, make, w=96, h=96):
    """The denoiser slot (Graphics/Denoiser/Denoiser.h:5-20: temporal accumulation with reprojection, history clamping, variance
    estimation, a-trous wavelet, bilateral pass): G-buffer AOVs and every denoised frame of a short sequence — static camera, then
    a camera move (reprojection), then a reset — are bit-identical to the oracle's; and it denoises: the (tone-mapped) error
    against a many-sample reference drops."""
    scene = pkg.scenes.make_scene("cornell", small=True)
    a, b = make(), orc_mod.Oracle(pkg)
    scene.upload(a)
    scene.upload(b)
    flags = R | T | D | J | pkg.GBUFFER
    depth = 5
    tm = lambda img: img / (1.0 + img)
    u0 = scene.uniform(a, w, h, 0, depth)
    truth = tm(b.render_frame(u0, b.opts(w, h, 256, R | T | D | J))[..., :3].astype(np.float64))
    err = lambda img: float(np.sqrt(((tm(img[..., :3].astype(np.float64)) - truth) ** 2).mean()))
    dop = a.denoise_opts(iterations=4, sigma_n_log2=5, sigma_z=0.05, sigma_l=4.0, clamp_gamma=0.0, max_history=32.0, flags=pkg.DENOISE_BILATERAL)
    err_noisy, err_dn = [], []
    for frame in range(6):
        u = scene.uniform(a, w, h, frame * 7 + 1, depth)
        if frame >= 4:
            u.viewInverse[3] += 0.05 * (frame - 3)  # the eye moves: history is found through reprojection
        ia = a.render_frame(u, a.opts(w, h, 1, flags))
        ib = b.render_frame(u, b.opts(w, h, 1, flags))
        assert np.array_equal(ia.view(np.uint32), ib.view(np.uint32))
        if frame == 0:
            for kind in (pkg.AOV_POSITION, pkg.AOV_NORMAL):
                ga, gb = a.get_aov(kind, w, h), b.get_aov(kind, w, h)
                assert np.array_equal(ga.view(np.uint32), gb.view(np.uint32)), kind
            hit = a.get_aov(pkg.AOV_INST_ID, w, h) != pkg.AOV_MISS
            nrm = a.get_aov(pkg.AOV_NORMAL, w, h)
            assert np.allclose(np.linalg.norm(nrm[hit][:, :3], axis=1), 1.0, atol=1e-5) and (nrm[~hit] == 0).all()
        da = a.denoise(u, dop, w, h)
        db = b.denoise(u, dop, w, h)
        assert np.array_equal(da.view(np.uint32), db.view(np.uint32)), frame
        assert (da[..., 3] == 1.0).all() and np.isfinite(da).all()
        if frame < 4:
            err_noisy.append(err(ia))
            err_dn.append(err(da))
    assert all(d < 0.65 * n for d, n in zip(err_dn, err_noisy)), (err_dn, err_noisy)
    assert a.get_stats().launches_denoise == 6 and a.get_stats().ms_denoise >= 0.0
    # temporal accumulation alone (clamped history, no wavelet passes): identical too, and the error keeps falling
    dop2 = a.denoise_opts(iterations=0, clamp_gamma=4.0, max_history=8.0, flags=pkg.DENOISE_RESET)
    acc = []
    for frame in range(5):
        u = scene.uniform(a, w, h, 100 + frame, depth)
        for api in (a, b):
            api.render_frame(u, api.opts(w, h, 1, flags))
        da, db = a.denoise(u, dop2, w, h), b.denoise(u, dop2, w, h)
        dop2.flags = 0
        assert np.array_equal(da.view(np.uint32), db.view(np.uint32)), frame
        acc.append(err(da))
    assert acc[4] < 0.8 * acc[0], acc
    # reset drops the history: with no filter either, the result is the frame itself
    dop3 = a.denoise_opts(iterations=0, flags=pkg.DENOISE_RESET)
    ia = a.render_frame(u, a.opts(w, h, 1, flags))
    da = a.denoise(u, dop3, w, h)
    assert np.array_equal(da[..., :3].view(np.uint32), ia[..., :3].view(np.uint32))
    # a frame without the G-buffer cannot be denoised; bad options are rejected
    a.render_frame(u, a.opts(w, h, 1, R))
    with pytest.raises(pkg.BrtError):
        a.denoise(u, dop, w, h)
    a.render_frame(u, a.opts(w, h, 1, flags))
    with pytest.raises(pkg.BrtError):
        a.denoise(u, a.denoise_opts(iterations=9), w, h)


def light_bvh(pkg, orc_mod, make):
    """Light BVH (LightBVHNode, RT/Scene.h:123-130; announced at SH/raytracing.slang:76): the tree is the oracle's, node for node;
    frames lit by ONE importance-sampled light per hit are bit-identical to the oracle's — with far more lights than the loop
    allows — and the estimator is unbiased: many samples converge to the full loop over the lights."""
    LB = pkg.LIGHT_BVH
    scene = pkg.scenes.make_scene("cornell", small=True)
    rng = np.random.default_rng(5)
    extra = [((rng.random(3) * 1.6 - 0.8).tolist(), (0.2 + 0.8 * rng.random(3)).tolist(), float(0.02 + 0.2 * rng.random())) for _ in range(11)]
    a, b = make(), orc_mod.Oracle(pkg)
    for api in (a, b):
        scene.upload(api, build=False)
        for pos, col, inten in extra:
            api.light_create(pos, col, inten)
        api.scene_build()
    n_lights = len(scene.lights) + len(extra)
    assert n_lights <= 16
    ta, tb = a.get_light_bvh(), b.get_light_bvh()
    assert len(ta) == len(tb) == 2 * n_lights - 1
    for x, y in zip(ta, tb):
        assert bytes(x) == bytes(y)
    leaves = sorted(-1 - n.childIndex for n in ta if n.childIndex < 0)
    assert leaves == list(range(n_lights))  # every light in exactly one leaf
    for n in ta:
        if n.childIndex >= 0:
            l, r = ta[n.childIndex], ta[n.childIndex + 1]
            assert abs(n.totalFlux - (l.totalFlux + r.totalFlux)) <= 1e-5 * n.totalFlux
            assert all(n.bBoxMin[k] <= min(l.bBoxMin[k], r.bBoxMin[k]) and n.bBoxMax[k] >= max(l.bBoxMax[k], r.bBoxMax[k]) for k in range(3))
    w, h = 64, 64
    u = scene.uniform(a, w, h, 3, 3)
    for flags, spp in ((LB, 1), (LB | R | T | D | J, 2)):
        r = compare_frames(pkg, a, b, u, w, h, flags, spp)
        assert r["bit_exact"] and r["id_agreement"] == 1.0, flags
        assert r["stats"].rays_occlusion == r["ref_stats"].rays_occlusion <= r["stats"].rays_closest  # at most one shadow ray per hit
    # unbiased: 512 sampled frames average to the loop over all lights (direct light only, no other randomness)
    full = a.render_frame(u, a.opts(w, h, 1, 0))[..., :3].astype(np.float64)
    est = a.render_frame(u, a.opts(w, h, 512, LB))[..., :3].astype(np.float64)
    tm = lambda x: x / (1.0 + x)
    assert np.sqrt(((tm(est) - tm(full)) ** 2).mean()) < 0.02 * max(tm(full).mean(), 1e-6) + 0.004
    one = a.render_frame(u, a.opts(w, h, 1, LB))[..., :3].astype(np.float64)
    assert np.sqrt(((tm(one) - tm(full)) ** 2).mean()) > 4 * np.sqrt(((tm(est) - tm(full)) ** 2).mean())  # and it is an estimator
    # more lights than BRT_MAX_LIGHTS: the loop refuses, the light BVH renders (and still matches the oracle)
    for api in (a, b):
        for k in range(40):
            api.light_create(((k % 8) * 0.2 - 0.7, -0.6 + 0.03 * k, (k // 8) * 0.3 - 0.6), (1.0, 0.9 - 0.01 * k, 0.5 + 0.01 * k), 0.05)
        api.scene_build()
    with pytest.raises(pkg.BrtError):
        a.render_frame(u, a.opts(w, h, 1, 0))
    r = compare_frames(pkg, a, b, u, w, h, LB | D | J, 2)
    assert r["bit_exact"]
    assert len(a.get_light_bvh()) == 2 * (n_lights + 40) - 1
    # SPOT / DIRECTIONAL lights (the shader's constant-direction branch, SH/light.slang:33-36) go into their own subtree under the root:
    # cone angle 0, importance independent of the shading point; tree and frames still equal the oracle's, the estimator stays unbiased
    a2, b2 = make(), orc_mod.Oracle(pkg)
    for api in (a2, b2):
        scene.upload(api, build=False)
        for k, (pos, col, inten) in enumerate(extra[:6]):
            api.light_create(pos, col, inten)
        api.light_create((0, 0, 0), (0.3, 0.25, 0.2), 0.4, type=2)
        api.light_create((0.5, 0.5, 0.5), (0.1, 0.2, 0.3), 0.3, type=1)
        api.light_create((0.1, -0.2, 0.3), (0.2, 0.1, 0.1), 0.2, type=2)
        api.scene_build()
    ta, tb = a2.get_light_bvh(), b2.get_light_bvh()
    n2 = len(scene.lights) + 6 + 3
    assert len(ta) == len(tb) == 2 * n2 - 1 and all(bytes(x) == bytes(y) for x, y in zip(ta, tb))
    assert ta[0].childIndex == 1 and ta[1].coneAngle > 3.0 and ta[2].coneAngle == 0.0 and abs(ta[2].coneAxis[0] + 0.9938837) < 1e-6
    assert sorted(-1 - n.childIndex for n in ta if n.childIndex < 0) == list(range(n2))
    for flags, spp in ((LB, 1), (LB | R | T | D | J, 2)):
        r = compare_frames(pkg, a2, b2, u, w, h, flags, spp)
        assert r["bit_exact"] and r["id_agreement"] == 1.0, flags
    full = a2.render_frame(u, a2.opts(w, h, 1, 0))[..., :3].astype(np.float64)
    est = a2.render_frame(u, a2.opts(w, h, 512, LB))[..., :3].astype(np.float64)
    assert np.sqrt(((tm(est) - tm(full)) ** 2).mean()) < 0.02 * max(tm(full).mean(), 1e-6) + 0.004
    # only constant-direction lights: one subtree, no root split
    a3, b3 = make(), orc_mod.Oracle(pkg)
    for api in (a3, b3):
        for mesh in scene.meshes:
            api.sphere_create(mesh[1], mesh[2]) if mesh[0] == "sphere" else api.mesh_create(mesh[1], mesh[2])
        for md in scene.materials:
            api.material_create(md.color, md.metallic, md.roughness, md.specular, **md.extra)
        api.light_create((0, 0, 0), (0.5, 0.5, 0.5), 0.5, type=2)
        api.light_create((1, 0, 0), (0.2, 0.5, 0.1), 0.25, type=1)
        for mesh, mat, x in scene.instances:
            api.instance_create(mesh, mat, x)
        api.scene_build()
    ta, tb = a3.get_light_bvh(), b3.get_light_bvh()
    assert len(ta) == len(tb) == 3 and all(bytes(x) == bytes(y) for x, y in zip(ta, tb)) and ta[0].coneAngle == 0.0
    assert compare_frames(pkg, a3, b3, u, w, h, LB, 1)["bit_exact"]


def denoise_in_flight(pkg, orc_mod, make, w=96, h=96):
    """BRT_RENDER_DENOISE: the denoiser stages run on the frame's own stream right after the resolve, also for frames in flight on
    rotating slots (their shared history is handed from frame to frame in submission order); the images handed back — linear
    float and the 8-bit present format — equal the oracle's render + denoise (+ present conversion) of the same sequence."""
    scene = pkg.scenes.make_scene("cornell", small=True)
    a, b = make(), orc_mod.Oracle(pkg)
    scene.upload(a)
    scene.upload(b)
    base = R | T | D | J
    dop = a.denoise_opts(iterations=3, sigma_n_log2=4, sigma_z=0.05, sigma_l=4.0, clamp_gamma=2.0, max_history=16.0, flags=pkg.DENOISE_BILATERAL)
    a.denoise_configure(dop)
    n = 7
    unis = []
    for k in range(n):
        u = scene.uniform(a, w, h, 3 * k + 1, 4)
        u.viewInverse[3] += 0.02 * k
        unis.append(u)
    outs = [np.zeros((h, w, 4), np.float32) for _ in range(n)]
    for k in range(n):
        a.render_frame_async(unis[k], a.opts(w, h, 1, base | pkg.DENOISE), k % 3, outs[k].ctypes.data)
    for slot in range(3):
        a.frame_wait(slot)
    for k in range(n):
        b.render_frame(unis[k], b.opts(w, h, 1, base | pkg.GBUFFER))
        ref = b.denoise(unis[k], dop, w, h)
        assert np.array_equal(outs[k].view(np.uint32), ref.view(np.uint32)), k
    st = a.get_stats()
    assert st.launches_denoise == 5 and st.ms_denoise >= 0.0
    # synchronous call, present format: the denoised frame converted on the GPU
    u = scene.uniform(a, w, h, 40, 4)
    img8 = a.render_frame(u, a.opts(w, h, 1, base | pkg.DENOISE | pkg.render_format(pkg.FORMAT_BGRA8_SRGB)))
    b.render_frame(u, b.opts(w, h, 1, base | pkg.GBUFFER))
    ref = b.denoise(u, dop, w, h)
    c = np.clip(ref[..., :3].astype(np.float64), 0.0, 1.0)
    srgb = np.rint(np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(c, 1 / 2.4) - 0.055) * 255.0)
    assert img8.dtype == np.uint8 and np.abs(img8[..., [2, 1, 0]].astype(np.float64) - srgb).max() <= 1
    assert (img8[..., [2, 1, 0]] == srgb).mean() > 0.99 and (img8[..., 3] == 255).all()
    # the separate entry point continues the same history
    a.render_frame(unis[0], a.opts(w, h, 1, base | pkg.GBUFFER))
    b.render_frame(unis[0], b.opts(w, h, 1, base | pkg.GBUFFER))
    assert np.array_equal(a.denoise(unis[0], dop, w, h).view(np.uint32), b.denoise(unis[0], dop, w, h).view(np.uint32))


def golden_frames(pkg, make):
    """The committed golden fixtures (tests/golden/oracle_frames.npz, generator make_golden.py) reproduced by the implementation under
    test WITHOUT the oracle in the loop: radiance, primitive and instance ids bit for bit — including the light-BVH frame, the 8-bit
    present frame and the third frame of the denoised sequence."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    g = np.load(os.path.join(here, "golden", "oracle_frames.npz"))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden

    class _Mod:  # make_golden.render_all builds its contexts through `orc_mod.Oracle(pkg)`: hand it the implementation under test
        @staticmethod
        def Oracle(_pkg):
            return make()

    for name, (img, prim, inst) in make_golden.render_all(pkg, _Mod).items():
        assert np.array_equal(g[name + "_prim"], prim), name
        assert np.array_equal(g[name + "_inst"], inst), name
        assert np.array_equal(g[name + "_img"].view(np.uint32), np.ascontiguousarray(img).view(np.uint32)), name
