"""Host logic: camera / uniform maths (Graphics/Camera.cpp:8-17,71-95; RT/RTApp.cpp:44-49), scene generators."""
import numpy as np


def test_camera_uniform_matches_oracle_and_analytic(pkg, orc_mod):
    lib = pkg.load()
    B = pkg.binding
    import ctypes as C
    lib.brt_camera_uniform.restype = None
    lib.brt_camera_uniform.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float, C.c_float,
                                       C.c_uint32, C.c_uint32, C.POINTER(B.Uniform)]
    olib = orc_mod.load()
    olib.orc_camera_uniform.restype = None
    olib.orc_camera_uniform.argtypes = lib.brt_camera_uniform.argtypes
    rng = np.random.default_rng(0)
    for k in range(50):
        pos = (rng.normal(size=3) * 5).astype(np.float32)
        rot = ((rng.random(3) - 0.5) * 2.5).astype(np.float32)
        fovy, aspect = float(np.float32(0.5 + rng.random())), float(np.float32(0.7 + rng.random()))
        a, b = B.Uniform(), B.Uniform()
        args = ((C.c_float * 3)(*pos), (C.c_float * 3)(*rot), fovy, aspect, 0.001, 100000.0, k, 2)
        lib.brt_camera_uniform(*args, C.byref(a))
        olib.orc_camera_uniform(*args, C.byref(b))
        va, vb = np.array(a.viewInverse).reshape(4, 4), np.array(b.viewInverse).reshape(4, 4)
        pa, pb = np.array(a.projInverse).reshape(4, 4), np.array(b.projInverse).reshape(4, 4)
        assert np.allclose(va, vb, rtol=2e-7, atol=1e-7) and np.allclose(pa, pb, rtol=2e-7, atol=1e-6)
        assert (a.frame, a.depthMax) == (k, 2)
        # origin = V^-1 (0,0,0,1) = camera position (SH/raytracing.slang:105)
        assert np.allclose(va[:3, 3], pos, atol=1e-5)
        # rotation part is orthonormal
        assert np.allclose(va[:3, :3] @ va[:3, :3].T, np.eye(3), atol=1e-5)
        # P^-1 (x, y, 1, 1).xyz = (x a tan(f/2), y tan(f/2), 1)
        t = np.tan(fovy / 2)
        v = pa @ np.array([0.3, -0.6, 1.0, 1.0])
        assert np.allclose(v[:3], [0.3 * aspect * t, -0.6 * t, 1.0], rtol=1e-5)
    # yaw 0: +Z forward (Graphics/Camera.cpp:42)
    a = B.Uniform()
    lib.brt_camera_uniform((C.c_float * 3)(0, 0, -2), (C.c_float * 3)(0, 0, 0), 1.0, 1.0, 0.001, 1e5, 0, 2, C.byref(a))
    assert np.allclose(np.array(a.viewInverse).reshape(4, 4), [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, -2], [0, 0, 0, 1]])


def test_scene_generators(pkg):
    S = pkg.scenes
    c = S.cornell()
    assert c.triangles() == 32 and sum(1 for m in c.meshes if m[0] == "sphere") == 2
    v, i = S.icosphere(2)
    assert len(i) // 3 == 320 and np.allclose(np.linalg.norm(v[:, :3], axis=1), 1.0, atol=1e-6)
    hv, hi = S.heightfield(16)
    assert len(hi) // 3 == 512 and hi.max() < len(hv)
    assert S.terrain_icospheres(64, 2, 24).triangles() == 2 * 64 * 64 + 24 * 320
    assert S.instanced_lattice(3, 2).triangles() == 28 * 320
    # deterministic
    assert np.array_equal(S.heightfield(16)[0], hv)
    for name, cfg in S.CONFIGS.items():
        assert cfg["width"] * cfg["height"] > 0 and cfg["depth_max"] >= 1


def test_camera_handle_inputs(pkg, orc_mod):
    """Camera::handleInputs (Graphics/Camera.cpp:26-61): product host code == oracle restatement == the rule written out."""
    import ctypes as C
    lib, olib = pkg.load(), orc_mod.load()
    sig = [C.c_uint32, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    for f in (lib.brt_camera_handle_inputs, olib.orc_camera_handle_inputs):
        f.restype, f.argtypes = None, sig
    rng = np.random.default_rng(7)
    pos = np.array([0.0, 0.0, -2.0], np.float32)  # RTApp::RTApp, RT/RTApp.cpp:25
    rot = np.zeros(3, np.float32)
    for k in range(400):
        keys = int(rng.integers(0, 1024)) if k % 5 else 0
        dt = float(np.float32(rng.random() * 0.05))
        pa, ra = (C.c_float * 3)(*pos), (C.c_float * 3)(*rot)
        pb, rb = (C.c_float * 3)(*pos), (C.c_float * 3)(*rot)
        lib.brt_camera_handle_inputs(keys, dt, pa, ra)
        olib.orc_camera_handle_inputs(keys, dt, pb, rb)
        assert list(pa) == list(pb) and list(ra) == list(rb), (k, keys)
        # the rule, in float64
        key = lambda b: bool(keys & b)
        rx = key(256) - key(512)
        ry = key(64) - key(128)
        r = rot.astype(np.float64)
        if rx or ry:
            n = np.hypot(rx, ry)
            r[0] += 1.5 * dt * rx / n
            r[1] += 1.5 * dt * ry / n
        r[0] = np.clip(r[0], -1.5, 1.5)
        r[1] = np.mod(r[1], 2 * np.pi)
        fwd = np.array([np.sin(r[1]), 0.0, np.cos(r[1])])
        right = np.array([fwd[2], 0.0, -fwd[0]])
        up = np.array([0.0, -1.0, 0.0])
        mv = fwd * (key(4) - key(8)) + right * (key(2) - key(1)) + up * (key(16) - key(32))
        p = pos.astype(np.float64)
        if np.dot(mv, mv) > 1.2e-7:
            p += 3.0 * dt * mv / np.linalg.norm(mv)
        assert np.allclose(np.array(pa), p, atol=1e-5) and np.allclose(np.array(ra), r, atol=1e-5), (k, keys)
        pos, rot = np.array(pa, np.float32), np.array(ra, np.float32)
        assert -1.5 <= rot[0] <= 1.5 and 0.0 <= rot[1] < 2 * np.pi + 1e-6
    # no key: nothing moves
    pa, ra = (C.c_float * 3)(1, 2, 3), (C.c_float * 3)(0.1, 0.2, 0.3)
    lib.brt_camera_handle_inputs(0, 0.016, pa, ra)
    assert np.allclose(list(pa), [1, 2, 3]) and np.allclose(list(ra), [0.1, 0.2, 0.3])
