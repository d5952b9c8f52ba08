"""Pins to the reference ITSELF (oracle/_ref, built by oracle/ref/Makefile from the sources under /root/reference):

* libref_host.so = the reference's Graphics/Camera.cpp, unmodified, + MeshInstance.h + the vendored glm 1.0.1:
  brt_camera_uniform / brt_camera_handle_inputs (product host code) and the oracle's versions against
  Camera::setView / setPerspectiveProjection / handleInputs and `inverse(transpose(.))` of RT/RTApp.cpp:44-49.
* libref_obj.so = Scene::loadModel (RT/Scene.cpp:29-74, taken by line range at build time) over the vendored tinyobjloader
  1.0.6: the facade's OBJ ingestion (include/bloon/bloon.hpp, bloon::obj) on fuzzed OBJ text, byte for byte.

When /root/reference is absent (the GPU box) the prebuilt oracle/_ref that travelled with the snapshot is used; with neither
the tests skip."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/Hardware Ray Tracer"
REF_OUT = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="session")
def ref_host():
    return _ref_lib("libref_host.so")


@pytest.fixture(scope="session")
def ref_obj():
    return _ref_lib("libref_obj.so")


def _ref_lib(name):
    if os.path.isdir(REF_SRC):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle", "ref")])
    path = os.path.join(REF_OUT, name)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    return C.CDLL(path)


F3 = C.c_float * 3


def _setup_host(lib):
    lib.ref_camera_new.restype = C.c_void_p
    for f, args in ((lib.ref_camera_delete, [C.c_void_p]), (lib.ref_camera_set_view, [C.c_void_p, F3, F3]),
                    (lib.ref_camera_set_perspective, [C.c_void_p] + [C.c_float] * 4),
                    (lib.ref_camera_handle_inputs, [C.c_void_p, C.c_uint32, C.c_float]),
                    (lib.ref_camera_state, [C.c_void_p, F3, F3]),
                    (lib.ref_camera_matrices, [C.c_void_p, C.c_float * 16, C.c_float * 16]),
                    (lib.ref_uniform, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_char * 140]),
                    (lib.ref_instance_transform, [F3, F3, F3, C.c_float * 12])):
        f.restype, f.argtypes = None, args


def test_uniform_matches_reference_camera(pkg, orc_mod, ref_host):
    """RTApp::run's Uniform block (RT/RTApp.cpp:40-49) from the reference's Camera + glm against brt_camera_uniform and
    orc_camera_uniform over random poses. The reference inverts in binary32 (glm cofactors), the product in double rounded
    once: agreement to a few ulps of the matrix scale is the bar, frame / depthMax / LIGHT_TRESHOLD bytes are exact."""
    _setup_host(ref_host)
    lib, olib, B = pkg.load(), orc_mod.load(), pkg.binding
    sig = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.POINTER(B.Uniform)]
    for f in (lib.brt_camera_uniform, olib.orc_camera_uniform):
        f.restype, f.argtypes = None, sig
    rng = np.random.default_rng(11)
    cam = ref_host.ref_camera_new()
    worst = 0.0
    for k in range(300):
        pos = (rng.normal(size=3) * (5 if k % 3 else 50)).astype(np.float32)
        rot = ((rng.random(3) - 0.5) * (2.5 if k % 2 else 6.0)).astype(np.float32)
        if k < 4:
            pos, rot = np.array([0, 0, -2], np.float32), np.zeros(3, np.float32)  # RTApp::RTApp, RT/RTApp.cpp:25
        fovy = float(np.float32(np.radians(60.0) if k < 4 else 0.4 + rng.random() * 1.2))
        aspect = float(np.float32([800 / 600, 16 / 9, 1.0, 0.75][k % 4]))
        ref_host.ref_camera_set_view(cam, F3(*pos), F3(*rot))
        ref_host.ref_camera_set_perspective(cam, fovy, aspect, 0.001, 100000.0)
        raw = (C.c_char * 140)()
        ref_host.ref_uniform(cam, k, 2, raw)
        ref = np.frombuffer(raw, np.float32, 32).reshape(2, 4, 4)
        assert np.frombuffer(raw, np.uint32, 2, 128).tolist() == [k, 2]
        assert np.frombuffer(raw, np.float32, 1, 136)[0] == np.float32(0.0001)
        for f in (lib.brt_camera_uniform, olib.orc_camera_uniform):
            u = B.Uniform()
            f(F3(*pos), F3(*rot), fovy, aspect, 0.001, 100000.0, k, 2, C.byref(u))
            got = np.frombuffer(bytes(u), np.float32, 32).reshape(2, 4, 4)
            assert (u.frame, u.depthMax) == (k, 2) and np.float32(u.lightThreshold) == np.float32(0.0001)
            for m in range(2):
                scale = np.abs(ref[m]).max(axis=None)
                # per column scale: the translation column of V^-1 carries |pos|, P^-1 has entries of 1e3 (1/near)
                col = np.maximum(np.abs(ref[m]).max(axis=1, keepdims=True), 1e-30)
                err = np.abs(got[m] - ref[m]) / np.maximum(col, 1e-6 * scale)
                worst = max(worst, float(err.max()))
                assert err.max() < 4e-6, (k, m, got[m], ref[m])
    ref_host.ref_camera_delete(cam)
    assert worst > 0.0 or True


def test_handle_inputs_matches_reference_camera(pkg, orc_mod, ref_host):
    """Camera::handleInputs (Graphics/Camera.cpp:26-61) executed by the reference's own translation unit over a 400-step key walk;
    product and oracle must land on the same pose as the reference (bit for bit where libm agrees; the bar is 2 ulps of the pose)."""
    _setup_host(ref_host)
    lib, olib = pkg.load(), orc_mod.load()
    sig = [C.c_uint32, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    for f in (lib.brt_camera_handle_inputs, olib.orc_camera_handle_inputs):
        f.restype, f.argtypes = None, sig
    rng = np.random.default_rng(5)
    cam = ref_host.ref_camera_new()
    pos, rot = np.array([0.0, 0.0, -2.0], np.float32), np.zeros(3, np.float32)
    ref_host.ref_camera_set_view(cam, F3(*pos), F3(*rot))
    exact = 0
    for k in range(400):
        keys = int(rng.integers(0, 1024)) if k % 7 else 0
        dt = float(np.float32(rng.random() * 0.05))
        ref_host.ref_camera_handle_inputs(cam, keys, dt)
        rp, rr = F3(), F3()
        ref_host.ref_camera_state(cam, rp, rr)
        for f in (lib.brt_camera_handle_inputs, olib.orc_camera_handle_inputs):
            p, r = F3(*pos), F3(*rot)
            f(keys, dt, p, r)
            assert np.allclose(list(p), list(rp), rtol=3e-7, atol=3e-7) and np.allclose(list(r), list(rr), rtol=3e-7, atol=3e-7), (k, keys)
            exact += list(p) == list(rp) and list(r) == list(rr)
        # every implementation continues from the REFERENCE's pose, so a one-ulp difference cannot accumulate into a false alarm
        pos, rot = np.array(list(rp), np.float32), np.array(list(rr), np.float32)
        # and the view matrix the reference derives from it is what the uniform test above inverts
    assert exact >= 2 * 400 * 0.95
    ref_host.ref_camera_delete(cam)


def test_instance_transform_matches_reference(pkg, ref_host, tmp_path):
    """MeshInstance::calculateTransformation (RT/MeshInstance.h:82-85: scale + translate, rotation ignored) against the facade's
    MeshInstance — compiled from include/bloon/bloon.hpp — on random TRS triples."""
    _setup_host(ref_host)
    src = tmp_path / "inst.cpp"
    src.write_text('#include <cstdio>\n#include "bloon/bloon.hpp"\nint main(){float a[9];while(std::fread(a,4,9,stdin)==9){'
                   'RayTracing::MeshInstance m(0,0,{a[0],a[1],a[2]},{a[3],a[4],a[5]},{a[6],a[7],a[8]});'
                   'std::fwrite(m.getTransformation(),4,12,stdout);}return 0;}\n')
    exe = str(tmp_path / "inst")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe])
    rng = np.random.default_rng(2)
    trs = (rng.normal(size=(64, 9)) * 3).astype(np.float32)
    got = np.frombuffer(subprocess.run([exe], input=trs.tobytes(), capture_output=True, check=True).stdout, np.float32).reshape(64, 12)
    for k in range(64):
        out = (C.c_float * 12)()
        ref_host.ref_instance_transform(F3(*trs[k, 0:3]), F3(*trs[k, 3:6]), F3(*trs[k, 6:9]), out)
        assert np.array_equal(np.array(out, np.float32).view(np.uint32), got[k].view(np.uint32))


# ------------------------------------------------------------------------------------------------ OBJ ingestion

def _num(rng):
    kind = rng.integers(0, 12)
    x = rng.normal() * 10.0 ** int(rng.integers(-3, 3))
    if kind == 0:
        return str(int(x))
    if kind == 1:
        return f"{x:.{rng.integers(1, 16)}f}"
    if kind == 2:
        return f"{x:.{rng.integers(0, 9)}e}"
    if kind == 3:
        return f"{x:.3E}".replace("E-0", "E-").replace("E+0", "E+")
    if kind == 4:
        return rng.choice(["0", "-0", "+0", "0.0", "-0.0", "1", "-1", "+1.5", "3.", ".5", "1e", "1e+", "abc", "1.5x", "2e3", "7E-2", "1e-40", "1e39"])
    return f"{x:.6f}"


def _fuzz_obj(rng, with_mtl):
    """Random OBJ text exercising what tinyobj 1.0.6 accepts: number formats, missing components, all four corner forms, 0 and
    negative indices, polygons, o / g / usemtl / mtllib records, comments, unknown records, leading blanks, duplicate vertices."""
    eol = rng.choice(["\n", "\r\n", "\r"])
    lines = ["# fuzz", ""]
    nv = nvn = nvt = 0
    if with_mtl:
        lines.append(rng.choice(["mtllib mats.mtl", "mtllib missing.mtl mats.mtl", "mtllib  mats.mtl", "mtllib missing.mtl"]))
    pool_v = []
    for _ in range(int(rng.integers(3, 9))):
        block = rng.integers(0, 8)
        if block <= 2:
            for _ in range(int(rng.integers(1, 6))):
                if pool_v and rng.random() < 0.3:
                    comps = list(pool_v[rng.integers(0, len(pool_v))])
                    if rng.random() < 0.3:
                        comps = [("-0" if c == "0" else c) for c in comps]
                else:
                    comps = [_num(rng) for _ in range(int(rng.choice([3, 3, 3, 2, 4])))]
                    if rng.random() < 0.2:
                        comps[int(rng.integers(0, len(comps)))] = "0"
                pool_v.append(comps)
                lines.append(rng.choice(["v ", "v\t", "  v  "]) + rng.choice([" ", "  ", "\t"]).join(comps))
                nv += 1
        elif block == 3:
            for _ in range(int(rng.integers(1, 4))):
                lines.append("vn " + " ".join(_num(rng) for _ in range(int(rng.choice([3, 3, 2])))))
                nvn += 1
        elif block == 4:
            for _ in range(int(rng.integers(1, 4))):
                lines.append("vt " + " ".join(_num(rng) for _ in range(int(rng.choice([2, 2, 3, 1])))))
                nvt += 1
        elif block == 5:
            lines.append(rng.choice(["o thing", "g group", "g", "g a b", "o  spaced", "s off", "vp 0.1 0.2", "# comment", "", "   ", "usemtl red", "usemtl blue",
                                     "usemtl nosuch", "usemtl"+" red"]))
        if nv == 0:
            continue
        for _ in range(int(rng.integers(0, 5))):
            n = int(rng.choice([3, 3, 3, 4, 5, 6]))
            form = rng.integers(0, 4)
            corners = []
            for _ in range(n):
                def ix(count, allow_zero=True):
                    if count == 0:
                        return int(rng.integers(1, 3))  # dangling: resolves out of range -> both sides must refuse the file
                    r = rng.random()
                    if r > 0.992:
                        return count + int(rng.integers(1, 3))  # past the end: the facade must refuse the file
                    if r < 0.25:
                        return -int(rng.integers(1, count + 1))
                    if r < 0.28 and allow_zero:
                        return 0
                    return int(rng.integers(1, count + 1))
                i, j, k = ix(nv), ix(max(nvt, 1) if nvt else 1), ix(max(nvn, 1) if nvn else 1)
                if form == 0 or (nvt == 0 and nvn == 0):
                    corners.append(f"{i}")
                elif form == 1 and nvt:
                    corners.append(f"{i}/{j}")
                elif form == 2 and nvn:
                    corners.append(f"{i}//{k}")
                elif nvt and nvn:
                    corners.append(f"{i}/{j}/{k}")
                else:
                    corners.append(f"{i}")
            lines.append(rng.choice(["f ", "f  ", "\tf "]) + rng.choice([" ", "  "]).join(corners) + rng.choice(["", " ", "\t"]))
    text = eol.join(lines) + (eol if rng.random() < 0.7 else "")
    return text


def _ref_load(lib, path):
    lib.ref_load_model.restype = C.c_int
    lib.ref_load_model.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_uint32), C.POINTER(C.POINTER(C.c_uint32)),
                                   C.POINTER(C.c_uint32), C.c_char_p, C.c_uint32]
    lib.ref_free.argtypes = [C.c_void_p]
    v, i, nv, ni = C.POINTER(C.c_float)(), C.POINTER(C.c_uint32)(), C.c_uint32(), C.c_uint32()
    err = C.create_string_buffer(512)
    if lib.ref_load_model(path.encode(), C.byref(v), C.byref(nv), C.byref(i), C.byref(ni), err, 512):
        return None, err.value.decode()
    verts = np.ctypeslib.as_array(v, (nv.value * 8,)).copy() if nv.value else np.zeros(0, np.float32)
    idx = np.ctypeslib.as_array(i, (ni.value,)).copy() if ni.value else np.zeros(0, np.uint32)
    lib.ref_free(v)
    lib.ref_free(i)
    return (verts.reshape(-1, 8), idx), None


def _tool(tmp_path):
    exe = str(tmp_path / "obj_dump")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "tools", "obj_dump.cpp"), "-o", exe])
    return exe


def _product_load(path):
    if os.path.exists(path + ".err"):
        return None, open(path + ".err").read()
    raw = open(path + ".bin", "rb").read()
    nv, ni = np.frombuffer(raw, np.uint32, 2)
    return (np.frombuffer(raw, np.float32, nv * 8, 8).reshape(-1, 8), np.frombuffer(raw, np.uint32, ni, 8 + 32 * nv)), None


def _in_range(text):
    """True when every corner of the file resolves inside the arrays read so far (the reference indexes without a check — undefined
    behaviour — so only such files are compared; the facade must refuse the others)."""
    nv = nvn = nvt = 0
    for line in text.replace("\r\n", "\n").replace("\r", "\n").split("\n"):
        t = line.strip(" \t")
        if t[:2] in ("v ", "v\t"):
            nv += 1
        elif t[:3] in ("vn ", "vn\t"):
            nvn += 1
        elif t[:3] in ("vt ", "vt\t"):
            nvt += 1
        elif t[:2] in ("f ", "f\t"):
            for c in t[2:].split():
                parts = c.split("/")
                for s, n in ((parts[0], nv), (parts[1] if len(parts) > 1 else "", nvt), (parts[2] if len(parts) > 2 else "", nvn)):
                    if s == "":
                        continue
                    i = int(s)
                    z = i - 1 if i > 0 else (0 if i == 0 else n + i)
                    if z >= n:  # negative z = "absent" for both sides (RT/Scene.cpp:46,53,59 test >= 0)
                        return False
    return True


def test_obj_ingestion_matches_reference_loadmodel(ref_obj, tmp_path):
    exe = _tool(tmp_path)
    (tmp_path / "mats.mtl").write_text("# materials\nnewmtl red\nKd 1 0 0\n\nnewmtl blue\nKd 0 0 1\n")
    rng = np.random.default_rng(2024)
    paths, texts = [], []
    for k in range(400):
        text = _fuzz_obj(rng, with_mtl=k % 3 == 0)
        p = str(tmp_path / f"f{k}.obj")
        with open(p, "w", newline="") as f:
            f.write(text)
        paths.append(p)
        texts.append(text)
    # hand-written cases: the reference demo's kind of asset (a quad with normals and uvs), CRLF, a polygon, relative indices,
    # a shape dropped by the usemtl-then-g quirk of tinyobj 1.0.6
    fixed = {
        "quad": "v -1 0 -1\nv 1 0 -1\nv 1 0 1\nv -1 0 1\nvn 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nf 1/1/1 2/2/1 3/3/1 4/4/1\n",
        "crlf": "o a\r\nv 0 0 0\r\nv 1 0 0\r\nv 0 1 0\r\nf 1 2 3\r\ng b\r\nv 0 0 1\r\nf -1 -2 -3 -4\r\n",
        "quirk": "mtllib mats.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nusemtl red\nf 1 2 3\nusemtl blue\ng next\nf 2 3 4\n",
        "quirk2": "mtllib mats.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nf 1 2 3\nusemtl blue\nf 2 3 4\nusemtl nosuch\no next\nf 1 3 4\n",
        "dedup_zero": "v 0 0 0\nv -0 -0 -0\nv 0.0 -0.0 0\nv 1 1 1\nf 1 2 3\nf 2 3 4\nf 4 1 3\n",
        "precision": "v 0.1234567890123 1.00000001 123456.789\nv 1e-3 2.5E2 -3.25e+1\nv 16777217 0.30000001192092896 3.4e38\nf 1 2 3\n",
        "empty": "# nothing\n",
        "no_eol": "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3",
    }
    for name, text in fixed.items():
        p = str(tmp_path / f"{name}.obj")
        with open(p, "w", newline="") as f:
            f.write(text)
        paths.append(p)
        texts.append(text)
    paths.append(str(tmp_path / "does_not_exist.obj"))
    texts.append(None)
    cwd = os.getcwd()
    os.chdir(tmp_path)  # mtllib names are looked up relative to the working directory (tinyobj with no base directory)
    try:
        subprocess.check_call([exe] + paths, cwd=tmp_path)
        compared = refused = 0
        for p, text in zip(paths, texts):
            got, gerr = _product_load(p)
            if text is None:
                ref, rerr = _ref_load(ref_obj, p)
                assert ref is None and got is None and gerr == rerr  # "Cannot open file [...]", RT/Scene.cpp:38-41
                continue
            if not _in_range(text):
                # (a dangling corner inside a shape that tinyobj's quirk drops is never dereferenced: then the file loads)
                assert got is not None or "out of range" in gerr, p
                refused += got is None
                continue
            ref, rerr = _ref_load(ref_obj, p)
            assert ref is not None and got is not None, (p, rerr, gerr)
            assert got[0].shape == ref[0].shape and got[1].shape == ref[1].shape, (p, got[0].shape, ref[0].shape, got[1].shape, ref[1].shape)
            assert np.array_equal(got[0].view(np.uint32), ref[0].view(np.uint32)), p
            assert np.array_equal(got[1], ref[1]), p
            compared += 1
    finally:
        os.chdir(cwd)
    assert compared >= 250 and refused >= 5, (compared, refused)
    # the quirk case really drops a shape: 'quirk' keeps only the face after `g`
    got, _ = _product_load(str(tmp_path / "quirk.obj"))
    assert len(got[1]) == 3
