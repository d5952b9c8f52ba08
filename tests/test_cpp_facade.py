"""The C++ host facade (include/bloon/bloon.hpp: Scene / Pipeline / Camera / MeshInstance with the reference's
names) compiles against the C ABI; on a GPU the reference's demo scene rendered through it — including
Scene::loadModel's OBJ path with the Y flip and vertex de-duplication — equals the oracle's frame."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "hardware-ray-tracer_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "rtapp_demo")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "rtapp_demo.cpp"),
                           "-L" + LIBDIR, "-lbrt", "-Wl,-rpath," + LIBDIR, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe])
    return exe


def test_facade_compiles_and_fails_loudly_without_gpu(pkg, tmp_path):
    pkg.load()
    exe = _build(tmp_path)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, str(tmp_path / "out")], capture_output=True, text=True)
        assert r.returncode != 0 and "failed to create the ray tracing device" in r.stderr  # std::runtime_error caught in main, as HRT/main.cpp:9-13


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["sync", "async"])
def test_rtapp_demo_matches_oracle(pkg, orc_mod, tmp_path, mode):
    """sync: Pipeline::traceRays per frame; async: two frames in flight (Pipeline::submitFrame / waitFrame)."""
    exe = _build(tmp_path)
    out = subprocess.run([exe, str(tmp_path / "out"), "2"] + (["async"] if mode == "async" else []), capture_output=True, text=True, check=True).stdout
    assert "rtapp_demo: 800x600" in out
    img = np.fromfile(tmp_path / "out.rgba32f", dtype=np.float32).reshape(600, 800, 4)
    scene = pkg.scenes.rtapp_demo()
    orc = orc_mod.Oracle(pkg)
    scene.upload(orc)
    u = scene.uniform(orc, 800, 600, frame=1, depth_max=2)  # the second frame used imageIndex 1
    ref = orc.render_frame(u, orc.opts(800, 600))
    assert np.array_equal(img.view(np.uint32), ref.view(np.uint32))
    assert (img[..., :3] > 0).mean() > 0.05


@pytest.mark.gpu
def test_rtapp_demo_post_path(pkg, orc_mod, tmp_path):
    """The facade's Camera::handleInputs, Extensions::Denoiser and rebuildRenderOutput(format, extent): the demo walks the camera for three
    frames, denoises each and renders a B8G8R8A8_UNORM frame; camera pose, denoised frame and present image equal the oracle's."""
    exe = _build(tmp_path)
    out = subprocess.run([exe, str(tmp_path / "out"), "3", "post"], capture_output=True, text=True, check=True).stdout
    assert "rtapp_demo post:" in out
    w, h = 800, 600
    scene = pkg.scenes.rtapp_demo()
    orc = orc_mod.Oracle(pkg)
    scene.upload(orc)
    host = pkg.Context(device=0)  # for the product's host-side camera maths (what the facade calls)
    pos, rot = np.array([0.0, 0.0, -2.0], np.float32), np.zeros(3, np.float32)
    dop = orc.denoise_opts(iterations=4, sigma_n_log2=5, sigma_z=0.05, sigma_l=4.0, clamp_gamma=0.0, max_history=32.0, flags=pkg.DENOISE_BILATERAL)
    for frame in range(3):
        pos, rot = orc.camera_handle_inputs(4 | 64, 1.0 / 60.0, pos, rot)  # BRT_KEY_MOVE_FORWARD | BRT_KEY_LOOK_RIGHT
        u = host.camera_uniform(pos, rot, 1.0471975512, w / h, 0.001, 100000.0, frame, 2)
        orc.render_frame(u, orc.opts(w, h, 1, pkg.GBUFFER))
        den = orc.denoise(u, dop, w, h)
    vals = [float(x) for x in out.split("camera")[1].split(",")[0].replace("rot", "").split()]
    assert np.allclose(vals, list(pos) + list(rot), atol=2e-6)
    got = np.fromfile(tmp_path / "out.denoised.rgba32f", dtype=np.float32).reshape(h, w, 4)
    assert np.array_equal(got.view(np.uint32), den.view(np.uint32))
    ref8 = orc.render_frame(u, orc.opts(w, h, 1, pkg.render_format(pkg.FORMAT_BGRA8_UNORM)))
    got8 = np.fromfile(tmp_path / "out.bgra8", dtype=np.uint8).reshape(h, w, 4)
    assert np.array_equal(got8, ref8)
