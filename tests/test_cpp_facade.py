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
