"""The C-ABI library loads and exports every symbol include/brt.h declares; PODs have the reference's byte
layouts; without a GPU the library fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol(pkg):
    lib = pkg.load()
    header = open(os.path.join(ROOT, "include", "brt.h")).read()
    declared = sorted(set(re.findall(r"\b(brt_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(pkg.binding.BRT_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_flag_constants_match_header(pkg):
    """BRT_CFG_* / BRT_RENDER_* values of the ctypes view against the #defines of include/brt.h."""
    header = open(os.path.join(ROOT, "include", "brt.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+BRT_(CFG_[A-Z_]+|RENDER_[A-Z_]+)\s+(\d+)u\b", header)}
    B = pkg.binding
    cfg = {k: v for k, v in defs.items() if k.startswith("CFG_")}
    assert set(cfg) >= {"CFG_COUNTERS", "CFG_NO_TREELET", "CFG_TREELET_ON_REBUILD", "CFG_NO_OVERLAP", "CFG_NO_GRAPH", "CFG_GREEDY_COLLAPSE"}
    for name, value in cfg.items():
        assert getattr(B, name) == value, name
    for name in ("BOUNCE_REFLECT", "BOUNCE_REFRACT", "BOUNCE_DIFFUSE", "JITTER", "SKY", "GBUFFER", "LIGHT_BVH", "DENOISE", "FAST_SHADING"):
        assert getattr(B, name) == defs["RENDER_" + name], name
    values = sorted(cfg.values())
    assert values == [1 << i for i in range(len(values))]  # distinct bits, none skipped


def test_pod_layouts(pkg):
    B = pkg.binding
    assert C.sizeof(B.Vertex) == 32 and B.Vertex.normal.offset == 12 and B.Vertex.uv.offset == 24
    assert C.sizeof(B.Material) == 52 and B.Material.metallic.offset == 16 and B.Material.clearCoatGloss.offset == 48
    assert C.sizeof(B.Light) == 32 and B.Light.color.offset == 12 and B.Light.intensity.offset == 24 and B.Light.type.offset == 28
    assert C.sizeof(B.Uniform) == 140 and B.Uniform.projInverse.offset == 64 and B.Uniform.frame.offset == 128
    assert B.Uniform.depthMax.offset == 132 and B.Uniform.lightThreshold.offset == 136
    assert C.sizeof(B.Sky) == 88
    # RT/Scene.h:84-88 (24 B, materialId at 16) and RT/Scene.h:106-121 (10 x u64)
    assert C.sizeof(B.InstanceInfo) == 24 and B.InstanceInfo.indexAddress.offset == 8 and B.InstanceInfo.materialId.offset == 16
    assert C.sizeof(B.SceneBufferInfo) == 80 and B.SceneBufferInfo.lCount.offset == 32 and B.SceneBufferInfo.sBuf.offset == 48


def test_no_cpu_fallback(pkg):
    """Without a CUDA device brt_create must fail (BRT_ERR_CUDA); with one it must succeed."""
    import torch
    if torch.cuda.is_available():
        ctx = pkg.Context(device=0)
        ctx.close()
    else:
        with pytest.raises(pkg.BrtError, match="status 2"):
            pkg.Context(device=0)


def test_product_does_not_reference_oracle():
    """Nothing under the package or include/ may import, link or name the oracle or the emulation shim."""
    pkg_dir = os.path.join(ROOT, "hardware-ray-tracer_b200")
    for base, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(base, f)).read()
                assert "liboracle" not in text and not re.search(r"\borc_", text), os.path.join(base, f)
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), os.path.join(base, f)
                assert "libbrt_emu" not in text, os.path.join(base, f)


def test_tile_buffer_bytes(pkg):
    lib = pkg.load()
    lib.brt_tile_buffer_bytes.restype = C.c_size_t
    lib.brt_tile_buffer_bytes.argtypes = [C.c_uint32] * 3
    assert lib.brt_tile_buffer_bytes(1920, 1080, 1) == 60 * 34 * 1024 * 16
    assert lib.brt_tile_buffer_bytes(1920, 1080, 8) == ((60 * 34 + 7) // 8) * 1024 * 16
    assert lib.brt_tile_buffer_bytes(31, 1, 4) == 1024 * 16


def test_package_asks_for_more_hardware_work_queues():
    """Importing the package sets CUDA_DEVICE_MAX_CONNECTIONS (unless the user already did): frames in flight + the fused exchange use more
    streams than CUDA's default 8 hardware queues (profiles/r2_configs.md)."""
    import importlib
    import os
    importlib.import_module("hardware-ray-tracer_b200")
    assert int(os.environ["CUDA_DEVICE_MAX_CONNECTIONS"]) >= 8
