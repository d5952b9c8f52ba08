"""Structure of the BLAS the CUDA builder writes, node by node (brt_debug_get_blas): every primitive in exactly one leaf slot, every
quantised child box contains what is below it, empty slots unreachable, permuted masks consistent. Device-only builder code is what this
adds over tests/test_emu_parity.py::test_blas_structure: the warp-cooperative collapse of narrow levels, the thread-per-item collapse of
wide ones, the warp-cooperative treelet pass, the radix sort with both tile sizes. (Runs last: the file name sorts after the parity files.)"""
import numpy as np
import pytest

import parity_cases as pc

pytestmark = pytest.mark.gpu


@pytest.fixture()
def make(pkg):
    return lambda flags=0: pkg.Context(device=0, flags=flags)


def test_blas_structure_small(pkg, make):
    pc.blas_structure(pkg, make)


def test_blas_structure_131k_triangles(pkg, make):
    """A 256 x 256 heightfield: its last collapse level has more than 8192 items (thread per item), the upper ones fewer (warp per item)."""
    v, idx = pkg.scenes.heightfield(256)
    for flags in (0, pkg.CFG_NO_TREELET):
        a = make(flags)
        mesh = a.mesh_create(v, idx)
        a.instance_create(mesh, a.material_create((0.5, 0.5, 0.5)), pkg.scenes.xform())
        a.scene_build()
        nodes, tris = a.debug_get_blas(mesh)
        fill, depth = pc.check_blas_structure(nodes, tris, len(idx) // 3)
        assert fill > 6.5 and depth <= 12
        assert a.get_stats().bvh_nodes >= len(nodes)


def test_degenerate_extents(pkg, orc_mod, make):
    """Planar grid, 30000 : 1 strip, 200 coincident triangles (equal Morton codes) through the device's sort / hierarchy / collapse."""
    pc.degenerate_extents(pkg, orc_mod, make)
