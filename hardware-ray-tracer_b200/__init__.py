"""B200-native hot path of the Bloon RT Engine: Python view of the C ABI (include/brt.h).

The product is lib/libbrt.so (hand-written sm_100a CUDA behind a C ABI). This package only loads
it with ctypes; there is no CPU fallback — `load()` raises when the library is missing and
`brt_create` fails when no CUDA device is usable.
"""
import ctypes
import os

# A frame in flight uses two CUDA streams, the fused multi-GPU exchange adds a stream whose head is a kernel that WAITS (device-side
# completion flags). With the default of 8 hardware work queues, streams share queues and everything queued behind such a wait stalls
# with it: 0.16 ms per C5 frame on the receiving rank (profiles/r2_configs.md). Only honoured if CUDA is not initialised yet in this process.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _binding as binding
from ._binding import *  # noqa: F401,F403
from . import scenes  # noqa: F401


def TiledFrame(*args, **kwargs):
    """Tile-parallel multi-GPU frame (multigpu.TiledFrame); torch is only imported when this is used."""
    import importlib
    return importlib.import_module(__name__ + ".multigpu").TiledFrame(*args, **kwargs)

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libbrt.so")

_lib = None


def load():
    """dlopen lib/libbrt.so (built by __graft_entry__.build() / csrc/Makefile). Fails loudly."""
    global _lib
    if _lib is None:
        path = os.environ.get("BRT_LIB", LIB_PATH)  # developer override: an alternative build of the same library
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        _lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    return _lib


def Context(device=0, tile_rank=0, tile_world=1, flags=0):
    """A `brt_context` on CUDA device `device`."""
    return binding.SceneApi(load(), "brt_", device, tile_rank, tile_world, flags)
