"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes loader for oracle/liboracle.so.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The struct definitions come from the product package's _binding module (they are the include/brt.h
PODs); nothing in the product imports this file.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
        C = ctypes
        _lib.orc_kat_hash.restype, _lib.orc_kat_hash.argtypes = C.c_uint32, [C.c_uint32] * 3
        _lib.orc_kat_pcg.restype, _lib.orc_kat_pcg.argtypes = C.c_uint32, [C.POINTER(C.c_uint32)]
        _lib.orc_kat_rand.restype, _lib.orc_kat_rand.argtypes = C.c_float, [C.POINTER(C.c_uint32)]
        _lib.orc_kat_log2.restype, _lib.orc_kat_log2.argtypes = C.c_float, [C.c_float]
        _lib.orc_kat_sincos.restype, _lib.orc_kat_sincos.argtypes = None, [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        f3 = C.POINTER(C.c_float)
        _lib.orc_kat_brdf.restype, _lib.orc_kat_brdf.argtypes = None, [C.c_void_p, f3, f3, f3, f3]
        _lib.orc_kat_sample_vndf.restype, _lib.orc_kat_sample_vndf.argtypes = None, [C.c_void_p, f3, f3, C.c_float, C.c_float, f3, f3]
        _lib.orc_kat_sample_cosine.restype, _lib.orc_kat_sample_cosine.argtypes = None, [C.c_float, C.c_float, f3, f3]
        _lib.orc_kat_intersect_tri.restype = C.c_int
        _lib.orc_kat_intersect_tri.argtypes = [f3, f3, C.c_float, C.c_float, f3, f3, f3, f3]
    return _lib


def Oracle(pkg, brute_force=False, tile_rank=0, tile_world=1, threads=0):
    """An oracle context with the same Python surface as pkg.Context()."""
    flags = 0x80000000 if brute_force else 0
    C = ctypes
    extra = {"set_threads": (C.c_int, [C.c_void_p, C.c_uint32]), "get_threads": (C.c_uint32, [C.c_void_p])}
    api = pkg.binding.SceneApi(load(), "orc_", 0, tile_rank, tile_world, flags, extra_signatures=extra)
    if threads:
        api._f("set_threads")(api.ctx, threads)
    return api
