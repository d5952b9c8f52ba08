"""Generates tests/golden/math_kat.json: bit patterns of the oracle's deterministic elementary functions (det_log2, det_exp2,
det_sincos — explicit polynomials with explicit fma), of the sRGB transfer function and of the UNORM8 conversion at
fixed arguments. They pin the arithmetic contract across compilers and machines: both the oracle and (through the parity tests) the
CUDA kernels must keep producing exactly these bits. An independent float64 evaluation in tests/test_oracle_kat.py bounds their error.
Run:  python tests/golden/make_math_kat.py
"""
import ctypes as C
import json
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402


def bits(f):
    return "%08x" % struct.unpack("<I", struct.pack("<f", f))[0]


def main():
    lib = ob.load()
    for name in ("orc_kat_exp2", "orc_kat_srgb", "orc_kat_log2"):
        getattr(lib, name).restype, getattr(lib, name).argtypes = C.c_float, [C.c_float]
    lib.orc_kat_unorm8.restype, lib.orc_kat_unorm8.argtypes = C.c_uint32, [C.c_float]
    out = {"exp2": [], "log2": [], "srgb": [], "sincos": [], "unorm8": []}
    for x in (-87.3, -10.25, -1.5, -0.49, 0.0, 0.3, 0.5, 1.0, 2.75, 19.999, 63.1):
        out["exp2"].append({"x": bits(x), "y": bits(lib.orc_kat_exp2(x))})
    for x in (1e-6, 0.001, 0.3, 0.70710678, 1.0, 1.41421354, 1.5, 2.0, 10.0, 12345.678):
        out["log2"].append({"x": bits(x), "y": bits(lib.orc_kat_log2(x))})
    for x in (0.0, 0.001, 0.0031308, 0.004, 0.05, 0.18, 0.5, 0.9, 1.0):
        out["srgb"].append({"x": bits(x), "y": bits(lib.orc_kat_srgb(x))})
    for x in (0.0, 0.1, 0.785398, 1.5707963, 3.0, 3.1415927, 4.5, 6.2831853):
        s, c = C.c_float(), C.c_float()
        lib.orc_kat_sincos(x, C.byref(s), C.byref(c))
        out["sincos"].append({"x": bits(x), "s": bits(s.value), "c": bits(c.value)})
    for x in (-1.0, 0.0, 0.001, 0.00196, 0.00588, 0.0098, 0.25, 0.5, 0.7, 0.998, 1.0, 2.0):
        out["unorm8"].append({"x": bits(x), "q": int(lib.orc_kat_unorm8(x))})
    json.dump(out, open(os.path.join(HERE, "math_kat.json"), "w"), indent=1)
    print("wrote math_kat.json:", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
