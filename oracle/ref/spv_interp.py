"""A small SPIR-V interpreter for the reference's shipped shader binary (TEST INFRASTRUCTURE, generation time only).

The reference loads `shaders/raytracing.slang.spv` at run time (`Graphics/RayTracing/RTPipeline.cpp:168`); that binary
is the one artefact under /root/reference that holds the device hot path in executable form (entry points `rgenMain`,
`rchitMain`, `rmissMain`, `rmissShadowMain`; slangc inlined every helper: BRDF, processLight, calculateColor, the
geometry fetch). This module executes those entry points on the CPU so that `tests/golden/make_spv_kat.py` can write
reference-held golden frames (`tests/golden/spv_kat.json`). It is never imported by the product, the tests read only
the committed JSON (the GPU box has no /root/reference).

Scope: exactly the opcodes the binary uses (~70 core opcodes + 8 GLSL.std.450 instructions, listed in `_run`); anything
else raises. Execution model:

* binary32 arithmetic is numpy.float32, one rounding per SPIR-V instruction, no contraction; `OpDot`, matrix products
  accumulate left to right. Sqrt / Normalize / Length / Pow / Log2 are evaluated in float64 and rounded once (SPIR-V leaves
  their last ulps to the implementation; the tests compare with a tolerance that covers it).
* Uniform / StorageBuffer / PhysicalStorageBuffer pointers address one flat byte memory and follow the module's own
  Offset / ArrayStride / MatrixStride / RowMajor / ColMajor decorations, so the host writes its structs byte for byte
  as the C++ application does (`RT/RTPipeline.cpp:44-47`, `RT/Scene.cpp:313-403`).
* `OpTraceRayKHR` calls back into the harness (the driver / RT cores have no source in the reference): the harness decides
  hit or miss, fills the built-ins and runs the closest-hit or miss entry point on the caller's payload.
"""
import copy
import math
import struct

import numpy as np

F32 = np.float32

# storage classes
SC_UNIFORM_CONSTANT, SC_INPUT, SC_UNIFORM, SC_FUNCTION, SC_STORAGE_BUFFER = 0, 1, 2, 7, 12
SC_RAY_PAYLOAD, SC_HIT_ATTRIBUTE, SC_INCOMING_PAYLOAD, SC_PHYSICAL = 5338, 5339, 5342, 5349
# decorations
DEC_BLOCK, DEC_ROW_MAJOR, DEC_COL_MAJOR, DEC_ARRAY_STRIDE, DEC_MATRIX_STRIDE, DEC_BUILTIN = 2, 4, 5, 6, 7, 11
DEC_BINDING, DEC_DESCRIPTOR_SET, DEC_OFFSET = 33, 34, 35

BUILTIN_NAMES = {5319: "LaunchIdKHR", 5320: "LaunchSizeKHR", 6: "InstanceId", 7: "PrimitiveId",
                 5327: "InstanceCustomIndexKHR", 5330: "ObjectToWorldKHR", 5331: "WorldToObjectKHR",
                 5322: "WorldRayDirectionKHR", 5321: "WorldRayOriginKHR", 5326: "RayTmaxKHR"}


def _str(words):
    b = b"".join(struct.pack("<I", x) for x in words)
    return b.split(b"\0")[0].decode()


class Memory:
    """Flat little-endian byte memory addressed by 64-bit 'device addresses'."""

    def __init__(self):
        self.blocks = []  # (base, bytearray)
        self.next = 0x10000

    def alloc(self, data: bytes, align=256):
        base = (self.next + align - 1) // align * align
        self.blocks.append((base, bytearray(data)))
        self.next = base + len(data) + 64
        return base

    def _find(self, addr, n):
        for base, buf in self.blocks:
            if base <= addr and addr + n <= base + len(buf):
                return buf, addr - base
        raise MemoryError(f"out-of-bounds access of {n} bytes at {addr:#x}")

    def read(self, addr, n):
        buf, off = self._find(addr, n)
        return bytes(buf[off:off + n])

    def write(self, addr, data):
        buf, off = self._find(addr, len(data))
        buf[off:off + len(data)] = data


class Module:
    def __init__(self, path):
        raw = open(path, "rb").read()
        w = struct.unpack("<%dI" % (len(raw) // 4), raw)
        if w[0] != 0x07230203:
            raise ValueError("not a SPIR-V module")
        self.version = w[1]
        self.types = {}      # id -> tuple describing the type
        self.consts = {}     # id -> value
        self.names = {}
        self.decor = {}      # id -> {decoration: operands}
        self.mdecor = {}     # (id, member) -> {decoration: operands}
        self.entry = {}      # name -> (function id, execution model)
        self.globals = {}    # id -> (pointer type id, storage class)
        self.functions = {}  # id -> list of (op, operands) with labels resolved
        self.ext = {}
        i = 5
        cur = None
        while i < len(w):
            op, n = w[i] & 0xFFFF, w[i] >> 16
            a = w[i + 1:i + n]
            i += n
            if cur is not None:
                if op == 56:  # FunctionEnd
                    cur = None
                else:
                    cur.append((op, a))
                continue
            if op == 5:
                self.names[a[0]] = _str(a[1:])
            elif op == 11:
                self.ext[a[0]] = _str(a[1:])
            elif op == 15:
                name = _str(a[2:])
                self.entry[name] = (a[1], a[0])
            elif op == 71:
                self.decor.setdefault(a[0], {})[a[1]] = a[2:]
            elif op == 72:
                self.mdecor.setdefault((a[0], a[1]), {})[a[2]] = a[3:]
            elif op == 19:
                self.types[a[0]] = ("void",)
            elif op == 20:
                self.types[a[0]] = ("bool",)
            elif op == 21:
                self.types[a[0]] = ("int", a[1], a[2])
            elif op == 22:
                self.types[a[0]] = ("float", a[1])
            elif op == 23:
                self.types[a[0]] = ("vec", a[1], a[2])
            elif op == 24:
                self.types[a[0]] = ("mat", a[1], a[2])
            elif op == 25:
                self.types[a[0]] = ("image",)
            elif op == 28:
                self.types[a[0]] = ("array", a[1], a[2])  # length is a constant id, resolved lazily
            elif op == 29:
                self.types[a[0]] = ("rtarray", a[1])
            elif op == 30:
                self.types[a[0]] = ("struct", tuple(a[1:]))
            elif op == 32:
                self.types[a[0]] = ("ptr", a[1], a[2])
            elif op == 33:
                self.types[a[0]] = ("fn",)
            elif op == 39:
                pass  # forward pointer; the OpTypePointer follows
            elif op == 5341:
                self.types[a[0]] = ("accel",)
            elif op == 41:
                self.consts[a[1]] = True
            elif op == 42:
                self.consts[a[1]] = False
            elif op == 43:
                self.consts[a[1]] = self._scalar_const(a[0], a[2:])
            elif op == 44:
                self.consts[a[1]] = [self.consts[x] for x in a[2:]]
            elif op == 46:  # ConstantNull
                self.consts[a[1]] = self.zero(a[0])
            elif op == 59:
                self.globals[a[1]] = (a[0], a[2])
            elif op == 54:
                cur = []
                self.functions[a[1]] = cur
            elif op in (3, 4, 6, 10, 14, 16, 17, 7, 8, 330, 331):
                pass
            else:
                raise NotImplementedError(f"module-level opcode {op}")

    def _scalar_const(self, tid, words):
        t = self.types[tid]
        if t[0] == "float":
            if t[1] != 32:
                raise NotImplementedError("only binary32 constants")
            return F32(struct.unpack("<f", struct.pack("<I", words[0]))[0])
        if t[0] == "int":
            v = words[0] | (words[1] << 32 if len(words) > 1 else 0)
            return self.wrap(v, t)
        raise NotImplementedError(t)

    @staticmethod
    def wrap(v, t):
        bits, signed = t[1], t[2]
        v &= (1 << bits) - 1
        if signed and v >> (bits - 1):
            v -= 1 << bits
        return v

    def zero(self, tid):
        t = self.types[tid]
        k = t[0]
        if k == "float":
            return F32(0)
        if k == "int":
            return 0
        if k == "bool":
            return False
        if k == "vec":
            return [self.zero(t[1]) for _ in range(t[2])]
        if k == "mat":
            return [self.zero(t[1]) for _ in range(t[2])]
        if k == "array":
            return [self.zero(t[1]) for _ in range(self.consts[t[2]])]
        if k == "struct":
            return [self.zero(m) for m in t[1]]
        if k == "ptr":
            return ("P", 0, t[2])
        if k in ("accel", "image"):
            return None
        raise NotImplementedError(t)

    # ---- byte layout of types behind Uniform / StorageBuffer / PhysicalStorageBuffer pointers ----
    def size_of(self, tid, mdec=None):
        t = self.types[tid]
        k = t[0]
        if k in ("float", "int"):
            return t[1] // 8
        if k == "vec":
            return self.size_of(t[1]) * t[2]
        if k == "ptr":
            return 8
        if k == "mat":
            stride = mdec[DEC_MATRIX_STRIDE][0]
            rows = self.types[t[1]][2]
            return stride * (rows if DEC_ROW_MAJOR in mdec else t[2])
        if k == "array":
            return self.decor[tid][DEC_ARRAY_STRIDE][0] * self.consts[t[2]]
        if k == "struct":
            end = 0
            for m, mt in enumerate(t[1]):
                md = self.mdecor.get((tid, m), {})
                end = max(end, md[DEC_OFFSET][0] + self.size_of(mt, md))
            return end
        raise NotImplementedError(t)

    def read(self, mem, addr, tid, mdec=None):
        t = self.types[tid]
        k = t[0]
        if k == "float":
            return F32(struct.unpack("<f", mem.read(addr, 4))[0])
        if k == "int":
            n = t[1] // 8
            return self.wrap(int.from_bytes(mem.read(addr, n), "little"), t)
        if k == "ptr":
            return ("P", int.from_bytes(mem.read(addr, 8), "little"), t[2])
        if k == "vec":
            s = self.size_of(t[1])
            return [self.read(mem, addr + s * c, t[1]) for c in range(t[2])]
        if k == "mat":
            stride = mdec[DEC_MATRIX_STRIDE][0]
            cols, rows = t[2], self.types[t[1]][2]
            st = self.types[t[1]][1]
            if DEC_ROW_MAJOR in mdec:  # row r stored contiguously at r*stride
                return [[self.read(mem, addr + r * stride + 4 * c, st) for r in range(rows)] for c in range(cols)]
            return [[self.read(mem, addr + c * stride + 4 * r, st) for r in range(rows)] for c in range(cols)]
        if k == "array":
            stride = self.decor[tid][DEC_ARRAY_STRIDE][0]
            return [self.read(mem, addr + stride * e, t[1]) for e in range(self.consts[t[2]])]
        if k == "struct":
            out = []
            for m, mt in enumerate(t[1]):
                md = self.mdecor.get((tid, m), {})
                out.append(self.read(mem, addr + md[DEC_OFFSET][0], mt, md))
            return out
        raise NotImplementedError(t)

    def disassemble(self):
        lines = []
        for name, (fid, model) in self.entry.items():
            lines.append(f"; entry {name} (model {model}) %{fid}")
            for op, a in self.functions[fid]:
                lines.append(f"  op{op} " + " ".join(str(x) for x in a))
        return "\n".join(lines)


class TraceCall:
    """What the interpreter hands to the harness at an OpTraceRayKHR."""

    def __init__(self, flags, mask, sbt_offset, sbt_stride, miss_index, origin, tmin, direction, tmax, payload_ref):
        self.flags, self.mask, self.sbt_offset, self.sbt_stride, self.miss_index = flags, mask, sbt_offset, sbt_stride, miss_index
        self.origin, self.tmin, self.direction, self.tmax = origin, tmin, direction, tmax
        self.payload_ref = payload_ref  # one-element list holding the caller's payload value (mutable)


class Interpreter:
    def __init__(self, module: Module, memory: Memory):
        self.m = module
        self.mem = memory
        self.bindings = {}   # (set, binding) -> device address (buffers) or any object
        self.image_writes = []
        self.trace = None    # callable(TraceCall)
        self.executed = 0

    # ------------------------------------------------------------------ pointers
    def _load(self, p):
        if p[0] == "L":
            v = p[1][0]
            for ix in p[2]:
                v = v[ix]
            return copy.deepcopy(v)
        _, addr, tid, mdec = p if len(p) == 4 else (*p, None)
        return self.m.read(self.mem, addr, tid, mdec)

    def _store(self, p, val):
        if p[0] != "L":
            raise NotImplementedError("store through a physical pointer (the shaders never do)")
        val = copy.deepcopy(val)
        if not p[2]:
            p[1][0] = val
            return
        v = p[1][0]
        for ix in p[2][:-1]:
            v = v[ix]
        v[p[2][-1]] = val

    def _chain(self, p, indices, ptr_elem=None):
        m = self.m
        if p[0] == "L":
            if ptr_elem is not None:
                raise NotImplementedError("OpPtrAccessChain on a logical pointer")
            return ("L", p[1], p[2] + tuple(int(i) for i in indices))
        addr, tid = p[1], p[2]
        mdec = p[3] if len(p) == 4 else None
        if ptr_elem is not None:  # element stride comes from the pointer TYPE's ArrayStride decoration
            addr += int(ptr_elem[0]) * ptr_elem[1]
        for ix in indices:
            ix = int(ix)
            t = m.types[tid]
            k = t[0]
            if k == "struct":
                md = m.mdecor.get((tid, ix), {})
                addr += md[DEC_OFFSET][0]
                tid, mdec = t[1][ix], md
            elif k == "array":
                addr += m.decor[tid][DEC_ARRAY_STRIDE][0] * ix
                tid = t[1]
            elif k == "vec":
                addr += m.size_of(t[1]) * ix
                tid, mdec = t[1], None
            elif k == "mat":
                raise NotImplementedError("access chain into a matrix in memory")
            else:
                raise NotImplementedError(t)
        return ("P", addr, tid, mdec)

    # ------------------------------------------------------------------ arithmetic helpers
    @staticmethod
    def _map(f, *xs):
        if isinstance(xs[0], list):
            return [Interpreter._map(f, *ys) for ys in zip(*xs)]
        return f(*xs)

    @staticmethod
    def _dot(a, b):
        acc = F32(a[0] * b[0])
        for x, y in zip(a[1:], b[1:]):
            acc = F32(acc + F32(x * y))
        return acc

    @staticmethod
    def _r(x):  # float64 -> binary32, once
        return F32(x)

    def _glsl(self, inst, xs):
        r, f64 = self._r, np.float64
        if inst == 31:   # Sqrt
            return self._map(lambda x: r(math.sqrt(f64(x))) if x >= 0 else F32(np.nan), xs[0])
        if inst == 30:   # Log2
            return self._map(lambda x: r(math.log2(f64(x))) if x > 0 else (F32(-np.inf) if x == 0 else F32(np.nan)), xs[0])
        if inst == 26:   # Pow
            return self._map(lambda x, y: r(math.pow(f64(x), f64(y))), xs[0], xs[1])
        if inst == 40:   # FMax
            return self._map(lambda x, y: y if x < y else x, xs[0], xs[1])
        if inst == 37:   # FMin
            return self._map(lambda x, y: y if y < x else x, xs[0], xs[1])
        if inst == 43:   # FClamp = min(max(x, lo), hi)
            return self._map(lambda x, lo, hi: min(max(x, lo), hi), xs[0], xs[1], xs[2])
        if inst == 46:   # FMix = x*(1-a) + y*a
            return self._map(lambda x, y, a: F32(F32(x * F32(F32(1) - a)) + F32(y * a)), xs[0], xs[1], xs[2])
        if inst == 66:   # Length
            return r(math.sqrt(sum(f64(x) * f64(x) for x in xs[0])))
        if inst == 69:   # Normalize
            n = math.sqrt(sum(f64(x) * f64(x) for x in xs[0]))
            return [r(f64(x) / n) if n > 0 else F32(np.nan) for x in xs[0]]
        raise NotImplementedError(f"GLSL.std.450 instruction {inst}")

    # ------------------------------------------------------------------ execution
    def run(self, entry, builtins=None, incoming_payload=None, hit_attribute=None):
        """Runs one invocation of an entry point. `builtins`: name -> value. `incoming_payload`: one-element list that
        is the caller's payload (read and written in place). Returns nothing; effects are image writes and the payload."""
        m = self.m
        fid, _ = m.entry[entry]
        body = m.functions[fid]
        labels = {a[0]: i for i, (op, a) in enumerate(body) if op == 248}
        val = dict(m.consts)
        builtins = builtins or {}
        own_payloads = {}
        for gid, (ptid, sc) in m.globals.items():
            pointee = m.types[ptid][2]
            dec = m.decor.get(gid, {})
            if sc == SC_INPUT:
                name = BUILTIN_NAMES[dec[DEC_BUILTIN][0]]
                val[gid] = ("L", [copy.deepcopy(builtins[name])] if name in builtins else [None], ())
            elif sc in (SC_UNIFORM, SC_STORAGE_BUFFER):
                key = (dec[DEC_DESCRIPTOR_SET][0], dec[DEC_BINDING][0])
                val[gid] = ("P", self.bindings[key], pointee, None)
            elif sc == SC_UNIFORM_CONSTANT:
                val[gid] = ("L", [(dec[DEC_DESCRIPTOR_SET][0], dec[DEC_BINDING][0])], ())
            elif sc == SC_RAY_PAYLOAD:
                own_payloads[gid] = [m.zero(pointee)]
                val[gid] = ("L", own_payloads[gid], ())
            elif sc == SC_INCOMING_PAYLOAD:
                val[gid] = ("L", incoming_payload if incoming_payload is not None else [m.zero(pointee)], ())
            elif sc == SC_HIT_ATTRIBUTE:
                val[gid] = ("L", [copy.deepcopy(hit_attribute) if hit_attribute is not None else m.zero(pointee)], ())
            else:
                raise NotImplementedError(f"global storage class {sc}")

        T = m.types
        W = Module.wrap
        pc = 0
        with np.errstate(all="ignore"):
            while True:
                op, a = body[pc]
                pc += 1
                self.executed += 1
                if op == 248 or op == 246 or op == 247 or op == 55:  # Label, LoopMerge, SelectionMerge, FunctionParameter
                    continue
                if op == 59:    # Variable (Function storage)
                    init = val[a[3]] if len(a) > 3 else m.zero(T[a[0]][2])
                    val[a[1]] = ("L", [copy.deepcopy(init)], ())
                elif op == 61:  # Load
                    val[a[1]] = self._load(val[a[2]])
                elif op == 62:  # Store
                    self._store(val[a[0]], val[a[1]])
                elif op in (65, 66):  # (InBounds)AccessChain
                    val[a[1]] = self._chain(val[a[2]], [val[x] for x in a[3:]])
                elif op == 67:  # PtrAccessChain: first index steps over elements of the pointee type
                    stride = m.decor[a[0]][DEC_ARRAY_STRIDE][0]
                    val[a[1]] = self._chain(val[a[2]], [val[x] for x in a[4:]], ptr_elem=(val[a[3]], stride))
                elif op == 79:  # VectorShuffle
                    both = list(val[a[2]]) + list(val[a[3]])
                    val[a[1]] = [both[c] for c in a[4:]]
                elif op == 80:  # CompositeConstruct (vectors may be built from scalars and vectors)
                    t = T[a[0]]
                    parts = [val[x] for x in a[2:]]
                    if t[0] == "vec":
                        flat = []
                        for p in parts:
                            flat.extend(p if isinstance(p, list) else [p])
                        val[a[1]] = flat
                    else:
                        val[a[1]] = copy.deepcopy(parts)
                elif op == 81:  # CompositeExtract
                    v = val[a[2]]
                    for ix in a[3:]:
                        v = v[ix]
                    val[a[1]] = copy.deepcopy(v)
                elif op == 82:  # CompositeInsert
                    comp = copy.deepcopy(val[a[3]])
                    v = comp
                    for ix in a[4:-1]:
                        v = v[ix]
                    v[a[-1]] = copy.deepcopy(val[a[2]])
                    val[a[1]] = comp
                elif op == 12:  # ExtInst
                    if m.ext[a[2]] != "GLSL.std.450":
                        raise NotImplementedError(m.ext[a[2]])
                    val[a[1]] = self._glsl(a[3], [val[x] for x in a[4:]])
                elif op == 112:  # ConvertUToF
                    t = None
                    val[a[1]] = self._map(lambda x: F32(x), val[a[2]])
                elif op == 111:  # ConvertSToF
                    val[a[1]] = self._map(lambda x: F32(x), val[a[2]])
                elif op in (109, 110):  # ConvertFToU / ConvertFToS (round toward zero)
                    rt = T[a[0]]
                    st = rt if rt[0] == "int" else T[rt[1]]
                    val[a[1]] = self._map(lambda x: W(int(x), st), val[a[2]])
                elif op in (113, 114):  # UConvert / SConvert
                    rt = T[a[0]]
                    st = rt if rt[0] == "int" else T[rt[1]]
                    if op == 113:
                        src_t = None
                        val[a[1]] = self._map(lambda x: W(x & ((1 << 64) - 1) if x >= 0 else x + (1 << 64), st), val[a[2]])
                    else:
                        val[a[1]] = self._map(lambda x: W(x, st), val[a[2]])
                elif op == 120:  # ConvertUToPtr
                    val[a[1]] = ("P", int(val[a[2]]) & ((1 << 64) - 1), T[a[0]][2], None)
                elif op == 117:  # ConvertPtrToU
                    val[a[1]] = val[a[2]][1]
                elif op == 124:  # Bitcast
                    val[a[1]] = self._bitcast(a[0], val[a[2]])
                elif op == 127:
                    val[a[1]] = self._map(lambda x: F32(-x), val[a[2]])
                elif op in (128, 130, 132):  # IAdd ISub IMul
                    rt = T[a[0]]
                    st = rt if rt[0] == "int" else T[rt[1]]
                    f = {128: lambda x, y: W(x + y, st), 130: lambda x, y: W(x - y, st), 132: lambda x, y: W(x * y, st)}[op]
                    val[a[1]] = self._map(f, val[a[2]], val[a[3]])
                elif op in (129, 131, 133, 136):  # FAdd FSub FMul FDiv
                    f = {129: lambda x, y: F32(x + y), 131: lambda x, y: F32(x - y), 133: lambda x, y: F32(x * y),
                         136: lambda x, y: F32(np.divide(x, y, dtype=F32))}[op]
                    val[a[1]] = self._map(f, val[a[2]], val[a[3]])
                elif op == 142:  # VectorTimesScalar
                    s = val[a[3]]
                    val[a[1]] = [F32(x * s) for x in val[a[2]]]
                elif op == 143:  # MatrixTimesScalar
                    s = val[a[3]]
                    val[a[1]] = [[F32(x * s) for x in col] for col in val[a[2]]]
                elif op == 144:  # VectorTimesMatrix: result[c] = dot(v, column c)
                    v, mat = val[a[2]], val[a[3]]
                    val[a[1]] = [self._dot(v, col) for col in mat]
                elif op == 145:  # MatrixTimesVector: result = sum_c column c * v[c]
                    mat, v = val[a[2]], val[a[3]]
                    rows = len(mat[0])
                    val[a[1]] = [self._dot([mat[c][r] for c in range(len(mat))], v) for r in range(rows)]
                elif op == 148:
                    val[a[1]] = self._dot(val[a[2]], val[a[3]])
                elif op == 169:  # Select
                    c, x, y = val[a[2]], val[a[3]], val[a[4]]
                    if isinstance(c, list):
                        val[a[1]] = [xx if cc else yy for cc, xx, yy in zip(c, x, y)]
                    else:
                        val[a[1]] = copy.deepcopy(x if c else y)
                elif op in (170, 171):
                    f = (lambda x, y: x == y) if op == 170 else (lambda x, y: x != y)
                    val[a[1]] = self._map(f, val[a[2]], val[a[3]])
                elif op in (172, 174, 176, 178):  # unsigned comparisons
                    def u(x):
                        return x if x >= 0 else x + (1 << 64)
                    f = {172: lambda x, y: u(x) > u(y), 174: lambda x, y: u(x) >= u(y),
                         176: lambda x, y: u(x) < u(y), 178: lambda x, y: u(x) <= u(y)}[op]
                    val[a[1]] = self._map(f, val[a[2]], val[a[3]])
                elif op in (173, 175, 177, 179):
                    f = {173: lambda x, y: x > y, 175: lambda x, y: x >= y, 177: lambda x, y: x < y, 179: lambda x, y: x <= y}[op]
                    val[a[1]] = self._map(f, val[a[2]], val[a[3]])
                elif op in (180, 182, 184, 186, 188, 190):  # ordered float comparisons (False on NaN)
                    f = {180: lambda x, y: bool(x == y), 182: lambda x, y: bool(x < y or x > y), 184: lambda x, y: bool(x < y),
                         186: lambda x, y: bool(x > y), 188: lambda x, y: bool(x <= y), 190: lambda x, y: bool(x >= y)}[op]
                    val[a[1]] = self._map(f, val[a[2]], val[a[3]])
                elif op in (166, 167):
                    f = (lambda x, y: x or y) if op == 166 else (lambda x, y: x and y)
                    val[a[1]] = self._map(f, val[a[2]], val[a[3]])
                elif op == 168:
                    val[a[1]] = self._map(lambda x: not x, val[a[2]])
                elif op == 249:  # Branch
                    pc = labels[a[0]]
                elif op == 250:  # BranchConditional
                    pc = labels[a[1] if val[a[0]] else a[2]]
                elif op == 251:  # Switch
                    sel = val[a[0]]
                    target = a[1]
                    for k in range(2, len(a), 2):
                        if a[k] == sel:
                            target = a[k + 1]
                    pc = labels[target]
                elif op == 253:  # Return
                    return
                elif op == 99:   # ImageWrite
                    self.image_writes.append((tuple(int(x) for x in val[a[1]]), [float(x) for x in val[a[2]]]))
                elif op == 4445:  # TraceRayKHR
                    pay = val[a[10]]
                    if pay[0] != "L" or pay[2]:
                        raise NotImplementedError("payload must be a whole variable")
                    self.trace(TraceCall(val[a[1]], val[a[2]], val[a[3]], val[a[4]], val[a[5]],
                                         [F32(x) for x in val[a[6]]], F32(val[a[7]]), [F32(x) for x in val[a[8]]], F32(val[a[9]]), pay[1]))
                else:
                    raise NotImplementedError(f"opcode {op} in {entry}")

    def _bitcast(self, tid, v):
        t = self.m.types[tid]
        if isinstance(v, tuple):  # pointer -> pointer / integer
            if t[0] == "ptr":
                return ("P", v[1], t[2], None)
            if t[0] == "int":
                return Module.wrap(v[1], t)
            raise NotImplementedError(t)
        if t[0] == "ptr":
            return ("P", int(v) & ((1 << 64) - 1), t[2], None)
        if isinstance(v, list):
            return [self._bitcast(t[1], x) for x in v]
        if t[0] == "float":
            return F32(struct.unpack("<f", struct.pack("<I", v & 0xFFFFFFFF))[0])
        if t[0] == "int":
            if isinstance(v, (np.floating, float)):
                return Module.wrap(struct.unpack("<I", struct.pack("<f", v))[0], t)
            return Module.wrap(v, t)
        raise NotImplementedError(t)


if __name__ == "__main__":
    import sys
    print(Module(sys.argv[1]).disassemble())
