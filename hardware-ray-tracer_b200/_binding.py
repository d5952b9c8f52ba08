"""ctypes view of include/brt.h (the C ABI of lib/libbrt.so, the CUDA product).

`SceneApi` takes the library handle and the symbol prefix as arguments, so test code can drive any
library that exports the same ABI under another prefix through the same Python surface; this module
itself only knows include/brt.h.
"""
import ctypes as C

import numpy as np

u32, i32, u64, f32 = C.c_uint32, C.c_int32, C.c_uint64, C.c_float


class Vertex(C.Structure):  # RT/Scene.h:28-31
    _fields_ = [("pos", f32 * 3), ("normal", f32 * 3), ("uv", f32 * 2)]


class Material(C.Structure):  # RT/Scene.h:50-62
    _fields_ = [("color", f32 * 3)] + [(n, f32) for n in (
        "subsurface", "metallic", "roughness", "specular", "specularTint", "anisotropic", "sheen",
        "sheenTint", "clearCoat", "clearCoatGloss")]


class Light(C.Structure):  # RT/Scene.h:70-75
    _fields_ = [("pos", f32 * 3), ("color", f32 * 3), ("intensity", f32), ("type", C.c_uint8), ("pad_", C.c_uint8 * 3)]


class Uniform(C.Structure):  # RT/RTPipeline.h:24-30
    _fields_ = [("viewInverse", f32 * 16), ("projInverse", f32 * 16), ("frame", u32), ("depthMax", u32), ("lightThreshold", f32)]


class Sky(C.Structure):  # RT/Scene.h:90-104
    _fields_ = [(n, f32 * 3) for n in ("skyColor", "horizonColor", "groundColor", "sunDirection", "upDirection")] + [
        (n, f32) for n in ("brightness", "horizonSize", "angularSize", "glowIntensity", "glowSharpness", "glowSize", "lightRadiance")]


class Config(C.Structure):
    _fields_ = [("struct_size", u32), ("device", i32), ("tile_rank", u32), ("tile_world", u32), ("flags", u32)]


class LightBvhNode(C.Structure):  # RT/Scene.h:123-130
    _fields_ = [("bBoxMin", f32 * 3), ("bBoxMax", f32 * 3), ("totalFlux", f32), ("coneAxis", f32 * 3), ("coneAngle", f32), ("childIndex", i32)]


class DenoiseOpts(C.Structure):
    _fields_ = [("struct_size", u32), ("flags", u32), ("iterations", u32), ("sigma_n_log2", u32), ("sigma_z", f32), ("sigma_l", f32),
                ("clamp_gamma", f32), ("max_history", f32)]


class RenderOpts(C.Structure):
    _fields_ = [(n, u32) for n in ("width", "height", "spp", "flags", "crop_x0", "crop_y0", "crop_w", "crop_h")]


class Stats(C.Structure):
    _fields_ = [(n, u64) for n in (
        "rays_closest", "rays_occlusion", "nodes_visited_closest", "prims_tested_closest", "spheres_tested_closest",
        "nodes_visited_occlusion", "prims_tested_occlusion", "spheres_tested_occlusion")] + [
        (n, f32) for n in ("ms_raygen", "ms_trace_closest", "ms_shade", "ms_trace_occlusion", "ms_accumulate", "ms_resolve", "ms_total")] + [
        (n, u32) for n in ("launches_trace_closest", "launches_trace_occlusion", "launches_total")] + [
        (n, f32) for n in ("ms_blas_build", "ms_tlas_build", "ms_cull")] + [("blas_built", u32)] + [
        (n, u64) for n in ("total_triangles", "bvh_nodes", "bvh_bytes")] + [
        ("sah_cost", f32), ("sah_cost_lbvh", f32), ("instances_visible", u32), ("instances_total", u32),
        ("ms_denoise", f32), ("launches_denoise", u32)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


assert C.sizeof(Vertex) == 32 and C.sizeof(Material) == 52 and C.sizeof(Light) == 32
assert C.sizeof(Uniform) == 140 and C.sizeof(Sky) == 88 and C.sizeof(LightBvhNode) == 48

CFG_COUNTERS = 1
CFG_NO_TREELET = 2
CFG_TREELET_ON_REBUILD = 4
CFG_NO_OVERLAP = 8
CFG_NO_GRAPH = 16
CFG_GREEDY_COLLAPSE = 32
CFG_HIT_SORT = 64
CFG_TREELET_PASSES_3 = 128
BOUNCE_REFLECT, BOUNCE_REFRACT, BOUNCE_DIFFUSE, JITTER, SKY, GBUFFER, LIGHT_BVH, DENOISE = 1, 2, 4, 8, 16, 32, 64, 128
FAST_SHADING = 2048
AOV_POSITION, AOV_NORMAL = 3, 4
DENOISE_RESET, DENOISE_BILATERAL = 1, 2
FORMAT_RGBA32F, FORMAT_RGBA8_UNORM, FORMAT_BGRA8_UNORM, FORMAT_RGBA8_SRGB, FORMAT_BGRA8_SRGB = 0, 1, 2, 3, 4
FORMAT_SHIFT, FORMAT_MASK = 8, 0x700


def render_format(fmt):
    """flag bits selecting the output format of the render entry points (BRT_RENDER_FORMAT)"""
    return fmt << FORMAT_SHIFT
AOV_PRIM_ID, AOV_INST_ID, AOV_HIT_T = 0, 1, 2
AOV_MISS = 0xFFFFFFFF

# every symbol include/brt.h declares (tests/test_abi.py checks the .so exports each of them)
BRT_SYMBOLS = [
    "brt_create", "brt_destroy", "brt_last_error", "brt_set_stream", "brt_mesh_create", "brt_mesh_update_vertices",
    "brt_sphere_create", "brt_material_create", "brt_material_set_transmission", "brt_light_create", "brt_sky_set",
    "brt_instance_create", "brt_instance_set_transform", "brt_instance_set_material", "brt_instance_destroy",
    "brt_scene_build", "brt_smart_cull", "brt_get_visibility", "brt_render_frame", "brt_render_frame_tiles",
    "brt_tile_buffer_bytes", "brt_untile", "brt_device_image", "brt_get_aov", "brt_get_stats", "brt_trace_rays",
    "brt_camera_uniform", "brt_debug_sort_pairs", "brt_gather_image_export", "brt_gather_image_open",
    "brt_render_frame_peers", "brt_gather_image", "brt_render_frame_async", "brt_frame_wait", "brt_frame_stream", "brt_camera_handle_inputs", "brt_denoise", "brt_denoised_image", "brt_denoise_configure", "brt_render_frame_peers_async", "brt_get_light_bvh",
    "brt_debug_get_blas", "brt_gather_configure", "brt_gather_wait", "brt_gather_release", "brt_gather_copy_to_host", "brt_gather_timed_out", "brt_debug_l2_read_gbs", "brt_get_scene_info_buffer", "brt_get_tlas",
]


class InstanceInfo(C.Structure):  # RT/Scene.h:84-88
    _fields_ = [("vertexAddress", C.c_uint64), ("indexAddress", C.c_uint64), ("materialId", C.c_uint32), ("pad", C.c_uint32)]


class SceneBufferInfo(C.Structure):  # RT/Scene.h:106-121
    _fields_ = [(n, C.c_uint64) for n in ("mBuf", "mStride", "lBuf", "lStride", "lCount", "vStride", "sBuf", "sStride", "skyBuf", "skyStride")]


class AccelInfo(C.Structure):  # RT/Scene.h:77-82 + counts
    _fields_ = [("handle", C.c_uint64), ("buffer", C.c_uint64), ("memory", C.c_uint64), ("address", C.c_uint64), ("n_nodes", C.c_uint32), ("n_instances", C.c_uint32)]


class BrtError(RuntimeError):
    """Non-zero status from the C ABI (the reference throws std::runtime_error, Graphics/Definitions.h:5)."""


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


class SceneApi:
    """One `brt_context` of a library exporting the include/brt.h ABI under `prefix`."""

    def __init__(self, lib, prefix="brt_", device=0, tile_rank=0, tile_world=1, flags=0, extra_signatures=None):
        self.lib, self.prefix = lib, prefix
        self._extra = extra_signatures or {}
        self._declare()
        cfg = Config(C.sizeof(Config), device, tile_rank, tile_world, flags)
        self.ctx = C.c_void_p()
        rc = self._f("create")(C.byref(cfg), C.byref(self.ctx))
        if rc != 0:
            msg = self._f("last_error")(None)
            raise BrtError(f"{prefix}create failed (status {rc}): {msg.decode() if msg else ''}")

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _declare(self):
        vp, P = C.c_void_p, C.POINTER
        sig = {
            "create": (C.c_int, [P(Config), P(vp)]),
            "destroy": (None, [vp]),
            "last_error": (C.c_char_p, [vp]),
            "mesh_create": (C.c_int, [vp, P(Vertex), u32, P(u32), u32, P(u32)]),
            "mesh_update_vertices": (C.c_int, [vp, u32, P(Vertex), u32]),
            "sphere_create": (C.c_int, [vp, P(f32), f32, P(u32)]),
            "material_create": (C.c_int, [vp, P(Material), P(u32)]),
            "material_set_transmission": (C.c_int, [vp, u32, f32, f32]),
            "light_create": (C.c_int, [vp, P(Light), P(u32)]),
            "sky_set": (C.c_int, [vp, P(Sky)]),
            "instance_create": (C.c_int, [vp, u32, u32, P(f32), P(u32)]),
            "instance_set_transform": (C.c_int, [vp, u32, P(f32)]),
            "instance_set_material": (C.c_int, [vp, u32, u32]),
            "instance_destroy": (C.c_int, [vp, u32]),
            "scene_build": (C.c_int, [vp]),
            "smart_cull": (C.c_int, [vp, P(Uniform), u32, u32, f32, f32, P(u32)]),
            "get_visibility": (C.c_int, [vp, P(C.c_uint8), u32]),
            "render_frame": (C.c_int, [vp, P(Uniform), P(RenderOpts), vp]),
            "get_aov": (C.c_int, [vp, C.c_int, vp]),
            "get_stats": (C.c_int, [vp, P(Stats)]),
            "trace_rays": (C.c_int, [vp, P(f32), u32, C.c_int, P(u32)]),
            "debug_sort_pairs": (C.c_int, [vp, P(u32), P(u32), u32, C.c_int]),
            "camera_uniform": (None, [P(f32), P(f32), f32, f32, f32, f32, u32, u32, P(Uniform)]),
            "camera_handle_inputs": (None, [u32, f32, P(f32), P(f32)]),
            "denoise": (C.c_int, [vp, P(Uniform), P(DenoiseOpts), vp]),
            "get_light_bvh": (C.c_int, [vp, P(LightBvhNode), u32, P(u32)]),
        }
        device_side = {  # entry points that take device pointers / streams, or that only the product has
            "debug_get_blas": (C.c_int, [vp, u32, vp, u32, vp, u32, P(u32), P(u32)]),
            "set_stream": (C.c_int, [vp, vp]),
            "render_frame_tiles": (C.c_int, [vp, P(Uniform), P(RenderOpts), vp]),
            "tile_buffer_bytes": (C.c_size_t, [u32, u32, u32]),
            "untile": (C.c_int, [vp, vp, u32, u32, u32, vp]),
            "device_image": (vp, [vp]),
            "gather_image_export": (C.c_int, [vp, u32, u32, vp]),
            "gather_image_open": (C.c_int, [vp, vp, u32]),
            "render_frame_peers": (C.c_int, [vp, P(Uniform), P(RenderOpts)]),
            "gather_image": (vp, [vp, u32]),
            "gather_configure": (C.c_int, [vp, u32, u32]),
            "gather_wait": (C.c_int, [vp, u32, vp]),
            "gather_release": (C.c_int, [vp, u32, vp]),
            "gather_copy_to_host": (C.c_int, [vp, u32, vp, vp]),
            "gather_timed_out": (C.c_int, [vp]),
            "debug_l2_read_gbs": (C.c_int, [vp, C.c_size_t, u32, P(C.c_float)]),
            "get_scene_info_buffer": (C.c_int, [vp, P(SceneBufferInfo), P(C.c_uint64)]),
            "get_tlas": (C.c_int, [vp, P(AccelInfo)]),
            "render_frame_peers_async": (C.c_int, [vp, P(Uniform), P(RenderOpts), u32]),
            "denoise_configure": (C.c_int, [vp, P(DenoiseOpts)]),
            "render_frame_async": (C.c_int, [vp, P(Uniform), P(RenderOpts), u32, vp]),
            "frame_wait": (C.c_int, [vp, u32]),
            "frame_stream": (vp, [vp, u32]),
        }
        sig.update({k: v for k, v in device_side.items() if hasattr(self.lib, self.prefix + k)})
        sig.update(self._extra)
        for name, (res, args) in sig.items():
            fn = self._f(name)
            fn.restype, fn.argtypes = res, args

    def _ck(self, rc):
        if rc != 0:
            msg = self._f("last_error")(self.ctx)
            raise BrtError(f"status {rc}: {msg.decode() if msg else ''}")

    def close(self):
        if self.ctx:
            self._f("destroy")(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene -----------------------------------------------------------------------------
    def mesh_create(self, vertices, indices):
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 8)
        i = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        out = u32()
        self._ck(self._f("mesh_create")(self.ctx, _ptr(v, Vertex), v.shape[0], _ptr(i, u32), i.shape[0], C.byref(out)))
        return out.value

    def mesh_update_vertices(self, mesh_id, vertices):
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 8)
        self._ck(self._f("mesh_update_vertices")(self.ctx, mesh_id, _ptr(v, Vertex), v.shape[0]))

    def sphere_create(self, center, radius):
        c = (f32 * 3)(*center)
        out = u32()
        self._ck(self._f("sphere_create")(self.ctx, c, radius, C.byref(out)))
        return out.value

    def material_create(self, color, metallic=0.0, roughness=1.0, specular=0.5, **kw):
        """Scene::createMaterial defaults (RT/Scene.cpp:80-86, RT/Scene.h:55): specular 0.5, rest 0."""
        m = Material()
        m.color[:] = color
        m.metallic, m.roughness, m.specular = metallic, roughness, specular
        for k, val in kw.items():
            setattr(m, k, val)
        out = u32()
        self._ck(self._f("material_create")(self.ctx, C.byref(m), C.byref(out)))
        return out.value

    def material_set_transmission(self, mat_id, transmission, ior):
        self._ck(self._f("material_set_transmission")(self.ctx, mat_id, transmission, ior))

    def light_create(self, pos, color, intensity, type=0):
        l = Light()
        l.pos[:] = pos
        l.color[:] = color
        l.intensity, l.type = intensity, type
        out = u32()
        self._ck(self._f("light_create")(self.ctx, C.byref(l), C.byref(out)))
        return out.value

    def sky_set(self, sky):
        self._ck(self._f("sky_set")(self.ctx, C.byref(sky)))

    def instance_create(self, mesh_id, material_id, xform3x4):
        x = np.ascontiguousarray(xform3x4, dtype=np.float32).reshape(12)
        out = u32()
        self._ck(self._f("instance_create")(self.ctx, mesh_id, material_id, _ptr(x, f32), C.byref(out)))
        return out.value

    def instance_set_transform(self, inst_id, xform3x4):
        x = np.ascontiguousarray(xform3x4, dtype=np.float32).reshape(12)
        self._ck(self._f("instance_set_transform")(self.ctx, inst_id, _ptr(x, f32)))

    def instance_set_material(self, inst_id, mat_id):
        self._ck(self._f("instance_set_material")(self.ctx, inst_id, mat_id))

    def instance_destroy(self, inst_id):
        self._ck(self._f("instance_destroy")(self.ctx, inst_id))

    def scene_build(self):
        self._ck(self._f("scene_build")(self.ctx))

    def smart_cull(self, uniform, width, height, threshold_px2, hysteresis):
        out = u32()
        self._ck(self._f("smart_cull")(self.ctx, C.byref(uniform), width, height, threshold_px2, hysteresis, C.byref(out)))
        return out.value

    def get_visibility(self, n):
        out = np.zeros(n, dtype=np.uint8)
        self._ck(self._f("get_visibility")(self.ctx, _ptr(out, C.c_uint8), n))
        return out

    # ---- render ----------------------------------------------------------------------------
    def camera_uniform(self, pos, rot, fovy, aspect, znear=0.001, zfar=100000.0, frame=0, depth_max=2):
        u = Uniform()
        self._f("camera_uniform")((f32 * 3)(*pos), (f32 * 3)(*rot), fovy, aspect, znear, zfar, frame, depth_max, C.byref(u))
        return u

    def camera_handle_inputs(self, keys, dt, position, rotation):
        """Camera::handleInputs (Graphics/Camera.cpp:26-61): returns the updated (position, rotation)."""
        p, r = (f32 * 3)(*position), (f32 * 3)(*rotation)
        self._f("camera_handle_inputs")(keys, dt, p, r)
        return np.array(p, dtype=np.float32), np.array(r, dtype=np.float32)

    @staticmethod
    def opts(width, height, spp=1, flags=0, crop=None):
        o = RenderOpts(width, height, spp, flags, 0, 0, 0, 0)
        if crop:
            o.crop_x0, o.crop_y0, o.crop_w, o.crop_h = crop
        return o

    def render_frame(self, uniform, opts, out=None, want_image=True):
        """Returns the image as an (h, w, 4) array — float32, or uint8 when opts.flags carries an 8-bit BRT_RENDER_FORMAT —
        (or None with want_image=False)."""
        ptr = None
        if want_image:
            if out is None:
                out = np.zeros((opts.height, opts.width, 4), dtype=np.uint8 if opts.flags & FORMAT_MASK else np.float32)
            ptr = out.ctypes.data_as(C.c_void_p)
        self._ck(self._f("render_frame")(self.ctx, C.byref(uniform), C.byref(opts), ptr))
        return out

    def render_frame_ptr(self, uniform, opts, host_ptr):
        self._ck(self._f("render_frame")(self.ctx, C.byref(uniform), C.byref(opts), C.c_void_p(host_ptr)))

    @staticmethod
    def denoise_opts(iterations=4, sigma_n_log2=5, sigma_z=0.02, sigma_l=4.0, clamp_gamma=0.0, max_history=32.0, flags=0):
        return DenoiseOpts(C.sizeof(DenoiseOpts), flags, iterations, sigma_n_log2, sigma_z, sigma_l, clamp_gamma, max_history)

    def denoise(self, uniform, dopts, width, height, want_image=True):
        """Extensions::Denoiser::denoise (Graphics/Denoiser/Denoiser.h:5-20) on the frame rendered last (BRT_RENDER_GBUFFER)."""
        out = np.zeros((height, width, 4), dtype=np.float32) if want_image else None
        self._ck(self._f("denoise")(self.ctx, C.byref(uniform), C.byref(dopts), out.ctypes.data_as(C.c_void_p) if want_image else None))
        return out

    def denoise_configure(self, dopts):
        """options of the denoiser stages run by frames rendered with the DENOISE flag"""
        self._ck(self._f("denoise_configure")(self.ctx, C.byref(dopts)))

    def get_light_bvh(self):
        """The light BVH (RT/Scene.h:123-130) as a list of LightBvhNode."""
        n = u32()
        self._ck(self._f("get_light_bvh")(self.ctx, None, 0, C.byref(n)))
        arr = (LightBvhNode * max(n.value, 1))()
        self._ck(self._f("get_light_bvh")(self.ctx, arr, n.value, C.byref(n)))
        return list(arr)[: n.value]

    def debug_get_blas(self, mesh_id):
        """(nodes, tris) of a mesh's BLAS: uint32 array [n_nodes, 20] (80-byte compressed 8-wide nodes) and float32 array [n_tris, 12]."""
        nn, nt = u32(), u32()
        self._ck(self._f("debug_get_blas")(self.ctx, mesh_id, None, 0, None, 0, C.byref(nn), C.byref(nt)))
        nodes = np.zeros((nn.value, 20), np.uint32)
        tris = np.zeros((nt.value, 12), np.float32)
        self._ck(self._f("debug_get_blas")(self.ctx, mesh_id, nodes.ctypes.data_as(C.c_void_p), nn.value, tris.ctypes.data_as(C.c_void_p), nt.value,
                                           C.byref(nn), C.byref(nt)))
        return nodes, tris

    def get_aov(self, kind, width, height):
        if kind in (AOV_POSITION, AOV_NORMAL):
            out = np.zeros((height, width, 4), dtype=np.float32)
            self._ck(self._f("get_aov")(self.ctx, kind, out.ctypes.data_as(C.c_void_p)))
            return out
        out = np.zeros((height, width), dtype=np.float32 if kind == AOV_HIT_T else np.uint32)
        self._ck(self._f("get_aov")(self.ctx, kind, out.ctypes.data_as(C.c_void_p)))
        return out

    def get_stats(self):
        s = Stats()
        self._ck(self._f("get_stats")(self.ctx, C.byref(s)))
        return s

    def trace_rays(self, rays, closest=True):
        r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        out = np.zeros((r.shape[0], 4), dtype=np.uint32)
        self._ck(self._f("trace_rays")(self.ctx, _ptr(r, f32), r.shape[0], 1 if closest else 0, _ptr(out, u32)))
        return out

    def debug_sort_pairs(self, keys, vals, bits=32):
        k = np.ascontiguousarray(keys, dtype=np.uint32).copy()
        v = np.ascontiguousarray(vals, dtype=np.uint32).copy()
        self._ck(self._f("debug_sort_pairs")(self.ctx, _ptr(k, u32), _ptr(v, u32), k.shape[0], bits))
        return k, v

    # ---- device-side entry points ------------------------------------------------------------
    def set_stream(self, stream_ptr):
        self._ck(self._f("set_stream")(self.ctx, C.c_void_p(stream_ptr)))

    def render_frame_tiles(self, uniform, opts, d_tiles_ptr):
        self._ck(self._f("render_frame_tiles")(self.ctx, C.byref(uniform), C.byref(opts), C.c_void_p(d_tiles_ptr)))

    def tile_buffer_bytes(self, width, height, world):
        return self._f("tile_buffer_bytes")(width, height, world)

    def untile(self, d_all_ptr, width, height, world, d_rgba_ptr):
        self._ck(self._f("untile")(self.ctx, C.c_void_p(d_all_ptr), width, height, world, C.c_void_p(d_rgba_ptr)))

    def device_image(self):
        return self._f("device_image")(self.ctx)

    # frames in flight (VK/SwapChain.h:8: MAX_FRAMES_IN_FLIGHT = 2)
    def render_frame_async(self, uniform, opts, slot, host_ptr=None):
        """Enqueues the frame on `slot` (after waiting for the slot's previous frame) and returns; host_ptr = address of a
        (pinned) RGBA32F host buffer or None."""
        self._ck(self._f("render_frame_async")(self.ctx, C.byref(uniform), C.byref(opts), slot, C.c_void_p(host_ptr) if host_ptr else None))

    def frame_wait(self, slot):
        self._ck(self._f("frame_wait")(self.ctx, slot))

    def frame_stream(self, slot):
        return self._f("frame_stream")(self.ctx, slot) or 0  # NULL = the legacy default stream

    # fused resolve + exchange over peer memory
    def gather_image_export(self, width, height):
        h = C.create_string_buffer(64)
        self._ck(self._f("gather_image_export")(self.ctx, width, height, h))
        return h.raw

    def gather_image_open(self, handles):
        blob = b"".join(handles)
        self._ck(self._f("gather_image_open")(self.ctx, C.c_char_p(blob), len(handles)))

    def render_frame_peers(self, uniform, opts):
        self._ck(self._f("render_frame_peers")(self.ctx, C.byref(uniform), C.byref(opts)))

    def render_frame_peers_async(self, uniform, opts, slot):
        self._ck(self._f("render_frame_peers_async")(self.ctx, C.byref(uniform), C.byref(opts), slot))

    def gather_image(self, slot):
        return self._f("gather_image")(self.ctx, slot)

    def gather_configure(self, root_only, root=0):
        self._ck(self._f("gather_configure")(self.ctx, 1 if root_only else 0, root))

    def gather_wait(self, slot, stream_ptr=None):
        self._ck(self._f("gather_wait")(self.ctx, slot, C.c_void_p(stream_ptr) if stream_ptr else None))

    def gather_release(self, slot, stream_ptr=None):
        self._ck(self._f("gather_release")(self.ctx, slot, C.c_void_p(stream_ptr) if stream_ptr else None))

    def gather_copy_to_host(self, slot, host_ptr, stream_ptr=None):
        self._ck(self._f("gather_copy_to_host")(self.ctx, slot, C.c_void_p(host_ptr), C.c_void_p(stream_ptr) if stream_ptr else None))

    def debug_l2_read_gbs(self, nbytes=64 << 20, iters=10):
        out = C.c_float()
        self._ck(self._f("debug_l2_read_gbs")(self.ctx, nbytes, iters, C.byref(out)))
        return out.value

    def get_scene_info_buffer(self):
        out, dptr = SceneBufferInfo(), C.c_uint64()
        self._ck(self._f("get_scene_info_buffer")(self.ctx, C.byref(out), C.byref(dptr)))
        return out, dptr.value

    def get_tlas(self):
        out = AccelInfo()
        self._ck(self._f("get_tlas")(self.ctx, C.byref(out)))
        return out

    def gather_timed_out(self):
        return self._f("gather_timed_out")(self.ctx)
