"""Deterministic procedural scenes for the BASELINE.json configs.

Everything is generated on the host with numpy from the reference's own integer hash
(SH/random.slang:2-12) so the CUDA product and the CPU oracle receive byte-identical input.
World conventions follow Graphics/Camera.cpp: Y points DOWN, +Z is forward at yaw 0.
"""
from dataclasses import dataclass, field

import numpy as np

from . import _binding as B


# ---- the reference's hash, vectorised (SH/random.slang:2-12) --------------------------------
def hash3(x, y, z):
    x = np.asarray(x, dtype=np.uint32)
    y = np.asarray(y, dtype=np.uint32)
    z = np.asarray(z, dtype=np.uint32)
    p0, p1, p2, p3 = np.uint32(2246822519), np.uint32(3266489917), np.uint32(668265263), np.uint32(374761393)
    with np.errstate(over="ignore"):
        h = z + p3 + x * p1
        h = p2 * ((h << np.uint32(17)) | (h >> np.uint32(15)))
        h = h + y * p1
        h = p2 * ((h << np.uint32(17)) | (h >> np.uint32(15)))
        h = p0 * (h ^ (h >> np.uint32(15)))
        h = p1 * (h ^ (h >> np.uint32(13)))
        return h ^ (h >> np.uint32(16))


@dataclass
class MaterialDesc:
    color: tuple
    metallic: float = 0.0
    roughness: float = 1.0
    specular: float = 0.5
    extra: dict = field(default_factory=dict)
    transmission: float = 0.0
    ior: float = 1.5


@dataclass
class SceneDesc:
    name: str
    meshes: list = field(default_factory=list)      # ("tri", vertices[n,8] f32, indices u32) | ("sphere", center, radius)
    materials: list = field(default_factory=list)   # MaterialDesc
    lights: list = field(default_factory=list)      # (pos, color, intensity[, type])
    instances: list = field(default_factory=list)   # (mesh, material, xform 3x4)
    cam_pos: tuple = (0.0, 0.0, -2.0)
    cam_rot: tuple = (0.0, 0.0, 0.0)
    fovy: float = float(np.radians(np.float32(60.0)))
    znear: float = 0.001
    zfar: float = 100000.0

    def triangles(self):
        per_mesh = [0 if m[0] == "sphere" else len(m[2]) // 3 for m in self.meshes]
        return sum(per_mesh[i[0]] for i in self.instances)

    def upload(self, api, build=True):
        """Replays the scene through the Scene API (mesh -> material -> light -> instance -> build)."""
        ids = []
        for m in self.meshes:
            ids.append(api.sphere_create(m[1], m[2]) if m[0] == "sphere" else api.mesh_create(m[1], m[2]))
        for md in self.materials:
            mid = api.material_create(md.color, md.metallic, md.roughness, md.specular, **md.extra)
            if md.transmission > 0.0:
                api.material_set_transmission(mid, md.transmission, md.ior)
        for light in self.lights:  # (pos, color, intensity[, type]); Scene::createLight always makes POINT lights (RT/Scene.cpp:88-97)
            api.light_create(*light)
        for mesh, mat, x in self.instances:
            api.instance_create(ids[mesh], mat, x)
        if build:
            api.scene_build()
        return api

    def uniform(self, api, width, height, frame=0, depth_max=2):
        return api.camera_uniform(self.cam_pos, self.cam_rot, self.fovy, width / height, self.znear, self.zfar, frame, depth_max)


def xform(scale=(1, 1, 1), translate=(0, 0, 0), yaw=0.0):
    """Row-major 3x4 object->world. yaw == 0 gives MeshInstance::calculateTransformation (RT/MeshInstance.h:82-85)."""
    c, s = np.float32(np.cos(np.float32(yaw))), np.float32(np.sin(np.float32(yaw)))
    r = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float32)
    m = np.zeros((3, 4), dtype=np.float32)
    m[:, :3] = r * np.asarray(scale, dtype=np.float32)[None, :]
    m[:, 3] = translate
    return m


def _quad(p0, p1, p2, p3):
    """Two triangles, flat normal; 4 private vertices."""
    p = np.array([p0, p1, p2, p3], dtype=np.float32)
    n = np.cross(p[1] - p[0], p[3] - p[0])
    n = (n / np.linalg.norm(n)).astype(np.float32)
    v = np.zeros((4, 8), dtype=np.float32)
    v[:, 0:3] = p
    v[:, 3:6] = n
    v[:, 6:8] = [[0, 0], [1, 0], [1, 1], [0, 1]]
    return v, np.array([0, 1, 2, 0, 2, 3], dtype=np.uint32)


def _merge(parts):
    vs, is_, base = [], [], 0
    for v, i in parts:
        vs.append(v)
        is_.append(i + np.uint32(base))
        base += len(v)
    return np.concatenate(vs).astype(np.float32), np.concatenate(is_).astype(np.uint32)


def _open_cube():
    """Unit cube [-0.5,0.5]^3 without its bottom (y = +0.5 is down): 5 faces = 10 triangles."""
    a = 0.5
    f = [
        _quad((-a, -a, -a), (a, -a, -a), (a, -a, a), (-a, -a, a)),    # top (y = -0.5 is up)
        _quad((-a, -a, -a), (-a, a, -a), (a, a, -a), (a, -a, -a)),    # front z-
        _quad((-a, -a, a), (a, -a, a), (a, a, a), (-a, a, a)),        # back z+
        _quad((-a, -a, -a), (-a, -a, a), (-a, a, a), (-a, a, -a)),    # left
        _quad((a, -a, -a), (a, a, -a), (a, a, a), (a, -a, a)),        # right
    ]
    return _merge(f)


def cornell():
    """C1: Cornell box in [-1,1]^3, 32 triangles + 2 spheres, one point light under the ceiling."""
    s = SceneDesc("cornell")
    white = _merge([
        _quad((-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)),       # back wall z=+1
        _quad((-1, 1, -1), (1, 1, -1), (1, 1, 1), (-1, 1, 1)),       # floor y=+1
        _quad((-1, -1, -1), (1, -1, -1), (1, -1, 1), (-1, -1, 1)),   # ceiling y=-1
        _quad((-0.25, -0.99, -0.25), (0.25, -0.99, -0.25), (0.25, -0.99, 0.25), (-0.25, -0.99, 0.25)),  # light quad
    ])
    left = _quad((-1, -1, -1), (-1, -1, 1), (-1, 1, 1), (-1, 1, -1))
    right = _quad((1, -1, -1), (1, 1, -1), (1, 1, 1), (1, -1, 1))
    cube = _open_cube()
    s.meshes = [("tri",) + white, ("tri",) + left, ("tri",) + right, ("tri",) + cube,
                ("sphere", (0.0, 0.0, 0.0), 0.3), ("sphere", (0.0, 0.0, 0.0), 0.3)]
    s.materials = [
        MaterialDesc((0.73, 0.73, 0.73)),                      # 0 white
        MaterialDesc((0.65, 0.05, 0.05)),                      # 1 red
        MaterialDesc((0.12, 0.45, 0.15)),                      # 2 green
        MaterialDesc((0.9, 0.9, 0.9), metallic=1.0, roughness=0.05),   # 3 mirror
        MaterialDesc((1.0, 1.0, 1.0), roughness=0.05, transmission=1.0, ior=1.5),  # 4 glass
    ]
    s.lights = [((0.0, -0.9, 0.0), (1.0, 1.0, 1.0), 2.0)]
    ident = xform()
    s.instances = [
        (0, 0, ident), (1, 1, ident), (2, 2, ident),
        (3, 0, xform((0.6, 0.6, 0.6), (0.35, 0.7, -0.25), yaw=-0.3)),      # short box
        (3, 0, xform((0.6, 1.2, 0.6), (-0.35, 0.4, 0.35), yaw=0.3)),       # tall box
        (4, 3, xform((1, 1, 1), (0.35, 0.1, -0.25))),                       # mirror sphere on the short box
        (5, 4, xform((1, 1, 1), (-0.45, 0.7, -0.45))),                      # glass sphere on the floor
    ]
    s.cam_pos, s.cam_rot = (0.0, 0.0, -3.4), (0.0, 0.0, 0.0)
    return s


def icosphere(subdiv):
    """Unit icosphere, 20 * 4^subdiv triangles, smooth normals (= positions)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = np.array(v, dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array(f, dtype=np.int64)
    for _ in range(subdiv):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        es = np.sort(e, axis=1)
        key = es[:, 0] * (len(v) + 1) + es[:, 1]
        uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
        mid = v[es[first, 0]] + v[es[first, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid])
        n = len(f)
        a, b, c = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        f = np.concatenate([np.stack([f[:, 0], a, c], 1), np.stack([f[:, 1], b, a], 1), np.stack([f[:, 2], c, b], 1), np.stack([a, b, c], 1)])
    out = np.zeros((len(v), 8), dtype=np.float32)
    out[:, 0:3] = v
    out[:, 3:6] = v
    out[:, 6] = 0.5 + np.arctan2(v[:, 2], v[:, 0]) / (2 * np.pi)
    out[:, 7] = 0.5 - np.arcsin(np.clip(v[:, 1], -1, 1)) / np.pi
    return out, f.astype(np.uint32).reshape(-1)


def _value_noise(x, z, seed):
    xi, zi = np.floor(x).astype(np.int64), np.floor(z).astype(np.int64)
    fx, fz = x - xi, z - zi
    sx, sz = fx * fx * (3 - 2 * fx), fz * fz * (3 - 2 * fz)

    def lat(i, j):
        return hash3((i & 0xFFFFFFFF).astype(np.uint32), (j & 0xFFFFFFFF).astype(np.uint32), np.uint32(seed)).astype(np.float64) / 4294967296.0

    a, b, c, d = lat(xi, zi), lat(xi + 1, zi), lat(xi, zi + 1), lat(xi + 1, zi + 1)
    return (a * (1 - sx) + b * sx) * (1 - sz) + (c * (1 - sx) + d * sx) * sz


def heightfield(n=512, extent=16.0, amplitude=1.6, seed=0xB100B200):
    """n x n quads (2 n^2 triangles) over [-extent/2, extent/2]^2, y = -fbm (Y is down), smooth normals."""
    g = np.arange(n + 1, dtype=np.float64)
    gx, gz = np.meshgrid(g, g, indexing="xy")
    u, w = gx / n * 6.0, gz / n * 6.0
    h = np.zeros_like(u)
    amp, freq = 1.0, 1.0
    for o in range(5):
        h += amp * _value_noise(u * freq, w * freq, (seed + o) & 0xFFFFFFFF)
        amp *= 0.5
        freq *= 2.0
    h = (h / 1.9375 - 0.5) * amplitude
    x = (gx / n - 0.5) * extent
    z = (gz / n - 0.5) * extent
    y = -h
    dx = extent / n
    dydx = np.gradient(y, dx, axis=1)
    dydz = np.gradient(y, dx, axis=0)
    nrm = np.stack([dydx, -np.ones_like(y), dydz], axis=-1)  # points up (-Y)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    v = np.zeros(((n + 1) * (n + 1), 8), dtype=np.float32)
    v[:, 0], v[:, 1], v[:, 2] = x.ravel(), y.ravel(), z.ravel()
    v[:, 3:6] = nrm.reshape(-1, 3)
    v[:, 6], v[:, 7] = (gx / n).ravel(), (gz / n).ravel()
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    v00 = (j * (n + 1) + i).ravel()
    v10, v01, v11 = v00 + 1, v00 + n + 1, v00 + n + 2
    idx = np.stack([v00, v10, v11, v00, v11, v01], axis=1).astype(np.uint32).reshape(-1)
    return v, idx


def _terrain_height(v, n, extent, px, pz):
    i = int(round((px / extent + 0.5) * n))
    j = int(round((pz / extent + 0.5) * n))
    return float(v[j * (n + 1) + i, 1])


def terrain_icospheres(n=512, subdiv=5, n_spheres=24, all_diffuse=False):
    """C2/C3/C5: heightfield (2 n^2 tris) + n_spheres instances of one icosphere (20*4^subdiv tris each).
    Defaults: 524,288 + 24 * 20,480 = 1,015,808 triangles."""
    s = SceneDesc("terrain_icospheres")
    extent = 16.0
    hv, hi = heightfield(n, extent)
    sv, si = icosphere(subdiv)
    s.meshes = [("tri", hv, hi), ("tri", sv, si)]
    mats = [
        MaterialDesc((0.55, 0.5, 0.42)),                                        # 0 terrain, diffuse
        MaterialDesc((0.95, 0.93, 0.88), metallic=1.0, roughness=0.05),        # 1 polished metal
        MaterialDesc((0.9, 0.6, 0.3), metallic=1.0, roughness=0.3),            # 2 brushed copper
        MaterialDesc((0.8, 0.8, 0.85), metallic=1.0, roughness=0.6),           # 3 rough metal
        MaterialDesc((0.7, 0.2, 0.2), metallic=0.0, roughness=1.0),            # 4 red diffuse
        MaterialDesc((0.2, 0.3, 0.7), metallic=0.0, roughness=0.4),            # 5 blue plastic
        MaterialDesc((0.3, 0.7, 0.3), metallic=0.0, roughness=0.7),            # 6 green diffuse
        MaterialDesc((1.0, 1.0, 1.0), roughness=0.05, transmission=1.0, ior=1.5),  # 7 glass
    ]
    if all_diffuse:
        mats = [MaterialDesc(m.color, 0.0, 1.0) for m in mats]
    s.materials = mats
    # lights: positions / colours of RT/RTApp.cpp:9-11 scaled to the scene (intensity 2 -> 2 * 40)
    s.lights = [((7.0, -6.0, 0.0), (0.0, 0.0, 1.0), 80.0), ((-7.0, -6.0, 0.0), (0.0, 1.0, 0.0), 80.0),
                ((0.0, -6.0, -7.0), (1.0, 0.0, 0.0), 80.0)]
    s.instances = [(0, 0, xform())]
    cols = 6
    for k in range(n_spheres):
        gx, gz = k % cols, k // cols
        jx = (int(hash3(k, 1, 77)) / 4294967296.0 - 0.5) * 0.8
        jz = (int(hash3(k, 2, 77)) / 4294967296.0 - 0.5) * 0.8
        r = 0.45 + 0.35 * (int(hash3(k, 3, 77)) / 4294967296.0)
        px = (gx - (cols - 1) / 2) * 2.2 + jx
        pz = (gz - 1.5) * 2.4 + jz + 1.0
        py = _terrain_height(hv, n, extent, px, pz) - r * 0.95
        s.instances.append((1, 1 + k % 7, xform((r, r, r), (px, py, pz))))
    s.cam_pos, s.cam_rot = (0.0, -3.2, -10.5), (-0.22, 0.0, 0.0)
    return s


def instanced_lattice(grid=8, subdiv=5, animated=True):
    """C4: one icosphere BLAS x grid^3 instances on a jittered lattice (scale + translate only, as the
    reference) + one separately animated icosphere mesh. Defaults: 512 * 20,480 = 10.49 M triangles."""
    s = SceneDesc("instanced_lattice")
    sv, si = icosphere(subdiv)
    s.meshes = [("tri", sv, si)]
    s.materials = [MaterialDesc((0.7, 0.7, 0.7)), MaterialDesc((0.8, 0.3, 0.2)), MaterialDesc((0.2, 0.4, 0.8), metallic=0.0, roughness=0.5)]
    s.lights = [((0.0, -30.0, -30.0), (1.0, 1.0, 1.0), 2500.0)]
    k = 0
    for iz in range(grid):
        for iy in range(grid):
            for ix in range(grid):
                j = [(int(hash3(k, a, 99)) / 4294967296.0 - 0.5) for a in range(4)]
                r = 0.6 + 0.5 * (j[3] + 0.5)
                p = ((ix - (grid - 1) / 2) * 3.0 + j[0], (iy - (grid - 1) / 2) * 3.0 + j[1], (iz - (grid - 1) / 2) * 3.0 + j[2] + 14.0)
                s.instances.append((0, k % 3, xform((r, r, r), p)))
                k += 1
    if animated:
        av, ai = icosphere(subdiv)
        s.meshes.append(("tri", av, ai))
        s.instances.append((1, 1, xform((2.0, 2.0, 2.0), (0.0, 0.0, -2.0))))
    s.cam_pos, s.cam_rot = (0.0, 0.0, -12.0), (0.0, 0.0, 0.0)
    return s


def animate_icosphere(base_vertices, frame):
    """C4 per-frame vertex animation: radial sine wave, phase = frame."""
    v = base_vertices.copy()
    p = base_vertices[:, 0:3].astype(np.float64)
    d = 1.0 + 0.15 * np.sin(6.0 * p[:, 1] + 0.7 * frame)
    v[:, 0:3] = (p * d[:, None]).astype(np.float32)
    return v


def rtapp_demo():
    """The reference's hard-coded demo (RT/RTApp.cpp:3-26): two instances of a unit plane, two metallic
    materials, three coloured point lights, camera at (0,0,-2). models/Plane.obj is not in the repo, so a
    2 x 2 quad in the XZ plane stands in."""
    s = SceneDesc("rtapp_demo")
    s.meshes = [("tri",) + _quad((-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1))]
    s.materials = [MaterialDesc((1, 1, 1), metallic=1.0), MaterialDesc((1, 1, 1), metallic=1.0, roughness=0.0)]
    s.lights = [((1.0, 0.0, 0.0), (0.0, 0.0, 1.0), 2.0), ((-1.0, 0.0, 0.0), (0.0, 1.0, 0.0), 2.0), ((0.0, 0.0, -1.0), (1.0, 0.0, 0.0), 2.0)]
    s.instances = [(0, 1, xform((1, 1, 1), (0, -1, 0))), (0, 0, xform((4, 1, 4), (0, 1, 0)))]
    s.cam_pos = (0.0, 0.0, -2.0)
    return s


# render settings per BASELINE.json config
CONFIGS = {
    "c1": dict(scene="cornell", width=512, height=512, spp=1, depth_max=1, flags=0),
    "c2": dict(scene="terrain", width=1920, height=1080, spp=1, depth_max=3, flags=B.BOUNCE_REFLECT | B.BOUNCE_REFRACT),
    "c3": dict(scene="terrain", width=3840, height=2160, spp=16, depth_max=5,
               flags=B.BOUNCE_REFLECT | B.BOUNCE_REFRACT | B.BOUNCE_DIFFUSE | B.JITTER),
    "c4": dict(scene="lattice", width=1920, height=1080, spp=1, depth_max=1, flags=0),
    "c5": dict(scene="terrain_diffuse", width=3840, height=2160, spp=1, depth_max=9, flags=B.BOUNCE_DIFFUSE),
}


def make_scene(kind, small=False):
    if kind == "cornell":
        return cornell()
    if kind == "terrain":
        return terrain_icospheres(64, 2, 24) if small else terrain_icospheres()
    if kind == "terrain_diffuse":
        return terrain_icospheres(64, 2, 24, all_diffuse=True) if small else terrain_icospheres(all_diffuse=True)
    if kind == "lattice":
        return instanced_lattice(3, 2) if small else instanced_lattice()
    if kind == "rtapp":
        return rtapp_demo()
    raise ValueError(kind)
