// ORACLE — TEST INFRASTRUCTURE ONLY. Pinned to the reference's shader binary and host code where they exist (see shaders.hpp header).
//
// Scene container, two-level BVH queries, the path loop of SH/raytracing.slang and the C API of
// oracle.h. Shorthand: SH/ = reference shaders/, RT/ = reference Graphics/RayTracing/.
#include "oracle.h"

#include <atomic>
#include <cstdio>
#include <string>
#include <thread>

#include "bvh.hpp"
#include "shaders.hpp"
#include "denoise.hpp"

using namespace orc;

namespace {

struct Mesh {
  std::vector<brt_vertex> v;
  std::vector<uint32_t> idx;
  bool sphere = false;
  vec3 c = V3(0.0f);
  float r = 0.0f;
  vec3 lo = V3(0.0f), hi = V3(0.0f);  // object-space bounds (exact min/max of the vertex positions)
  BVH bvh;
  bool dirty = true;
};

struct Instance {
  uint32_t mesh, mat;
  float o2w[12], w2o[12];
  vec3 wlo, whi;  // world bounds
  bool visible = true;
};

struct MatExt {
  float transmission = 0.0f, ior = 1.5f;
};

struct Hit {
  float t, u, v;
  uint32_t inst, prim;
  bool hit;
};

}  // namespace

struct orc_context {
  std::vector<Mesh> meshes;
  std::vector<Instance> inst;
  std::vector<Material> mats;
  std::vector<MatExt> mat_ext;
  std::vector<brt_light> lights;
  std::vector<brt_light_bvh_node> light_bvh;  // RT/Scene.h:123-130
  brt_sky sky{};
  BVH tlas;                        // over the visible instances
  std::vector<uint32_t> tlas_ids;  // tlas primitive -> instance index
  bool built = false, brute = false;
  uint32_t threads = 0;
  uint32_t tile_rank = 0, tile_world = 1;
  std::string err;
  // last frame
  uint32_t fw = 0, fh = 0;
  std::vector<uint32_t> aov_prim, aov_inst;
  std::vector<float> aov_t;
  std::vector<Px4> aov_pos, aov_nrm, last_image;  // BRT_RENDER_GBUFFER + the linear frame (inputs of orc_denoise)
  bool has_gbuffer = false;
  uint32_t render_flags = 0;
  DenoiseState dn;
  brt_stats stats{};
};

namespace {

int fail(orc_context* c, int code, const char* msg) {
  if (c) c->err = msg;
  return code;
}

// row-major 3x4 affine inverse, computed in double and rounded once (DESIGN.md §3 "W2O")
void invert3x4(const float m[12], float out[12]) {
  double a = m[0], b = m[1], c = m[2], tx = m[3];
  double d = m[4], e = m[5], f = m[6], ty = m[7];
  double g = m[8], h = m[9], i = m[10], tz = m[11];
  double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  double det = a * A + b * B + c * C;
  double id = 1.0 / det;
  double r00 = A * id, r01 = -(b * i - c * h) * id, r02 = (b * f - c * e) * id;
  double r10 = B * id, r11 = (a * i - c * g) * id, r12 = -(a * f - c * d) * id;
  double r20 = C * id, r21 = -(a * h - b * g) * id, r22 = (a * e - b * d) * id;
  out[0] = (float)r00; out[1] = (float)r01; out[2] = (float)r02; out[3] = (float)(-(r00 * tx + r01 * ty + r02 * tz));
  out[4] = (float)r10; out[5] = (float)r11; out[6] = (float)r12; out[7] = (float)(-(r10 * tx + r11 * ty + r12 * tz));
  out[8] = (float)r20; out[9] = (float)r21; out[10] = (float)r22; out[11] = (float)(-(r20 * tx + r21 * ty + r22 * tz));
}

inline vec3 xform_point(const float m[12], vec3 p) {
  return V3(((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3], ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
            ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]);
}
inline vec3 xform_dir(const float m[12], vec3 p) {
  return V3((m[0] * p.x + m[1] * p.y) + m[2] * p.z, (m[4] * p.x + m[5] * p.y) + m[6] * p.z, (m[8] * p.x + m[9] * p.y) + m[10] * p.z);
}
// mul(WorldToObject4x3(), n): n_j = sum_i W2O(i,j) * n_i  (inverse-transpose, SH/raytracing.slang:150)
inline vec3 xform_normal(const float w2o[12], vec3 n) {
  return V3((w2o[0] * n.x + w2o[4] * n.y) + w2o[8] * n.z, (w2o[1] * n.x + w2o[5] * n.y) + w2o[9] * n.z,
            (w2o[2] * n.x + w2o[6] * n.y) + w2o[10] * n.z);
}

inline vec3 vpos(const brt_vertex& v) { return V3(v.pos[0], v.pos[1], v.pos[2]); }
inline vec3 vnrm(const brt_vertex& v) { return V3(v.normal[0], v.normal[1], v.normal[2]); }

void instance_world_bounds(const Mesh& m, Instance& in) {
  // centre/extent form: c' = M c, e'_j = sum_i |M_ji| e_i
  vec3 c = (m.lo + m.hi) * 0.5f, e = (m.hi - m.lo) * 0.5f;
  vec3 wc = xform_point(in.o2w, c);
  const float* q = in.o2w;
  vec3 we = V3((std::fabs(q[0]) * e.x + std::fabs(q[1]) * e.y) + std::fabs(q[2]) * e.z,
               (std::fabs(q[4]) * e.x + std::fabs(q[5]) * e.y) + std::fabs(q[6]) * e.z,
               (std::fabs(q[8]) * e.x + std::fabs(q[9]) * e.y) + std::fabs(q[10]) * e.z);
  in.wlo = wc - we;
  in.whi = wc + we;
}

void build_mesh(Mesh& m) {
  if (m.sphere) {
    m.lo = m.c - V3(m.r);
    m.hi = m.c + V3(m.r);
    m.dirty = false;
    return;
  }
  uint32_t nt = (uint32_t)(m.idx.size() / 3);
  std::vector<vec3> lo(nt), hi(nt);
  vec3 mlo = V3(INFINITY), mhi = V3(-INFINITY);
  for (uint32_t t = 0; t < nt; ++t) {
    vec3 a = vpos(m.v[m.idx[3 * t]]), b = vpos(m.v[m.idx[3 * t + 1]]), c = vpos(m.v[m.idx[3 * t + 2]]);
    lo[t] = V3(std::fmin(std::fmin(a.x, b.x), c.x), std::fmin(std::fmin(a.y, b.y), c.y), std::fmin(std::fmin(a.z, b.z), c.z));
    hi[t] = V3(std::fmax(std::fmax(a.x, b.x), c.x), std::fmax(std::fmax(a.y, b.y), c.y), std::fmax(std::fmax(a.z, b.z), c.z));
    mlo = V3(std::fmin(mlo.x, lo[t].x), std::fmin(mlo.y, lo[t].y), std::fmin(mlo.z, lo[t].z));
    mhi = V3(std::fmax(mhi.x, hi[t].x), std::fmax(mhi.y, hi[t].y), std::fmax(mhi.z, hi[t].z));
  }
  if (nt == 0) { mlo = V3(0.0f); mhi = V3(0.0f); }
  m.lo = mlo;
  m.hi = mhi;
  build_bvh(m.bvh, lo, hi, 4);
  m.dirty = false;
}

void build_tlas(orc_context* c) {
  std::vector<vec3> lo, hi;
  c->tlas_ids.clear();
  for (uint32_t i = 0; i < c->inst.size(); ++i) {
    Instance& in = c->inst[i];
    const Mesh& m = c->meshes[in.mesh];
    instance_world_bounds(m, in);
    if (!in.visible) continue;
    if (!m.sphere && m.idx.empty()) continue;
    lo.push_back(in.wlo);
    hi.push_back(in.whi);
    c->tlas_ids.push_back(i);
  }
  build_bvh(c->tlas, lo, hi, 1);
}

struct Counters {
  uint64_t rays_closest = 0, rays_occl = 0, nodes_c = 0, prims_c = 0, sph_c = 0, nodes_o = 0, prims_o = 0, sph_o = 0;
};

// lexicographic (t, inst, prim) order — the documented tie-break
inline bool better(float t, uint32_t ii, uint32_t prim, const Hit& best) {
  if (!best.hit || t < best.t) return true;
  return t == best.t && (ii < best.inst || (ii == best.inst && prim < best.prim));
}

// closest hit in one instance; updates `best` with the (t, inst, prim) lexicographic minimum
inline void closest_in_instance(const orc_context* c, uint32_t ii, const Ray& r, Hit& best, Counters& k) {
  const Instance& in = c->inst[ii];
  const Mesh& m = c->meshes[in.mesh];
  vec3 o = xform_point(in.w2o, r.o), d = xform_dir(in.w2o, r.d);
  if (m.sphere) {
    float t;
    k.sph_c++;
    // candidates with t == best.t must reach the tie-break, hence the one-ulp-open upper limit
    float lim = best.hit ? std::nextafter(best.t, INFINITY) : r.tmax;
    if (intersect_sphere(o, d, r.tmin, lim, m.c, m.r, t) && better(t, ii, 0u, best)) best = Hit{t, 0.0f, 0.0f, ii, 0u, true};
    return;
  }
  RayShear sh = make_shear(d);
  auto test = [&](uint32_t prim) {
    k.prims_c++;
    float t, u, v;
    vec3 a = vpos(m.v[m.idx[3 * prim]]), b = vpos(m.v[m.idx[3 * prim + 1]]), cc = vpos(m.v[m.idx[3 * prim + 2]]);
    float lim = best.hit ? std::nextafter(best.t, INFINITY) : r.tmax;
    if (intersect_tri(o, sh, r.tmin, lim, a, b, cc, t, u, v) && better(t, ii, prim, best)) best = Hit{t, u, v, ii, prim, true};
  };
  if (c->brute) {
    for (uint32_t p = 0; p < m.idx.size() / 3; ++p) test(p);
  } else {
    float tmax = best.hit ? best.t : r.tmax;
    // the traversal culls with entry <= tmax, so equal-t candidates in other leaves are still visited
    auto leaf = [&](uint32_t prim) {
      test(prim);
      if (best.hit) tmax = best.t;
      return false;
    };
    traverse(m.bvh, o, d, r.tmin, tmax, k.nodes_c, leaf);
  }
}

Hit trace_closest(const orc_context* c, const Ray& r, Counters& k) {
  Hit best{0.0f, 0.0f, 0.0f, BRT_AOV_MISS, BRT_AOV_MISS, false};
  k.rays_closest++;
  if (c->brute) {
    for (uint32_t id : c->tlas_ids) closest_in_instance(c, id, r, best, k);
    return best;
  }
  float tmax = r.tmax;
  auto leaf = [&](uint32_t p) {
    closest_in_instance(c, c->tlas_ids[p], r, best, k);
    if (best.hit) tmax = best.t;
    return false;
  };
  traverse(c->tlas, r.o, r.d, r.tmin, tmax, k.nodes_c, leaf);
  return best;
}

inline bool occluded_in_instance(const orc_context* c, uint32_t ii, const Ray& r, Counters& k) {
  const Instance& in = c->inst[ii];
  const Mesh& m = c->meshes[in.mesh];
  vec3 o = xform_point(in.w2o, r.o), d = xform_dir(in.w2o, r.d);
  if (m.sphere) {
    float t;
    k.sph_o++;
    return intersect_sphere(o, d, r.tmin, r.tmax, m.c, m.r, t);
  }
  RayShear sh = make_shear(d);
  auto test = [&](uint32_t prim) {
    k.prims_o++;
    float t, u, v;
    return intersect_tri(o, sh, r.tmin, r.tmax, vpos(m.v[m.idx[3 * prim]]), vpos(m.v[m.idx[3 * prim + 1]]), vpos(m.v[m.idx[3 * prim + 2]]), t, u, v);
  };
  if (c->brute) {
    for (uint32_t p = 0; p < m.idx.size() / 3; ++p)
      if (test(p)) return true;
    return false;
  }
  float tmax = r.tmax;
  return traverse(m.bvh, o, d, r.tmin, tmax, k.nodes_o, test);
}

// RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH | SKIP_CLOSEST_HIT_SHADER (SH/raytracing.slang:67)
bool trace_occluded(const orc_context* c, const Ray& r, Counters& k) {
  k.rays_occl++;
  if (c->brute) {
    for (uint32_t id : c->tlas_ids)
      if (occluded_in_instance(c, id, r, k)) return true;
    return false;
  }
  float tmax = r.tmax;
  return traverse(c->tlas, r.o, r.d, r.tmin, tmax, k.nodes_o, [&](uint32_t p) { return occluded_in_instance(c, c->tlas_ids[p], r, k); });
}

// ---- shading ------------------------------------------------------------------------------------
inline bool is_zero(vec3 v) { return v.x == 0.0f && v.y == 0.0f && v.z == 0.0f; }

// extension (BRT_RENDER_SKY): gradient from the SkyInfo the reference uploads but never reads
vec3 sky_color(const brt_sky& s, vec3 dir) {
  vec3 d = normalize(dir);
  float h = dot(d, V3(s.upDirection[0], s.upDirection[1], s.upDirection[2]));
  vec3 hor = V3(s.horizonColor[0], s.horizonColor[1], s.horizonColor[2]);
  vec3 c;
  if (h >= 0.0f)
    c = lerp3(hor, V3(s.skyColor[0], s.skyColor[1], s.skyColor[2]), clampf(h / s.horizonSize, 0.0f, 1.0f));
  else
    c = lerp3(hor, V3(s.groundColor[0], s.groundColor[1], s.groundColor[2]), clampf(-h / s.horizonSize, 0.0f, 1.0f));
  return c * s.brightness;
}

// calculateColor (SH/raytracing.slang:72-88) + processLight (SH/light.slang:23-39) + testShadow (:56-70).
// Deviation that cannot change the image: the shadow ray is skipped when the unshadowed
// contribution is exactly zero (0 * shadowFactor == 0 either way); it is then not counted as a ray.
// ---- light BVH (the reference declares LightBVHNode, RT/Scene.h:123-130, and names its purpose at SH/raytracing.slang:76;
// there is no code to follow: the rule is DESIGN.md §13) -------------------------------------------------------
// build: recursive median split on the widest axis of the positions, ties by light index; children adjacent, left subtree first
// Point lights and constant-direction lights (SPOT / DIRECTIONAL: SH/light.slang:33-36) never share a subtree; the cone fields tell them apart
// (point: axis (0,0,1), angle pi; constant direction: axis -normalize(0.9, -0.1, 0), angle 0, importance independent of the shading point).
void light_bvh_build_node(const std::vector<brt_light>& lights, std::vector<uint32_t>& order, std::vector<brt_light_bvh_node>& nodes, uint32_t node,
                          uint32_t first, uint32_t count, bool directional) {
  brt_light_bvh_node nd{};
  for (int k = 0; k < 3; ++k) { nd.bBoxMin[k] = INFINITY; nd.bBoxMax[k] = -INFINITY; }
  float flux = 0.0f;
  for (uint32_t i = first; i < first + count; ++i) {
    const brt_light& l = lights[order[i]];
    for (int k = 0; k < 3; ++k) {
      nd.bBoxMin[k] = std::fmin(nd.bBoxMin[k], l.pos[k]);
      nd.bBoxMax[k] = std::fmax(nd.bBoxMax[k], l.pos[k]);
    }
    flux = flux + std::fabs(l.intensity * ((0.2126f * l.color[0] + 0.7152f * l.color[1]) + 0.0722f * l.color[2]));
  }
  nd.totalFlux = flux;
  if (directional) {
    nd.coneAxis[0] = -0.993883729f; nd.coneAxis[1] = 0.110431522f; nd.coneAxis[2] = 0.0f;
    nd.coneAngle = 0.0f;
  } else {
    nd.coneAxis[2] = 1.0f;      // point lights radiate everywhere: cone = the whole sphere
    nd.coneAngle = 3.14159274f;
  }
  if (count == 1) {
    nd.childIndex = -1 - (int32_t)order[first];
    nodes[node] = nd;
    return;
  }
  int axis = 0;
  for (int k = 1; k < 3; ++k)
    if (nd.bBoxMax[k] - nd.bBoxMin[k] > nd.bBoxMax[axis] - nd.bBoxMin[axis]) axis = k;
  std::stable_sort(order.begin() + first, order.begin() + first + count, [&](uint32_t a, uint32_t b) {
    if (lights[a].pos[axis] != lights[b].pos[axis]) return lights[a].pos[axis] < lights[b].pos[axis];
    return a < b;
  });
  const uint32_t mid = count / 2, left = (uint32_t)nodes.size();
  nd.childIndex = (int32_t)left;
  nodes[node] = nd;
  nodes.resize(nodes.size() + 2);
  light_bvh_build_node(lights, order, nodes, left, first, mid, directional);
  light_bvh_build_node(lights, order, nodes, left + 1, first + mid, count - mid, directional);
}
void light_bvh_build(orc_context* c) {
  c->light_bvh.clear();
  if (c->lights.empty()) return;
  const uint32_t n = (uint32_t)c->lights.size();
  std::vector<uint32_t> order;
  for (uint32_t i = 0; i < n; ++i)
    if (c->lights[i].type == BRT_LIGHT_POINT) order.push_back(i);
  const uint32_t n_point = (uint32_t)order.size();
  for (uint32_t i = 0; i < n; ++i)
    if (c->lights[i].type != BRT_LIGHT_POINT) order.push_back(i);
  c->light_bvh.resize(1);
  if (n_point == 0 || n_point == n) {
    light_bvh_build_node(c->lights, order, c->light_bvh, 0, 0, n, n_point == 0);
    return;
  }
  c->light_bvh.resize(3);  // the root separates the kinds: left = point lights, right = constant-direction lights
  light_bvh_build_node(c->lights, order, c->light_bvh, 1, 0, n_point, false);
  light_bvh_build_node(c->lights, order, c->light_bvh, 2, n_point, n - n_point, true);
  brt_light_bvh_node root{};
  for (int k = 0; k < 3; ++k) {
    root.bBoxMin[k] = std::fmin(c->light_bvh[1].bBoxMin[k], c->light_bvh[2].bBoxMin[k]);
    root.bBoxMax[k] = std::fmax(c->light_bvh[1].bBoxMax[k], c->light_bvh[2].bBoxMax[k]);
  }
  root.totalFlux = c->light_bvh[1].totalFlux + c->light_bvh[2].totalFlux;
  root.coneAxis[2] = 1.0f;
  root.coneAngle = 3.14159274f;
  root.childIndex = 1;
  c->light_bvh[0] = root;
}
float light_bvh_importance(const brt_light_bvh_node& n, vec3 P) {
  if (n.coneAngle == 0.0f) return n.totalFlux;  // constant-direction lights: no falloff
  vec3 lo = V3(n.bBoxMin[0], n.bBoxMin[1], n.bBoxMin[2]), hi = V3(n.bBoxMax[0], n.bBoxMax[1], n.bBoxMax[2]);
  vec3 ctr = (lo + hi) * 0.5f, hd = (hi - lo) * 0.5f;
  vec3 d = P - ctr;
  return n.totalFlux / std::fmax(std::fmax(dot(d, d), dot(hd, hd)), 1e-12f);
}
// one light by stochastic descent; r is rescaled at every level; returns the light index and 1 / probability
uint32_t light_bvh_sample(const orc_context* c, vec3 P, float r, float& inv_pdf) {
  uint32_t node = 0;
  float pdf = 1.0f;
  for (int guard = 0; guard < 64; ++guard) {
    int child = c->light_bvh[node].childIndex;
    if (child < 0) break;
    float w0 = light_bvh_importance(c->light_bvh[child], P), w1 = light_bvh_importance(c->light_bvh[child + 1], P);
    float sum = w0 + w1;
    float p0 = sum > 0.0f ? w0 / sum : 0.5f;
    if (r < p0 || p0 >= 1.0f) {
      node = (uint32_t)child;
      pdf = pdf * p0;
      r = r / p0;
    } else {
      float q = 1.0f - p0;
      node = (uint32_t)child + 1u;
      pdf = pdf * q;
      r = (r - p0) / q;
    }
  }
  inv_pdf = 1.0f / pdf;
  return (uint32_t)(-1 - c->light_bvh[node].childIndex);
}

// one light of calculateColor's loop body (SH/raytracing.slang:77-84); scale = 1 in the loop, 1 / pdf for a sampled light
vec3 light_contribution(const orc_context* c, const brt_light& l, float scale, bool scaled, const Material& mat, vec3 normal, vec3 view, vec3 worldPos,
                        Counters& k) {
  vec3 ldir;
  float intensity = l.intensity;
  if (l.type == BRT_LIGHT_POINT) {
    ldir = V3(l.pos[0], l.pos[1], l.pos[2]) - worldPos;
    float d = length(ldir);
    intensity /= (d * d);
  } else {
    ldir = V3(0.9f, -0.1f, 0.0f);
  }
  if (intensity < kLIGHT_TRESHOLD) return V3(0.0f);
  vec3 L = normalize(ldir);
  vec3 color = BRDF(mat, normal, view, L);
  vec3 contrib = color * V3(l.color[0], l.color[1], l.color[2]) * intensity;
  if (scaled) contrib = contrib * scale;
  float shadow = 1.0f;
  if (!is_zero(contrib)) {
    Ray sr;
    sr.o = worldPos + normal * 0.0001f;
    sr.d = normalize(ldir);
    sr.tmin = 0.001f;
    sr.tmax = length(ldir);
    shadow = trace_occluded(c, sr, k) ? 0.0f : 1.0f;
  }
  return contrib * shadow;
}

vec3 calculate_color(const orc_context* c, const Material& mat, vec3 normal, vec3 view, vec3 worldPos, Counters& k, bool light_bvh, uint32_t& seed) {
  vec3 acc = V3(0.0f);
  if (light_bvh) {  // SH/raytracing.slang:76: "later be replaced with a light bounding volume hierarchy system"
    if (c->lights.empty()) return acc;
    float inv_pdf;
    uint32_t li = light_bvh_sample(c, worldPos, rnd(seed), inv_pdf);
    return acc + light_contribution(c, c->lights[li], inv_pdf, true, mat, normal, view, worldPos, k);
  }
  for (size_t i = 0; i < c->lights.size(); ++i) acc = acc + light_contribution(c, c->lights[i], 1.0f, false, mat, normal, view, worldPos, k);
  return acc;
}

struct PixelOut {
  vec3 c;
};

void render_pixel(orc_context* c, const brt_uniform& u, const brt_render_opts& o, uint32_t px, uint32_t py, Counters& k, float* rgba) {
  const float* Pi = u.projInverse;
  const float* Vi = u.viewInverse;
  vec3 csum = V3(0.0f);
  const bool any_bounce = (o.flags & (BRT_RENDER_BOUNCE_REFLECT | BRT_RENDER_BOUNCE_REFRACT | BRT_RENDER_BOUNCE_DIFFUSE)) != 0;
  for (uint32_t s = 0; s < o.spp; ++s) {
    uint32_t frame = u.frame + s;
    uint32_t seed = hash3(px, py, frame);  // SH/raytracing.slang:96
    float jx = 0.0f, jy = 0.0f;
    if (o.flags & BRT_RENDER_JITTER) {  // :97-98 (computed by the shader, then dropped at :100)
      if (frame == 0) { jx = 0.5f; jy = 0.5f; }
      else { jx = rnd(seed); jy = rnd(seed); }
    }
    // :100-107
    float cx = ((float)px + jx) / (float)o.width * 2.0f - 1.0f;
    float cy = ((float)py + jy) / (float)o.height * 2.0f - 1.0f;
    vec3 vc = V3(((Pi[0] * cx + Pi[1] * cy) + Pi[2] * 1.0f) + Pi[3] * 1.0f, ((Pi[4] * cx + Pi[5] * cy) + Pi[6] * 1.0f) + Pi[7] * 1.0f,
                 ((Pi[8] * cx + Pi[9] * cy) + Pi[10] * 1.0f) + Pi[11] * 1.0f);
    vec3 dn = normalize(vc);
    Ray ray;
    ray.o = V3(Vi[3], Vi[7], Vi[11]);
    ray.d = V3((Vi[0] * dn.x + Vi[1] * dn.y) + Vi[2] * dn.z, (Vi[4] * dn.x + Vi[5] * dn.y) + Vi[6] * dn.z, (Vi[8] * dn.x + Vi[9] * dn.y) + Vi[10] * dn.z);
    ray.tmin = 0.001f;
    ray.tmax = kINFINITE;
    vec3 weight = V3(1.0f);  // HitPayload.weight, widened to RGB for the diffuse-bounce extension
    uint32_t depth = 0;
    while (depth < u.depthMax) {  // :119-126
      vec3 prevWeight = weight;
      Hit h = trace_closest(c, ray, k);
      if (s == 0 && depth == 0) {
        size_t pi = (size_t)py * o.width + px;
        c->aov_prim[pi] = h.hit ? h.prim : BRT_AOV_MISS;
        c->aov_inst[pi] = h.hit ? h.inst : BRT_AOV_MISS;
        c->aov_t[pi] = h.hit ? h.t : 0.0f;
      }
      if (!h.hit) {  // rmissMain :172-176
        if (o.flags & BRT_RENDER_SKY) csum = csum + sky_color(c->sky, ray.d) * prevWeight;
        break;
      }
      // rchitMain :136-169
      const Instance& in = c->inst[h.inst];
      const Mesh& m = c->meshes[in.mesh];  // intent of SURVEY A.7.2: the hit instance's mesh
      vec3 pos, nrm;
      if (m.sphere) {
        vec3 oo = xform_point(in.w2o, ray.o), od = xform_dir(in.w2o, ray.d);
        pos = oo + od * h.t;
        nrm = (pos - m.c) * (1.0f / m.r);
      } else {
        float b0 = 1.0f - h.u - h.v, b1 = h.u, b2 = h.v;  // :137
        const brt_vertex& v0 = m.v[m.idx[3 * h.prim]];
        const brt_vertex& v1 = m.v[m.idx[3 * h.prim + 1]];
        const brt_vertex& v2 = m.v[m.idx[3 * h.prim + 2]];
        pos = (b0 * vpos(v0) + b1 * vpos(v1)) + b2 * vpos(v2);  // SH/objects.slang:35-41
        nrm = (b0 * vnrm(v0) + b1 * vnrm(v1)) + b2 * vnrm(v2);
      }
      vec3 worldPos = xform_point(in.o2w, pos);               // :149
      vec3 worldNormal = normalize(xform_normal(in.w2o, nrm));  // :150
      const Material& mat = c->mats[in.mat];
      vec3 N = normalize(worldNormal);  // :154
      vec3 V = ray.d;                   // :155
      bool flipped = false;
      if (dot(N, -V) < 0.0f) { N = -N; flipped = true; }  // :157-158
      if (s == 0 && depth == 0 && c->has_gbuffer) {
        size_t pi = (size_t)py * o.width + px;
        c->aov_pos[pi] = Px4{worldPos.x, worldPos.y, worldPos.z, 1.0f};
        c->aov_nrm[pi] = Px4{N.x, N.y, N.z, h.t};
      }
      vec3 color = calculate_color(c, mat, N, -V, worldPos, k, (o.flags & BRT_RENDER_LIGHT_BVH) != 0, seed);  // :160
      depth++;                                                   // :164
      csum = csum + color * prevWeight;                          // :122
      // bounce (SH/raytracing.slang:161-168). Verbatim: weight = 0 (indirect illumination deactivated).
      if (!any_bounce) { weight = V3(0.0f); break; }
      float r1 = rnd(seed), r2 = rnd(seed), r3 = rnd(seed);
      const MatExt& ext = c->mat_ext[in.mat];
      vec3 ndir = V3(0.0f);
      vec3 norg = worldPos + N * 0.001f;  // :165
      if ((o.flags & BRT_RENDER_BOUNCE_REFRACT) && ext.transmission > 0.0f) {
        // extension: smooth dielectric, Schlick Fresnel, stochastic reflect/refract
        float eta = flipped ? ext.ior : 1.0f / ext.ior;
        float cosi = dot(N, -V);
        float k2 = 1.0f - eta * eta * (1.0f - cosi * cosi);
        float f0 = square((1.0f - ext.ior) / (1.0f + ext.ior));
        float Fr = k2 < 0.0f ? 1.0f : schlickFresnel(f0, cosi);
        if (r3 < Fr) {
          ndir = reflect(V, N);
        } else {
          ndir = V * eta + N * (eta * cosi - std::sqrt(k2));
          norg = worldPos - N * 0.001f;
          weight = weight * (mat.color * ext.transmission);
        }
      } else {
        bool refl = (o.flags & BRT_RENDER_BOUNCE_REFLECT) != 0, diff = (o.flags & BRT_RENDER_BOUNCE_DIFFUSE) != 0;
        bool spec;
        if (refl && diff) spec = r3 < mat.metallic;
        else spec = refl;
        if (spec) {
          float pdf;
          ndir = sampleGGXVNDFSphericalCap(mat, V, N, vec2{r1, r2}, pdf);  // :166
          weight = weight * (diff ? pdf : mat.metallic * pdf);             // :167 (lobe probability cancels metallic)
        } else if (diff) {
          float pdf;
          ndir = toWorld(sampleCosineWeightedHemisphere(vec2{r1, r2}, pdf), N);  // SURVEY A.7.7 intent
          weight = weight * (refl ? mat.color : mat.color * (1.0f - mat.metallic));
        } else {
          weight = V3(0.0f);
        }
      }
      if (!(weight.x > 0.0f || weight.y > 0.0f || weight.z > 0.0f)) break;  // stop instead of tracing a weight-0 ray
      ray.o = norg;
      ray.d = ndir;
    }
  }
  vec3 out = V3(csum.x / (float)o.spp, csum.y / (float)o.spp, csum.z / (float)o.spp);  // :129
  float* p = rgba + 4 * ((size_t)py * o.width + px);
  p[0] = out.x; p[1] = out.y; p[2] = out.z; p[3] = 1.0f;  // :132
}

// 4x4 inverse in double (cofactors), row-major in/out
void invert4x4(const double m[16], double inv[16]) {
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  double id = 1.0 / det;
  for (int i = 0; i < 16; ++i) inv[i] *= id;
}

bool owns_tile(const orc_context* c, uint32_t px, uint32_t py, uint32_t width) {
  if (c->tile_world <= 1) return true;
  uint32_t tiles_x = (width + 31) / 32;
  uint32_t tile = (py / 32) * tiles_x + (px / 32);
  // groups of `world` consecutive tiles, ranks rotated by the group index (csrc/render_kernels.cuh: tile_of_rank)
  return (tile % c->tile_world + c->tile_world - (tile / c->tile_world) % c->tile_world) % c->tile_world == c->tile_rank;
}

}  // namespace

extern "C" {

int orc_create(const brt_config* cfg, orc_context** out) {
  if (!out) return BRT_ERR_INVALID;
  orc_context* c = new orc_context();
  if (cfg) {
    c->brute = (cfg->flags & ORC_CFG_BRUTE_FORCE) != 0;
    c->tile_rank = cfg->tile_rank;
    c->tile_world = cfg->tile_world ? cfg->tile_world : 1;
  }
  c->threads = std::thread::hardware_concurrency();
  if (!c->threads) c->threads = 1;
  *out = c;
  return BRT_OK;
}
void orc_destroy(orc_context* c) { delete c; }
const char* orc_last_error(const orc_context* c) { return c ? c->err.c_str() : "null context"; }
int orc_set_threads(orc_context* c, uint32_t n) {
  c->threads = n ? n : std::max(1u, std::thread::hardware_concurrency());
  return BRT_OK;
}
uint32_t orc_get_threads(const orc_context* c) { return c->threads; }

int orc_mesh_create(orc_context* c, const brt_vertex* v, uint32_t nv, const uint32_t* idx, uint32_t ni, uint32_t* id) {
  if (!c || (!v && nv) || (!idx && ni) || ni % 3) return fail(c, BRT_ERR_INVALID, "mesh_create: bad arguments");
  for (uint32_t i = 0; i < ni; ++i)
    if (idx[i] >= nv) return fail(c, BRT_ERR_INVALID, "mesh_create: index out of range");
  Mesh m;
  m.v.assign(v, v + nv);
  m.idx.assign(idx, idx + ni);
  c->meshes.push_back(std::move(m));
  if (id) *id = (uint32_t)c->meshes.size() - 1;
  c->built = false;
  return BRT_OK;
}
int orc_mesh_update_vertices(orc_context* c, uint32_t mesh, const brt_vertex* v, uint32_t nv) {
  if (!c || mesh >= c->meshes.size() || c->meshes[mesh].sphere || nv != c->meshes[mesh].v.size())
    return fail(c, BRT_ERR_INVALID, "mesh_update_vertices: bad mesh id or vertex count");
  c->meshes[mesh].v.assign(v, v + nv);
  c->meshes[mesh].dirty = true;
  c->built = false;
  return BRT_OK;
}
int orc_sphere_create(orc_context* c, const float center[3], float radius, uint32_t* id) {
  if (!c || !center || !(radius > 0.0f)) return fail(c, BRT_ERR_INVALID, "sphere_create: bad arguments");
  Mesh m;
  m.sphere = true;
  m.c = V3(center[0], center[1], center[2]);
  m.r = radius;
  c->meshes.push_back(std::move(m));
  if (id) *id = (uint32_t)c->meshes.size() - 1;
  c->built = false;
  return BRT_OK;
}
int orc_material_create(orc_context* c, const brt_material* m, uint32_t* id) {
  if (!c || !m) return fail(c, BRT_ERR_INVALID, "material_create: null");
  Material mm;
  static_assert(sizeof(Material) == sizeof(brt_material), "material layout");
  std::memcpy(&mm, m, sizeof(mm));
  c->mats.push_back(mm);
  c->mat_ext.push_back(MatExt{});
  if (id) *id = (uint32_t)c->mats.size() - 1;
  return BRT_OK;
}
int orc_material_set_transmission(orc_context* c, uint32_t id, float tr, float ior) {
  if (!c || id >= c->mats.size()) return fail(c, BRT_ERR_INVALID, "material_set_transmission: bad id");
  c->mat_ext[id].transmission = tr;
  c->mat_ext[id].ior = ior;
  return BRT_OK;
}
int orc_light_create(orc_context* c, const brt_light* l, uint32_t* id) {
  if (!c || !l) return fail(c, BRT_ERR_INVALID, "light_create: null");
  c->lights.push_back(*l);
  if (id) *id = (uint32_t)c->lights.size() - 1;
  return BRT_OK;
}
int orc_sky_set(orc_context* c, const brt_sky* s) {
  if (!c || !s) return fail(c, BRT_ERR_INVALID, "sky_set: null");
  c->sky = *s;
  return BRT_OK;
}
int orc_instance_create(orc_context* c, uint32_t mesh, uint32_t mat, const float x[12], uint32_t* id) {
  if (!c || !x || mesh >= c->meshes.size() || mat >= c->mats.size()) return fail(c, BRT_ERR_INVALID, "instance_create: bad mesh/material id");
  Instance in;
  in.mesh = mesh;
  in.mat = mat;
  std::memcpy(in.o2w, x, sizeof(in.o2w));
  invert3x4(in.o2w, in.w2o);
  c->inst.push_back(in);
  if (id) *id = (uint32_t)c->inst.size() - 1;
  c->built = false;
  return BRT_OK;
}
int orc_instance_set_transform(orc_context* c, uint32_t id, const float x[12]) {
  if (!c || !x || id >= c->inst.size()) return fail(c, BRT_ERR_INVALID, "instance_set_transform: bad id");
  std::memcpy(c->inst[id].o2w, x, 48);
  invert3x4(c->inst[id].o2w, c->inst[id].w2o);
  c->built = false;
  return BRT_OK;
}
int orc_instance_set_material(orc_context* c, uint32_t id, uint32_t mat) {
  if (!c || id >= c->inst.size() || mat >= c->mats.size()) return fail(c, BRT_ERR_INVALID, "instance_set_material: bad id");
  c->inst[id].mat = mat;
  return BRT_OK;
}
int orc_instance_destroy(orc_context* c, uint32_t id) {  // RT/Scene.cpp:122-125 swap-remove
  if (!c || id >= c->inst.size()) return fail(c, BRT_ERR_INVALID, "instance_destroy: bad id");
  c->inst[id] = c->inst.back();
  c->inst.pop_back();
  c->built = false;
  return BRT_OK;
}
int orc_scene_build(orc_context* c) {
  if (!c) return BRT_ERR_INVALID;
  uint32_t rebuilt = 0;
  for (Mesh& m : c->meshes)
    if (m.dirty) { build_mesh(m); rebuilt++; }
  build_tlas(c);
  c->built = true;
  c->stats.blas_built = rebuilt;
  uint64_t tris = 0;
  for (const Instance& in : c->inst) tris += c->meshes[in.mesh].idx.size() / 3;
  c->stats.total_triangles = tris;
  c->stats.instances_total = (uint32_t)c->inst.size();
  c->stats.instances_visible = (uint32_t)c->tlas_ids.size();
  return BRT_OK;
}

// README.md:15-18 Smart Culling; rule defined in DESIGN.md §6 (no reference code exists).
int orc_smart_cull(orc_context* c, const brt_uniform* u, uint32_t w, uint32_t h, float thr, float hyst, uint32_t* visible) {
  if (!c || !u) return fail(c, BRT_ERR_INVALID, "smart_cull: null");
  (void)w;
  for (Mesh& m : c->meshes)
    if (m.dirty) build_mesh(m);
  const float* Vi = u->viewInverse;
  vec3 eye = V3(Vi[3], Vi[7], Vi[11]);
  vec3 fwd = V3(Vi[2], Vi[6], Vi[10]);
  float p11 = 1.0f / u->projInverse[5];
  float k = p11 * ((float)h * 0.5f);
  for (Instance& in : c->inst) {
    if (!(thr > 0.0f)) { in.visible = true; continue; }
    instance_world_bounds(c->meshes[in.mesh], in);
    vec3 ctr = (in.wlo + in.whi) * 0.5f;
    float r = length((in.whi - in.wlo) * 0.5f);
    float z = dot(ctr - eye, fwd);
    if (z - r <= 0.0f) { in.visible = true; continue; }  // touches the camera plane: keep
    float rp = (r * k) / z;
    float fp = kPI * (rp * rp);
    in.visible = in.visible ? (fp >= thr * (1.0f - hyst)) : (fp > thr * (1.0f + hyst));
  }
  build_tlas(c);
  c->built = true;
  c->stats.instances_total = (uint32_t)c->inst.size();
  c->stats.instances_visible = (uint32_t)c->tlas_ids.size();
  if (visible) {
    uint32_t n = 0;
    for (const Instance& in : c->inst) n += in.visible ? 1u : 0u;
    *visible = n;
  }
  return BRT_OK;
}
int orc_get_visibility(orc_context* c, uint8_t* out, uint32_t n) {
  if (!c || !out || n != c->inst.size()) return fail(c, BRT_ERR_INVALID, "get_visibility: size mismatch");
  for (uint32_t i = 0; i < n; ++i) out[i] = c->inst[i].visible ? 1 : 0;
  return BRT_OK;
}

// rebuildRenderOutput(format, extent) + copyImageToSwapchain (RT/RTPipeline.cpp:49-55, RT/RTApp.cpp:87-152): the linear frame
// stored in an 8-bit swapchain format, one packed texel per pixel
static void present_convert(const float* rgba, size_t npx, uint32_t format, uint8_t* out) {
  const bool srgb = format == BRT_FORMAT_R8G8B8A8_SRGB || format == BRT_FORMAT_B8G8R8A8_SRGB;
  const bool bgra = format == BRT_FORMAT_B8G8R8A8_UNORM || format == BRT_FORMAT_B8G8R8A8_SRGB;
  for (size_t i = 0; i < npx; ++i) {
    const float* px = rgba + 4 * i;
    uint32_t ch[3];
    for (int k = 0; k < 3; ++k) ch[k] = float_to_unorm8(srgb ? linear_to_srgb(px[k]) : px[k]);
    out[4 * i + 0] = (uint8_t)(bgra ? ch[2] : ch[0]);
    out[4 * i + 1] = (uint8_t)ch[1];
    out[4 * i + 2] = (uint8_t)(bgra ? ch[0] : ch[2]);
    out[4 * i + 3] = (uint8_t)float_to_unorm8(px[3]);
  }
}

int orc_render_frame(orc_context* c, const brt_uniform* u, const brt_render_opts* o, float* rgba_out) {
  if (!c || !u || !o || !rgba_out || !o->width || !o->height || !o->spp) return fail(c, BRT_ERR_INVALID, "render_frame: bad arguments");
  if (!c->built) return fail(c, BRT_ERR_STATE, "render_frame: scene not built");
  size_t npx = (size_t)o->width * o->height;
  const uint32_t format = (o->flags & BRT_RENDER_FORMAT_MASK) >> BRT_RENDER_FORMAT_SHIFT;
  if (format > BRT_FORMAT_B8G8R8A8_SRGB) return fail(c, BRT_ERR_INVALID, "render_frame: unknown BRT_RENDER_FORMAT");
  if (o->flags & BRT_RENDER_LIGHT_BVH) light_bvh_build(c);
  std::vector<float> linear;
  float* rgba = rgba_out;
  if (format != BRT_FORMAT_R32G32B32A32_SFLOAT) {
    linear.resize(npx * 4);
    rgba = linear.data();
  }
  c->fw = o->width;
  c->fh = o->height;
  c->aov_prim.assign(npx, BRT_AOV_MISS);
  c->aov_inst.assign(npx, BRT_AOV_MISS);
  c->aov_t.assign(npx, 0.0f);
  c->has_gbuffer = (o->flags & BRT_RENDER_GBUFFER) != 0;
  c->render_flags = o->flags;
  if (c->has_gbuffer) {
    c->aov_pos.assign(npx, Px4{0, 0, 0, 0});
    c->aov_nrm.assign(npx, Px4{0, 0, 0, 0});
  }
  std::memset(rgba, 0, npx * 16);
  uint32_t x0 = 0, y0 = 0, x1 = o->width, y1 = o->height;
  if (o->crop_w) {
    x0 = o->crop_x0; y0 = o->crop_y0;
    x1 = std::min(o->width, x0 + o->crop_w);
    y1 = std::min(o->height, y0 + o->crop_h);
  }
  std::atomic<uint32_t> next{y0};
  std::vector<Counters> ks(c->threads);
  auto worker = [&](uint32_t tid) {
    Counters& k = ks[tid];
    for (;;) {
      uint32_t y = next.fetch_add(1);
      if (y >= y1) break;
      for (uint32_t x = x0; x < x1; ++x)
        if (owns_tile(c, x, y, o->width)) render_pixel(c, *u, *o, x, y, k, rgba);
    }
  };
  std::vector<std::thread> th;
  for (uint32_t t = 1; t < c->threads; ++t) th.emplace_back(worker, t);
  worker(0);
  for (auto& t : th) t.join();
  Counters s;
  for (const Counters& k : ks) {
    s.rays_closest += k.rays_closest; s.rays_occl += k.rays_occl;
    s.nodes_c += k.nodes_c; s.prims_c += k.prims_c; s.sph_c += k.sph_c;
    s.nodes_o += k.nodes_o; s.prims_o += k.prims_o; s.sph_o += k.sph_o;
  }
  c->stats.rays_closest = s.rays_closest;
  c->stats.rays_occlusion = s.rays_occl;
  c->stats.nodes_visited_closest = s.nodes_c;
  c->stats.prims_tested_closest = s.prims_c;
  c->stats.spheres_tested_closest = s.sph_c;
  c->stats.nodes_visited_occlusion = s.nodes_o;
  c->stats.prims_tested_occlusion = s.prims_o;
  c->stats.spheres_tested_occlusion = s.sph_o;
  c->last_image.resize(npx);
  std::memcpy(c->last_image.data(), rgba, npx * 16);
  if (format != BRT_FORMAT_R32G32B32A32_SFLOAT) present_convert(rgba, npx, format, reinterpret_cast<uint8_t*>(rgba_out));
  return BRT_OK;
}

int orc_get_aov(orc_context* c, int kind, void* out) {
  if (!c || !out || !c->fw) return fail(c, BRT_ERR_STATE, "get_aov: no frame rendered");
  size_t n = (size_t)c->fw * c->fh;
  switch (kind) {
    case BRT_AOV_PRIM_ID: std::memcpy(out, c->aov_prim.data(), n * 4); break;
    case BRT_AOV_INST_ID: std::memcpy(out, c->aov_inst.data(), n * 4); break;
    case BRT_AOV_HIT_T: std::memcpy(out, c->aov_t.data(), n * 4); break;
    case BRT_AOV_POSITION:
    case BRT_AOV_NORMAL:
      if (!c->has_gbuffer) return fail(c, BRT_ERR_STATE, "get_aov: the frame was not rendered with BRT_RENDER_GBUFFER");
      std::memcpy(out, kind == BRT_AOV_POSITION ? (const void*)c->aov_pos.data() : (const void*)c->aov_nrm.data(), n * 16);
      break;
    default: return fail(c, BRT_ERR_INVALID, "get_aov: bad kind");
  }
  return BRT_OK;
}
// Extensions::Denoiser::denoise (Graphics/Denoiser/Denoiser.h:5-20), stages of oracle/denoise.hpp on the frame rendered last
int orc_denoise(orc_context* c, const brt_uniform* u, const brt_denoise_opts* d, float* rgba) {
  if (!c || !u || !d || d->struct_size != sizeof(brt_denoise_opts)) return fail(c, BRT_ERR_INVALID, "denoise: null / struct_size mismatch");
  if (d->iterations > 6 || d->sigma_n_log2 > 8) return fail(c, BRT_ERR_INVALID, "denoise: iterations <= 6, sigma_n_log2 <= 8");
  if (!c->fw || !c->has_gbuffer) return fail(c, BRT_ERR_STATE, "denoise: render the frame with BRT_RENDER_GBUFFER first");
  const uint32_t W = c->fw, H = c->fh;
  const size_t npx = (size_t)W * H;
  if (c->dn.w != W || c->dn.h != H || (d->flags & BRT_DENOISE_RESET)) c->dn.have = false;
  c->dn.w = W;
  c->dn.h = H;
  // world -> clip of this frame = (V^-1 P^-1)^-1 (the uniform carries the inverses, RT/RTApp.cpp:44-49)
  double vi[16], pi[16], a[16], vp[16];
  for (int i = 0; i < 16; ++i) { vi[i] = u->viewInverse[i]; pi[i] = u->projInverse[i]; }
  for (int r = 0; r < 4; ++r)
    for (int k = 0; k < 4; ++k) {
      double acc = 0.0;
      for (int j = 0; j < 4; ++j) acc += vi[4 * r + j] * pi[4 * j + k];
      a[4 * r + k] = acc;
    }
  invert4x4(a, vp);
  DenoiseState next;
  std::vector<Px4> work[2];
  dn_temporal(c->last_image.data(), c->aov_pos.data(), c->aov_nrm.data(), c->aov_inst.data(), c->dn, W, H, d->clamp_gamma,
              d->max_history >= 1.0f ? d->max_history : 1.0f, (c->render_flags & BRT_RENDER_JITTER) ? 0.0f : 0.5f, work[0], next);
  int cur = 0;
  const uint32_t passes = d->iterations + ((d->flags & BRT_DENOISE_BILATERAL) ? 1u : 0u);
  for (uint32_t it = 0; it < passes; ++it) {
    const bool bilateral = it >= d->iterations;
    dn_atrous(work[cur], c->aov_nrm.data(), c->aov_inst.data(), W, H, bilateral ? 1 : (1 << it), bilateral ? 1 : 2, d->sigma_z,
              d->sigma_l, d->sigma_n_log2, it + 1 == passes, work[cur ^ 1]);
    cur ^= 1;
  }
  if (passes == 0) {
    dn_atrous(work[0], c->aov_nrm.data(), c->aov_inst.data(), W, H, 1, 0, d->sigma_z, d->sigma_l, 0, true, work[1]);
    cur = 1;
  }
  next.w = W;
  next.h = H;
  next.have = true;
  next.nrm = c->aov_nrm;
  next.inst = c->aov_inst;
  next.t = c->aov_t;
  for (int i = 0; i < 16; ++i) next.vp[i] = (float)vp[i];
  next.eye[0] = u->viewInverse[3];
  next.eye[1] = u->viewInverse[7];
  next.eye[2] = u->viewInverse[11];
  c->dn = std::move(next);
  if (rgba) std::memcpy(rgba, work[cur].data(), npx * 16);
  return BRT_OK;
}

int orc_get_light_bvh(orc_context* c, brt_light_bvh_node* out, uint32_t max_nodes, uint32_t* n_nodes) {
  if (!c || !n_nodes) return fail(c, BRT_ERR_INVALID, "get_light_bvh: null");
  if (!c->built) return fail(c, BRT_ERR_STATE, "get_light_bvh: scene not built");
  light_bvh_build(c);
  *n_nodes = (uint32_t)c->light_bvh.size();
  if (out) std::memcpy(out, c->light_bvh.data(), std::min<size_t>(max_nodes, c->light_bvh.size()) * sizeof(brt_light_bvh_node));
  return BRT_OK;
}

int orc_get_stats(orc_context* c, brt_stats* out) {
  if (!c || !out) return BRT_ERR_INVALID;
  *out = c->stats;
  return BRT_OK;
}
int orc_trace_rays(orc_context* c, const float* rays, uint32_t n, int closest, uint32_t* out) {
  if (!c || !rays || !out) return fail(c, BRT_ERR_INVALID, "trace_rays: null");
  if (!c->built) return fail(c, BRT_ERR_STATE, "trace_rays: scene not built");
  std::atomic<uint32_t> next{0};
  auto worker = [&]() {
    Counters k;
    for (;;) {
      uint32_t b = next.fetch_add(256);
      if (b >= n) break;
      for (uint32_t i = b; i < std::min(n, b + 256); ++i) {
        Ray r;
        r.o = V3(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2]);
        r.tmin = rays[8 * i + 3];
        r.d = V3(rays[8 * i + 4], rays[8 * i + 5], rays[8 * i + 6]);
        r.tmax = rays[8 * i + 7];
        if (closest) {
          Hit h = trace_closest(c, r, k);
          std::memcpy(&out[4 * i], &h.t, 4);
          out[4 * i + 1] = h.prim;
          out[4 * i + 2] = h.inst;
          out[4 * i + 3] = h.hit ? 1u : 0u;
          if (!h.hit) out[4 * i] = 0;
        } else {
          out[4 * i] = out[4 * i + 1] = out[4 * i + 2] = 0;
          out[4 * i + 3] = trace_occluded(c, r, k) ? 1u : 0u;
        }
      }
    }
  };
  std::vector<std::thread> th;
  for (uint32_t t = 1; t < c->threads; ++t) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
  return BRT_OK;
}

int orc_debug_sort_pairs(orc_context*, uint32_t* keys, uint32_t* vals, uint32_t n, int bits) {
  std::vector<uint32_t> order(n);
  for (uint32_t i = 0; i < n; ++i) order[i] = i;
  const uint32_t mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return (keys[a] & mask) < (keys[b] & mask); });
  std::vector<uint32_t> k(n), v(n);
  for (uint32_t i = 0; i < n; ++i) { k[i] = keys[order[i]]; v[i] = vals[order[i]]; }
  std::copy(k.begin(), k.end(), keys);
  std::copy(v.begin(), v.end(), vals);
  return BRT_OK;
}

uint32_t orc_kat_hash(uint32_t x, uint32_t y, uint32_t z) { return hash3(x, y, z); }
uint32_t orc_kat_pcg(uint32_t* s) { return pcg(*s); }
float orc_kat_rand(uint32_t* s) { return rnd(*s); }
void orc_kat_brdf(const brt_material* m, const float N[3], const float V[3], const float L[3], float out[3]) {
  Material mm;
  std::memcpy(&mm, m, sizeof(mm));
  vec3 r = BRDF(mm, V3(N[0], N[1], N[2]), V3(V[0], V[1], V[2]), V3(L[0], L[1], L[2]));
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_kat_sample_vndf(const brt_material* m, const float V[3], const float N[3], float r1, float r2, float out[3], float* pdf) {
  Material mm;
  std::memcpy(&mm, m, sizeof(mm));
  vec3 r = sampleGGXVNDFSphericalCap(mm, V3(V[0], V[1], V[2]), V3(N[0], N[1], N[2]), vec2{r1, r2}, *pdf);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_kat_sample_cosine(float r1, float r2, float out[3], float* pdf) {
  vec3 r = sampleCosineWeightedHemisphere(vec2{r1, r2}, *pdf);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_kat_sincos(float x, float* s, float* c) { det_sincos(x, s, c); }
float orc_kat_log2(float x) { return det_log2(x); }
float orc_kat_exp2(float x) { return det_exp2(x); }
float orc_kat_srgb(float x) { return linear_to_srgb(x); }
uint32_t orc_kat_unorm8(float x) { return float_to_unorm8(x); }
// probability with which the light BVH descent ends at each light for a shading point P: the products of the branch
// probabilities along every root-to-leaf path, computed by exhaustive recursion (test infrastructure for the sampler's pdf)
int orc_kat_light_pdfs(orc_context* c, const float P[3], float* out, uint32_t n) {
  if (!c || !out || n != c->lights.size()) return BRT_ERR_INVALID;
  light_bvh_build(c);
  for (uint32_t i = 0; i < n; ++i) out[i] = 0.0f;
  if (c->light_bvh.empty()) return BRT_OK;
  struct Item { uint32_t node; float p; };
  std::vector<Item> st{{0u, 1.0f}};
  vec3 p = V3(P[0], P[1], P[2]);
  while (!st.empty()) {
    Item it = st.back();
    st.pop_back();
    int child = c->light_bvh[it.node].childIndex;
    if (child < 0) { out[-1 - child] = it.p; continue; }
    float w0 = light_bvh_importance(c->light_bvh[child], p), w1 = light_bvh_importance(c->light_bvh[child + 1], p);
    float sum = w0 + w1;
    float p0 = sum > 0.0f ? w0 / sum : 0.5f;
    st.push_back({(uint32_t)child, it.p * p0});
    st.push_back({(uint32_t)child + 1u, it.p * (1.0f - p0)});
  }
  return BRT_OK;
}
// one draw of the sampler: light index and 1 / pdf for random number r
uint32_t orc_kat_light_sample(orc_context* c, const float P[3], float r, float* inv_pdf) {
  light_bvh_build(c);
  return light_bvh_sample(c, V3(P[0], P[1], P[2]), r, *inv_pdf);
}
int orc_kat_intersect_tri(const float o[3], const float d[3], float tmin, float tmax, const float v0[3], const float v1[3], const float v2[3], float tuv[3]) {
  RayShear s = make_shear(V3(d[0], d[1], d[2]));
  return intersect_tri(V3(o[0], o[1], o[2]), s, tmin, tmax, V3(v0[0], v0[1], v0[2]), V3(v1[0], v1[1], v1[2]), V3(v2[0], v2[1], v2[2]), tuv[0], tuv[1], tuv[2]) ? 1 : 0;
}

// Camera::handleInputs (Graphics/Camera.cpp:26-61), GLFW polling replaced by a key mask (bits as BRT_KEY_*)
void orc_camera_handle_inputs(uint32_t keys, float dt, float position[3], float rotation[3]) {
  vec3 rotate = V3(0.0f);
  if (keys & BRT_KEY_LOOK_RIGHT) rotate.y += 1.f;                                   // :29
  if (keys & BRT_KEY_LOOK_LEFT) rotate.y -= 1.f;                                    // :30
  if (keys & BRT_KEY_LOOK_UP) rotate.x += 1.f;                                      // :32
  if (keys & BRT_KEY_LOOK_DOWN) rotate.x -= 1.f;                                    // :33
  vec3 rot = V3(rotation[0], rotation[1], rotation[2]);
  if (dot(rotate, rotate) > std::numeric_limits<float>::epsilon())                  // :35-36
    rot = rot + (rotate * (1.0f / std::sqrt(dot(rotate, rotate)))) * (1.5f * dt);
  rot.x = std::fmin(std::fmax(rot.x, -1.5f), 1.5f);                                 // :38
  const float two_pi = 6.28318530717958647692f;
  rot.y = rot.y - two_pi * std::floor(rot.y / two_pi);                              // :39 glm::mod
  float yaw = rot.y;
  const vec3 forwardDir = V3(std::sin(yaw), 0.f, std::cos(yaw));                    // :42
  const vec3 rightDir = V3(forwardDir.z, 0.f, -forwardDir.x);                       // :43
  const vec3 upDir = V3(0.0f, -1.0f, 0.0f);                                         // :44
  vec3 moveDir = V3(0.0f);
  if (keys & BRT_KEY_MOVE_FORWARD) moveDir = moveDir + forwardDir;                  // :48-58
  if (keys & BRT_KEY_MOVE_BACKWARD) moveDir = moveDir - forwardDir;
  if (keys & BRT_KEY_MOVE_RIGHT) moveDir = moveDir + rightDir;
  if (keys & BRT_KEY_MOVE_LEFT) moveDir = moveDir - rightDir;
  if (keys & BRT_KEY_MOVE_UP) moveDir = moveDir + upDir;
  if (keys & BRT_KEY_MOVE_DOWN) moveDir = moveDir - upDir;
  vec3 pos = V3(position[0], position[1], position[2]);
  if (dot(moveDir, moveDir) > std::numeric_limits<float>::epsilon())                // :60-61
    pos = pos + (moveDir * (1.0f / std::sqrt(dot(moveDir, moveDir)))) * (3.f * dt);
  position[0] = pos.x; position[1] = pos.y; position[2] = pos.z;
  rotation[0] = rot.x; rotation[1] = rot.y; rotation[2] = rot.z;
}

void orc_camera_uniform(const float pos[3], const float rot[3], float fovy, float aspect, float znear, float zfar, uint32_t frame,
                        uint32_t depth_max, brt_uniform* out) {
  // Camera::updateView (Graphics/Camera.cpp:71-95): Tait-Bryan Y-X-Z, rows u, v, w
  const float c3 = std::cos(rot[2]), s3 = std::sin(rot[2]);
  const float c2 = std::cos(rot[0]), s2 = std::sin(rot[0]);
  const float c1 = std::cos(rot[1]), s1 = std::sin(rot[1]);
  const float ux = c1 * c3 + s1 * s2 * s3, uy = c2 * s3, uz = c1 * s2 * s3 - c3 * s1;
  const float vx = c3 * s1 * s2 - c1 * s3, vy = c2 * c3, vz = c1 * c3 * s2 + s1 * s3;
  const float wx = c2 * s1, wy = -s2, wz = c1 * c2;
  const float tu = -(ux * pos[0] + uy * pos[1] + uz * pos[2]);
  const float tv = -(vx * pos[0] + vy * pos[1] + vz * pos[2]);
  const float tw = -(wx * pos[0] + wy * pos[1] + wz * pos[2]);
  double view[16] = {ux, uy, uz, tu, vx, vy, vz, tv, wx, wy, wz, tw, 0, 0, 0, 1};
  // Camera::setPerspectiveProjection (Graphics/Camera.cpp:8-17), math row-major
  const float t = std::tan(fovy / 2.0f);
  double proj[16] = {0};
  proj[0] = 1.0f / (aspect * t);
  proj[5] = 1.0f / t;
  proj[10] = zfar / (zfar - znear);
  proj[14] = 1.0f;
  proj[11] = -(zfar * znear) / (zfar - znear);
  // RTApp::run (RT/RTApp.cpp:44-49): glm::inverse(glm::transpose(M)) stored column-major
  // == M^-1 stored row-major
  double vi[16], pi[16];
  invert4x4(view, vi);
  invert4x4(proj, pi);
  for (int i = 0; i < 16; ++i) {
    out->viewInverse[i] = (float)vi[i];
    out->projInverse[i] = (float)pi[i];
  }
  out->frame = frame;
  out->depthMax = depth_max;
  out->lightThreshold = 0.0001f;
}

}  // extern "C"
