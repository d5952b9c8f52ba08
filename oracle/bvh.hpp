// ORACLE — TEST INFRASTRUCTURE ONLY (see shaders.hpp header). PARITY UNPINNED for this file: nothing in the reference can pin it.
//
// The reference has no BVH or intersection source: both live in the Vulkan driver / RT cores behind
// vkCmdBuildAccelerationStructuresKHR (RT/Scene.cpp:304) and TraceRay (SH/raytracing.slang:67,121).
// This file supplies the CPU stand-ins the oracle needs: a watertight ray/triangle test
// (Woop, Benthin, Wald, "Watertight Ray/Triangle Intersection", JCGT 2013 — published algorithm,
// restated), an analytic ray/sphere test (extension) and a plain binned-SAH binary BVH with a
// conservative slab test. The BVH is deliberately a different structure from the product's
// compressed 8-wide LBVH so that agreement between the two is evidence, not tautology.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "shaders.hpp"

namespace orc {

struct Ray {
  vec3 o;
  float tmin;
  vec3 d;
  float tmax;
};

// Per-ray constants of the watertight test. kz = index of the largest |d| component (ties: lowest
// index), kx = (kz+1)%3, ky = (kz+2)%3; no winding swap (no face culling in this renderer:
// VK_GEOMETRY_INSTANCE_TRIANGLE_CULL_DISABLE, RT/Scene.cpp:188).
struct RayShear {
  int kx, ky, kz;
  float Sx, Sy, Sz;
};
static inline float comp(vec3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
static inline RayShear make_shear(vec3 d) {
  RayShear s;
  float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
  s.kz = 0;
  float m = ax;
  if (ay > m) { s.kz = 1; m = ay; }
  if (az > m) { s.kz = 2; }
  s.kx = (s.kz + 1) % 3;
  s.ky = (s.kz + 2) % 3;
  float dz = comp(d, s.kz);
  s.Sx = comp(d, s.kx) / dz;
  s.Sy = comp(d, s.ky) / dz;
  s.Sz = 1.0f / dz;
  return s;
}

// returns true and (t,u,v) when tmin < t < tmax; u,v = barycentric weights of v1, v2
// (BuiltInTriangleIntersectionAttributes.barycentrics, SH/raytracing.slang:137)
static inline bool intersect_tri(const vec3& o, const RayShear& s, float tmin, float tmax, vec3 v0, vec3 v1, vec3 v2,
                                 float& t_out, float& u_out, float& v_out) {
  vec3 A = v0 - o, B = v1 - o, C = v2 - o;
  float Akz = comp(A, s.kz), Bkz = comp(B, s.kz), Ckz = comp(C, s.kz);
  float Ax = comp(A, s.kx) - s.Sx * Akz, Ay = comp(A, s.ky) - s.Sy * Akz;
  float Bx = comp(B, s.kx) - s.Sx * Bkz, By = comp(B, s.ky) - s.Sy * Bkz;
  float Cx = comp(C, s.kx) - s.Sx * Ckz, Cy = comp(C, s.ky) - s.Sy * Ckz;
  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
    V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
    W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  float det = (U + V) + W;
  if (det == 0.0f) return false;
  float Az = s.Sz * Akz, Bz = s.Sz * Bkz, Cz = s.Sz * Ckz;
  float T = (U * Az + V * Bz) + W * Cz;
  float rcp = 1.0f / det;
  float t = T * rcp;
  if (!(t > tmin && t < tmax)) return false;
  t_out = t;
  u_out = V * rcp;
  v_out = W * rcp;
  return true;
}

// extension: object-space sphere. Picks the nearer root inside (tmin, tmax), else the farther one.
static inline bool intersect_sphere(vec3 o, vec3 d, float tmin, float tmax, vec3 c, float r, float& t_out) {
  vec3 oc = o - c;
  float a = dot(d, d);
  float b = dot(oc, d);
  float cc = dot(oc, oc) - r * r;
  float disc = b * b - a * cc;
  if (!(disc >= 0.0f)) return false;
  float sq = std::sqrt(disc);
  float t0 = (-b - sq) / a;
  float t1 = (-b + sq) / a;
  if (t0 > tmin && t0 < tmax) { t_out = t0; return true; }
  if (t1 > tmin && t1 < tmax) { t_out = t1; return true; }
  return false;
}

// ---- binned SAH binary BVH over arbitrary AABBs --------------------------------------------------
struct BNode {
  vec3 lo;
  uint32_t left;   // internal: index of left child (right = left + 1); leaf: first slot in `order`
  vec3 hi;
  uint32_t count;  // 0 = internal
};
struct BVH {
  std::vector<BNode> nodes;
  std::vector<uint32_t> order;
  bool empty() const { return nodes.empty(); }
};

static inline float half_area(vec3 lo, vec3 hi) {
  vec3 e = hi - lo;
  return e.x * e.y + e.y * e.z + e.z * e.x;
}

static inline void build_bvh(BVH& bvh, const std::vector<vec3>& plo, const std::vector<vec3>& phi, uint32_t max_leaf) {
  const uint32_t n = (uint32_t)plo.size();
  bvh.nodes.clear();
  bvh.order.resize(n);
  for (uint32_t i = 0; i < n; ++i) bvh.order[i] = i;
  if (n == 0) return;
  bvh.nodes.reserve(2 * (size_t)n);
  std::vector<vec3> cen(n);
  for (uint32_t i = 0; i < n; ++i) cen[i] = (plo[i] + phi[i]) * 0.5f;
  struct Task { uint32_t node, first, count; };
  std::vector<Task> stack;
  bvh.nodes.push_back(BNode{});
  stack.push_back({0, 0, n});
  const int NB = 16;
  while (!stack.empty()) {
    Task tk = stack.back();
    stack.pop_back();
    vec3 lo = V3(INFINITY), hi = V3(-INFINITY), clo = V3(INFINITY), chi = V3(-INFINITY);
    for (uint32_t i = tk.first; i < tk.first + tk.count; ++i) {
      uint32_t p = bvh.order[i];
      lo = V3(std::fmin(lo.x, plo[p].x), std::fmin(lo.y, plo[p].y), std::fmin(lo.z, plo[p].z));
      hi = V3(std::fmax(hi.x, phi[p].x), std::fmax(hi.y, phi[p].y), std::fmax(hi.z, phi[p].z));
      clo = V3(std::fmin(clo.x, cen[p].x), std::fmin(clo.y, cen[p].y), std::fmin(clo.z, cen[p].z));
      chi = V3(std::fmax(chi.x, cen[p].x), std::fmax(chi.y, cen[p].y), std::fmax(chi.z, cen[p].z));
    }
    BNode& nd = bvh.nodes[tk.node];
    nd.lo = lo;
    nd.hi = hi;
    if (tk.count <= max_leaf) {
      nd.left = tk.first;
      nd.count = tk.count;
      continue;
    }
    // best binned split over the three axes
    int best_axis = -1, best_bin = -1;
    float best_cost = INFINITY;
    for (int ax = 0; ax < 3; ++ax) {
      float c0 = comp(clo, ax), c1 = comp(chi, ax);
      if (!(c1 > c0)) continue;
      float scale = (float)NB / (c1 - c0);
      vec3 blo[NB], bhi[NB];
      uint32_t bcnt[NB];
      for (int b = 0; b < NB; ++b) { blo[b] = V3(INFINITY); bhi[b] = V3(-INFINITY); bcnt[b] = 0; }
      for (uint32_t i = tk.first; i < tk.first + tk.count; ++i) {
        uint32_t p = bvh.order[i];
        int b = std::min(NB - 1, std::max(0, (int)((comp(cen[p], ax) - c0) * scale)));
        bcnt[b]++;
        blo[b] = V3(std::fmin(blo[b].x, plo[p].x), std::fmin(blo[b].y, plo[p].y), std::fmin(blo[b].z, plo[p].z));
        bhi[b] = V3(std::fmax(bhi[b].x, phi[p].x), std::fmax(bhi[b].y, phi[p].y), std::fmax(bhi[b].z, phi[p].z));
      }
      float rarea[NB];
      uint32_t rcnt[NB];
      vec3 l = V3(INFINITY), h = V3(-INFINITY);
      uint32_t c = 0;
      for (int b = NB - 1; b > 0; --b) {
        l = V3(std::fmin(l.x, blo[b].x), std::fmin(l.y, blo[b].y), std::fmin(l.z, blo[b].z));
        h = V3(std::fmax(h.x, bhi[b].x), std::fmax(h.y, bhi[b].y), std::fmax(h.z, bhi[b].z));
        c += bcnt[b];
        rarea[b] = c ? half_area(l, h) : 0.0f;
        rcnt[b] = c;
      }
      l = V3(INFINITY); h = V3(-INFINITY); c = 0;
      for (int b = 0; b < NB - 1; ++b) {
        l = V3(std::fmin(l.x, blo[b].x), std::fmin(l.y, blo[b].y), std::fmin(l.z, blo[b].z));
        h = V3(std::fmax(h.x, bhi[b].x), std::fmax(h.y, bhi[b].y), std::fmax(h.z, bhi[b].z));
        c += bcnt[b];
        if (c == 0 || rcnt[b + 1] == 0) continue;
        float cost = half_area(l, h) * (float)c + rarea[b + 1] * (float)rcnt[b + 1];
        if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = b; }
      }
    }
    uint32_t mid;
    if (best_axis >= 0) {
      float c0 = comp(clo, best_axis), c1 = comp(chi, best_axis);
      float scale = (float)NB / (c1 - c0);
      auto it = std::partition(bvh.order.begin() + tk.first, bvh.order.begin() + tk.first + tk.count, [&](uint32_t p) {
        int b = std::min(NB - 1, std::max(0, (int)((comp(cen[p], best_axis) - c0) * scale)));
        return b <= best_bin;
      });
      mid = (uint32_t)(it - bvh.order.begin());
    } else {
      mid = tk.first + tk.count / 2;  // all centroids coincide: split the list in half
    }
    if (mid == tk.first || mid == tk.first + tk.count) mid = tk.first + tk.count / 2;
    uint32_t left = (uint32_t)bvh.nodes.size();
    bvh.nodes.push_back(BNode{});
    bvh.nodes.push_back(BNode{});
    bvh.nodes[tk.node].left = left;
    bvh.nodes[tk.node].count = 0;
    stack.push_back({left, tk.first, mid - tk.first});
    stack.push_back({left + 1, mid, tk.first + tk.count - mid});
  }
}

// Conservative slab test. The primitive tests round differently from the slab arithmetic and accept
// rays that pass a few ulps (of the coordinates involved) outside the exact geometry, so the box is
// widened in SPACE by 2^-20 of the largest coordinate magnitude in play (a pad on t alone is not
// enough for slabs the ray is nearly parallel to), and the interval by 2^-20 relative in t.
static inline bool slab(const BNode& n, vec3 o, vec3 idir, float tmin, float tmax, float& entry) {
  const float pad = 9.5367431640625e-07f;
  float mag = std::fmax(std::fmax(std::fabs(o.x), std::fabs(o.y)), std::fabs(o.z));
  mag = std::fmax(mag, std::fmax(std::fmax(std::fabs(n.lo.x), std::fabs(n.lo.y)), std::fabs(n.lo.z)));
  mag = std::fmax(mag, std::fmax(std::fmax(std::fabs(n.hi.x), std::fabs(n.hi.y)), std::fabs(n.hi.z)));
  const float e = mag * pad;
  float x0 = ((n.lo.x - e) - o.x) * idir.x, x1 = ((n.hi.x + e) - o.x) * idir.x;
  float y0 = ((n.lo.y - e) - o.y) * idir.y, y1 = ((n.hi.y + e) - o.y) * idir.y;
  float z0 = ((n.lo.z - e) - o.z) * idir.z, z1 = ((n.hi.z + e) - o.z) * idir.z;
  float tn = std::fmax(std::fmax(std::fmin(x0, x1), std::fmin(y0, y1)), std::fmin(z0, z1));
  float tf = std::fmin(std::fmin(std::fmax(x0, x1), std::fmax(y0, y1)), std::fmax(z0, z1));
  // (multiplicative so that an infinite entry/exit stays infinite instead of turning into inf - inf = NaN)
  float e0 = std::fmax(tn > 0.0f ? tn * (1.0f - pad) : tn * (1.0f + pad), tmin);
  float e1 = std::fmin(tf > 0.0f ? tf * (1.0f + pad) : tf * (1.0f - pad), tmax);
  entry = e0;
  return e0 <= e1;
}

// Generic traversal: calls leaf(prim_index) for every primitive whose leaf box the ray may touch;
// `tmax` is a reference the callback may shrink; the callback returns true to stop (any-hit).
template <class F>
static inline bool traverse(const BVH& bvh, vec3 o, vec3 d, float tmin, float& tmax, uint64_t& nodes_visited, F&& leaf) {
  if (bvh.empty()) return false;
  vec3 idir = V3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
  uint32_t stack[128];
  int sp = 0;
  float e;
  if (!slab(bvh.nodes[0], o, idir, tmin, tmax, e)) return false;
  stack[sp++] = 0;
  while (sp) {
    const BNode& n = bvh.nodes[stack[--sp]];
    float en;
    if (!slab(n, o, idir, tmin, tmax, en)) continue;  // tmax may have shrunk since the push
    nodes_visited++;
    if (n.count) {
      for (uint32_t i = 0; i < n.count; ++i)
        if (leaf(bvh.order[n.left + i])) return true;
      continue;
    }
    float e0, e1;
    bool h0 = slab(bvh.nodes[n.left], o, idir, tmin, tmax, e0);
    bool h1 = slab(bvh.nodes[n.left + 1], o, idir, tmin, tmax, e1);
    if (h0 && h1) {
      if (e0 <= e1) { stack[sp++] = n.left + 1; stack[sp++] = n.left; }
      else { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
    } else if (h0) stack[sp++] = n.left;
    else if (h1) stack[sp++] = n.left + 1;
  }
  return false;
}

}  // namespace orc
