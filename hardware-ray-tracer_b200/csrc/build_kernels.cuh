// GPU BVH builder, replacing vkCmdBuildAccelerationStructuresKHR (RT/Scene.cpp:163-174,176-255,
// 256-311; the driver's builder has no source in the reference) and filling the LBVH placeholder
// of Scene::prepareRendering (RT/Scene.cpp:135-138).
//
// Pipeline (one launch each, no host synchronisation in between):
//   prim_bounds  : primitive AABBs (+ exact mesh bounds via warp-reduced atomics)
//   morton       : 32-bit extent-adaptive Morton code of the AABB centre inside the scene bounds
//   radix sort   : (code, primitive) pairs                                  (radix_sort.cuh)
//   hierarchy    : Karras 2012 — one thread per internal node, binary radix tree over the sorted codes
//   refit        : bottom-up AABBs + subtree primitive counts, atomic arrival counters
//   treelet      : SAH treelet restructuring (treelet.cuh, optional)
//   collapse     : level-synchronous conversion into the compressed 8-wide layout of device_types.cuh
#pragma once
#include "device_types.cuh"
#include "vecmath.cuh"

namespace brt {

// scratch words shared by the build kernels of one BVH
struct BuildGlobals {
  uint32_t bounds[6];        // ordered-uint min xyz / max xyz of the padded primitive boxes
  uint32_t exact[6];         // ordered-uint min xyz / max xyz of the exact primitive boxes (mesh bounds)
  uint32_t node_count;       // 8-wide nodes allocated so far
  uint32_t prim_count;       // primitive records allocated so far
  uint32_t levels;           // number of collapse levels that had work
  uint32_t overflow;         // set when a capacity was exceeded
  uint32_t level_count[64];  // work items per collapse level
  uint32_t pad[4];
};

// ---- warp-aggregated float min/max into ordered uints -------------------------------------------------
BRT_HD void reduce_bounds(uint32_t* dst, f3 lo, f3 hi) {
  uint32_t l[3] = {float_to_ordered(lo.x), float_to_ordered(lo.y), float_to_ordered(lo.z)};
  uint32_t h[3] = {float_to_ordered(hi.x), float_to_ordered(hi.y), float_to_ordered(hi.z)};
#ifdef BRT_EMU
  for (int k = 0; k < 3; ++k) {
    atomic_min(dst + k, l[k]);
    atomic_max(dst + 3 + k, h[k]);
  }
#else
  const unsigned mask = __activemask();
  const int leader = __ffs(mask) - 1;
  const int lane = threadIdx.x & 31;
  for (int k = 0; k < 3; ++k) {
    uint32_t a = __reduce_min_sync(mask, l[k]);
    uint32_t b = __reduce_max_sync(mask, h[k]);
    if (lane == leader) {
      atomicMin(dst + k, a);
      atomicMax(dst + 3 + k, b);
    }
  }
#endif
}

// Every primitive box is padded by 2^-20 of its largest coordinate magnitude so that a primitive
// test that accepts a ray passing a few ulps outside the exact geometry is never culled by a box.
BRT_HD void pad_box(f3& lo, f3& hi) {
  float m = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))), fmaxf(fabsf(lo.z), fabsf(hi.z)));
  float e = fmaxf(m * 9.5367431640625e-07f, 1e-30f);
  lo = F3(lo.x - e, lo.y - e, lo.z - e);
  hi = F3(hi.x + e, hi.y + e, hi.z + e);
}

// ---- primitive bounds: triangles ---------------------------------------------------------------------
struct TriBoundsParams {
  uint32_t count;
  const uint32_t* count_ptr;
  const float* vertices;    // stride 8 floats (brt_vertex)
  const uint32_t* indices;  // 3 per triangle
  float4* prim_lo;          // xyz = padded min, w = unused
  float4* prim_hi;
  BuildGlobals* g;
};
BRT_HD void tri_bounds_body(const TriBoundsParams& p, uint32_t i) {
  const uint32_t i0 = p.indices[3 * i], i1 = p.indices[3 * i + 1], i2 = p.indices[3 * i + 2];
  const f3 a = F3(p.vertices[8 * (size_t)i0], p.vertices[8 * (size_t)i0 + 1], p.vertices[8 * (size_t)i0 + 2]);
  const f3 b = F3(p.vertices[8 * (size_t)i1], p.vertices[8 * (size_t)i1 + 1], p.vertices[8 * (size_t)i1 + 2]);
  const f3 c = F3(p.vertices[8 * (size_t)i2], p.vertices[8 * (size_t)i2 + 1], p.vertices[8 * (size_t)i2 + 2]);
  f3 lo = fmin3(fmin3(a, b), c), hi = fmax3(fmax3(a, b), c);
  reduce_bounds(p.g->exact, lo, hi);
  pad_box(lo, hi);
  reduce_bounds(p.g->bounds, lo, hi);
  p.prim_lo[i] = make_float4(lo.x, lo.y, lo.z, 0.0f);
  p.prim_hi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
}

// ---- primitive bounds: instances (TLAS) ---------------------------------------------------------------
// world box of an instance = |M| applied to the mesh box in centre/extent form
struct InstBoundsParams {
  uint32_t count;
  const uint32_t* count_ptr;
  const InstShade* shade;     // indexed by instance id
  const uint32_t* inst_ids;   // TLAS primitive -> instance id
  const float4* mesh_bounds;  // 2 per mesh: exact lo, hi
  float4* prim_lo;
  float4* prim_hi;
  BuildGlobals* g;
};
BRT_HD void inst_bounds_body(const InstBoundsParams& p, uint32_t i) {
  const InstShade& s = p.shade[p.inst_ids[i]];
  f3 lo, hi;
  instance_world_box(s.o2w, xyz(p.mesh_bounds[2 * s.mesh]), xyz(p.mesh_bounds[2 * s.mesh + 1]), lo, hi);
  // the transform itself rounds: pad generously relative to the box magnitude
  pad_box(lo, hi);
  pad_box(lo, hi);
  reduce_bounds(p.g->bounds, lo, hi);
  p.prim_lo[i] = make_float4(lo.x, lo.y, lo.z, 0.0f);
  p.prim_hi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
}

// ---- Morton codes ------------------------------------------------------------------------------------
#define BRT_MORTON_BITS 32
struct MortonParams {
  uint32_t count;
  const uint32_t* count_ptr;
  const float4* prim_lo;
  const float4* prim_hi;
  const BuildGlobals* g;
  uint32_t* keys;
  uint32_t* vals;
};
BRT_HD void morton_body(const MortonParams& p, uint32_t i) {
  const f3 blo = F3(ordered_to_float(p.g->bounds[0]), ordered_to_float(p.g->bounds[1]), ordered_to_float(p.g->bounds[2]));
  const f3 bhi = F3(ordered_to_float(p.g->bounds[3]), ordered_to_float(p.g->bounds[4]), ordered_to_float(p.g->bounds[5]));
  const f3 c = (xyz(p.prim_lo[i]) + xyz(p.prim_hi[i])) * 0.5f;
  const f3 ext = bhi - blo;
  const float fx = ext.x > 0.0f ? (c.x - blo.x) / ext.x : 0.0f;
  const float fy = ext.y > 0.0f ? (c.y - blo.y) / ext.y : 0.0f;
  const float fz = ext.z > 0.0f ? (c.z - blo.z) / ext.z : 0.0f;
  // 32 code bits, most significant first; each bit halves the axis whose cells are currently the longest (ties: x, y, z), so a
  // flat or elongated mesh does not waste a third of its bits on an axis that has nothing to separate: a 16 x 1.6 x 16 terrain gets
  // 12 + 8 + 12 bits instead of 10 + 10 + 10 (for a cube this is the usual x, y, z interleave)
  const uint32_t qx = (uint32_t)fminf(fmaxf(fx * 65536.0f, 0.0f), 65535.0f);
  const uint32_t qy = (uint32_t)fminf(fmaxf(fy * 65536.0f, 0.0f), 65535.0f);
  const uint32_t qz = (uint32_t)fminf(fmaxf(fz * 65536.0f, 0.0f), 65535.0f);
  float cx = ext.x, cy = ext.y, cz = ext.z;
  int bx = 15, by = 15, bz = 15;  // next bit of each 16-bit coordinate
  uint32_t code = 0;
  for (int b = 0; b < BRT_MORTON_BITS; ++b) {
    uint32_t bit;
    if (cx >= cy && cx >= cz && bx >= 0) { bit = (qx >> bx) & 1u; bx--; cx *= 0.5f; }
    else if (cy >= cz && by >= 0) { bit = (qy >> by) & 1u; by--; cy *= 0.5f; }
    else if (bz >= 0) { bit = (qz >> bz) & 1u; bz--; cz *= 0.5f; }
    else if (bx >= 0) { bit = (qx >> bx) & 1u; bx--; cx *= 0.5f; }
    else { bit = (qy >> by) & 1u; by--; cy *= 0.5f; }
    code = (code << 1) | bit;
  }
  p.keys[i] = code;
  p.vals[i] = i;
}

// ---- Karras hierarchy --------------------------------------------------------------------------------
struct HierarchyParams {
  uint32_t count;  // number of internal nodes = n - 1
  const uint32_t* count_ptr;
  uint32_t n;      // primitives
  const uint32_t* keys;  // sorted
  BNode* nodes;          // 2n-1
  uint32_t* parent;      // 2n-1
};
BRT_HD int lcp(const uint32_t* keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const uint32_t a = keys[i], b = keys[j];
  return a != b ? clz32(a ^ b) : 32 + clz32((uint32_t)i ^ (uint32_t)j);
}
BRT_HD void hierarchy_body(const HierarchyParams& p, uint32_t ui) {
  const int n = (int)p.n, i = (int)ui;
  const int d = (lcp(p.keys, n, i, i + 1) - lcp(p.keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = lcp(p.keys, n, i, i - d);
  int lmax = 2;
  while (lcp(p.keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (lcp(p.keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = lcp(p.keys, n, i, j);
  int s = 0, t = l;
  do {
    t = (t + 1) / 2;
    if (lcp(p.keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int gamma = i + s * d + (d < 0 ? -1 : 0);
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  const uint32_t n_int = p.n - 1;
  const uint32_t left = lo == gamma ? n_int + (uint32_t)gamma : (uint32_t)gamma;
  const uint32_t right = hi == gamma + 1 ? n_int + (uint32_t)(gamma + 1) : (uint32_t)(gamma + 1);
  p.nodes[i].lo.w = u2f(left);
  p.nodes[i].hi.w = u2f(right);
  p.parent[left] = ui;
  p.parent[right] = ui;
  if (i == 0) p.parent[0] = BRT_MISS;
}

#define BRT_SAH_CI 1.2f  // cost of visiting an internal node
#define BRT_SAH_CT 1.0f  // cost of one primitive test
#define BRT_WCOST_STRIDE 8
BRT_HD float box_area(f3 lo, f3 hi) {
  const f3 e = hi - lo;
  return 2.0f * ((e.x * e.y + e.y * e.z) + e.z * e.x);
}
BRT_HD void wide_cost_node(float* wcost, unsigned long long* wplan, uint32_t par, uint32_t l, uint32_t r, uint32_t n_int, float area);

// ---- bottom-up refit ---------------------------------------------------------------------------------
// One walk from every leaf towards the root (the second thread to arrive at a node owns it) fills the boxes and the
// primitive counts and, when asked to, the SAH cost of every subtree (the statistic of brt_get_stats) and the cost table
// of the collapse (wide_cost_node): a build without treelet restructuring — a re-built dynamic mesh, the per-frame TLAS —
// has its final topology here, and each separate bottom-up pass costs a chain of ~depth x 2 us of dependent atomics.
struct RefitParams {
  uint32_t count;  // n leaves
  const uint32_t* count_ptr;
  const uint32_t* vals;  // sorted position -> primitive
  const float4* prim_lo;
  const float4* prim_hi;
  BNode* nodes;
  const uint32_t* parent;
  uint32_t* arrive;     // n-1, zeroed
  uint32_t* sub_count;  // 2n-1: primitives below each node
  float* cost;          // 2n-1 or null: SAH cost of every subtree
  float* wcost;         // n-1 rows or null: cost table of the collapse
  unsigned long long* wplan;  // n-1: the choices that realise it (with wcost)
};
BRT_HD void refit_body(const RefitParams& p, uint32_t i) {
  const uint32_t n_int = p.count - 1;
  const uint32_t prim = p.vals[i];
  float4 lo = p.prim_lo[prim], hi = p.prim_hi[prim];
  lo.w = u2f(prim);
  hi.w = u2f(BRT_MISS);
  p.nodes[n_int + i].lo = lo;
  p.nodes[n_int + i].hi = hi;
  p.sub_count[n_int + i] = 1;
  if (p.cost) p.cost[n_int + i] = BRT_SAH_CT * box_area(xyz(lo), xyz(hi));
  uint32_t cur = n_int + i;
  for (;;) {
    const uint32_t par = p.parent[cur];
    if (par == BRT_MISS) break;
    if (atomic_add_acq_rel(&p.arrive[par], 1u) == 0u) break;  // the sibling subtree is not finished yet
    volatile BNode* nd = p.nodes + par;
    const uint32_t l = f2u(nd->lo.w), r = f2u(nd->hi.w);
    volatile const BNode* a = p.nodes + l;
    volatile const BNode* b = p.nodes + r;
    nd->lo.x = fminf(a->lo.x, b->lo.x);
    nd->lo.y = fminf(a->lo.y, b->lo.y);
    nd->lo.z = fminf(a->lo.z, b->lo.z);
    nd->hi.x = fmaxf(a->hi.x, b->hi.x);
    nd->hi.y = fmaxf(a->hi.y, b->hi.y);
    nd->hi.z = fmaxf(a->hi.z, b->hi.z);
    volatile uint32_t* sc = p.sub_count;
    sc[par] = sc[l] + sc[r];
    if (p.cost || p.wcost) {
      const float area = box_area(F3(nd->lo.x, nd->lo.y, nd->lo.z), F3(nd->hi.x, nd->hi.y, nd->hi.z));
      if (p.cost) {
        volatile float* c = p.cost;
        c[par] = BRT_SAH_CI * area + c[l] + c[r];
      }
      if (p.wcost) wide_cost_node(p.wcost, p.wplan, par, l, r, n_int, area);
    }
    cur = par;
  }
}


// ---- cost table for the collapse: which binary nodes become 8-wide nodes ---------------------------------
// The collapse decides which binary nodes become 8-wide nodes. Every triangle is its own leaf slot, so the primitive
// tests a ray performs do not depend on that choice; what does is the number of wide nodes it visits, in expectation
// proportional to the sum of their surface areas. That sum is minimised exactly by a dynamic programme over the binary
// tree (after Ylitie, Karras & Laine, HPG 2017, §3.1): for every internal binary node n and i = 1..7
//   C(n, 1) = area(n) + D(n, 8)                      n becomes a wide node, its subtree is spread over 8 slots
//   C(n, i) = min(D(n, i), C(n, i - 1))              n's subtree occupies at most i slots of an ancestor's wide node
//   D(n, j) = min over 0 < k < j of C(left, k) + C(right, j - k),      C(leaf, .) = 0
// One bottom-up pass (arrival counters, as the refit) stores C(n, 1..7) and, per node, the choices that realise D(n, 2..8)
// (how many slots each child really uses: 0 = it is a primitive, 1 = it becomes a wide node, b > 1 = it is opened further
// over b slots), 6 bits per j in one 64-bit word; the collapse only follows them (collapse_select). (Opening the child with the largest area until 8 slots are full — the previous rule — left the nodes
// half empty: 4.0 of 8 slots on the 1M-triangle scene, 9.3 node visits per primary ray against 7.1 now.)
struct WideCostParams {
  uint32_t count;  // n leaves
  const uint32_t* count_ptr;
  const BNode* nodes;
  const uint32_t* parent;
  uint32_t* arrive;  // n-1, zeroed before the pass
  float* wcost;      // BRT_WCOST_STRIDE floats per internal node: C(n, 1..7)
  unsigned long long* wplan;  // per internal node: slots used by the left / right child for j = 2..8
};
// D(n, j) for j = 2..8 from the children's tables (index i-1 holds C(., i)); returns the arg-min split in `split[j]`
BRT_HD void wide_distribute(const float* L, const float* R, float* D, int* split) {
  for (int j = 2; j <= 8; ++j) {
    float best = INFINITY;
    int bk = 1;
    for (int k = 1; k < j; ++k) {
      const int a = k < 7 ? k : 7, b = (j - k) < 7 ? (j - k) : 7;
      const float c = L[a - 1] + R[b - 1];
      if (c < best) { best = c; bk = k; }
    }
    D[j] = best;
    if (split) split[j] = bk;
  }
}
BRT_HD void wide_cost_load(const float* wcost, uint32_t id, uint32_t n_int, float* out) {
  volatile const float* w = wcost + (size_t)id * BRT_WCOST_STRIDE;
  for (int k = 0; k < 7; ++k) out[k] = id < n_int ? w[k] : 0.0f;
}
// slots a child really uses when it is offered b: none of its own if it is a primitive (0 = leaf slot), fewer if the extra
// ones buy nothing (C(c, b) = C(c, b - 1)), 1 = it becomes a wide node
BRT_HD int wide_effective(const float* T, bool leaf, int b) {
  if (leaf) return 0;
  if (b > 7) b = 7;
  while (b > 1 && T[b - 1] == T[b - 2]) b--;
  return b;
}
// C(par, 1..7) and the plan of par from the finished tables of its children l, r
BRT_HD void wide_cost_node(float* wcost, unsigned long long* wplan, uint32_t par, uint32_t l, uint32_t r, uint32_t n_int, float area) {
  float L[7], R[7], D[9];
  int split[9];
  wide_cost_load(wcost, l, n_int, L);
  wide_cost_load(wcost, r, n_int, R);
  wide_distribute(L, R, D, split);
  unsigned long long plan = 0ull;
  for (int j = 2; j <= 8; ++j) {
    const unsigned bl = (unsigned)wide_effective(L, l >= n_int, split[j]), br = (unsigned)wide_effective(R, r >= n_int, j - split[j]);
    plan |= (unsigned long long)(bl | (br << 3)) << (6 * (j - 2));
  }
  wplan[par] = plan;
  volatile float* w = wcost + (size_t)par * BRT_WCOST_STRIDE;
  float c = area + D[8];
  w[0] = c;
  for (int k = 2; k <= 7; ++k) {
    c = fminf(D[k], c);
    w[k - 1] = c;
  }
}
BRT_HD void wide_cost_body(const WideCostParams& p, uint32_t i) {
  const uint32_t n_int = p.count - 1;
  uint32_t cur = n_int + i;
  for (;;) {
    const uint32_t par = p.parent[cur];
    if (par == BRT_MISS) break;
    if (atomic_add_acq_rel(&p.arrive[par], 1u) == 0u) break;  // the sibling subtree is not finished yet
    volatile const BNode* nd = p.nodes + par;
    const uint32_t l = f2u(nd->lo.w), r = f2u(nd->hi.w);
    wide_cost_node(p.wcost, p.wplan, par, l, r, n_int, box_area(F3(nd->lo.x, nd->lo.y, nd->lo.z), F3(nd->hi.x, nd->hi.y, nd->hi.z)));
    cur = par;
  }
}

// ---- collapse to the compressed 8-wide layout ----------------------------------------------------------
struct CollapseParams {
  uint32_t count;
  const uint32_t* count_ptr;  // = &g->level_count[level]
  uint32_t level;
  uint32_t n;                 // primitives
  uint32_t max_leaf;          // primitives per leaf slot: 1 (the node format has one leaf bit per slot)
  const BNode* nodes;
  const uint32_t* sub_count;
  const float* wcost;         // C(n, 1..7) of wide_cost_body; null: greedy largest-area opening
  const unsigned long long* wplan;  // the choices behind wcost
  const uint2* queue_in;      // (binary node id, wide node index)
  uint2* queue_out;
  uint32_t queue_cap;
  BuildGlobals* g;
  Node8* out_nodes;
  uint32_t node_cap;
  // leaf payload: triangles
  const float* vertices;
  const uint32_t* indices;
  TriRec* out_tris;
  // leaf payload: instances
  const InstRec* src_inst;  // indexed by TLAS primitive
  InstRec* out_inst;
};

BRT_HD void emit_prim(const CollapseParams& p, uint32_t prim, uint32_t dst) {
  if (p.out_tris) {
    const uint32_t i0 = p.indices[3 * (size_t)prim], i1 = p.indices[3 * (size_t)prim + 1], i2 = p.indices[3 * (size_t)prim + 2];
    TriRec r;
    r.v0 = make_float4(p.vertices[8 * (size_t)i0], p.vertices[8 * (size_t)i0 + 1], p.vertices[8 * (size_t)i0 + 2], u2f(prim));
    r.v1 = make_float4(p.vertices[8 * (size_t)i1], p.vertices[8 * (size_t)i1 + 1], p.vertices[8 * (size_t)i1 + 2], 0.0f);
    r.v2 = make_float4(p.vertices[8 * (size_t)i2], p.vertices[8 * (size_t)i2 + 1], p.vertices[8 * (size_t)i2 + 2], 0.0f);
    p.out_tris[dst] = r;
  } else {
    p.out_inst[dst] = p.src_inst[prim];
  }
}

// smallest biased exponent e such that the 8-bit grid p + [0,255] * 2^(e-127) reaches `hi`
BRT_HD uint32_t grid_exponent(float p, float hi) {
  const float extent = hi - p;
  uint32_t e = 1u;
  if (extent > 0.0f) {
    const float step = extent / 255.0f;
    e = ((f2u(step) >> 23) & 0xffu) + 1u;  // floor(log2(step)) + 1
    if (e < 1u) e = 1u;
    if (e > 254u) e = 254u;
    while (e < 254u && add_rd(p, 255.0f * u2f(e << 23)) < hi) e++;  // rounding of extent / step
  }
  return e;
}
// largest q in [0,255] with p + q*s <= v, evaluated with the same rounding the traversal may see
BRT_HD uint32_t quant_lo(float p, float s, float v) {
  float q = floorf((v - p) / s);
  q = fminf(fmaxf(q, 0.0f), 255.0f);
  while (q > 0.0f && add_ru(p, q * s) > v) q -= 1.0f;
  return (uint32_t)q;
}
BRT_HD uint32_t quant_hi(float p, float s, float v) {
  float q = ceilf((v - p) / s);
  q = fminf(fmaxf(q, 0.0f), 255.0f);
  while (q < 255.0f && add_rd(p, q * s) < v) q += 1.0f;
  return (uint32_t)q;
}

// Which binary nodes become the (up to 8) children of the wide node made from binary node `root`: ch[0..n), bit k of
// `inner`: child k becomes a wide node itself (otherwise it is a leaf slot). Returns n.
BRT_HD int collapse_select(const CollapseParams& p, uint32_t root, uint32_t* ch, uint32_t& inner) {
  const uint32_t n_int = p.n - 1;
  int n = 1;
  ch[0] = root;
  inner = (root < n_int && p.sub_count[root] > p.max_leaf) ? 1u : 0u;
  if (!inner) return n;
  if (p.wcost) {
    // follow the dynamic programme's plan: spread the subtree over the 8 slots with the splits that realise D(root, 8)
    uint32_t st_id[8];
    int st_budget[8];
    int sp = 0;
    n = 0;
    inner = 0u;
    st_id[sp] = root;
    st_budget[sp] = 8;
    sp++;
    while (sp) {
      --sp;
      const uint32_t id = st_id[sp];
      const int j = st_budget[sp];
      const BNode nd = p.nodes[id];
      const unsigned bits = (unsigned)(p.wplan[id] >> (6 * (j - 2))) & 63u;
      const uint32_t c2[2] = {f2u(nd.lo.w), f2u(nd.hi.w)};
      const int budget[2] = {(int)(bits & 7u), (int)(bits >> 3)};
      for (int s = 0; s < 2; ++s) {
        const uint32_t c = c2[s];
        const int b = budget[s];  // 0: a primitive (leaf slot), 1: becomes a wide node, else: opened over b slots
        if (b <= 1) {
          ch[n] = c;
          if (b == 1) inner |= 1u << n;
          n++;
        } else {
          st_id[sp] = c;
          st_budget[sp] = b;
          sp++;
        }
      }
    }
  } else {
    // greedy: always open the expandable child with the largest surface area
    float area[8];  // < 0: not expandable (becomes a leaf slot)
    area[0] = 1.0f;
    while (n < 8) {
      int best = -1;
      float ba = 0.0f;
      for (int k = 0; k < n; ++k)
        if (area[k] >= 0.0f && (best < 0 || area[k] > ba)) { best = k; ba = area[k]; }
      if (best < 0) break;
      const BNode nd = p.nodes[ch[best]];
      const uint32_t c2[2] = {f2u(nd.lo.w), f2u(nd.hi.w)};
      for (int s = 0; s < 2; ++s) {
        const uint32_t c = c2[s];
        const int dst = s == 0 ? best : n;
        ch[dst] = c;
        if (c < n_int && p.sub_count[c] > p.max_leaf) {
          const BNode cn = p.nodes[c];
          area[dst] = box_area(xyz(cn.lo), xyz(cn.hi));
        } else {
          area[dst] = -1.0f;
        }
      }
      n++;
    }
    inner = 0u;
    for (int k = 0; k < n; ++k)
      if (area[k] >= 0.0f) inner |= 1u << k;
  }
  return n;
}

BRT_HD void collapse_body(const CollapseParams& p, uint32_t item) {
  const uint2 work = p.queue_in[item];
  const uint32_t n_int = p.n - 1;
  uint32_t ch[8];
  uint32_t inner_mask;
  const int n = collapse_select(p, work.x, ch, inner_mask);
  float area[8];  // < 0: not expandable (becomes a leaf slot)
  for (int k = 0; k < 8; ++k) area[k] = (inner_mask >> k) & 1u ? 1.0f : -1.0f;
  // child boxes, node box
  f3 clo[8], chi[8];
  f3 nlo = F3(INFINITY), nhi = F3(-INFINITY);
  for (int k = 0; k < n; ++k) {
    const BNode cn = p.nodes[ch[k]];
    clo[k] = xyz(cn.lo);
    chi[k] = xyz(cn.hi);
    nlo = fmin3(nlo, clo[k]);
    nhi = fmax3(nhi, chi[k]);
  }
  // octant-ordered slot assignment: greedily give each (child, slot) pair with the highest
  // projection of the child's offset onto the slot's diagonal
  const f3 nc = (nlo + nhi) * 0.5f;
  int slot_of[8];
  int child_in[8];
  for (int k = 0; k < 8; ++k) { slot_of[k] = -1; child_in[k] = -1; }
  for (int round = 0; round < n; ++round) {
    float bs = -INFINITY;
    int bk = -1, bsl = -1;
    for (int k = 0; k < n; ++k) {
      if (slot_of[k] >= 0) continue;
      const f3 off = (clo[k] + chi[k]) * 0.5f - nc;
      for (int s = 0; s < 8; ++s) {
        if (child_in[s] >= 0) continue;
        const float sc = ((s & 1) ? off.x : -off.x) + ((s & 2) ? off.y : -off.y) + ((s & 4) ? off.z : -off.z);
        if (sc > bs) { bs = sc; bk = k; bsl = s; }
      }
    }
    if (bk < 0) {  // NaN boxes: fall back to first free
      for (int k = 0; k < n && bk < 0; ++k) if (slot_of[k] < 0) bk = k;
      for (int s = 0; s < 8 && bsl < 0; ++s) if (child_in[s] < 0) bsl = s;
    }
    slot_of[bk] = bsl;
    child_in[bsl] = bk;
  }
  // allocation
  uint32_t n_inner = 0, n_prims = 0;
  for (int s = 0; s < 8; ++s) {
    const int k = child_in[s];
    if (k < 0) continue;
    if (area[k] >= 0.0f) n_inner++;
    else n_prims += p.sub_count[ch[k]];
  }
  uint32_t child_base = 0, prim_base = 0;
  if (n_inner) child_base = atomic_add(&p.g->node_count, n_inner);
  if (n_prims) prim_base = atomic_add(&p.g->prim_count, n_prims);
  uint32_t queue_base = 0;
  if (n_inner) {
    queue_base = atomic_add(&p.g->level_count[p.level + 1], n_inner);
    if (child_base + n_inner > p.node_cap || queue_base + n_inner > p.queue_cap) {
      p.g->overflow = 1u;
      n_inner = 0;  // drop the subtree rather than write out of bounds; the host reports the error
    }
  }
  // grid
  const uint32_t ex = grid_exponent(nlo.x, nhi.x), ey = grid_exponent(nlo.y, nhi.y), ez = grid_exponent(nlo.z, nhi.z);
  const float sx = u2f(ex << 23), sy = u2f(ey << 23), sz = u2f(ez << 23);
  uint32_t qlx[8], qly[8], qlz[8], qhx[8], qhy[8], qhz[8];
  uint32_t imask = 0, leafmask = 0, inner_i = 0, prim_off = 0;
  for (int s = 0; s < 8; ++s) {
    qlx[s] = qly[s] = qlz[s] = 255u; qhx[s] = qhy[s] = qhz[s] = 0u;
    const int k = child_in[s];
    if (k < 0) continue;
    if (area[k] >= 0.0f) {
      if (n_inner == 0) continue;  // overflow: slot left empty
      imask |= 1u << s;
      p.queue_out[queue_base + inner_i] = make_uint2(ch[k], child_base + inner_i);
      inner_i++;
    } else {
      // a leaf slot holds exactly one primitive (max_leaf == 1): its record index is prim_base + the rank of
      // the slot among the node's leaf slots; the walk below tolerates a larger subtree only to stay in bounds
      const uint32_t cnt = p.sub_count[ch[k]];
      uint32_t st[4];
      int sp = 0;
      st[sp++] = ch[k];
      uint32_t w = 0;
      while (sp) {
        const uint32_t id = st[--sp];
        const BNode bn = p.nodes[id];
        if (id >= n_int) {
          emit_prim(p, f2u(bn.lo.w), prim_base + prim_off + w);
          w++;
        } else {
          st[sp++] = f2u(bn.hi.w);
          st[sp++] = f2u(bn.lo.w);
        }
      }
      leafmask |= 1u << s;
      prim_off += cnt;
    }
    qlx[s] = quant_lo(nlo.x, sx, clo[k].x); qly[s] = quant_lo(nlo.y, sy, clo[k].y); qlz[s] = quant_lo(nlo.z, sz, clo[k].z);
    qhx[s] = quant_hi(nlo.x, sx, chi[k].x); qhy[s] = quant_hi(nlo.y, sy, chi[k].y); qhz[s] = quant_hi(nlo.z, sz, chi[k].z);
  }
  auto pack4 = [](const uint32_t* b) { return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24); };
  Node8 out;
  out.q[0] = make_uint4(f2u(nlo.x), f2u(nlo.y), f2u(nlo.z), ex | (ey << 8) | (ez << 16) | (imask << 24));
  // byte x of q1.z / q1.w = inner / leaf mask with the two low bits of every slot index XORed with x: the traversal
  // selects the byte for its ray octant instead of permuting bits (traverse.cuh)
  uint32_t iperm = 0, lperm = 0;
  for (uint32_t x = 0; x < 4; ++x)
    for (uint32_t j = 0; j < 8; ++j) {
      iperm |= ((imask >> (j ^ x)) & 1u) << (8 * x + j);
      lperm |= ((leafmask >> (j ^ x)) & 1u) << (8 * x + j);
    }
  out.q[1] = make_uint4(child_base, prim_base, iperm, lperm);
  out.q[2] = make_uint4(pack4(qlx), pack4(qlx + 4), pack4(qly), pack4(qly + 4));
  out.q[3] = make_uint4(pack4(qlz), pack4(qlz + 4), pack4(qhx), pack4(qhx + 4));
  out.q[4] = make_uint4(pack4(qhy), pack4(qhy + 4), pack4(qhz), pack4(qhz + 4));
  if (work.y < p.node_cap) p.out_nodes[work.y] = out;
  if (item == 0) atomic_max(&p.g->levels, p.level + 1u);
}

#if !defined(BRT_EMU) || defined(BRT_EMU_WARP)  // (BRT_EMU_WARP: the test-only lock-step warp of tests/emu/warp_emu.h)
// ---- warp-cooperative collapse (device only) ------------------------------------------------------------
// collapse_body with the 32 lanes of a warp sharing ONE work item: a single thread spends ~50 us on an item (the children's
// boxes, the 64 (child, slot) scores of the octant-ordered assignment, 48 quantisations with directed rounding and up to eight
// dependent index -> vertex fetches for the triangle records, one after the other), and a build has one level-synchronous
// step per tree level, so small builds (the per-frame TLAS, a re-built dynamic BLAS) and the top levels of large ones
// were bound by that latency. Here lane k owns child k, then lane s owns slot s; the choice of the children
// (collapse_select) is evaluated redundantly by all lanes. Same node as collapse_body produces for the same item.
__device__ __forceinline__ float warp_min_f(float v) { return ordered_to_float(__reduce_min_sync(0xffffffffu, float_to_ordered(v))); }
__device__ __forceinline__ float warp_max_f(float v) { return ordered_to_float(__reduce_max_sync(0xffffffffu, float_to_ordered(v))); }

__device__ __forceinline__ void collapse_warp(const CollapseParams& p, uint32_t item, unsigned lane) {
  const unsigned FULL = 0xffffffffu;
  const uint2 work = p.queue_in[item];
  const uint32_t n_int = p.n - 1;
  uint32_t ch[8];
  uint32_t inner_mask;
  const int n = collapse_select(p, work.x, ch, inner_mask);  // warp-uniform
  // lane k < n: box of child k
  uint32_t my_ch = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if ((int)lane == k) my_ch = ch[k];
  f3 blo = F3(INFINITY), bhi = F3(-INFINITY);
  if ((int)lane < n) {
    const BNode cn = p.nodes[my_ch];
    blo = xyz(cn.lo);
    bhi = xyz(cn.hi);
  }
  const f3 nlo = F3(warp_min_f(blo.x), warp_min_f(blo.y), warp_min_f(blo.z));
  const f3 nhi = F3(warp_max_f(bhi.x), warp_max_f(bhi.y), warp_max_f(bhi.z));
  // octant-ordered slot assignment: lane L scores the pairs (child L >> 2, slots 2 (L & 3) and 2 (L & 3) + 1); every round takes the
  // pair with the highest score, ties to the lowest (child, slot) — the order collapse_body scans them in
  const f3 nc = (nlo + nhi) * 0.5f;
  const int pk = (int)(lane >> 2);
  const f3 off = F3(__shfl_sync(FULL, (blo.x + bhi.x) * 0.5f - nc.x, pk), __shfl_sync(FULL, (blo.y + bhi.y) * 0.5f - nc.y, pk),
                    __shfl_sync(FULL, (blo.z + bhi.z) * 0.5f - nc.z, pk));
  float score[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int sl = (int)(lane & 3u) * 2 + e;
    score[e] = ((sl & 1) ? off.x : -off.x) + ((sl & 2) ? off.y : -off.y) + ((sl & 4) ? off.z : -off.z);
  }
  uint32_t child_done = 0, slot_used = 0;
  int my_child = -1;  // lane s < 8: the child in slot s
  for (int round = 0; round < n; ++round) {
    float bs = -INFINITY;
    int bi = 64;  // pair index = child * 8 + slot
    if (pk < n && !((child_done >> pk) & 1u)) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int sl = (int)(lane & 3u) * 2 + e;
        if (!((slot_used >> sl) & 1u) && score[e] > bs) { bs = score[e]; bi = pk * 8 + sl; }
      }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const float os = __shfl_xor_sync(FULL, bs, d);
      const int oi = __shfl_xor_sync(FULL, bi, d);
      if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
    }
    int bk, bsl;
    if (bi < 64) {
      bk = bi >> 3;
      bsl = bi & 7;
    } else {  // NaN / infinite boxes: first free child, first free slot
      bk = __ffs(~child_done & 0xffu) - 1;
      bsl = __ffs(~slot_used & 0xffu) - 1;
    }
    child_done |= 1u << bk;
    slot_used |= 1u << bsl;
    if ((int)lane == bsl) my_child = bk;
  }
  // lane s < 8 owns slot s from here on
  const bool owner = lane < 8u && my_child >= 0;
  const int src = owner ? my_child : 0;
  const f3 clo = F3(__shfl_sync(FULL, blo.x, src), __shfl_sync(FULL, blo.y, src), __shfl_sync(FULL, blo.z, src));
  const f3 chi = F3(__shfl_sync(FULL, bhi.x, src), __shfl_sync(FULL, bhi.y, src), __shfl_sync(FULL, bhi.z, src));
  const uint32_t cid = __shfl_sync(FULL, my_ch, src);
  const bool is_inner = owner && ((inner_mask >> my_child) & 1u);
  const bool is_leaf = owner && !is_inner;
  const uint32_t cnt = is_leaf ? p.sub_count[cid] : 0u;
  const uint32_t inner_slots = __ballot_sync(FULL, is_inner);
  uint32_t n_inner = (uint32_t)__popc(inner_slots);
  const uint32_t n_prims = __reduce_add_sync(FULL, cnt);
  uint32_t prim_off = 0;  // primitives in the leaf slots below this one
#pragma unroll
  for (int s2 = 0; s2 < 8; ++s2) {
    const uint32_t c2 = __shfl_sync(FULL, cnt, s2);
    if ((int)lane > s2) prim_off += c2;
  }
  uint32_t child_base = 0, prim_base = 0, queue_base = 0, drop = 0;
  if (lane == 0) {
    if (n_inner) child_base = atomic_add(&p.g->node_count, n_inner);
    if (n_prims) prim_base = atomic_add(&p.g->prim_count, n_prims);
    if (n_inner) {
      queue_base = atomic_add(&p.g->level_count[p.level + 1], n_inner);
      if (child_base + n_inner > p.node_cap || queue_base + n_inner > p.queue_cap) {
        p.g->overflow = 1u;
        drop = 1u;  // drop the subtree rather than write out of bounds; the host reports the error
      }
    }
  }
  child_base = __shfl_sync(FULL, child_base, 0);
  prim_base = __shfl_sync(FULL, prim_base, 0);
  queue_base = __shfl_sync(FULL, queue_base, 0);
  drop = __shfl_sync(FULL, drop, 0);
  const bool keep_inner = is_inner && !drop;
  const uint32_t imask = __ballot_sync(FULL, keep_inner) & 0xffu;
  const uint32_t leafmask = __ballot_sync(FULL, is_leaf) & 0xffu;
  // grid (warp-uniform), quantised bounds per slot
  const uint32_t ex = grid_exponent(nlo.x, nhi.x), ey = grid_exponent(nlo.y, nhi.y), ez = grid_exponent(nlo.z, nhi.z);
  const float sx = u2f(ex << 23), sy = u2f(ey << 23), sz = u2f(ez << 23);
  uint32_t q[6] = {255u, 255u, 255u, 0u, 0u, 0u};  // lo xyz, hi xyz
  if (keep_inner || is_leaf) {
    q[0] = quant_lo(nlo.x, sx, clo.x); q[1] = quant_lo(nlo.y, sy, clo.y); q[2] = quant_lo(nlo.z, sz, clo.z);
    q[3] = quant_hi(nlo.x, sx, chi.x); q[4] = quant_hi(nlo.y, sy, chi.y); q[5] = quant_hi(nlo.z, sz, chi.z);
  }
  if (keep_inner) {
    const uint32_t inner_i = (uint32_t)__popc(imask & ((1u << lane) - 1u));
    p.queue_out[queue_base + inner_i] = make_uint2(cid, child_base + inner_i);
  }
  if (is_leaf) {
    // a leaf slot holds exactly one primitive (max_leaf == 1); the walk tolerates a larger subtree only to stay in bounds
    uint32_t st[4];
    int sp = 0;
    st[sp++] = cid;
    uint32_t w = 0;
    while (sp) {
      const uint32_t id = st[--sp];
      const BNode bn = p.nodes[id];
      if (id >= n_int) {
        emit_prim(p, f2u(bn.lo.w), prim_base + prim_off + w);
        w++;
      } else {
        st[sp++] = f2u(bn.hi.w);
        st[sp++] = f2u(bn.lo.w);
      }
    }
  }
  // pack: word h (slots 4h .. 4h+3) of bound b
  uint32_t words[12];
#pragma unroll
  for (int b = 0; b < 6; ++b)
#pragma unroll
    for (int h = 0; h < 2; ++h)
      words[2 * b + h] = __reduce_or_sync(FULL, (lane < 8u && (int)(lane >> 2) == h) ? q[b] << (8u * (lane & 3u)) : 0u);
  // byte x of iperm / lperm = inner / leaf mask with the two low bits of every slot index XORed with x (traverse.cuh)
  const uint32_t pj = (lane & 7u) ^ (lane >> 3);
  const uint32_t iperm = __ballot_sync(FULL, (imask >> pj) & 1u);
  const uint32_t lperm = __ballot_sync(FULL, (leafmask >> pj) & 1u);
  if (work.y < p.node_cap && lane < 5u) {
    uint4 v;
    if (lane == 0) v = make_uint4(f2u(nlo.x), f2u(nlo.y), f2u(nlo.z), ex | (ey << 8) | (ez << 16) | (imask << 24));
    else if (lane == 1) v = make_uint4(child_base, prim_base, iperm, lperm);
    else if (lane == 2) v = make_uint4(words[0], words[1], words[2], words[3]);    // lo x, lo y
    else if (lane == 3) v = make_uint4(words[4], words[5], words[6], words[7]);    // lo z, hi x
    else v = make_uint4(words[8], words[9], words[10], words[11]);                 // hi y, hi z
    p.out_nodes[work.y].q[lane] = v;
  }
  if (item == 0 && lane == 0) atomic_max(&p.g->levels, p.level + 1u);
}
#endif

}  // namespace brt
