// Data layouts in HBM shared by the builder, the traversal kernels and the shading kernels.
// Everything the traversal fetches is a multiple of 16 bytes and 16-byte aligned so that every load
// is one LDG.128.
#pragma once
#include "compat.cuh"

namespace brt {

// ---- compressed 8-wide BVH node, 80 bytes = 5 x 16 B (layout after Ylitie, Karras, Laine 2017) -------
//  q0: p.x, p.y, p.z (float, origin of the local grid), [ex | ey<<8 | ez<<16 | imask<<24]
//      e* = biased binary32 exponent of the per-axis grid step, imask = slots holding inner nodes
//  q1: child_base (index of the first inner child, children are contiguous in slot order),
//      prim_base  (index of the first primitive record; a leaf slot holds ONE primitive, record index =
//                  prim_base + number of leaf slots below it),
//      iperm, lperm — byte x (x = 0..3) = the inner / leaf slot mask with bit j taken from slot j ^ x, i.e. the
//                  mask in the order a ray with (7 - octant) & 3 == x visits the slots of each half
//  q2: qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7]
//  q3: qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7]
//  q4: qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7]
//  child box = p + q * 2^(e-127) per axis; the builder guarantees it contains the child's real box.
//  Slot s of a node is the child that lies towards (s&1 ? +x : -x, s&2 ? +y : -y, s&4 ? +z : -z), so a
//  ray visits slots in the order of descending (s ^ (7 - octant)) to go front to back.
struct __align__(16) Node8 {
  uint4 q[5];
};
static_assert(sizeof(Node8) == 80, "BVH8 node must be 80 bytes");

// ---- triangle record, 48 bytes = 3 x 16 B, stored in BVH leaf order ---------------------------------
//  v0.xyz | primitive index (bits), v1.xyz | 0, v2.xyz | 0   — original vertices in the mesh's own
//  vertex order (the watertight test and the barycentrics refer to them)
struct __align__(16) TriRec {
  float4 v0, v1, v2;
};
static_assert(sizeof(TriRec) == 48, "triangle record must be 48 bytes");

// ---- instance record fetched by the traversal at a TLAS leaf, 96 bytes = 6 x 16 B --------------------
struct __align__(16) InstRec {
  float4 w2o[3];          // world -> object, row-major 3x4
  const Node8* nodes;     // BLAS nodes (unused for a sphere)
  const TriRec* tris;     // BLAS primitive records
  float4 sphere;          // kind == 1: centre xyz, radius
  uint32_t kind;          // 0 triangles, 1 analytic sphere
  uint32_t inst_id;       // index in the scene's instance table (InstanceIndex())
  uint32_t pad0, pad1;
};
static_assert(sizeof(InstRec) == 96, "instance record must be 96 bytes");

// ---- per-instance shading record (read by the shade kernel only), 144 bytes ---------------------------
struct __align__(16) InstShade {
  float4 o2w[3];
  float4 w2o[3];
  const float* vertices;    // mesh vertex buffer, stride 32 B (RT/Scene.h:28-31)
  const uint32_t* indices;  // mesh index buffer
  float4 sphere;
  uint32_t kind, material, mesh, pad;
};
static_assert(sizeof(InstShade) == 144, "shading record must be 144 bytes");

// ---- binary LBVH node used during the build ---------------------------------------------------------
// ids < n_internal are internal nodes, ids >= n_internal are leaves (sorted position = id - n_internal)
struct __align__(16) BNode {
  float4 lo;  // xyz = bounds min, w = bits(left child id)   | leaf: bits(primitive index)
  float4 hi;  // xyz = bounds max, w = bits(right child id)
};

struct Hit {
  float t, u, v;
  uint32_t inst, prim;  // inst == 0xffffffff: miss
};

// device-side light record, 32 bytes == brt_light (RT/Scene.h:64-75)
struct __align__(16) LightRec {
  float4 pos_colr;        // pos xyz, color.r
  float color_g, color_b, intensity;
  uint32_t type;          // low byte = type
};
static_assert(sizeof(LightRec) == 32, "light record must be 32 bytes");

#define BRT_STACK_SIZE 64       // traversal stack entries (8 B each)
#define BRT_STACK_ALLOC (BRT_STACK_SIZE + 5)  // + the world ray and its slab-test constants parked behind the stack (traverse.cuh)
#define BRT_MAX_TREE_LEVELS 28  // TLAS + deepest BLAS levels the traversal stack is guaranteed to hold (see traverse.cuh)
#define BRT_MAX_LIGHTS 16
#define BRT_MISS 0xffffffffu
#define BRT_TILE 32        // image tile edge in pixels (multi-GPU partition)

}  // namespace brt
