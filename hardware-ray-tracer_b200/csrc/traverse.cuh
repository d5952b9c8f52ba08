// Software replacement for TraceRay (SH/raytracing.slang:67,121; RT cores on the reference's
// hardware): two-level traversal of compressed 8-wide BVHs, one ray per thread.
//
// Algorithm: stack of 8-byte "groups" after Ylitie, Karras & Laine (HPG 2017). A node group is
// (child_base, hits<<24 | imask): the not-yet-visited inner children of one node, visited in
// descending bit order, which the octant trick turns into front-to-back order. A leaf group is
// (prim_base, hit bits | leaf slot mask << 8): the leaf slots (one primitive each) whose box the ray hit. TLAS leaves are instances: the
// ray is transformed to object space (t is preserved, the direction is not renormalised) and the
// BLAS is traversed on the same stack; when the stack drops back to the entry level the world ray
// is restored.
//
// Tie-break: the closest hit is the lexicographic minimum of (t, instance, primitive),
// so the result does not depend on traversal order; boxes are culled with `entry <= best t` so that
// equal-t candidates are still visited.
#pragma once
#include "device_types.cuh"
#include "intersect.cuh"

namespace brt {

struct TraceCounters {
  uint32_t nodes, prims, spheres;
};

struct RayBox {   // per-ray constants of the quantised slab test, valid for one coordinate space
  f3 idir;        // 1/d with |d| clamped away from zero
  uint32_t octinv;  // 7 - octant, octant bit k = (d_k < 0)
};
BRT_HD RayBox make_raybox(f3 d) {
  RayBox rb;
  const float tiny = 8.2718061e-25f;  // 2^-80
  float dx = fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x);
  float dy = fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y);
  float dz = fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z);
  // the slab test is conservative by construction (slack term), so the approximate reciprocal (1 ulp, one MUFU)
  // is good enough here; the primitive tests keep IEEE division
  rb.idir = F3(fast_rcp(dx), fast_rcp(dy), fast_rcp(dz));
  uint32_t oct = (dx < 0.0f ? 1u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 4u : 0u);
  rb.octinv = 7u - oct;
  return rb;
}

// PRMT: result byte i = byte (nibble i of sel) of the 8-byte pool {a: 0..3, b: 4..7}
BRT_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef BRT_EMU
  const uint64_t pool = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((pool >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
  return r;
#else
  return __byte_perm(a, b, sel);
#endif
}

// One byte of w -> the float 32768 + byte with ONE byte permute: 0x47000000 is 2^15 and byte 1 of a binary32 with that
// exponent has weight 1. (An I2F.U8 conversion would run on the quarter-rate XU pipe; ncu showed it as the busiest
// pipe of the first version of this kernel.) sel = 0x7604 | (byte index << 4); the byte index is per-ray data, see below.
BRT_HD float magic_byte(uint32_t w, uint32_t sel) { return u2f(byte_perm(w, 0x47000000u, sel)); }

// Tests the ray against the 8 quantised child boxes of one node.
// out: G  = (child_base, inner hits << 24 | inner slot mask)        hits in visiting order: bit p <-> slot p ^ octinv
//      Gt = (prim_base, leaf hits | leaf slot mask << 8)            (descending p = front to back)
//
// Front-to-back order is "descending slot ^ octinv". The XOR is applied to the DATA instead of to the result bits:
// the two low bits of octinv pick which byte of each 4-slot word a PRMT converts (a per-ray selector, no extra
// instruction), the node stores its slot masks pre-permuted for the four possible values (iperm / lperm), and bit 2
// swaps the two 4-slot halves of the result. The eight comparisons then land on constant bit positions.
BRT_HD void intersect_node(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4, const RayBox& rb, f3 o, float tmin,
                           float tmax, uint2& G, uint2& Gt) {
  const float px = u2f(n0.x) - o.x, py = u2f(n0.y) - o.y, pz = u2f(n0.z) - o.z;
  const float sx = u2f((n0.w & 0xffu) << 23), sy = u2f(((n0.w >> 8) & 0xffu) << 23), sz = u2f(((n0.w >> 16) & 0xffu) << 23);
  const float idx = sx * rb.idir.x, idy = sy * rb.idir.y, idz = sz * rb.idir.z;
  // Rounding slack. The slab arithmetic below and the (differently rounded) primitive tests both
  // carry errors of a few ulps of the largest coordinate difference involved; widening every slab
  // by 2^-21 of that magnitude keeps the box test conservative with respect to the primitive tests.
  // The 2^-8 grid-step term covers the rounding of the (origin - 32768 * step) constants below.
  const float mag = fmaxf(fmaxf(fabsf(px), fabsf(py)), fabsf(pz)) + 256.0f * fmaxf(fmaxf(sx, sy), sz);
  const float slack = mag * 4.76837158e-07f;
#ifndef BRT_SLACK_STEPS
#define BRT_SLACK_STEPS 0.00390625f  // 2^-8 grid steps (experiment: larger values emulate a less precise slab test, profiles/r2_traversal.md)
#endif
  const float ex = fma_rn(fabsf(idx), BRT_SLACK_STEPS, slack * fabsf(rb.idir.x));
  const float ey = fma_rn(fabsf(idy), BRT_SLACK_STEPS, slack * fabsf(rb.idir.y));
  const float ez = fma_rn(fabsf(idz), BRT_SLACK_STEPS, slack * fabsf(rb.idir.z));
  const float ox = px * rb.idir.x, oy = py * rb.idir.y, oz = pz * rb.idir.z;
  // t = (32768 + q) * step + (origin - 32768 * step)
  const float ox0 = fma_rn(-32768.0f, idx, ox - ex), ox1 = fma_rn(-32768.0f, idx, ox + ex);
  const float oy0 = fma_rn(-32768.0f, idy, oy - ey), oy1 = fma_rn(-32768.0f, idy, oy + ey);
  const float oz0 = fma_rn(-32768.0f, idz, oz - ez), oz1 = fma_rn(-32768.0f, idz, oz + ez);
  const bool nx = rb.idir.x < 0.0f, ny = rb.idir.y < 0.0f, nz = rb.idir.z < 0.0f;
  const uint32_t x = rb.octinv & 3u;
  const uint32_t sel0 = 0x7604u | (x << 4), sel1 = sel0 ^ 0x10u, sel2 = sel0 ^ 0x20u, sel3 = sel0 ^ 0x30u;
  // The eight hit bits are collected in the mantissa of 2^23 with FSET + FFMA: the comparisons, the min / max and the byte permutes
  // all run on the ALU pipe, the busiest one of this kernel (ncu: 65-69 %), the FMA pipe has room (29 %).
  float hitf = 8388608.0f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t lox4 = h ? n2.y : n2.x, loy4 = h ? n2.w : n2.z, loz4 = h ? n3.y : n3.x;
    const uint32_t hix4 = h ? n3.w : n3.z, hiy4 = h ? n4.y : n4.x, hiz4 = h ? n4.w : n4.z;
    const uint32_t nearx4 = nx ? hix4 : lox4, farx4 = nx ? lox4 : hix4;
    const uint32_t neary4 = ny ? hiy4 : loy4, fary4 = ny ? loy4 : hiy4;
    const uint32_t nearz4 = nz ? hiz4 : loz4, farz4 = nz ? loz4 : hiz4;
#define BRT_SLOT(I, SEL)                                                                                              \
  {                                                                                                                   \
    const float t0x = fma_rn(magic_byte(nearx4, SEL), idx, ox0), t1x = fma_rn(magic_byte(farx4, SEL), idx, ox1);       \
    const float t0y = fma_rn(magic_byte(neary4, SEL), idy, oy0), t1y = fma_rn(magic_byte(fary4, SEL), idy, oy1);       \
    const float t0z = fma_rn(magic_byte(nearz4, SEL), idz, oz0), t1z = fma_rn(magic_byte(farz4, SEL), idz, oz1);       \
    const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));                                                        \
    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tmax));                                                        \
    hitf = fma_rn(tn <= tf ? 1.0f : 0.0f, (float)(1u << (4 * h + I)), hitf);                                          \
  }
    BRT_SLOT(0, sel0) BRT_SLOT(1, sel1) BRT_SLOT(2, sel2) BRT_SLOT(3, sel3)
#undef BRT_SLOT
  }
  // byte 0: inner hits, byte 1: leaf hits (empty slots are in neither mask)
  const uint32_t hits = f2u(hitf) & 0xffu;
  uint32_t v = (hits * 0x0101u) & byte_perm(n1.z, n1.w, 0x40u | (x * 0x11u));
  if (rb.octinv & 4u) v = ((v & 0x0f0fu) << 4) | ((v >> 4) & 0x0f0fu);
  G = make_uint2(n1.x, (v << 24) | (n0.w >> 24));
  Gt = make_uint2(n1.y, (v >> 8) | ((n1.w & 0xffu) << 8));
}

// One ray's traversal as a resumable state machine: init(), then step() until it returns true.
// The CUDA kernel (brt_api.cu) keeps one Traversal per lane and refills finished lanes with new rays;
// the host emulation simply loops.
// ANY: occlusion query (RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH | SKIP_CLOSEST_HIT_SHADER) — ends at
// the first primitive with tmin < t < tmax. Otherwise: closest-hit query, `best` is filled.
template <bool ANY, bool COUNT>
struct Traversal {
  const Node8* tlas;
  const InstRec* insts;
  float tmin;
  f3 co;    // origin in the current space (world, or the object space of cur_inst)
  RayBox rb;
  RayShear sh;
  const Node8* nodes;
  const TriRec* tris;
  uint32_t cur_inst;
  uint2 G, Gt;
  int sp, blas_sp;  // blas_sp >= 0 while inside a BLAS: stack height at entry
  bool found;
  Hit best;
  // The stack (uint2[BRT_STACK_ALLOC]) is a separate local array owned by the caller: keeping the dynamically indexed
  // array out of this struct lets the compiler hold every other member in registers. Its last three entries hold the
  // WORLD ray (origin, direction): it is only needed when an instance is entered or left, so it lives in local memory
  // instead of six registers (the kernel runs at the register limit of 7 blocks per SM).
  static BRT_HDM void store_world(uint2* __restrict__ stack, f3 o_, f3 d_) {
    stack[BRT_STACK_SIZE + 0] = make_uint2(f2u(o_.x), f2u(o_.y));
    stack[BRT_STACK_SIZE + 1] = make_uint2(f2u(o_.z), f2u(d_.x));
    stack[BRT_STACK_SIZE + 2] = make_uint2(f2u(d_.y), f2u(d_.z));
  }
  static BRT_HDM void load_world(const uint2* __restrict__ stack, f3& o_, f3& d_) {
    const uint2 a = stack[BRT_STACK_SIZE + 0], b = stack[BRT_STACK_SIZE + 1], c = stack[BRT_STACK_SIZE + 2];
    o_ = F3(u2f(a.x), u2f(a.y), u2f(b.x));
    d_ = F3(u2f(b.y), u2f(c.x), u2f(c.y));
  }
  // ... and so do the world ray's slab-test constants: leaving a BLAS reloads them (two 8-byte loads) instead of recomputing
  // three reciprocals and the octant
  static BRT_HDM void store_world_raybox(uint2* __restrict__ stack, const RayBox& rb_) {
    stack[BRT_STACK_SIZE + 3] = make_uint2(f2u(rb_.idir.x), f2u(rb_.idir.y));
    stack[BRT_STACK_SIZE + 4] = make_uint2(f2u(rb_.idir.z), rb_.octinv);
  }
  BRT_HDM void restore_world(const uint2* __restrict__ stack) {
    const uint2 a = stack[BRT_STACK_SIZE + 0], b = stack[BRT_STACK_SIZE + 1];
    const uint2 e = stack[BRT_STACK_SIZE + 3], f = stack[BRT_STACK_SIZE + 4];
    co = F3(u2f(a.x), u2f(a.y), u2f(b.x));
    rb.idir = F3(u2f(e.x), u2f(e.y), u2f(f.x));
    rb.octinv = f.y;
  }

  // Experiment (profiles/r2_traversal.md): -DBRT_SMEM_STACK=K keeps the K bottom entries of every lane's stack in shared memory
  // (column threadIdx.x of a [K][128] array of 8-byte words: conflict-free), deeper entries stay in local memory.
#if defined(BRT_SMEM_STACK) && !defined(BRT_EMU)
  uint2* sst;  // this lane's column
  BRT_HDM void push(uint2* __restrict__ stack, uint2 v) {
    if (sp < BRT_SMEM_STACK) sst[sp * 128] = v;
    else stack[sp] = v;
    ++sp;
  }
  BRT_HDM uint2 pop(const uint2* __restrict__ stack) {
    --sp;
    return sp < BRT_SMEM_STACK ? sst[sp * 128] : stack[sp];
  }
#else
  BRT_HDM void push(uint2* __restrict__ stack, uint2 v) { stack[sp++] = v; }
  BRT_HDM uint2 pop(const uint2* __restrict__ stack) { return stack[--sp]; }
#endif

  // Experiment (profiles/r2_traversal.md): -DBRT_NODE_PREFETCH stages the NEXT node through shared memory. As soon as a node test has
  // produced its hit groups the node visited next is known (the nearest hit inner child, else the nearest child of the group on top
  // of the stack): its 80 bytes are requested with five 16-byte cp.async into this lane's column of a [5][128] shared array while the
  // lane runs its primitive tests, and the next step reads them with LDS.128 instead of waiting for five dependent LDG.128.
#if defined(BRT_NODE_PREFETCH) && !defined(BRT_EMU)
  uint4* snode;         // this lane's column of the staging array
  const Node8* pf_node;  // node whose words are (being) staged, or null
  BRT_HDM void prefetch_node(const Node8* np) {
    asm volatile("cp.async.wait_all;" ::: "memory");  // an older request into the same column must have landed
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(snode);
#pragma unroll
    for (int j = 0; j < 5; ++j)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s0 + (unsigned)j * 2048u), "l"(&np->q[j]) : "memory");
    pf_node = np;
  }
  // the node the next step will visit, given the groups as they stand (null: a leaf group, or nothing left)
  BRT_HDM const Node8* next_node(const uint2* __restrict__ stack) const {
    uint2 g = G;
    const Node8* base = nodes;
    if (!(g.y & 0xff000000u)) {
      if (sp == 0) return nullptr;
      if (blas_sp >= 0 && sp == blas_sp) base = tlas;  // the BLAS is exhausted: the pop returns to the TLAS
      g = stack[sp - 1];
      if (!(g.y & 0xff000000u)) return nullptr;
    }
    const int bit = 31 - clz32(g.y);
    const uint32_t slot = (uint32_t)(bit - 24) ^ ((blas_sp >= 0 && base == tlas) ? f2u_octinv_world(stack) : rb.octinv);
    const uint32_t rel = popc((g.y & ~(1u << bit)) & ~(0xffffffffu << slot));
    return base + (uint32_t)(g.x + rel);
  }
  static BRT_HDM uint32_t f2u_octinv_world(const uint2* __restrict__ stack) { return stack[BRT_STACK_SIZE + 4].y; }
#endif

  BRT_HDM void init(uint2* __restrict__ stack, const Node8* tlas_, const InstRec* insts_, f3 o_, f3 d_, float tmin_, float tmax_) {
    tlas = tlas_;
    insts = insts_;
    store_world(stack, o_, d_);
    tmin = tmin_;
    best.t = tmax_;
    best.u = 0.0f;
    best.v = 0.0f;
    best.inst = BRT_MISS;
    best.prim = BRT_MISS;
    found = false;
    sp = 0;
    blas_sp = -1;
#if defined(BRT_NODE_PREFETCH) && !defined(BRT_EMU)
    pf_node = nullptr;
#endif
    co = o_;
    rb = make_raybox(d_);
    store_world_raybox(stack, rb);
    // (no shear constants for the world ray: triangles are only tested inside a BLAS, whose entry computes them for the object-space ray)
    sh.kz = 0;
    sh.Sx = sh.Sy = sh.Sz = 0.0f;
    nodes = tlas_;
    tris = nullptr;
    cur_inst = 0;
    G = make_uint2(0u, tlas_ ? 0x80000000u : 0u);
    Gt = make_uint2(0u, 0u);
  }

  // Tests ONE pending leaf slot of Gt (a triangle, or an instance: sphere test / BLAS entry). Returns true when that ends the ray
  // (an occlusion query found its hit).
  BRT_HDM bool leaf_visit(uint2* __restrict__ stack, TraceCounters& ctr) {
    // hit bit p <-> slot p ^ octinv; a leaf slot's record is prim_base + its rank among the node's leaf slots
    const uint32_t slot = (uint32_t)(ffs32(Gt.y) - 1) ^ rb.octinv;
    Gt.y &= Gt.y - 1u;
    const uint32_t rank = (uint32_t)popc((Gt.y >> 8) & ~(0xffffffffu << slot));
    if (blas_sp >= 0) {
      const TriRec* tr = tris + (uint32_t)(Gt.x + rank);
      const float4 a = ldg4(&tr->v0), b = ldg4(&tr->v1), c = ldg4(&tr->v2);
      if (COUNT) ctr.prims++;
      float t, u, v;
      // (an occlusion query ends at its first hit, so `found` is still false here: the interval test is a compile-time choice)
      if (intersect_tri(co, sh, tmin, best.t, ANY ? false : found, xyz(a), xyz(b), xyz(c), t, u, v)) {
        found = true;  // (for ANY this is the answer)
        if (ANY) return true;
        const uint32_t prim = f2u(a.w);
        if (best.inst == BRT_MISS || t < best.t || cur_inst < best.inst || (cur_inst == best.inst && prim < best.prim)) {
          best.t = t; best.u = u; best.v = v; best.inst = cur_inst; best.prim = prim;
        }
      }
    } else {
      const InstRec* ir = insts + (uint32_t)(Gt.x + rank);
      const float4 m0 = ldg4(&ir->w2o[0]), m1 = ldg4(&ir->w2o[1]), m2 = ldg4(&ir->w2o[2]);
      const uint4 tail = ldg4(reinterpret_cast<const uint4*>(&ir->kind));
      const float4 m[3] = {m0, m1, m2};
      f3 o, d;
      load_world(stack, o, d);
      const f3 oo = xform_point(m, o), od = xform_dir(m, d);
      if (tail.x == 1u) {  // analytic sphere
        const float4 s = ldg4(&ir->sphere);
        if (COUNT) ctr.spheres++;
        float t;
        if (intersect_sphere(oo, od, tmin, best.t, ANY ? false : found, xyz(s), s.w, t)) {
          found = true;
          if (ANY) return true;
          if (best.inst == BRT_MISS || t < best.t || tail.y < best.inst) {
            best.t = t; best.u = 0.0f; best.v = 0.0f; best.inst = tail.y; best.prim = 0u;
          }
        }
      } else {
        // enter the BLAS: keep the TLAS continuation on the stack
        if (Gt.y & 0xffu) push(stack, Gt);
        if (G.y & 0xff000000u) push(stack, G);
        blas_sp = sp;
        const uint4 ptrs = ldg4(reinterpret_cast<const uint4*>(&ir->nodes));
        nodes = reinterpret_cast<const Node8*>(((uint64_t)ptrs.y << 32) | ptrs.x);
        tris = reinterpret_cast<const TriRec*>(((uint64_t)ptrs.w << 32) | ptrs.z);
        cur_inst = tail.y;
        co = oo;
        rb = make_raybox(od);
        sh = make_shear(od);
        G = make_uint2(0u, 0x80000000u);
        Gt = make_uint2(0u, 0u);
      }
    }
    return false;
  }

  // ---- deferred-leaf mode (the warp-synchronous loop of k_trace, bounce rounds) ---------------------------------------------------
  // In an incoherent wavefront a lane reaches a leaf slot only every fourth node step, so the primitive tests of step() ran with 3 of 32
  // lanes and took 37 % of the kernel's issue slots (ncu source view, profiles/r2_traversal.md). Here a lane PARKS its leaf hits in Gt
  // and keeps visiting nodes; the warp votes, and once enough lanes have a parked leaf (or nobody has node work left) all of them test
  // one primitive each in the same pass. Results do not depend on the order of the tests (closest hit = lexicographic minimum).
  BRT_HDM bool can_node() const { return (G.y & 0xff000000u) != 0u; }
  BRT_HDM bool leaf_pending() const { return (Gt.y & 0xffu) != 0u; }
  BRT_HDM void node_visit(uint2* __restrict__ stack, TraceCounters& ctr) {
    const int bit = 31 - clz32(G.y);
    G.y &= ~(1u << bit);
    if (G.y & 0xff000000u) push(stack, G);
    const uint32_t slot = (uint32_t)(bit - 24) ^ rb.octinv;
    const uint32_t rel = popc(G.y & ~(0xffffffffu << slot));
    if (COUNT) ctr.nodes++;
    const Node8* np = nodes + (uint32_t)(G.x + rel);
    uint2 nGt;
    intersect_node(ldg4(&np->q[0]), ldg4(&np->q[1]), ldg4(&np->q[2]), ldg4(&np->q[3]), ldg4(&np->q[4]), rb, co, tmin, best.t, G, nGt);
    if (nGt.y & 0xffu) {
      if (Gt.y & 0xffu) push(stack, Gt);  // an older parked group goes to the stack (it is popped before anything below it)
      Gt = nGt;
    }
  }
  // both registers empty: leave an exhausted BLAS, fetch the next group from the stack; true when nothing is left
  BRT_HDM bool advance(uint2* __restrict__ stack) {
    if ((G.y & 0xff000000u) || (Gt.y & 0xffu)) return false;
    if (blas_sp >= 0 && sp == blas_sp) {
      blas_sp = -1;
      nodes = tlas;
      restore_world(stack);
    }
    if (sp == 0) return true;
    const uint2 g = pop(stack);
    if (g.y & 0xff000000u) {
      G = g;
    } else {
      Gt = g;
      G = make_uint2(0u, 0u);
    }
    return false;
  }

  // One round: visit at most one node, then the pending leaf group. Returns true when the ray is done.
  BRT_HDM bool step(uint2* __restrict__ stack, TraceCounters& ctr) {
    if (G.y & 0xff000000u) {
      const int bit = 31 - clz32(G.y);
      G.y &= ~(1u << bit);
      if (G.y & 0xff000000u) push(stack, G);
      const uint32_t slot = (uint32_t)(bit - 24) ^ rb.octinv;
      const uint32_t rel = popc(G.y & ~(0xffffffffu << slot));
      if (COUNT) ctr.nodes++;
      const Node8* np = nodes + (uint32_t)(G.x + rel);  // 32-bit index: one IMAD.WIDE
#if defined(BRT_NODE_PREFETCH) && !defined(BRT_EMU)
      uint4 n0, n1, n2, n3, n4;
      if (pf_node == np) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        n0 = snode[0]; n1 = snode[128]; n2 = snode[256]; n3 = snode[384]; n4 = snode[512];
      } else {
        n0 = ldg4(&np->q[0]); n1 = ldg4(&np->q[1]); n2 = ldg4(&np->q[2]); n3 = ldg4(&np->q[3]); n4 = ldg4(&np->q[4]);
      }
      intersect_node(n0, n1, n2, n3, n4, rb, co, tmin, best.t, G, Gt);
      if (!(Gt.y & 0xffu) || blas_sp >= 0) {  // (an instance leaf may enter a BLAS and change everything: no request then)
        const Node8* nx = next_node(stack);
        if (nx) prefetch_node(nx);
      }
#else
      intersect_node(ldg4(&np->q[0]), ldg4(&np->q[1]), ldg4(&np->q[2]), ldg4(&np->q[3]), ldg4(&np->q[4]), rb, co, tmin, best.t, G, Gt);
#endif
    } else {
      Gt = G;
      G = make_uint2(0u, 0u);
    }

    while (Gt.y & 0xffu) {
      if (leaf_visit(stack, ctr)) return true;
    }

    if (!(G.y & 0xff000000u)) {
      if (blas_sp >= 0 && sp == blas_sp) {  // BLAS exhausted: back to the world ray
        blas_sp = -1;
        nodes = tlas;
        restore_world(stack);
      }
      if (sp == 0) return true;
      G = pop(stack);
    }
    return false;
  }
};

}  // namespace brt
