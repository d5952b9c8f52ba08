// Wavefront restatement of the reference's shader pipeline (SH/ = reference shaders/):
//   raygen      rgenMain prologue                      SH/raytracing.slang:92-112
//   trace       TraceRay closest / occlusion           SH/raytracing.slang:121 / :67   (traverse.cuh)
//   shade       rchitMain + calculateColor + rmissMain SH/raytracing.slang:135-176, :72-88, light.slang:23-39
//   accumulate  `c += payload.color * weight`          SH/raytracing.slang:122
//   resolve     `outImage = float4(c / SAMPLES, 1)`    SH/raytracing.slang:129-132
// One path per pixel per sample; the per-pixel loop `while (depth < depthMax)` (:119-126) becomes one
// trace/shade/occlude/accumulate round per depth over a compacted queue of live paths.
#pragma once
#include "../../include/brt.h"
#include "device_types.cuh"
#include "shading.cuh"
#include "traverse.cuh"

namespace brt {

struct PathQueue {  // structure of arrays, one slot per live path
  float4* o;        // origin xyz, tmin
  float4* d;        // direction xyz, tmax
  float4* w;        // HitPayload.weight widened to RGB (the diffuse-bounce extension tints per channel)
  uint32_t* px;     // path id: tile-order slot of the pixel (low 26 bits) | sample-in-batch (high 6 bits); BRT_MISS for padding
  uint32_t* seed;   // PCG state (SH/random.slang)
};

struct FrameCounters {  // device-side, closest-hit / shade chain
  uint32_t n_paths[2];  // live paths in queue 0 / 1
  uint32_t work_closest;  // dynamic-fetch cursor of the closest-hit kernel
  uint32_t pad;
};
struct ShadowCounters {  // device-side, one per round parity: the occlusion / accumulate chain of a round runs on a second
  uint32_t work_occl;    // stream, overlapped with the next round's closest-hit traversal
  uint32_t n_items;      // path slots of the round (what k_accumulate walks)
  uint32_t pad[2];
  uint32_t n_shadow[BRT_MAX_LIGHTS];  // shadow rays queued per light (segment l of the shadow queue)
};
struct FrameStats {  // device-side, reset per frame
  unsigned long long rays_closest, rays_occlusion;
  unsigned long long nodes_c, prims_c, spheres_c, nodes_o, prims_o, spheres_o;
};

struct TileMap {
  uint32_t width, height;
  uint32_t tiles_x, n_tiles;        // over the whole image
  uint32_t tile_rank, tile_world;
  uint32_t crop_x0, crop_y0, crop_x1, crop_y1;  // traced window [x0,x1) x [y0,y1)
};
#define BRT_MAX_PEERS 16
#define BRT_SLOT_BITS 26
#define BRT_SLOT_MASK 0x03ffffffu
#define BRT_MAX_SAMPLE_BATCH 63u  // sample-in-batch must stay below 63 so that an id never equals BRT_MISS
BRT_HD uint32_t compact5(uint32_t v) {  // even bits of a 10-bit Morton code -> 5 bits
  v &= 0x155u;
  v = (v | (v >> 1)) & 0x133u;
  v = (v | (v >> 2)) & 0x10fu;
  v = (v | (v >> 4)) & 0x01fu;
  return v;
}
// Which rank owns which 32x32 tile: the tiles (row-major ids) are dealt in groups of `world` consecutive ones, and inside group k the ranks are
// rotated by k: the rank's k-th tile is  k * world + (rank + k) % world.  (Plain tile_id % world == rank puts a rank's tiles into COLUMNS
// whenever the number of tile columns is a multiple of world — 4K has 120, so on 8 GPUs every rank rendered vertical stripes and the
// slowest rank of C3 had 2.2 % more work than the mean; rotated, the stripes become diagonals.)
BRT_HD uint32_t tile_of_rank(uint32_t k, uint32_t rank, uint32_t world) { return k * world + (rank + k) % world; }
BRT_HD uint32_t rank_of_tile(uint32_t tile, uint32_t world) { return (tile % world + world - (tile / world) % world) % world; }  // (its k is tile / world)
// path slot -> pixel. Slots are laid out tile by tile (this rank's k-th tile: tile_of_rank),
// Morton order inside the 32x32 tile, so one warp covers an 8x4 pixel block.
BRT_HD bool slot_to_pixel(const TileMap& m, uint32_t slot, uint32_t& x, uint32_t& y) {
  const uint32_t tile = tile_of_rank(slot >> 10, m.tile_rank, m.tile_world);
  if (tile >= m.n_tiles) return false;
  const uint32_t j = slot & 1023u;
  x = (tile % m.tiles_x) * BRT_TILE + compact5(j);
  y = (tile / m.tiles_x) * BRT_TILE + compact5(j >> 1);
  return x >= m.crop_x0 && x < m.crop_x1 && y >= m.crop_y0 && y < m.crop_y1;
}

BRT_HD uint32_t spread5(uint32_t v) {  // 5 bits -> even bits of a 10-bit Morton code
  v &= 0x1fu;
  v = (v | (v << 4)) & 0x10fu;
  v = (v | (v << 2)) & 0x133u;
  v = (v | (v << 1)) & 0x155u;
  return v;
}
// pixel -> path slot of the rank that owns its tile (inverse of slot_to_pixel)
BRT_HD uint32_t pixel_to_slot(const TileMap& m, uint32_t x, uint32_t y) {
  const uint32_t tile = (y / BRT_TILE) * m.tiles_x + x / BRT_TILE;
  return (tile / m.tile_world) * 1024u + (spread5(x % BRT_TILE) | (spread5(y % BRT_TILE) << 1));
}

// ---- raygen --------------------------------------------------------------------------------------
// Values that change from frame to frame (camera, frame counter, sky) are read through a pointer instead of being kernel parameters:
// a captured CUDA graph of the frame can then be replayed with new ones (brt_api.cu).
struct FrameConsts {
  float Vi[16], Pi[16];  // brt_uniform.viewInverse / projInverse, read as row-major M^-1 (RT/RTApp.cpp:44-49)
  uint32_t frame;        // uniform.frame
  uint32_t pad[3];
  brt_sky sky;
  uint32_t gather_seq;   // fused multi-GPU exchange: number of this frame on its gather image (1, 2, ...), see GatherFlags
  uint32_t pad2[3];
  float scene_lo[4];     // world bounds of the scene's instances and 64 / extent per axis: the grid of the hit sort (below)
  float scene_inv[4];
};

struct RaygenParams {
  uint32_t count;
  const uint32_t* count_ptr;
  TileMap map;
  const FrameConsts* fc;
  uint32_t sample0;      // index of the first sample of this batch: sample s of the batch uses RNG frame fc->frame + sample0 + s
  uint32_t flags;
  uint32_t cap;          // path slots per sample; the batch traces count = cap * samples paths at once
  PathQueue q;
};
BRT_HD void raygen_body(const RaygenParams& p, uint32_t i) {
  const uint32_t sib = i / p.cap, slot = i - sib * p.cap;  // sample in batch, tile-order slot
  uint32_t px, py;
  if (!slot_to_pixel(p.map, slot, px, py)) {
    p.q.px[i] = BRT_MISS;
    return;
  }
  const uint32_t frame = p.fc->frame + p.sample0 + sib;
  uint32_t seed = hash3(px, py, frame);  // :96
  float jx = 0.0f, jy = 0.0f;
  if (p.flags & BRT_RENDER_JITTER) {  // :97-98 (the shader computes it and then drops it at :100)
    if (frame == 0u) { jx = 0.5f; jy = 0.5f; }
    else { jx = rnd(seed); jy = rnd(seed); }
  }
  const float* Pi = p.fc->Pi;
  const float* Vi = p.fc->Vi;
  const float cx = ((float)px + jx) / (float)p.map.width * 2.0f - 1.0f;  // :100
  const float cy = ((float)py + jy) / (float)p.map.height * 2.0f - 1.0f;
  const f3 vc = F3(((Pi[0] * cx + Pi[1] * cy) + Pi[2] * 1.0f) + Pi[3] * 1.0f, ((Pi[4] * cx + Pi[5] * cy) + Pi[6] * 1.0f) + Pi[7] * 1.0f,
                   ((Pi[8] * cx + Pi[9] * cy) + Pi[10] * 1.0f) + Pi[11] * 1.0f);  // :101
  const f3 dn = normalize(vc);
  const f3 dir = F3((Vi[0] * dn.x + Vi[1] * dn.y) + Vi[2] * dn.z, (Vi[4] * dn.x + Vi[5] * dn.y) + Vi[6] * dn.z,
                    (Vi[8] * dn.x + Vi[9] * dn.y) + Vi[10] * dn.z);  // :104
  p.q.o[i] = make_float4(Vi[3], Vi[7], Vi[11], 0.001f);               // :105-106
  p.q.d[i] = make_float4(dir.x, dir.y, dir.z, BRT_INFINITE);           // :107
  p.q.w[i] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);                      // :110
  p.q.px[i] = slot | (sib << BRT_SLOT_BITS);
  p.q.seed[i] = seed;
}

// ---- trace ---------------------------------------------------------------------------------------
struct TraceParams {
  uint32_t count;
  const uint32_t* count_ptr;
  const Node8* tlas;
  const InstRec* insts;
  const float4* o;
  const float4* d;
  const uint32_t* px;    // closest: padding slots are skipped; may be null
  float4* hit;           // closest: t, u, v, bits(prim)
  uint32_t* hit_inst;    // closest: instance or BRT_MISS
  const uint32_t* target;  // occlusion: index into contrib
  float4* contrib;         // occlusion: zeroed when the ray is blocked
  uint32_t* work;          // dynamic-fetch cursor (device kernels)
  FrameStats* stats;
  // optional segmented queue: n_segs segments of `seg_stride` slots, segment k holds seg_counts[k] rays.
  // (The shadow queue has one segment per light so that neighbouring lanes trace towards the same light.)
  const uint32_t* seg_counts;
  uint32_t n_segs, seg_stride;
  uint32_t refill_lanes;  // device kernel: idle lanes fetch new rays when fewer than this many lanes are still traversing
  uint32_t defer_lanes;   // device kernel: > 0 = deferred-leaf mode, a pass of parked primitive tests runs once this many lanes have one
};
BRT_HD uint32_t trace_total(const TraceParams& p) {
  if (p.seg_counts) {
    uint32_t n = 0;
    for (uint32_t k = 0; k < p.n_segs; ++k) n += p.seg_counts[k];
    return n;
  }
  return p.count_ptr ? *p.count_ptr : p.count;
}
// work item (0 .. trace_total) -> queue slot
BRT_HD uint32_t trace_slot(const TraceParams& p, uint32_t w) {
  if (!p.seg_counts) return w;
  uint32_t k = 0;
  for (; k + 1 < p.n_segs; ++k) {
    const uint32_t c = p.seg_counts[k];
    if (w < c) break;
    w -= c;
  }
  return k * p.seg_stride + w;
}
// load ray i into a traversal; false for a padding slot (its miss is written right away)
template <bool ANY, bool COUNT>
BRT_HD bool trace_load(const TraceParams& p, uint32_t i, Traversal<ANY, COUNT>& t, uint2* __restrict__ stack) {
  if (!ANY && p.px && p.px[i] == BRT_MISS) {
    p.hit_inst[i] = BRT_MISS;
    return false;
  }
  const float4 o = p.o[i], d = p.d[i];
  t.init(stack, p.tlas, p.insts, F3(o.x, o.y, o.z), F3(d.x, d.y, d.z), o.w, d.w);
  return true;
}
template <bool ANY, bool COUNT>
BRT_HD void trace_store(const TraceParams& p, uint32_t i, const Traversal<ANY, COUNT>& t) {
  if (ANY) {
    if (t.found) p.contrib[p.target[i]] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // shadow factor 0 (SH/raytracing.slang:69)
  } else {
    p.hit[i] = make_float4(t.best.t, t.best.u, t.best.v, u2f(t.best.prim));
    p.hit_inst[i] = t.best.inst;
  }
}

// ---- shade ---------------------------------------------------------------------------------------
struct ShadeParams {
  uint32_t count;
  const uint32_t* count_ptr;
  PathQueue cur, next;
  const float4* hit;
  const uint32_t* hit_inst;
  FrameCounters* ctr;
  ShadowCounters* sctr;
  float4* aux;             // per slot: path weight rgb + pixel (bits), read by k_accumulate
  uint32_t next_slot;      // which n_paths[] entry counts `next`
  uint32_t cap;            // slots per queue == stride of contrib per light
  const InstShade* inst;
  const float* materials;  // 13 floats each (brt_material)
  const float2* mat_ext;   // transmission, ior
  const LightRec* lights;
  const float4* light_bvh;  // BRT_RENDER_LIGHT_BVH: brt_light_bvh_node[ ] as 3 x float4 per node
  uint32_t n_lights;
  float4* contrib;         // [n_lights (>= 1)][cap]
  float4* s_o;             // shadow queue
  float4* s_d;
  uint32_t* s_target;
  uint32_t flags;
  uint32_t last_round;     // no bounce is generated in the last round of the depth loop
  uint32_t write_aov;      // round 0 of the first batch: sample 0 writes the primary-hit AOVs
  TileMap map;
  uint32_t* aov_prim;
  uint32_t* aov_inst;
  float* aov_t;
  float4* aov_pos;  // BRT_RENDER_GBUFFER: world position (w = 1) and shading normal (w = hit distance) of the primary hit, else null
  float4* aov_nrm;
  const FrameConsts* fc;  // sky
  const uint32_t* order;  // bounce rounds: path slots in the order of their hit positions (hit sort, below); null = queue order
  uint32_t ahead;         // CUDA kernels: the head of slot i + ahead is requested into the L2 while slot i is shaded (0 = off)
};

// Light BVH sampling (RT/Scene.h:123-130, SH/raytracing.slang:76; rule in DESIGN.md §13): stochastic descent from the root, a child
// is taken with probability ~ totalFlux / max(squared distance to its box centre, squared half-diagonal of its box); the single
// random number is rescaled at every level. Returns the light index, inv_pdf = 1 / (product of the branch probabilities).
BRT_HD float light_node_importance(const float4 a, const float4 b, const float4 c, f3 P) {  // a = min.xyz, max.x   b = max.yz, flux, cone.x   c = cone.yz, angle, child
  if (c.z == 0.0f) return b.z;  // a subtree of constant-direction lights (cone angle 0): no falloff, the same from everywhere
  const f3 lo = F3(a.x, a.y, a.z), hi = F3(a.w, b.x, b.y);
  const f3 ctr = (lo + hi) * 0.5f, hd = (hi - lo) * 0.5f;
  const f3 d = P - ctr;
  return b.z / fmaxf(fmaxf(dot(d, d), dot(hd, hd)), 1e-12f);
}
BRT_HD uint32_t sample_light_bvh(const float4* __restrict__ nodes, f3 P, float r, float& inv_pdf) {
  uint32_t node = 0;
  float pdf = 1.0f;
  for (int guard = 0; guard < 64; ++guard) {
    const int child = (int)f2u(nodes[3 * (size_t)node + 2].w);
    if (child < 0) break;
    const float w0 = light_node_importance(nodes[3 * (size_t)child], nodes[3 * (size_t)child + 1], nodes[3 * (size_t)child + 2], P);
    const float w1 = light_node_importance(nodes[3 * (size_t)child + 3], nodes[3 * (size_t)child + 4], nodes[3 * (size_t)child + 5], P);
    const float sum = w0 + w1;
    const float p0 = sum > 0.0f ? w0 / sum : 0.5f;
    if (r < p0 || p0 >= 1.0f) {
      node = (uint32_t)child;
      pdf = pdf * p0;
      r = r / p0;
    } else {
      const float q = 1.0f - p0;
      node = (uint32_t)child + 1u;
      pdf = pdf * q;
      r = (r - p0) / q;
    }
  }
  inv_pdf = 1.0f / pdf;
  return (uint32_t)(-1 - (int)f2u(nodes[3 * (size_t)node + 2].w));
}

// Warp-aggregated append: the lanes that execute this together AND target the same counter share one atomic.
// (Grouping by address matters: with independent thread scheduling, lanes that are in different iterations of
// the per-light loop — different counters — can be converged at this instruction; a plain __activemask()
// aggregate would then hand them slots of the wrong queue segment.)
BRT_HD uint32_t append_slot(uint32_t* counter) {
#ifdef BRT_EMU
  return atomic_add(counter, 1u);
#else
  const unsigned peers = __match_any_sync(__activemask(), (unsigned long long)counter);
  const int leader = __ffs(peers) - 1;
  const int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  return base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
#endif
}

// extension (BRT_RENDER_SKY): gradient from the SkyInfo the reference uploads but never reads (RT/Scene.cpp:333-355)
BRT_HD f3 sky_color(const brt_sky& s, f3 dir) {
  const f3 d = normalize(dir);
  const float h = dot(d, F3(s.upDirection[0], s.upDirection[1], s.upDirection[2]));
  const f3 hor = F3(s.horizonColor[0], s.horizonColor[1], s.horizonColor[2]);
  f3 c;
  if (h >= 0.0f)
    c = lerp3(hor, F3(s.skyColor[0], s.skyColor[1], s.skyColor[2]), clampf(h / s.horizonSize, 0.0f, 1.0f));
  else
    c = lerp3(hor, F3(s.groundColor[0], s.groundColor[1], s.groundColor[2]), clampf(-h / s.horizonSize, 0.0f, 1.0f));
  return c * s.brightness;
}

// Shading is split in two so that the CUDA kernel can compact the hits of a block before the expensive part (brt_api.cu):
//   shade_prologue  every path slot: bookkeeping, primary-hit AOVs, and the whole miss shader; returns true for a hit
//   shade_hit       rchitMain: geometry fetch, BRDF x lights, shadow rays, bounce
// The head of a path slot: everything the prologue reads. With BRT_SHADE_HOIST_HEAD it is fetched together BEFORE the first branch (fetched
// where they are used, px -> branch -> hit_inst -> branch -> hit / ray are three dependent trips to DRAM at five blocks per SM: 46 % of
// the stall samples of k_shade_primary are on those waits, profiles/r2_ncu_summary.md §4); hit_inst / hit of a padding slot
// (px == BRT_MISS) are in bounds and unused.
struct PathHead {
  uint32_t px, inst_id;
  float4 hit, w;
};
// The queues are streamed once, at addresses known long in advance, by a kernel that keeps ~20 warps per SM busy with arithmetic: the head
// (and ray) of the slot that is `ahead` slots further on — roughly what the resident blocks reach next — is pulled into the L2 now.
BRT_HD void shade_prefetch(const ShadeParams& p, uint32_t i) {
#ifndef BRT_EMU
  prefetch_l2(p.cur.px + i);
  prefetch_l2(p.hit_inst + i);
  prefetch_l2(p.hit + i);
  prefetch_l2(p.cur.w + i);
  prefetch_l2(p.cur.o + i);
  prefetch_l2(p.cur.d + i);
  prefetch_l2(p.cur.seed + i);
#endif
}
// Both hoists are OFF by default. In the serial schedule they take 7 % off the shade kernels of C3 (16.6 -> 15.4 ms per frame), but with the
// head hoist the frame time of the PRODUCT schedule (two overlapping chains) became bistable per process and box: 121.7 ms in some
// processes, 126.5 ms in others, steady within a process, against a steady 122.9 ms without it (profiles/r2_ncu_summary.md §5).
#ifndef BRT_SHADE_HOIST_HEAD
#define BRT_SHADE_HOIST_HEAD 0
#endif
#ifndef BRT_SHADE_HOIST_HIT
#define BRT_SHADE_HOIST_HIT 0
#endif
BRT_HD PathHead shade_load(const ShadeParams& p, uint32_t i) {
  PathHead h;
  h.px = p.cur.px[i];
  h.w = p.cur.w[i];
#if BRT_SHADE_HOIST_HEAD
  h.inst_id = p.hit_inst[i];
  h.hit = p.hit[i];
#else
  h.inst_id = BRT_MISS;
  h.hit = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (h.px != BRT_MISS) {
    h.inst_id = p.hit_inst[i];
    h.hit = p.hit[i];
  }
#endif
  return h;
}

BRT_HD bool shade_prologue(const ShadeParams& p, uint32_t i, const PathHead& h) {
  const uint32_t px = h.px;
  if (i == 0) p.sctr->n_items = p.count_ptr ? *p.count_ptr : p.count;
  p.aux[i] = make_float4(h.w.x, h.w.y, h.w.z, u2f(px));
  if (px == BRT_MISS) return false;
  const bool lbvh = (p.flags & BRT_RENDER_LIGHT_BVH) != 0u;  // one stochastically chosen light per hit instead of the loop
  const uint32_t n_slots = (p.n_lights && !lbvh) ? p.n_lights : 1u;
  const uint32_t inst_id = h.inst_id;
  const float4 hit = h.hit;
  if (p.write_aov && (px >> BRT_SLOT_BITS) == 0u) {
    uint32_t x, y;
    slot_to_pixel(p.map, px & BRT_SLOT_MASK, x, y);
    const size_t pix = (size_t)y * p.map.width + x;
    p.aov_prim[pix] = inst_id == BRT_MISS ? BRT_MISS : f2u(hit.w);
    p.aov_inst[pix] = inst_id;
    p.aov_t[pix] = inst_id == BRT_MISS ? 0.0f : hit.x;
  }
  if (inst_id == BRT_MISS) {  // rmissMain :172-176
    f3 c = F3(0.0f);
    if (p.flags & BRT_RENDER_SKY) {
      const float4 rd = p.cur.d[i];
      c = sky_color(p.fc->sky, F3(rd.x, rd.y, rd.z));
    }
    p.contrib[i] = make_float4(c.x, c.y, c.z, 0.0f);
    for (uint32_t l = 1; l < n_slots; ++l) p.contrib[(size_t)l * p.cap + i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    return false;
  }
  return true;
}

BRT_HD void shade_hit(const ShadeParams& p, uint32_t i, uint32_t px, uint32_t inst_id, float4 hit) {
  const bool lbvh = (p.flags & BRT_RENDER_LIGHT_BVH) != 0u;
  const uint32_t n_slots = (p.n_lights && !lbvh) ? p.n_lights : 1u;
  // everything that depends on (i, inst_id) only is requested here, ahead of the dependent fetches (indices -> vertices)
  const float4 rd = p.cur.d[i];
#if BRT_SHADE_HOIST_HIT
  uint32_t seed = p.cur.seed[i];
#endif
  const f3 ray_d = F3(rd.x, rd.y, rd.z);
  // rchitMain :136-169
  const InstShade& in = p.inst[inst_id];
  const float4 o2w[3] = {in.o2w[0], in.o2w[1], in.o2w[2]};
  const float4 w2o[3] = {in.w2o[0], in.w2o[1], in.w2o[2]};
  const uint32_t mat_id = in.material;
  Material mat;
#define BRT_LOAD_MATERIAL                                                                                                     \
  {                                                                                                                           \
    const float* m = p.materials + 13 * (size_t)mat_id; /* SH/objects.slang:56-59 */                                          \
    mat.color = F3(m[0], m[1], m[2]);                                                                                         \
    mat.subsurface = m[3]; mat.metallic = m[4]; mat.roughness = m[5]; mat.specular = m[6]; mat.specularTint = m[7];           \
    mat.anisotropic = m[8]; mat.sheen = m[9]; mat.sheenTint = m[10]; mat.clearCoat = m[11]; mat.clearCoatGloss = m[12];       \
  }
#if BRT_SHADE_HOIST_HIT
  BRT_LOAD_MATERIAL
#endif
  f3 pos, nrm;
  if (in.kind == 1u) {
    const float4 ro = p.cur.o[i];  // only the analytic sphere needs the origin
    const f3 oo = xform_point(w2o, F3(ro.x, ro.y, ro.z)), od = xform_dir(w2o, ray_d);
    pos = oo + od * hit.x;
    nrm = (pos - xyz(in.sphere)) * (1.0f / in.sphere.w);
  } else {
    const uint32_t prim = f2u(hit.w);
    const float b0 = 1.0f - hit.y - hit.z, b1 = hit.y, b2 = hit.z;  // :137
    const uint32_t i0 = in.indices[3 * (size_t)prim], i1 = in.indices[3 * (size_t)prim + 1], i2 = in.indices[3 * (size_t)prim + 2];  // SH/objects.slang:30-33
    const float4* v0 = reinterpret_cast<const float4*>(in.vertices + 8 * (size_t)i0);
    const float4* v1 = reinterpret_cast<const float4*>(in.vertices + 8 * (size_t)i1);
    const float4* v2 = reinterpret_cast<const float4*>(in.vertices + 8 * (size_t)i2);
    const float4 a0 = ldg4(v0), a1 = ldg4(v0 + 1), b0_ = ldg4(v1), b1_ = ldg4(v1 + 1), c0 = ldg4(v2), c1 = ldg4(v2 + 1);
    pos = (b0 * F3(a0.x, a0.y, a0.z) + b1 * F3(b0_.x, b0_.y, b0_.z)) + b2 * F3(c0.x, c0.y, c0.z);  // SH/objects.slang:35-41
    nrm = (b0 * F3(a0.w, a1.x, a1.y) + b1 * F3(b0_.w, b1_.x, b1_.y)) + b2 * F3(c0.w, c1.x, c1.y);
  }
  const f3 worldPos = xform_point(o2w, pos);                       // :149
  const f3 worldNormal = normalize(xform_normal(w2o, nrm));        // :150
#if !BRT_SHADE_HOIST_HIT
  BRT_LOAD_MATERIAL
#endif
#undef BRT_LOAD_MATERIAL
  f3 N = normalize(worldNormal);  // :154
  const f3 V = ray_d;             // :155
  bool flipped = false;
  if (dot(N, -V) < 0.0f) { N = -N; flipped = true; }  // :157-158
  if (p.write_aov && p.aov_pos && (px >> BRT_SLOT_BITS) == 0u) {
    uint32_t x, y;
    slot_to_pixel(p.map, px & BRT_SLOT_MASK, x, y);
    const size_t pix = (size_t)y * p.map.width + x;
    p.aov_pos[pix] = make_float4(worldPos.x, worldPos.y, worldPos.z, 1.0f);
    p.aov_nrm[pix] = make_float4(N.x, N.y, N.z, hit.x);
  }
  const Frame fr = make_frame(N);
  const BrdfSetup bs = brdf_setup(mat);
  const f3 Vout = -V;
  // calculateColor :72-88 — the unshadowed contribution of each light goes to contrib[l][slot]; lights
  // whose contribution is not exactly zero get a shadow ray (a zero times the shadow factor is zero
  // either way), which blanks the entry when it is blocked.
#if !BRT_SHADE_HOIST_HIT
  uint32_t seed = p.cur.seed[i];
#endif
  for (uint32_t l = 0; l < n_slots; ++l) {
    f3 contrib = F3(0.0f);
    uint32_t li = l;
    float inv_pdf = 1.0f;
    if (lbvh && p.n_lights) li = sample_light_bvh(p.light_bvh, worldPos, rnd(seed), inv_pdf);
    if (li < p.n_lights) {
      const LightRec lr = p.lights[li];
      f3 ldir;
      float intensity = lr.intensity;
      if ((lr.type & 0xffu) == BRT_LIGHT_POINT) {  // SH/light.slang:23-39
        ldir = F3(lr.pos_colr.x, lr.pos_colr.y, lr.pos_colr.z) - worldPos;
        const float dist = length(ldir);
        intensity /= (dist * dist);
      } else {
        ldir = F3(0.9f, -0.1f, 0.0f);
      }
      if (!(intensity < BRT_LIGHT_TRESHOLD)) {  // :79
        const f3 L = normalize(ldir);
        const f3 color = BRDF(mat, bs, fr, Vout, L);
        contrib = color * F3(lr.pos_colr.w, lr.color_g, lr.color_b) * intensity;  // :83
        if (lbvh) contrib = contrib * inv_pdf;
        if (!(contrib.x == 0.0f && contrib.y == 0.0f && contrib.z == 0.0f)) {
          const f3 so = worldPos + N * 0.0001f;  // testShadow :56-70
          const uint32_t k = l * p.cap + append_slot(&p.sctr->n_shadow[l]);
          p.s_o[k] = make_float4(so.x, so.y, so.z, 0.001f);
          p.s_d[k] = make_float4(L.x, L.y, L.z, length(ldir));
          p.s_target[k] = l * p.cap + i;
        }
      }
    }
    p.contrib[(size_t)l * p.cap + i] = make_float4(contrib.x, contrib.y, contrib.z, 0.0f);
  }
  // bounce (:161-168). With no BOUNCE flag this is the reference verbatim: weight = 0, the path ends.
  const bool any_bounce = (p.flags & (BRT_RENDER_BOUNCE_REFLECT | BRT_RENDER_BOUNCE_REFRACT | BRT_RENDER_BOUNCE_DIFFUSE)) != 0u;
  if (!any_bounce || p.last_round) return;
  const float r1 = rnd(seed), r2 = rnd(seed), r3 = rnd(seed);
  const float4 w4 = p.cur.w[i];
  f3 weight = F3(w4.x, w4.y, w4.z);
  const float2 ext = p.mat_ext[mat_id];  // transmission, ior
  f3 ndir = F3(0.0f);
  f3 norg = worldPos + N * 0.001f;  // :165
  if ((p.flags & BRT_RENDER_BOUNCE_REFRACT) && ext.x > 0.0f) {
    // extension: smooth dielectric, Schlick Fresnel, stochastic reflect / refract
    const float eta = flipped ? ext.y : 1.0f / ext.y;
    const float cosi = dot(N, -V);
    const float k2 = 1.0f - eta * eta * (1.0f - cosi * cosi);
    const float f0 = square((1.0f - ext.y) / (1.0f + ext.y));
    const float Fr = k2 < 0.0f ? 1.0f : schlickFresnel(f0, cosi);
    if (r3 < Fr) {
      ndir = reflect(V, N);
    } else {
      ndir = V * eta + N * (eta * cosi - sqrtf(k2));
      norg = worldPos - N * 0.001f;
      weight = weight * (mat.color * ext.x);
    }
  } else {
    const bool refl = (p.flags & BRT_RENDER_BOUNCE_REFLECT) != 0u, diff = (p.flags & BRT_RENDER_BOUNCE_DIFFUSE) != 0u;
    const bool spec = (refl && diff) ? (r3 < mat.metallic) : refl;
    if (spec) {
      float pdf;
      ndir = sampleGGXVNDFSphericalCap(mat, bs.ap, V, fr, r1, r2, pdf);  // :166
      weight = weight * (diff ? pdf : mat.metallic * pdf);                // :167
    } else if (diff) {
      ndir = toWorld(sampleCosineWeightedHemisphere(r1, r2), fr);  // SH/sampler.slang:53-65 + toWorld
      weight = weight * (refl ? mat.color : mat.color * (1.0f - mat.metallic));
    } else {
      weight = F3(0.0f);
    }
  }
  if (!(weight.x > 0.0f || weight.y > 0.0f || weight.z > 0.0f)) return;  // no weight-0 rays
  const uint32_t k = append_slot(&p.ctr->n_paths[p.next_slot]);
  p.next.o[k] = make_float4(norg.x, norg.y, norg.z, 0.001f);
  p.next.d[k] = make_float4(ndir.x, ndir.y, ndir.z, BRT_INFINITE);
  p.next.w[k] = make_float4(weight.x, weight.y, weight.z, 0.0f);
  p.next.px[k] = px;
  p.next.seed[k] = seed;
}

BRT_HD void shade_body(const ShadeParams& p, uint32_t i) {
  const PathHead h = shade_load(p, i);
  if (shade_prologue(p, i, h)) shade_hit(p, i, h.px, h.inst_id, h.hit);
}

// ---- hit sort (bounce rounds) -------------------------------------------------------------------------
// OPT-IN (BRT_CFG_HIT_SORT), measured and found slower on B200 (profiles/r2_traversal.md). After the first bounce the paths of a
// wavefront hit the scene at unrelated places: the shadow rays and the next bounce rays that the shade kernel emits then start at
// unrelated origins. With this flag the hits of a bounce round are counting-sorted by the cell of their world position (64^3 grid over the
// scene, Morton order, misses last) and the shade kernel walks the paths in that order: neighbouring threads shade neighbouring
// surface points, and both queues it appends to come out in runs of rays with neighbouring origins. Nothing but the ORDER of
// work changes — every result is addressed by its path slot —, so frames stay bit-identical.
//   k_hit_keys  key = cell of the hit, rank = arrival number inside the cell (one warp-aggregated atomic per distinct cell and warp)
//   k_bins_scan exclusive scan of the cell counters (brt_api.cu)
//   k_hit_order order[start(cell) + rank] = path slot
#define BRT_HITSORT_CELL_BITS 6
#define BRT_HITSORT_BINS ((1u << (3 * BRT_HITSORT_CELL_BITS)) + 1u)  // + one bin for misses and padding slots
struct HitSortParams {
  uint32_t count;
  const uint32_t* count_ptr;
  const float4* o;
  const float4* d;
  const uint32_t* px;
  const float4* hit;
  const uint32_t* hit_inst;
  const FrameConsts* fc;
  uint32_t* key;    // per path slot
  uint32_t* rank;   // per path slot
  uint32_t* bins;   // BRT_HITSORT_BINS counters, zeroed before k_hit_keys, scanned before k_hit_order
  uint32_t* order;  // out
};
BRT_HD uint32_t spread_bits3(uint32_t v) {  // 10 bits -> every third bit
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
BRT_HD uint32_t hit_sort_key(const HitSortParams& p, uint32_t i) {
  if (p.px[i] == BRT_MISS || p.hit_inst[i] == BRT_MISS) return BRT_HITSORT_BINS - 1u;
  const float4 o = p.o[i], d = p.d[i];
  const float t = p.hit[i].x;
  const float top = (float)((1u << BRT_HITSORT_CELL_BITS) - 1u);
  const float cx = fminf(fmaxf((o.x + d.x * t - p.fc->scene_lo[0]) * p.fc->scene_inv[0], 0.0f), top);
  const float cy = fminf(fmaxf((o.y + d.y * t - p.fc->scene_lo[1]) * p.fc->scene_inv[1], 0.0f), top);
  const float cz = fminf(fmaxf((o.z + d.z * t - p.fc->scene_lo[2]) * p.fc->scene_inv[2], 0.0f), top);
  return spread_bits3((uint32_t)cx) | (spread_bits3((uint32_t)cy) << 1) | (spread_bits3((uint32_t)cz) << 2);
}
BRT_HD void hit_order_body(const HitSortParams& p, uint32_t i) { p.order[p.bins[p.key[i]] + p.rank[i]] = i; }

// ---- accumulate ------------------------------------------------------------------------------------
// `c += payload.color * weight` (SH/raytracing.slang:122), deferred: round r of sample-in-batch s of a slot writes its
// term into rad[s][r][slot]; k_sum_samples then adds the terms of a slot in the shader's order (samples in order, depths
// in order). That keeps the sum bit-identical to the sequential loop while a whole batch of samples shares one wavefront.
struct AccumParams {
  uint32_t count;
  const uint32_t* count_ptr;
  const float4* aux;  // weight rgb, path id
  const float4* contrib;
  uint32_t n_slots, cap;   // light slots; stride of contrib per light (= slots of the whole batch)
  uint32_t slots, rounds, round;  // path slots per sample, depth rounds per sample, this round
  float4* rad;        // [sample in batch][round][slot]
  uint32_t* alive;    // [sample in batch][slot]: rounds that wrote a term so far (rad is NOT cleared: k_sum_samples reads only these)
};
BRT_HD void accumulate_body(const AccumParams& p, uint32_t i) {
  const float4 w = p.aux[i];
  const uint32_t id = f2u(w.w);
  if (id == BRT_MISS) return;
  const float4 c0 = p.contrib[i];
  f3 c = F3(c0.x, c0.y, c0.z);
  for (uint32_t l = 1; l < p.n_slots; ++l) {
    const float4 cl = p.contrib[(size_t)l * p.cap + i];
    c = c + F3(cl.x, cl.y, cl.z);  // :84, light order
  }
  const uint32_t sib = id >> BRT_SLOT_BITS, slot = id & BRT_SLOT_MASK;
  p.rad[((size_t)sib * p.rounds + p.round) * p.slots + slot] = make_float4(c.x * w.x, c.y * w.y, c.z * w.z, 0.0f);
  p.alive[(size_t)sib * p.slots + slot] = p.round + 1u;  // a path is in the queue of every round up to its last: rounds arrive in order
}

struct SumSamplesParams {
  uint32_t count;  // slots
  const uint32_t* count_ptr;
  uint32_t samples, rounds;
  const float4* rad;
  const uint32_t* alive;  // rounds of (sample, slot) that wrote a term
  float4* accum;   // per slot running sum over all samples so far
};
BRT_HD void sum_samples_body(const SumSamplesParams& p, uint32_t i) {
  float4 a = p.accum[i];
  for (uint32_t s = 0; s < p.samples; ++s) {
    const uint32_t n = p.alive[(size_t)s * p.count + i];  // terms of dead rounds would be + 0: skipped, never written, never cleared
    for (uint32_t r = 0; r < n; ++r) {
      const float4 t = p.rad[((size_t)s * p.rounds + r) * p.count + i];
      a.x = a.x + t.x;  // :122
      a.y = a.y + t.y;
      a.z = a.z + t.z;
    }
  }
  p.accum[i] = a;
}

// float4 -> one packed 8-bit texel of a present format (RT/RTPipeline.cpp:49-55), byte 0 first in memory
BRT_HD uint32_t pack_present(float4 c, uint32_t format) {
  const bool srgb = format == BRT_FORMAT_R8G8B8A8_SRGB || format == BRT_FORMAT_B8G8R8A8_SRGB;
  const bool bgra = format == BRT_FORMAT_B8G8R8A8_UNORM || format == BRT_FORMAT_B8G8R8A8_SRGB;
  const uint32_t r = float_to_unorm8(srgb ? linear_to_srgb(c.x) : c.x);
  const uint32_t g = float_to_unorm8(srgb ? linear_to_srgb(c.y) : c.y);
  const uint32_t b = float_to_unorm8(srgb ? linear_to_srgb(c.z) : c.z);
  const uint32_t a = float_to_unorm8(c.w);
  return (bgra ? b : r) | (g << 8) | ((bgra ? r : b) << 16) | (a << 24);
}

// ---- resolve ---------------------------------------------------------------------------------------
// Fused multi-GPU exchange, completion without the host. Every rank's gather allocation ends in one GatherFlags block that the
// OTHER ranks write through peer memory:
//   arrive[img][src]   = n  once rank `src` has stored all its pixels of the n-th frame on gather image `img` into this rank's image
//                           (written by the last block of src's resolve kernel, release at system scope after every block's fence)
//   consumed[img][dst] = n  once receiver `dst` has finished reading the n-th frame of image `img` (written by dst's release kernel):
//                           frame n + 1 of that image waits for it on the device before its resolve kernel may overwrite the pixels
// A receiver waits for arrive[img][*] >= n with a one-warp kernel on the stream that reads the image (k_gather_wait_arrive). No
// collective, no host barrier, no stream synchronisation per frame.
struct GatherFlags {
  uint32_t arrive[BRT_GATHER_IMAGES][BRT_MAX_PEERS];
  uint32_t consumed[BRT_GATHER_IMAGES][BRT_MAX_PEERS];
  uint32_t timeout;  // set by a wait that gave up (a rank died or never submitted the frame)
  uint32_t pad[3];
};
struct ResolveParams {
  uint32_t count;  // slots (tiles owned * 1024)
  const uint32_t* count_ptr;
  TileMap map;
  float spp;
  const float4* accum;  // per slot
  float4* image;   // full frame, row major
  float4* tiles;   // this rank's tiles packed tile-major, row-major inside a tile (may be null)
  void* peers[BRT_MAX_PEERS];  // fused exchange: the receivers' gather images (peer memory over NVLink), n_peers of them
  GatherFlags* peer_flags[BRT_MAX_PEERS];  // ... and their flag blocks
  uint32_t n_peers;
  uint32_t format;       // of the gather images: BRT_FORMAT_R32G32B32A32_SFLOAT or an 8-bit present format (4 bytes per pixel over NVLink)
  uint32_t img, rank;    // gather image this frame stores into; this rank
  uint32_t* done;        // block arrival counter of k_resolve_peers
  const FrameConsts* fc; // gather_seq
};
BRT_HD void resolve_body(const ResolveParams& p, uint32_t i) {
  // unlike the path slots this walks the tile row by row so that both stores coalesce
  const uint32_t tile = tile_of_rank(i >> 10, p.map.tile_rank, p.map.tile_world);
  if (tile >= p.map.n_tiles) return;
  const uint32_t lx = i & 31u, ly = (i >> 5) & 31u;
  const uint32_t x = (tile % p.map.tiles_x) * BRT_TILE + lx, y = (tile / p.map.tiles_x) * BRT_TILE + ly;
  float4 out = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  const bool inside = x < p.map.width && y < p.map.height;
  if (inside && x >= p.map.crop_x0 && x < p.map.crop_x1 && y >= p.map.crop_y0 && y < p.map.crop_y1) {
    const float4 a = p.accum[(i & ~1023u) | spread5(lx) | (spread5(ly) << 1)];
    out = make_float4(a.x / p.spp, a.y / p.spp, a.z / p.spp, 1.0f);  // :129-132
  }
  if (inside) p.image[(size_t)y * p.map.width + x] = out;
  if (p.tiles) p.tiles[i] = out;
  if (inside && p.n_peers) {  // st.global to mapped peer pointers
    const size_t pix = (size_t)y * p.map.width + x;
    if (p.format == BRT_FORMAT_R32G32B32A32_SFLOAT) {
      for (uint32_t k = 0; k < p.n_peers; ++k) static_cast<float4*>(p.peers[k])[pix] = out;
    } else {
      const uint32_t texel = pack_present(out, p.format);
      for (uint32_t k = 0; k < p.n_peers; ++k) static_cast<uint32_t*>(p.peers[k])[pix] = texel;
    }
  }
}

// ---- present: the render output in the swapchain's format (RT/RTPipeline.cpp:49-55, RT/RTApp.cpp:87-152) ---------
struct PresentParams {
  uint32_t count;  // width * height
  const uint32_t* count_ptr;
  uint32_t format;  // BRT_FORMAT_*
  const float4* image;
  uint32_t* image8;  // one packed texel per pixel, byte 0 first in memory
};
BRT_HD void present_body(const PresentParams& p, uint32_t i) {
  p.image8[i] = pack_present(p.image[i], p.format);
}

// ---- un-tile after the framebuffer gather (root rank) ----------------------------------------------------
struct UntileParams {
  uint32_t count;  // width * height
  const uint32_t* count_ptr;
  uint32_t width, height, tiles_x, tile_world, slots_per_rank;
  const float4* all;  // tile_world packed buffers back to back
  float4* image;
};
BRT_HD void untile_body(const UntileParams& p, uint32_t i) {
  const uint32_t x = i % p.width, y = i / p.width;
  const uint32_t tile = (y / BRT_TILE) * p.tiles_x + x / BRT_TILE;
  const uint32_t rank = rank_of_tile(tile, p.tile_world), k = tile / p.tile_world;
  p.image[i] = p.all[(size_t)rank * p.slots_per_rank + (size_t)k * 1024u + (y % BRT_TILE) * BRT_TILE + (x % BRT_TILE)];
}

// ---- Smart Culling (README.md:15-18; rule defined in DESIGN.md §6) ---------------------------------------
struct CullParams {
  uint32_t count;
  const uint32_t* count_ptr;
  const InstShade* inst;
  const float4* mesh_bounds;
  float eye[3], fwd[3];
  float k;          // P[1][1] * height / 2
  float threshold, hysteresis;
  uint8_t* visible;  // in/out: previous state feeds the hysteresis
};
BRT_HD void cull_body(const CullParams& p, uint32_t i) {
  if (!(p.threshold > 0.0f)) { p.visible[i] = 1; return; }
  const InstShade& s = p.inst[i];
  f3 lo, hi;
  instance_world_box(s.o2w, xyz(p.mesh_bounds[2 * s.mesh]), xyz(p.mesh_bounds[2 * s.mesh + 1]), lo, hi);
  const f3 ctr = (lo + hi) * 0.5f;
  const float r = length((hi - lo) * 0.5f);
  const float z = dot(ctr - F3(p.eye[0], p.eye[1], p.eye[2]), F3(p.fwd[0], p.fwd[1], p.fwd[2]));
  if (z - r <= 0.0f) { p.visible[i] = 1; return; }  // touches the camera plane: keep
  const float rp = (r * p.k) / z;
  const float fp = BRT_PI * (rp * rp);  // footprint of the bounding sphere in px^2
  const bool was = p.visible[i] != 0;
  p.visible[i] = (was ? (fp >= p.threshold * (1.0f - p.hysteresis)) : (fp > p.threshold * (1.0f + p.hysteresis))) ? 1 : 0;
}

}  // namespace brt
