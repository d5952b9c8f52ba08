// libbrt.so — the C ABI of include/brt.h over the hand-written sm_100a kernels.
// Host side of the drop-in boundary: what the reference does in RayTracing::Scene (RT/Scene.cpp),
// RayTracing::Pipeline (RT/RTPipeline.cpp) and the uniform block of RTApp::run (RT/RTApp.cpp:44-49).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "../../include/brt.h"
#include "builder.h"
#include "render_kernels.cuh"
#include "shade_kernels.cuh"
#include "denoise.cuh"

namespace brt {

#ifndef BRT_EMU
void launch_shade_fast(const ShadeParams& sp, uint32_t grid, bool primary, cudaStream_t stream);  // shade_fast.cu
#endif

// ---- kernels ---------------------------------------------------------------------------------------
BRT_KERNEL_1D(k_raygen, RaygenParams, raygen_body)
#ifndef BRT_EMU
// hit sort of the bounce rounds (render_kernels.cuh): keys + per-cell arrival ranks, scan of the cell counters, scatter
__global__ void __launch_bounds__(256) k_hit_keys(const HitSortParams p) {
  const uint32_t n = p.count_ptr ? *p.count_ptr : p.count;
  const uint32_t stride = gridDim.x * blockDim.x;
  const unsigned lane = threadIdx.x & 31u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t key = hit_sort_key(p, i);
    const unsigned peers = __match_any_sync(__activemask(), key);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(&p.bins[key], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    p.key[i] = key;
    p.rank[i] = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
  }
}
// in-place exclusive scan of `count` counters by one block of 1024 threads
__global__ void __launch_bounds__(1024) k_bins_scan(uint32_t* __restrict__ bins, uint32_t count) {
  __shared__ uint32_t warp_sums[32];
  const uint32_t per = (count + 1023u) / 1024u;
  const uint32_t begin = min(threadIdx.x * per, count), end = min(begin + per, count);
  uint32_t sum = 0;
  for (uint32_t i = begin; i < end; ++i) sum += bins[i];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t incl = sum;
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
    if ((int)lane >= off) incl += v;
  }
  if (lane == 31u) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = warp_sums[lane];
    uint32_t wi = w;
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, wi, off);
      if ((int)lane >= off) wi += v;
    }
    warp_sums[lane] = wi - w;
  }
  __syncthreads();
  uint32_t run = warp_sums[warp] + incl - sum;
  for (uint32_t i = begin; i < end; ++i) {
    const uint32_t v = bins[i];
    bins[i] = run;
    run += v;
  }
}
BRT_KERNEL_1D(k_hit_order, HitSortParams, hit_order_body)
#endif
BRT_KERNEL_1D(k_accumulate, AccumParams, accumulate_body)
BRT_KERNEL_1D(k_sum_samples, SumSamplesParams, sum_samples_body)
BRT_KERNEL_1D(k_resolve, ResolveParams, resolve_body)
BRT_KERNEL_1D(k_untile, UntileParams, untile_body)
#ifndef BRT_EMU
// measurement aid (brt_debug_l2_read_gbs): streaming read of a buffer that fits the L2 with 16-byte loads, four in flight per thread
__global__ void __launch_bounds__(256) k_l2_read(const uint4* __restrict__ p, uint32_t n16, uint32_t* __restrict__ out) {
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint4 a = make_uint4(0, 0, 0, 0);
  for (; i + 3u * stride < n16; i += 4u * stride) {
    const uint4 v0 = ldg4(p + i), v1 = ldg4(p + i + stride), v2 = ldg4(p + i + 2u * stride), v3 = ldg4(p + i + 3u * stride);
    a.x ^= v0.x ^ v1.x ^ v2.x ^ v3.x; a.y ^= v0.y ^ v1.y ^ v2.y ^ v3.y;
    a.z ^= v0.z ^ v1.z ^ v2.z ^ v3.z; a.w ^= v0.w ^ v1.w ^ v2.w ^ v3.w;
  }
  for (; i < n16; i += stride) {
    const uint4 v = ldg4(p + i);
    a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w;
  }
  if ((a.x ^ a.y ^ a.z ^ a.w) == 0x9e3779b9u) *out = a.x;  // keeps the loads alive; practically never true
}
// ---- fused multi-GPU exchange: resolve + peer stores + device-side completion flags (GatherFlags, render_kernels.cuh) -------------
#ifndef BRT_GATHER_POLL_NS
#define BRT_GATHER_POLL_NS 100
#endif
#ifndef BRT_GATHER_TIMEOUT_NS
#define BRT_GATHER_TIMEOUT_NS 20000000000ull  // a wait gives up after 20 s (a rank died or never submitted the frame) instead of hanging the GPU
#endif
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// lane k < n spins until flags[k] has reached seq (sequence numbers compared modulo 2^32)
__device__ void wait_flags(const uint32_t* flags, uint32_t n, uint32_t seq, uint32_t* timeout_flag) {
  if (threadIdx.x < n) {
    const unsigned long long t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - seq) < 0) {
      if (global_ns() - t0 > BRT_GATHER_TIMEOUT_NS) {
        *timeout_flag = 1u;
        break;
      }
      __nanosleep(BRT_GATHER_POLL_NS);
    }
  }
}
// producer, inside the frame: the receivers [first, first + n) must have released the previous frame of this gather image
__global__ void __launch_bounds__(32) k_gather_wait_consumed(GatherFlags* mine, uint32_t img, const FrameConsts* fc, uint32_t first, uint32_t n) {
  wait_flags(&mine->consumed[img][first], n, fc->gather_seq - 1u, &mine->timeout);
}
// receiver, on the stream that reads the image: every rank's pixels of frame `seq` of this gather image have landed
__global__ void __launch_bounds__(32) k_gather_wait_arrive(GatherFlags* mine, uint32_t img, uint32_t seq, uint32_t n_src) {
  wait_flags(&mine->arrive[img][0], n_src, seq, &mine->timeout);
}
struct GatherReleaseParams {
  GatherFlags* peer_flags[BRT_MAX_PEERS];
  uint32_t n, img, seq, me;
};
// receiver: frame `seq` of this gather image has been read, the producers may overwrite it
__global__ void __launch_bounds__(32) k_gather_release(const GatherReleaseParams p) {
  if (threadIdx.x < p.n) st_release_sys(&p.peer_flags[threadIdx.x]->consumed[p.img][p.me], p.seq);
}
// resolve + exchange in one kernel: every owned pixel goes to the local image and, through NVLink peer stores, to the receivers' gather
// images; the block that finishes last publishes arrive[img][rank] = seq in every receiver's flag block.
__global__ void __launch_bounds__(256) k_resolve_peers(const ResolveParams p) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.count; i += stride) resolve_body(p, i);
  __threadfence_system();  // this thread's peer stores are ordered before the block's arrival below, at system scope
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(p.done, 1u);
    if (prev == gridDim.x - 1u) {
      __threadfence_system();
      *p.done = 0u;  // for the next frame of this slot (replayed graph)
      const uint32_t seq = p.fc->gather_seq;
      for (uint32_t k = 0; k < p.n_peers; ++k) st_release_sys(&p.peer_flags[k]->arrive[p.img][p.rank], seq);
    }
  }
}
// The same exchange with the transfer taken out of the resolve kernel (the default): k_resolve writes the local image and the packed tiles, then
// a SMALL grid (64 blocks per receiver) copies the packed tiles into the receivers' gather images, four independent 16-byte loads in flight per
// thread, and publishes the arrival flags. NVLink needs a few dozen blocks, not every thread slot of the GPU, and the other frames in flight keep
// the SMs. Measured on 2 and 8 GPUs it is exactly as fast as the one-kernel form at any grid size from 16 blocks up (profiles/r2_configs.md: what
// the exchange did cost was a shared hardware work queue, not this kernel); it is kept as the default because it is the form that leaves the SMs alone.
__device__ __forceinline__ void push_pixel(const ResolveParams& p, uint32_t i, const float4 v) {
  const uint32_t tile = tile_of_rank(i >> 10, p.map.tile_rank, p.map.tile_world);
  if (tile >= p.map.n_tiles) return;
  const uint32_t lx = i & 31u, ly = (i >> 5) & 31u;
  const uint32_t x = (tile % p.map.tiles_x) * BRT_TILE + lx, y = (tile / p.map.tiles_x) * BRT_TILE + ly;
  if (x >= p.map.width || y >= p.map.height) return;
  const size_t pix = (size_t)y * p.map.width + x;
  if (p.format == BRT_FORMAT_R32G32B32A32_SFLOAT) {
    for (uint32_t k = 0; k < p.n_peers; ++k) static_cast<float4*>(p.peers[k])[pix] = v;
  } else {
    const uint32_t texel = pack_present(v, p.format);
    for (uint32_t k = 0; k < p.n_peers; ++k) static_cast<uint32_t*>(p.peers[k])[pix] = texel;
  }
}
__global__ void __launch_bounds__(256) k_push_peers(const ResolveParams p) {
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3u * stride < p.count; i += 4u * stride) {
    const float4 v0 = p.tiles[i], v1 = p.tiles[i + stride], v2 = p.tiles[i + 2u * stride], v3 = p.tiles[i + 3u * stride];
    push_pixel(p, i, v0);
    push_pixel(p, i + stride, v1);
    push_pixel(p, i + 2u * stride, v2);
    push_pixel(p, i + 3u * stride, v3);
  }
  for (; i < p.count; i += stride) push_pixel(p, i, p.tiles[i]);
  __threadfence_system();  // this thread's peer stores are ordered before the block's arrival below, at system scope
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(p.done, 1u);
    if (prev == gridDim.x - 1u) {
      __threadfence_system();
      *p.done = 0u;  // for the next frame of this slot (replayed graph)
      const uint32_t seq = p.fc->gather_seq;
      for (uint32_t k = 0; k < p.n_peers; ++k) st_release_sys(&p.peer_flags[k]->arrive[p.img][p.rank], seq);
    }
  }
}
#endif
BRT_KERNEL_1D(k_present, PresentParams, present_body)
BRT_KERNEL_1D(k_dn_temporal, DnTemporalParams, dn_temporal_body)
BRT_KERNEL_1D(k_dn_atrous, DnAtrousParams, dn_atrous_body)
BRT_KERNEL_1D(k_cull, CullParams, cull_body)

#ifdef BRT_EMU
template <bool ANY, bool COUNT>
static void k_trace(const TraceParams p) {
  const uint32_t n = trace_total(p);
  TraceCounters c{0, 0, 0};
  unsigned long long rays = 0;
  Traversal<ANY, COUNT> t;
  uint2 stack[BRT_STACK_ALLOC];
  for (uint32_t w = 0; w < n; ++w) {
    const uint32_t i = trace_slot(p, w);
    if (!trace_load(p, i, t, stack)) continue;
    rays++;
    while (!t.step(stack, c)) {}
    trace_store(p, i, t);
  }
  if (ANY) { p.stats->rays_occlusion += rays; p.stats->nodes_o += c.nodes; p.stats->prims_o += c.prims; p.stats->spheres_o += c.spheres; }
  else { p.stats->rays_closest += rays; p.stats->nodes_c += c.nodes; p.stats->prims_c += c.prims; p.stats->spheres_c += c.spheres; }
}
#define BRT_LAUNCH_TRACE(ANY, COUNT, params, grid, stream) k_trace<ANY, COUNT>(params)
#else
#ifndef BRT_TRACE_MIN_BLOCKS
#define BRT_TRACE_MIN_BLOCKS 7
#endif
// Persistent warps with per-lane refill: every lane runs one ray's Traversal; when fewer than
// p.refill_lanes lanes of the warp are still busy (0 = only when all are done), the idle lanes claim the next rays of the queue from
// a global cursor (one warp-aggregated atomic) so that the SIMD lanes stay occupied however uneven the
// per-ray cost is. The first fetch of a warp is 32 consecutive rays = one 8x4 pixel block (coherent).
template <bool ANY, bool COUNT, bool DEFER>
__global__ void __launch_bounds__(128, BRT_TRACE_MIN_BLOCKS) k_trace(const TraceParams p) {
  const uint32_t n = trace_total(p);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt = (1u << lane) - 1u;
  TraceCounters c{0, 0, 0};
  uint32_t rays = 0;
  Traversal<ANY, COUNT> t;
  uint2 stack[BRT_STACK_ALLOC];
#ifdef BRT_SMEM_STACK
  __shared__ uint2 s_stack[BRT_SMEM_STACK * 128];
  t.sst = s_stack + threadIdx.x;
#endif
#ifdef BRT_NODE_PREFETCH
  __shared__ uint4 s_node[5 * 128];
  t.snode = s_node + threadIdx.x;
#endif
  bool active = false;
  bool exhausted = false;  // warp-uniform
  uint32_t ray = 0;
  for (;;) {
    if (!exhausted) {
      const unsigned idle = __ballot_sync(0xffffffffu, !active);
      const int n_idle = __popc(idle);
      const int leader = __ffs(idle) - 1;
      uint32_t base = 0;
      if ((int)lane == leader) base = atomicAdd(p.work, (uint32_t)n_idle);
      base = __shfl_sync(0xffffffffu, base, leader);
      if (!active) {
        const uint32_t w = base + (uint32_t)__popc(idle & lt);
        if (w < n) {
          const uint32_t i = trace_slot(p, w);
          if (trace_load(p, i, t, stack)) {
            active = true;
            ray = i;
            rays++;
          }
        }
      }
      exhausted = base + (uint32_t)n_idle >= n;
    }
    if (!__any_sync(0xffffffffu, active)) {
      if (exhausted) break;
      continue;  // only padding slots were fetched: fetch again
    }
    if (DEFER) {
      // Deferred-leaf mode (traverse.cuh): all 32 lanes stay in this loop; a lane visits nodes and parks its leaf hits, the warp votes,
      // and the parked primitive tests run together once p.defer_lanes lanes have one (or nobody has node work left).
      for (;;) {
        if (active && t.can_node()) t.node_visit(stack, c);
        const unsigned pend = __ballot_sync(0xffffffffu, active && t.leaf_pending());
        const unsigned cann = __ballot_sync(0xffffffffu, active && t.can_node());
        bool finished = false;
        if (__popc(pend) >= (int)p.defer_lanes || cann == 0u) {
          if (active && t.leaf_pending()) finished = t.leaf_visit(stack, c);
        }
        if (active && (finished || t.advance(stack))) {
          trace_store(p, ray, t);
          active = false;
        }
        const int n_active = __popc(__ballot_sync(0xffffffffu, active));
        if (n_active == 0 || (!exhausted && n_active < (int)p.refill_lanes)) break;
      }
    } else {
      if (active) {
        for (;;) {
          if (t.step(stack, c)) {
            trace_store(p, ray, t);
            active = false;
            break;
          }
          if (!exhausted && __popc(__activemask()) < (int)p.refill_lanes) break;
        }
      }
      __syncwarp();
    }
  }
  rays = __reduce_add_sync(0xffffffffu, rays);
  if (COUNT) {
    c.nodes = __reduce_add_sync(0xffffffffu, c.nodes);
    c.prims = __reduce_add_sync(0xffffffffu, c.prims);
    c.spheres = __reduce_add_sync(0xffffffffu, c.spheres);
  }
  if (lane == 0) {
    if (rays) atomicAdd(ANY ? &p.stats->rays_occlusion : &p.stats->rays_closest, (unsigned long long)rays);
    if (COUNT) {
      atomicAdd(ANY ? &p.stats->nodes_o : &p.stats->nodes_c, (unsigned long long)c.nodes);
      atomicAdd(ANY ? &p.stats->prims_o : &p.stats->prims_c, (unsigned long long)c.prims);
      atomicAdd(ANY ? &p.stats->spheres_o : &p.stats->spheres_c, (unsigned long long)c.spheres);
    }
  }
}
#define BRT_LAUNCH_TRACE(ANY, COUNT, params, grid, stream)                                   \
  do {                                                                                        \
    if ((params).defer_lanes) k_trace<ANY, COUNT, true><<<(grid), 128, 0, (stream)>>>(params); \
    else k_trace<ANY, COUNT, false><<<(grid), 128, 0, (stream)>>>(params);                     \
  } while (0)
#endif

// ---- host-side scene ---------------------------------------------------------------------------------
struct MeshData {
  bool sphere = false;
  float center[3] = {0, 0, 0};
  float radius = 0.0f;
  uint32_t n_vertices = 0, n_indices = 0;
  DevBuf vertices, indices, nodes, tris;
  bool dirty = true;
  bool dynamic = false;  // vertices were replaced after the first build: rebuilds favour build speed
  BuildResult blas;
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
};
struct InstanceData {
  uint32_t mesh, material;
  float o2w[12], w2o[12];
};

// row-major 3x4 affine inverse, computed in double and rounded once (the reference gets W2O from the driver)
static void invert3x4(const float m[12], float out[12]) {
  const double a = m[0], b = m[1], c = m[2], tx = m[3];
  const double d = m[4], e = m[5], f = m[6], ty = m[7];
  const double g = m[8], h = m[9], i = m[10], tz = m[11];
  const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  const double det = a * A + b * B + c * C;
  const double id = 1.0 / det;
  const double r00 = A * id, r01 = -(b * i - c * h) * id, r02 = (b * f - c * e) * id;
  const double r10 = B * id, r11 = (a * i - c * g) * id, r12 = -(a * f - c * d) * id;
  const double r20 = C * id, r21 = -(a * h - b * g) * id, r22 = (a * e - b * d) * id;
  out[0] = (float)r00; out[1] = (float)r01; out[2] = (float)r02; out[3] = (float)(-(r00 * tx + r01 * ty + r02 * tz));
  out[4] = (float)r10; out[5] = (float)r11; out[6] = (float)r12; out[7] = (float)(-(r10 * tx + r11 * ty + r12 * tz));
  out[8] = (float)r20; out[9] = (float)r21; out[10] = (float)r22; out[11] = (float)(-(r20 * tx + r21 * ty + r22 * tz));
}

struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  int cls = 0;
};

#ifndef BRT_REFILL_LANES_INCOHERENT
#define BRT_REFILL_LANES_INCOHERENT 16u
#endif
enum { CLS_RAYGEN = 0, CLS_CLOSEST, CLS_SHADE, CLS_OCCL, CLS_ACCUM, CLS_RESOLVE, CLS_COUNT };

}  // namespace brt

using namespace brt;

// host copy of what a device table holds (upload_if_changed)
struct CachedTable {
  std::vector<unsigned char> host;
  const void* dev = nullptr;
};

// Everything one frame in flight owns: its streams, wavefront buffers, output images and timing events. The reference keeps
// MAX_FRAMES_IN_FLIGHT = 2 frames in flight with a fence per frame (VK/SwapChain.h:8, VK/SwapChain.cpp:45-60,92-131) so that
// the tail of one frame overlaps the head of the next; brt_render_frame_async / brt_frame_wait expose the same here.
struct FrameSlot {
  cudaStream_t stream = nullptr;   // closest-hit chain
  cudaStream_t stream2 = nullptr;  // shadow chain (occlusion + accumulate), overlapped with the next round's traversal
  bool own_stream = false;
  cudaEvent_t ev_shade = nullptr, ev_acc[2] = {nullptr, nullptr};
  cudaEvent_t ev_head = nullptr;  // recorded when round 0 of the frame's first sample batch has been traced (brt_context::stagger)
  uint32_t cap = 0;    // path slots per sample (owned tiles * 1024)
  uint32_t batch = 1;  // samples traced together in one wavefront
  size_t frame_bytes = 0;
  uint64_t frame_key[4] = {0, 0, 0, 0};
  uint32_t frame_w = 0, frame_h = 0;
  DevBuf q_o[2], q_d[2], q_w[2], q_px[2], q_seed[2], d_hit, d_hit_inst, d_contrib[2], d_aux[2], s_o[2], s_d[2], s_target[2];
  DevBuf d_image8;  // the frame in an 8-bit present format (BRT_RENDER_FORMAT)
  DevBuf d_aov_pos, d_aov_nrm;  // BRT_RENDER_GBUFFER
  DevBuf d_denoised;            // output of the denoiser stages for this slot's frame (brt_denoise / BRT_RENDER_DENOISE)
  cudaEvent_t ev_dn[2] = {nullptr, nullptr};
  bool has_denoised = false;
  uint32_t dn_launches = 0;
  uint32_t gather_img = 0;  // fused multi-GPU exchange: the gather image this slot's frame stored into (= the slot's index)
  uint32_t gather_seq = 0;  // ... and that frame's sequence number on the image
  uint32_t gather_format = 0;
  DevBuf d_resolve_done;    // block arrival counter of k_resolve_peers
  bool has_gbuffer = false;
  uint32_t render_flags = 0;  // of the frame rendered last on this slot
  DevBuf d_sort_key, d_sort_rank, d_sort_order, d_sort_bins;  // hit sort of the bounce rounds
  DevBuf d_alive;  // rounds that contributed per (sample in batch, slot)
  DevBuf d_rad, d_accum, d_image, d_tiles, d_aov_prim, d_aov_inst, d_aov_t, d_counters, d_fstats;
  std::deque<EventPair> events;  // deque: references stay valid while the pool grows
  size_t events_used = 0;
  uint32_t launches = 0, l_closest = 0, l_occl = 0;
  FrameStats* fs_host = nullptr;  // pinned: device-side statistics of the frame, copied back on the frame's stream
  bool in_flight = false;
  bool ready = false;
  FrameConsts* h_consts = nullptr;  // pinned staging of the per-frame constants
  DevBuf d_consts;
  uint32_t generation = 0;          // bumped whenever the wavefront buffers are reallocated (invalidates the graph)
  uint64_t cleared_key[3] = {~0ull, 0, 0};  // (generation, traced window, size) the full-frame clears were last done for
#ifndef BRT_EMU
  // the whole frame, captured once per frame shape and replayed ([1]: the second gather image of the fused multi-GPU exchange)
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  uint64_t graph_key[2][18] = {{0}, {0}};
#endif  // streams / events / pinned block created
};

struct brt_context {
  int device = 0;
  int sm_count = 148;
  uint32_t tile_rank = 0, tile_world = 1, flags = 0;
  cudaStream_t stream = nullptr;  // scene work (uploads, builds, culling) and frame slot 0
  cudaEvent_t ev_t[3] = {nullptr, nullptr, nullptr};  // build / cull timing
  bool own_stream = false;
  std::string err;
  std::unique_ptr<Builder> builder;

  // scene (host mirrors)
  std::vector<std::unique_ptr<MeshData>> meshes;
  std::vector<InstanceData> instances;
  std::vector<uint8_t> visible;  // per instance, Smart Culling state
  std::vector<brt_material> materials;
  std::vector<float> mat_ext;    // transmission, ior per material
  std::vector<brt_light> lights;
  brt_sky sky{};
  bool built = false;
  bool tables_dirty = true, tlas_dirty = true;

  // scene (device)
  std::vector<brt_light_bvh_node> light_bvh;  // RT/Scene.h:123-130, built on the host by build_tables
  DevBuf d_light_bvh;
  DevBuf d_materials, d_mat_ext, d_lights, d_inst_shade, d_inst_src, d_inst_ids, d_mesh_bounds, d_visible;
  CachedTable t_light_bvh, t_materials, t_mat_ext, t_lights, t_mesh_bounds, t_inst_shade, t_instance_info, t_sky, t_scene_info, t_inst_src, t_inst_ids;
  DevBuf d_instance_info, d_sky, d_scene_info;  // the reference's InstanceInfo[] / SkyInfo / SceneBufferInfo (RT/Scene.h:84-121) for callers that bind them
  brt_scene_buffer_info scene_info{};
  DevBuf d_tlas_nodes, d_tlas_inst;
  float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};  // world box of all instances (grid of the hit sort)
  uint32_t tlas_count = 0;  // visible, non-empty instances in the TLAS
  BuildResult tlas{};

  // frames
  FrameSlot slots[BRT_FRAMES_IN_FLIGHT];
  cudaEvent_t prev_head = nullptr;        // ev_head of the frame submitted last: the next frame starts behind it (staggered frames)
  uint32_t last_slot = 0;                 // slot of the frame most recently waited for (brt_get_aov / brt_device_image read it)
  uint32_t target_wavefront = 16u << 20;  // paths per wavefront aimed for (BRT_WAVEFRONT_PATHS overrides, for tuning)
  bool stagger_first = true; // ... of the frame's first sample batch (false: of its last one); BRT_STAGGER_FIRST
  uint32_t stagger = 4;      // where in a frame the next frame in flight may start: 0 at once, 1 / 2 / 3 after round 0's closest-hit trace / shade / occlusion trace, 4 = by the frame's depth (BRT_STAGGER)
  uint32_t peer_grid = 0;    // blocks of the kernel that stores into the receivers' gather images (0 = default: 64, or a full grid for the fused kernel); BRT_PEER_GRID
  bool peer_push = true;     // resolve locally, then push the tiles with a small grid (false: one fused full-grid kernel); BRT_PEER_PUSH
  uint32_t shade_ahead = 0;  // tuning aid: slots ahead of which the shade kernels request path heads into the L2 (BRT_SHADE_AHEAD)
  uint32_t rays_per_warp = 0;  // 0 = full grid always; else bounce-round launches are sized for this many rays per warp (BRT_RAYS_PER_WARP)
  uint32_t refill_primary = 0, refill_bounce = BRT_REFILL_LANES_INCOHERENT;  // lanes still busy below which a warp refills its idle lanes
  uint32_t defer_primary = 0, defer_bounce = 0;  // deferred-leaf traversal: lanes with a parked primitive test that trigger a pass (0 = off)
  DevBuf d_rays, d_ray_out;  // brt_trace_rays staging
  // fused resolve + exchange: own gather image and the peers' (opened through cudaIpc)
  DevBuf d_gather;  // BRT_GATHER_IMAGES full frames back to back + one GatherFlags block (one allocation, one IPC handle): slot k uses image k
  uint32_t gather_w = 0, gather_h = 0, n_peers = 0;
  uint32_t gather_next = 0;                       // slot the next synchronous brt_render_frame_peers uses
  uint32_t gather_seq[BRT_GATHER_IMAGES] = {0};   // frames submitted so far per gather image (every rank counts the same)
  uint32_t gather_root_only = 0, gather_root = 0; // brt_gather_configure: who receives the pixels
  void* peer_images[BRT_MAX_PEERS] = {nullptr};
  bool peer_opened[BRT_MAX_PEERS] = {false};
  cudaStream_t gather_stream = nullptr;           // default stream of brt_gather_wait / release / copy_to_host
  // denoiser history (Graphics/Denoiser/Denoiser.h): accumulated colour + history length, luminance moments, last G-buffer, camera
  DevBuf dn_work[2], dn_hist_color[2], dn_hist_mom[2], dn_h_nrm, dn_h_inst, dn_h_t;
  uint32_t dn_w = 0, dn_h = 0, dn_parity = 0;
  brt_denoise_opts dn_opts{sizeof(brt_denoise_opts), BRT_DENOISE_BILATERAL, 4u, 5u, 0.05f, 4.0f, 0.0f, 32.0f};  // brt_denoise_configure
  cudaEvent_t dn_event = nullptr;  // recorded after the denoiser stages of a frame: the next frame's stages read its history
  bool dn_event_recorded = false;
  bool dn_have_history = false;
  float dn_prev_vp[16] = {0}, dn_prev_eye[3] = {0, 0, 0};
  brt_stats stats{};
};

// general 4x4 inverse in double (camera matrices)
static void invert4x4(const double m[16], double inv[16]) {
  const double s0 = m[0] * m[5] - m[4] * m[1], s1 = m[0] * m[6] - m[4] * m[2], s2 = m[0] * m[7] - m[4] * m[3];
  const double s3 = m[1] * m[6] - m[5] * m[2], s4 = m[1] * m[7] - m[5] * m[3], s5 = m[2] * m[7] - m[6] * m[3];
  const double c5 = m[10] * m[15] - m[14] * m[11], c4 = m[9] * m[15] - m[13] * m[11], c3 = m[9] * m[14] - m[13] * m[10];
  const double c2 = m[8] * m[15] - m[12] * m[11], c1 = m[8] * m[14] - m[12] * m[10], c0 = m[8] * m[13] - m[12] * m[9];
  const double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
  const double id = 1.0 / det;
  inv[0] = (m[5] * c5 - m[6] * c4 + m[7] * c3) * id;
  inv[1] = (-m[1] * c5 + m[2] * c4 - m[3] * c3) * id;
  inv[2] = (m[13] * s5 - m[14] * s4 + m[15] * s3) * id;
  inv[3] = (-m[9] * s5 + m[10] * s4 - m[11] * s3) * id;
  inv[4] = (-m[4] * c5 + m[6] * c2 - m[7] * c1) * id;
  inv[5] = (m[0] * c5 - m[2] * c2 + m[3] * c1) * id;
  inv[6] = (-m[12] * s5 + m[14] * s2 - m[15] * s1) * id;
  inv[7] = (m[8] * s5 - m[10] * s2 + m[11] * s1) * id;
  inv[8] = (m[4] * c4 - m[5] * c2 + m[7] * c0) * id;
  inv[9] = (-m[0] * c4 + m[1] * c2 - m[3] * c0) * id;
  inv[10] = (m[12] * s4 - m[13] * s2 + m[15] * s0) * id;
  inv[11] = (-m[8] * s4 + m[9] * s2 - m[11] * s0) * id;
  inv[12] = (-m[4] * c3 + m[5] * c1 - m[6] * c0) * id;
  inv[13] = (m[0] * c3 - m[1] * c1 + m[2] * c0) * id;
  inv[14] = (-m[12] * s3 + m[13] * s1 - m[14] * s0) * id;
  inv[15] = (m[8] * s3 - m[9] * s1 + m[10] * s0) * id;
}

namespace {

thread_local std::string g_create_error;

template <class F>
int guarded(brt_context* ctx, F&& f) {
  try {
    f();
    return BRT_OK;
  } catch (const CudaError& e) {
    if (ctx) ctx->err = e.what();
    return BRT_ERR_CUDA;
  } catch (const LimitError& e) {
    if (ctx) ctx->err = e.what();
    return BRT_ERR_LIMIT;
  } catch (const std::invalid_argument& e) {
    if (ctx) ctx->err = e.what();
    return BRT_ERR_INVALID;
  } catch (const std::logic_error& e) {
    if (ctx) ctx->err = e.what();
    return BRT_ERR_STATE;
  } catch (const std::exception& e) {
    if (ctx) ctx->err = e.what();
    return BRT_ERR_CUDA;
  }
}
[[noreturn]] void invalid(const char* m) { throw std::invalid_argument(m); }
[[noreturn]] void bad_state(const char* m) { throw std::logic_error(m); }

void upload(cudaStream_t s, DevBuf& buf, const void* src, size_t bytes) {
  buf.ensure(std::max<size_t>(bytes, 16));
  if (bytes) BRT_CUDA(cudaMemcpyAsync(buf.ptr(), src, bytes, cudaMemcpyHostToDevice, s));
}

// Scene tables are re-derived by every build; most of them do not change from one per-frame re-build to the next (C4: only the bounds of
// the animated mesh do). Each table keeps a host copy of what is on the device and is uploaded only when it differs.
void upload_if_changed(cudaStream_t s, DevBuf& buf, CachedTable& cache, const void* src, size_t bytes) {
  buf.ensure(std::max<size_t>(bytes, 16));
  if (cache.dev == buf.ptr() && cache.host.size() == bytes && (bytes == 0 || std::memcmp(cache.host.data(), src, bytes) == 0)) return;
  cache.host.assign(static_cast<const unsigned char*>(src), static_cast<const unsigned char*>(src) + bytes);
  cache.dev = buf.ptr();
  if (bytes) BRT_CUDA(cudaMemcpyAsync(buf.ptr(), cache.host.data(), bytes, cudaMemcpyHostToDevice, s));
}

uint32_t grid_for(const brt_context* c, uint32_t n, uint32_t block, uint32_t blocks_per_sm) {
  return std::max(1u, std::min(div_up(n, block), (uint32_t)c->sm_count * blocks_per_sm));
}

// ---- timing ----------------------------------------------------------------------------------------
EventPair& next_events(FrameSlot* c, int cls) {
  if (c->events_used == c->events.size()) {
    EventPair e;
    BRT_CUDA(cudaEventCreate(&e.a));
    BRT_CUDA(cudaEventCreate(&e.b));
    c->events.push_back(e);
  }
  EventPair& e = c->events[c->events_used++];
  e.cls = cls;
  return e;
}
// Events that must really be recorded when a captured frame graph is replayed (timing, the head event other slots wait for) are
// "external" event-record nodes; outside a capture the flag must not be used.
thread_local bool g_capturing = false;
void record_event_nothrow(cudaEvent_t e, cudaStream_t s) {
#ifdef BRT_EMU
  cudaEventRecord(e, s);
#else
  cudaEventRecordWithFlags(e, s, g_capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
#endif
}
void record_event(cudaEvent_t e, cudaStream_t s) {
#ifdef BRT_EMU
  BRT_CUDA(cudaEventRecord(e, s));
#else
  BRT_CUDA(cudaEventRecordWithFlags(e, s, g_capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
#endif
}
struct Timed {  // brackets one launch with events of class `cls` on the stream it is launched on
  cudaStream_t s;
  EventPair* e;
  // (not inside a captured frame graph: ~60 event-record nodes would cost more than the launch gaps the graph saves; a graph
  // frame reports its total time only, per-class times need BRT_CFG_NO_GRAPH / BRT_CFG_NO_OVERLAP)
  Timed(FrameSlot* ctx, int cls, cudaStream_t stream) : s(stream), e(g_capturing ? nullptr : &next_events(ctx, cls)) {
    if (e) record_event(e->a, s);
  }
  ~Timed() {
    if (e) record_event_nothrow(e->b, s);
  }
};

// ---- light BVH (RT/Scene.h:123-130) ---------------------------------------------------------
// Median split along the widest axis of the light positions, ties broken by light index; children are allocated in pairs
// (left, right = left + 1) and built depth-first, so the node order is fully determined by the light list.
// Point lights and the lights of the shader's constant-direction branch (SPOT / DIRECTIONAL, SH/light.slang:33-36: direction (0.9, -0.1, 0),
// no falloff) never share a subtree: when both kinds exist the root's left child holds the point lights, the right child the others.
// The cone fields tell the kinds apart: point lights radiate everywhere (axis (0,0,1), angle pi), a constant-direction subtree has
// angle 0 and the emission axis -normalize(0.9, -0.1, 0); its importance does not depend on the shading point.
void build_light_subtree(const std::vector<brt_light>& lights, std::vector<uint32_t>& order, std::vector<brt_light_bvh_node>& nodes, uint32_t root,
                         uint32_t first0, uint32_t count0, bool directional) {
  struct Task { uint32_t node, first, count; };
  std::vector<Task> stack;
  stack.push_back({root, first0, count0});
  while (!stack.empty()) {
    const Task t = stack.back();
    stack.pop_back();
    brt_light_bvh_node nd{};
    float flux = 0.0f;
    for (int k = 0; k < 3; ++k) { nd.bBoxMin[k] = INFINITY; nd.bBoxMax[k] = -INFINITY; }
    for (uint32_t i = t.first; i < t.first + t.count; ++i) {
      const brt_light& l = lights[order[i]];
      for (int k = 0; k < 3; ++k) {
        nd.bBoxMin[k] = std::fmin(nd.bBoxMin[k], l.pos[k]);
        nd.bBoxMax[k] = std::fmax(nd.bBoxMax[k], l.pos[k]);
      }
      flux = flux + std::fabs(l.intensity * ((0.2126f * l.color[0] + 0.7152f * l.color[1]) + 0.0722f * l.color[2]));
    }
    nd.totalFlux = flux;
    if (directional) {
      nd.coneAxis[0] = -0.993883729f; nd.coneAxis[1] = 0.110431522f; nd.coneAxis[2] = 0.0f;  // -normalize(0.9, -0.1, 0)
      nd.coneAngle = 0.0f;
    } else {
      nd.coneAxis[0] = 0.0f; nd.coneAxis[1] = 0.0f; nd.coneAxis[2] = 1.0f;
      nd.coneAngle = 3.14159274f;
    }
    if (t.count == 1) {
      nd.childIndex = -1 - (int32_t)order[t.first];
    } else {
      int axis = 0;
      float best = nd.bBoxMax[0] - nd.bBoxMin[0];
      for (int k = 1; k < 3; ++k)
        if (nd.bBoxMax[k] - nd.bBoxMin[k] > best) { best = nd.bBoxMax[k] - nd.bBoxMin[k]; axis = k; }
      std::sort(order.begin() + t.first, order.begin() + t.first + t.count, [&](uint32_t a, uint32_t b) {
        const float pa = lights[a].pos[axis], pb = lights[b].pos[axis];
        return pa < pb || (pa == pb && a < b);
      });
      const uint32_t mid = t.count / 2;
      const uint32_t left = (uint32_t)nodes.size();
      nd.childIndex = (int32_t)left;
      nodes.push_back(brt_light_bvh_node{});
      nodes.push_back(brt_light_bvh_node{});
      stack.push_back({left + 1, t.first + mid, t.count - mid});  // popped second: the left subtree is built first
      stack.push_back({left, t.first, mid});
    }
    nodes[t.node] = nd;
  }
}
void build_light_bvh(const std::vector<brt_light>& lights, std::vector<brt_light_bvh_node>& nodes) {
  nodes.clear();
  const uint32_t n = (uint32_t)lights.size();
  if (!n) return;
  nodes.reserve(2 * (size_t)n + 2);
  std::vector<uint32_t> order;
  for (uint32_t i = 0; i < n; ++i)
    if (lights[i].type == BRT_LIGHT_POINT) order.push_back(i);
  const uint32_t n_point = (uint32_t)order.size();
  for (uint32_t i = 0; i < n; ++i)
    if (lights[i].type != BRT_LIGHT_POINT) order.push_back(i);
  nodes.push_back(brt_light_bvh_node{});
  if (n_point == 0 || n_point == n) {
    build_light_subtree(lights, order, nodes, 0, 0, n, n_point == 0);
    return;
  }
  // both kinds: the root only separates them (left = point lights, right = constant-direction lights)
  nodes.push_back(brt_light_bvh_node{});
  nodes.push_back(brt_light_bvh_node{});
  build_light_subtree(lights, order, nodes, 1, 0, n_point, false);
  build_light_subtree(lights, order, nodes, 2, n_point, n - n_point, true);
  brt_light_bvh_node root{};
  for (int k = 0; k < 3; ++k) {
    root.bBoxMin[k] = std::fmin(nodes[1].bBoxMin[k], nodes[2].bBoxMin[k]);
    root.bBoxMax[k] = std::fmax(nodes[1].bBoxMax[k], nodes[2].bBoxMax[k]);
  }
  root.totalFlux = nodes[1].totalFlux + nodes[2].totalFlux;
  root.coneAxis[2] = 1.0f;
  root.coneAngle = 3.14159274f;
  root.childIndex = 1;
  nodes[0] = root;
}

// ---- scene tables ------------------------------------------------------------------------------------
void build_tables(brt_context* c) {
  cudaStream_t s = c->stream;
  build_light_bvh(c->lights, c->light_bvh);
  upload_if_changed(s, c->d_light_bvh, c->t_light_bvh, c->light_bvh.data(), c->light_bvh.size() * sizeof(brt_light_bvh_node));
  upload_if_changed(s, c->d_materials, c->t_materials, c->materials.data(), c->materials.size() * sizeof(brt_material));
  upload_if_changed(s, c->d_mat_ext, c->t_mat_ext, c->mat_ext.data(), c->mat_ext.size() * 4);
  upload_if_changed(s, c->d_lights, c->t_lights, c->lights.data(), c->lights.size() * sizeof(brt_light));
  // exact object-space box of every mesh (read by the TLAS builder and Smart Culling)
  std::vector<float> mb(std::max<size_t>(c->meshes.size(), 1) * 8, 0.0f);
  for (size_t mi = 0; mi < c->meshes.size(); ++mi)
    for (int k = 0; k < 3; ++k) {
      mb[8 * mi + k] = c->meshes[mi]->lo[k];
      mb[8 * mi + 4 + k] = c->meshes[mi]->hi[k];
    }
  upload_if_changed(s, c->d_mesh_bounds, c->t_mesh_bounds, mb.data(), mb.size() * 4);
  std::vector<InstShade> shade(c->instances.size());
  for (size_t i = 0; i < c->instances.size(); ++i) {
    const InstanceData& in = c->instances[i];
    const MeshData& m = *c->meshes[in.mesh];
    InstShade& r = shade[i];
    std::memset(&r, 0, sizeof(r));
    std::memcpy(r.o2w, in.o2w, 48);
    std::memcpy(r.w2o, in.w2o, 48);
    r.vertices = m.vertices.as<float>();
    r.indices = m.indices.as<uint32_t>();
    r.sphere = make_float4(m.center[0], m.center[1], m.center[2], m.radius);
    r.kind = m.sphere ? 1u : 0u;
    r.material = in.material;
    r.mesh = in.mesh;
  }
  upload_if_changed(s, c->d_inst_shade, c->t_inst_shade, shade.data(), shade.size() * sizeof(InstShade));
  // the reference's own tables (RT/Scene.cpp:357-403): InstanceInfo per instance, SkyInfo, SceneBufferInfo
  {
    std::vector<brt_instance_info> info(std::max<size_t>(c->instances.size(), 1));
    for (size_t i = 0; i < c->instances.size(); ++i) {
      const MeshData& m = *c->meshes[c->instances[i].mesh];
      info[i] = brt_instance_info{(uint64_t)m.vertices.ptr(), (uint64_t)m.indices.ptr(), c->instances[i].material, 0u};
    }
    upload_if_changed(s, c->d_instance_info, c->t_instance_info, info.data(), info.size() * sizeof(brt_instance_info));
    upload_if_changed(s, c->d_sky, c->t_sky, &c->sky, sizeof(brt_sky));
    c->scene_info = brt_scene_buffer_info{(uint64_t)c->d_materials.ptr(), sizeof(brt_material), (uint64_t)c->d_lights.ptr(), sizeof(brt_light),
                                          (uint64_t)c->lights.size(), sizeof(brt_vertex), (uint64_t)c->d_instance_info.ptr(), sizeof(brt_instance_info),
                                          (uint64_t)c->d_sky.ptr(), sizeof(brt_sky)};
    upload_if_changed(s, c->d_scene_info, c->t_scene_info, &c->scene_info, sizeof(c->scene_info));
  }
  // world box of all instances (centre / extent form of |M| box): the grid of the bounce rounds' hit sort
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (const InstanceData& in : c->instances) {
    const MeshData& m = *c->meshes[in.mesh];
    for (int r = 0; r < 3; ++r) {
      float ctr = in.o2w[4 * r + 3], ext = 0.0f;
      for (int k = 0; k < 3; ++k) {
        ctr += in.o2w[4 * r + k] * 0.5f * (m.lo[k] + m.hi[k]);
        ext += std::fabs(in.o2w[4 * r + k]) * 0.5f * (m.hi[k] - m.lo[k]);
      }
      lo[r] = std::fmin(lo[r], ctr - ext);
      hi[r] = std::fmax(hi[r], ctr + ext);
    }
  }
  for (int k = 0; k < 3; ++k) {
    c->scene_lo[k] = c->instances.empty() ? 0.0f : lo[k];
    c->scene_hi[k] = c->instances.empty() ? 0.0f : hi[k];
  }
  c->tables_dirty = false;
}

void build_tlas(brt_context* c) {
  cudaStream_t s = c->stream;
  std::vector<InstRec> src;
  std::vector<uint32_t> ids;
  for (size_t i = 0; i < c->instances.size(); ++i) {
    const InstanceData& in = c->instances[i];
    const MeshData& m = *c->meshes[in.mesh];
    if (!c->visible[i]) continue;
    if (!m.sphere && m.n_indices == 0) continue;
    InstRec r;
    std::memset(&r, 0, sizeof(r));
    std::memcpy(r.w2o, in.w2o, 48);
    r.nodes = m.nodes.as<Node8>();
    r.tris = m.tris.as<TriRec>();
    r.sphere = make_float4(m.center[0], m.center[1], m.center[2], m.radius);
    r.kind = m.sphere ? 1u : 0u;
    r.inst_id = (uint32_t)i;
    src.push_back(r);
    ids.push_back((uint32_t)i);
  }
  c->tlas_count = (uint32_t)src.size();
  c->tlas = BuildResult{};
  if (c->tlas_count) {
    upload_if_changed(s, c->d_inst_src, c->t_inst_src, src.data(), src.size() * sizeof(InstRec));
    upload_if_changed(s, c->d_inst_ids, c->t_inst_ids, ids.data(), ids.size() * 4);
    c->d_tlas_nodes.ensure((size_t)Builder::node_capacity(c->tlas_count) * sizeof(Node8));
    c->d_tlas_inst.ensure((size_t)c->tlas_count * sizeof(InstRec));
    c->builder->build_instances(s, c->d_inst_shade.as<InstShade>(), c->d_inst_ids.as<uint32_t>(), c->d_inst_src.as<InstRec>(), c->tlas_count,
                                c->d_mesh_bounds.as<float4>(), c->d_tlas_nodes.as<Node8>(), c->d_tlas_inst.as<InstRec>(), &c->tlas);
  }
  c->tlas_dirty = false;
}

uint32_t max_blas_levels(const brt_context* c) {
  uint32_t l = 0;
  for (const auto& m : c->meshes)
    if (!m->sphere) l = std::max(l, m->blas.levels);
  return l;
}

void refresh_scene_stats(brt_context* c) {
  uint64_t tris = 0, nodes = c->tlas.n_nodes, bytes = (uint64_t)c->tlas.n_nodes * 80 + (uint64_t)c->tlas_count * sizeof(InstRec);
  for (const InstanceData& in : c->instances) tris += c->meshes[in.mesh]->n_indices / 3;
  float sah = 0.0f, sah0 = 0.0f;
  uint32_t biggest = 0;
  for (const auto& m : c->meshes) {
    if (m->sphere || m->n_indices == 0) continue;
    nodes += m->blas.n_nodes;
    bytes += (uint64_t)m->blas.n_nodes * 80 + (uint64_t)(m->n_indices / 3) * 48;
    if (m->n_indices / 3 >= biggest) { biggest = m->n_indices / 3; sah = m->blas.sah_final; sah0 = m->blas.sah_lbvh; }
  }
  c->stats.total_triangles = tris;
  c->stats.bvh_nodes = nodes;
  c->stats.bvh_bytes = bytes;
  c->stats.sah_cost = sah;
  c->stats.sah_cost_lbvh = sah0;
  c->stats.instances_total = (uint32_t)c->instances.size();
  c->stats.instances_visible = c->tlas_count;
}

// with_tlas = false: the caller builds the TLAS itself right afterwards (brt_smart_cull, once the visibility is known)
void scene_build(brt_context* c, bool with_tlas = true) {
  cudaStream_t s = c->stream;
  cudaEvent_t e0 = c->ev_t[0], e1 = c->ev_t[1], e2 = c->ev_t[2];
  BRT_CUDA(cudaEventRecord(e0, s));
  uint32_t rebuilt = 0;
  for (size_t mi = 0; mi < c->meshes.size(); ++mi) {
    MeshData& m = *c->meshes[mi];
    if (!m.dirty) continue;
    if (m.sphere) {
      for (int k = 0; k < 3; ++k) { m.lo[k] = m.center[k] - m.radius; m.hi[k] = m.center[k] + m.radius; }
    } else if (m.n_indices == 0) {
      for (int k = 0; k < 3; ++k) m.lo[k] = m.hi[k] = 0.0f;
    } else {
      const uint32_t nt = m.n_indices / 3;
      m.nodes.ensure((size_t)Builder::node_capacity(nt) * sizeof(Node8));
      m.tris.ensure((size_t)nt * sizeof(TriRec));
      c->builder->build_triangles(s, m.vertices.as<float>(), m.indices.as<uint32_t>(), nt, m.nodes.as<Node8>(), m.tris.as<TriRec>(), nullptr,
                                  !(c->flags & BRT_CFG_NO_TREELET) && (!m.dynamic || (c->flags & BRT_CFG_TREELET_ON_REBUILD)), &m.blas);
      for (int k = 0; k < 3; ++k) { m.lo[k] = m.blas.lo[k]; m.hi[k] = m.blas.hi[k]; }
    }
    m.dirty = false;
    rebuilt++;
  }
  BRT_CUDA(cudaEventRecord(e1, s));
  build_tables(c);
  if (with_tlas) build_tlas(c);
  BRT_CUDA(cudaEventRecord(e2, s));
  BRT_CUDA(cudaStreamSynchronize(s));
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, e0, e1);
  c->stats.ms_blas_build = ms;
  cudaEventElapsedTime(&ms, e1, e2);
  c->stats.ms_tlas_build = ms;
  c->stats.blas_built = rebuilt;
  if (with_tlas && c->tlas.levels + max_blas_levels(c) > BRT_MAX_TREE_LEVELS) throw LimitError("scene_build: BVH deeper than the traversal stack allows");
  c->built = true;
  refresh_scene_stats(c);
}

// ---- frame ---------------------------------------------------------------------------------------------
TileMap make_tile_map(const brt_context* c, const brt_render_opts& o) {
  TileMap m;
  m.width = o.width;
  m.height = o.height;
  m.tiles_x = div_up(o.width, BRT_TILE);
  m.n_tiles = m.tiles_x * div_up(o.height, BRT_TILE);
  m.tile_rank = c->tile_rank;
  m.tile_world = c->tile_world;
  m.crop_x0 = 0; m.crop_y0 = 0; m.crop_x1 = o.width; m.crop_y1 = o.height;
  if (o.crop_w) {
    m.crop_x0 = o.crop_x0;
    m.crop_y0 = o.crop_y0;
    m.crop_x1 = std::min<uint64_t>(o.width, (uint64_t)o.crop_x0 + o.crop_w);
    m.crop_y1 = std::min<uint64_t>(o.height, (uint64_t)o.crop_y0 + o.crop_h);
  }
  return m;
}
uint32_t tiles_per_rank(uint32_t width, uint32_t height, uint32_t world) {
  const uint32_t n = div_up(width, BRT_TILE) * div_up(height, BRT_TILE);
  return div_up(n, world ? world : 1);
}

// Sizes the per-frame buffers. A wavefront holds `batch` samples of every owned pixel at once (cap slots per sample): the
// more samples share a launch, the less the latency-bound tail of every launch costs — this is what keeps the tile-parallel
// multi-GPU frames efficient, where each rank only has 1/N of the pixels.
void ensure_frame_buffers(brt_context* c, FrameSlot* f, const brt_render_opts& o, uint32_t rounds) {
  const bool lbvh = (o.flags & BRT_RENDER_LIGHT_BVH) != 0u;
  const uint32_t cap = tiles_per_rank(o.width, o.height, c->tile_world) * 1024u;
  const size_t npx = (size_t)o.width * o.height;
  const uint32_t L = lbvh ? 1u : std::max<uint32_t>(1, (uint32_t)c->lights.size());  // segments per path
  const uint32_t R = std::max(1u, rounds);
  const uint64_t key[4] = {((uint64_t)o.width << 32) | o.height, ((uint64_t)o.spp << 32) | R, L, c->target_wavefront};
  if (f->frame_bytes && std::memcmp(key, f->frame_key, sizeof(key)) == 0) return;  // same frame shape as last time
  std::memcpy(f->frame_key, key, sizeof(key));
  f->generation++;  // buffers below may move: a captured frame graph of this slot is stale
  // bytes per sample of the batch: two path queues, hits, and per round parity: contributions, weights, shadow queue; + radiance terms
  const size_t per_sample = (size_t)cap * (2 * 56 + 20 + 2 * (16 * L + 16 + 36 * L) + 16 * R);
  size_t free_b = 0, total_b = 0;
  BRT_CUDA(cudaMemGetInfo(&free_b, &total_b));
  size_t budget = std::min<size_t>((size_t)32 << 30, (free_b + f->frame_bytes) * 2 / 5);
  uint32_t batch = (uint32_t)std::max<size_t>(1, std::min<size_t>(budget / std::max<size_t>(per_sample, 1), BRT_MAX_SAMPLE_BATCH));
  // enough samples per wavefront to amortise the per-launch tails (about 16 M paths), no more: larger wavefronts only cost
  // memory traffic
  batch = std::min(batch, std::max(1u, div_up(c->target_wavefront, std::max(cap, 1u))));
  batch = std::min(batch, std::max(1u, o.spp));
  if ((uint64_t)cap * batch > 0xfffffff0ull / L) batch = std::max<uint64_t>(1, 0xfffffff0ull / L / cap);  // 32-bit shadow-queue targets
  const size_t capw = (size_t)cap * batch;
  for (int k = 0; k < 2; ++k) {
    f->q_o[k].ensure(capw * 16);
    f->q_d[k].ensure(capw * 16);
    f->q_w[k].ensure(capw * 16);
    f->q_px[k].ensure(capw * 4);
    f->q_seed[k].ensure(capw * 4);
    // per round parity: what the occlusion / accumulate chain of a round owns
    f->d_contrib[k].ensure(capw * 16 * L);
    f->d_aux[k].ensure(capw * 16);
    f->s_o[k].ensure(capw * 16 * L);
    f->s_d[k].ensure(capw * 16 * L);
    f->s_target[k].ensure(capw * 4 * L);
  }
  f->d_hit.ensure(capw * 16);
  f->d_hit_inst.ensure(capw * 4);
  f->d_rad.ensure(capw * 16 * R);
  f->d_alive.ensure(capw * 4);
  if (R > 1 && (c->flags & BRT_CFG_HIT_SORT)) {
    f->d_sort_key.ensure(capw * 4);
    f->d_sort_rank.ensure(capw * 4);
    f->d_sort_order.ensure(capw * 4);
    f->d_sort_bins.ensure((size_t)BRT_HITSORT_BINS * 4);
  }
  f->d_accum.ensure((size_t)cap * 16);
  f->d_image.ensure(npx * 16);
  f->d_tiles.ensure((size_t)cap * 16);
  f->d_aov_prim.ensure(npx * 4);
  f->d_aov_inst.ensure(npx * 4);
  f->d_aov_t.ensure(npx * 4);
  f->d_counters.ensure(sizeof(FrameCounters) + 2 * sizeof(ShadowCounters));
  f->d_fstats.ensure(sizeof(FrameStats));
  f->frame_bytes = per_sample * batch;
  f->cap = cap;
  f->batch = batch;
  f->frame_w = o.width;
  f->frame_h = o.height;
}

PathQueue queue_of(FrameSlot* f, int k) {
  return PathQueue{f->q_o[k].as<float4>(), f->q_d[k].as<float4>(), f->q_w[k].as<float4>(), f->q_px[k].as<uint32_t>(), f->q_seed[k].as<uint32_t>()};
}

// max_rays: upper bound of the queue this launch walks (the live count is on the device). A persistent warp claims 32 rays at a time and
// refills idle lanes from the cursor; when the queue holds less than a few claims per warp there is nothing to refill from and the
// lanes of a warp idle until its slowest ray is done — so small queues get a smaller grid (c->rays_per_warp claims' worth per warp).
template <bool ANY>
void launch_trace(brt_context* c, const TraceParams& p, cudaStream_t stream, uint64_t max_rays = ~0ull) {
  uint32_t grid = (uint32_t)c->sm_count * 8u;
  if (c->rays_per_warp && max_rays != ~0ull) {
    const uint64_t warps = std::max<uint64_t>(1, max_rays / c->rays_per_warp);
    grid = (uint32_t)std::max<uint64_t>((uint64_t)c->sm_count, std::min<uint64_t>(grid, (warps + 3) / 4));
  }
  if (c->flags & BRT_CFG_COUNTERS) BRT_LAUNCH_TRACE(ANY, true, p, grid, stream);
  else BRT_LAUNCH_TRACE(ANY, false, p, grid, stream);
  BRT_CHECK_LAUNCH();
}

// ---- denoiser stages (Graphics/Denoiser/Denoiser.h:5-20) ------------------------------------------------------
void check_denoise_opts(const brt_denoise_opts& d) {
  if (d.struct_size != sizeof(brt_denoise_opts)) invalid("denoise: struct_size mismatch");
  if (d.iterations > 6 || d.sigma_n_log2 > 8) invalid("denoise: iterations <= 6, sigma_n_log2 <= 8");
}

// Enqueues temporal accumulation with reprojection + history clamping + variance estimation (k_dn_temporal), the a-trous wavelet
// iterations and the bilateral pass (k_dn_atrous) for the frame of slot `f` on stream `s`; the result lands in f->d_denoised. The
// history (accumulated colour, moments, last G-buffer, camera) lives in the context: frames must be enqueued in display order, and
// the stages of a frame start behind those of the previous one (dn_event) whichever slot / stream that frame used.
void enqueue_denoise(brt_context* c, FrameSlot* f, const brt_uniform& u, const brt_denoise_opts& d, cudaStream_t s) {
  const uint32_t W = f->frame_w, H = f->frame_h;
  const size_t npx = (size_t)W * H;
  if (c->dn_w != W || c->dn_h != H || (d.flags & BRT_DENOISE_RESET)) c->dn_have_history = false;
  c->dn_w = W;
  c->dn_h = H;
  for (int k = 0; k < 2; ++k) {
    c->dn_work[k].ensure(npx * 16);
    c->dn_hist_color[k].ensure(npx * 16);
    c->dn_hist_mom[k].ensure(npx * 8);
  }
  c->dn_h_nrm.ensure(npx * 16);
  c->dn_h_inst.ensure(npx * 4);
  c->dn_h_t.ensure(npx * 4);
  f->d_denoised.ensure(npx * 16);
  // this frame's camera: world -> clip = P V = (V^-1 P^-1)^-1 from the two inverses the uniform carries (RT/RTApp.cpp:44-49)
  double vi[16], pi[16], a[16], vp[16];
  for (int i = 0; i < 16; ++i) { vi[i] = u.viewInverse[i]; pi[i] = u.projInverse[i]; }
  for (int r = 0; r < 4; ++r)
    for (int k = 0; k < 4; ++k) {
      double acc = 0.0;
      for (int j = 0; j < 4; ++j) acc += vi[4 * r + j] * pi[4 * j + k];
      a[4 * r + k] = acc;
    }
  invert4x4(a, vp);
  const GBuffer g{f->d_aov_pos.as<float4>(), f->d_aov_nrm.as<float4>(), f->d_aov_inst.as<uint32_t>(), f->d_aov_t.as<float>()};
  const GBuffer hg{nullptr, c->dn_h_nrm.as<float4>(), c->dn_h_inst.as<uint32_t>(), c->dn_h_t.as<float>()};
  const uint32_t par = c->dn_parity;
  const uint32_t grid = grid_for(c, (uint32_t)npx, 256, 8);
  uint32_t launches = 0;
  if (c->dn_event_recorded) BRT_CUDA(cudaStreamWaitEvent(s, c->dn_event, 0));
  BRT_CUDA(cudaEventRecord(f->ev_dn[0], s));
  {
    DnTemporalParams tp{};
    tp.count = (uint32_t)npx;
    tp.width = W;
    tp.height = H;
    tp.color = f->d_image.as<float4>();
    tp.g = g;
    tp.hg = hg;
    tp.h_color = c->dn_hist_color[par].as<float4>();
    tp.h_moments = c->dn_hist_mom[par].as<float2>();
    std::memcpy(tp.prev_vp, c->dn_prev_vp, sizeof(tp.prev_vp));
    std::memcpy(tp.prev_eye, c->dn_prev_eye, sizeof(tp.prev_eye));
    tp.pixel_offset = (f->render_flags & BRT_RENDER_JITTER) ? 0.0f : 0.5f;
    tp.have_history = c->dn_have_history ? 1u : 0u;
    tp.clamp_gamma = d.clamp_gamma;
    tp.max_history = d.max_history >= 1.0f ? d.max_history : 1.0f;
    tp.out = c->dn_work[0].as<float4>();
    tp.out_h_color = c->dn_hist_color[par ^ 1].as<float4>();
    tp.out_h_moments = c->dn_hist_mom[par ^ 1].as<float2>();
    BRT_LAUNCH_1D(k_dn_temporal, tp, grid, 256, s);
    BRT_CHECK_LAUNCH();
    launches++;
  }
  uint32_t cur = 0;
  const uint32_t passes = std::max(1u, d.iterations + ((d.flags & BRT_DENOISE_BILATERAL) ? 1u : 0u));
  for (uint32_t it = 0; it < passes; ++it) {
    const bool none = d.iterations == 0 && !(d.flags & BRT_DENOISE_BILATERAL);  // accumulation only: alpha = 1 through a zero-radius pass
    const bool bilateral = it >= d.iterations;
    const bool last = it + 1 == passes;
    DnAtrousParams ap{};
    ap.count = (uint32_t)npx;
    ap.width = W;
    ap.height = H;
    ap.step = bilateral ? 1 : (1 << it);
    ap.radius = none ? 0 : (bilateral ? 1 : 2);
    ap.sigma_z = d.sigma_z;
    ap.sigma_l = d.sigma_l;
    ap.sigma_n_log2 = d.sigma_n_log2;
    ap.final_pass = last ? 1u : 0u;
    ap.in = c->dn_work[cur].as<float4>();
    ap.g = g;
    ap.out = last ? f->d_denoised.as<float4>() : c->dn_work[cur ^ 1].as<float4>();  // the result belongs to the slot (its copy-out may
    BRT_LAUNCH_1D(k_dn_atrous, ap, grid, 256, s);                                    // still run when the next frame is filtered)
    BRT_CHECK_LAUNCH();
    launches++;
    cur ^= 1;
  }
  BRT_CUDA(cudaEventRecord(f->ev_dn[1], s));
  // this frame becomes the history of the next one
  BRT_CUDA(cudaMemcpyAsync(c->dn_h_nrm.ptr(), f->d_aov_nrm.ptr(), npx * 16, cudaMemcpyDeviceToDevice, s));
  BRT_CUDA(cudaMemcpyAsync(c->dn_h_inst.ptr(), f->d_aov_inst.ptr(), npx * 4, cudaMemcpyDeviceToDevice, s));
  BRT_CUDA(cudaMemcpyAsync(c->dn_h_t.ptr(), f->d_aov_t.ptr(), npx * 4, cudaMemcpyDeviceToDevice, s));
  BRT_CUDA(cudaEventRecord(c->dn_event, s));
  c->dn_event_recorded = true;
  for (int i = 0; i < 16; ++i) c->dn_prev_vp[i] = (float)vp[i];
  c->dn_prev_eye[0] = u.viewInverse[3];
  c->dn_prev_eye[1] = u.viewInverse[7];
  c->dn_prev_eye[2] = u.viewInverse[11];
  c->dn_have_history = true;
  c->dn_parity = par ^ 1;
  f->has_denoised = true;
  f->dn_launches = launches;
}

// Renders into d_image (and d_tiles); no host synchronisation except the final stats read-back.
//
// Two streams. The closest-hit chain  raygen -> { k_trace<closest> -> k_shade } per round  runs on the main stream; the
// shadow chain of round k  { k_trace<occlusion> -> k_accumulate }  runs on the second stream and overlaps with round
// k+1's closest-hit traversal (they are independent: shade(k) produced both the shadow rays of round k and the paths of
// round k+1). The buffers the shadow chain owns (contributions, weights/pixels, shadow queue, counters) are
// double-buffered by round parity; shade(k+2) waits for accumulate(k). Accumulation stays in round order on one stream,
// so the result is bit-identical to the serial schedule (BRT_CFG_NO_OVERLAP, used for per-kernel timing).
void render_frame_device(brt_context* c, FrameSlot* f, const brt_uniform& u, const brt_render_opts& o, void* d_tiles_out, bool to_peers = false) {
  if (!c->built) bad_state("render_frame: scene not built (call brt_scene_build)");
  if (!o.width || !o.height || !o.spp) invalid("render_frame: width, height and spp must be non-zero");
  const bool lbvh = (o.flags & BRT_RENDER_LIGHT_BVH) != 0u;
  if (!lbvh && c->lights.size() > BRT_MAX_LIGHTS) throw LimitError("render_frame: more than BRT_MAX_LIGHTS lights (use BRT_RENDER_LIGHT_BVH)");
  if (c->tables_dirty || c->tlas_dirty) {
    // the tables are uploaded (and the TLAS built) on the context's stream; frames of slots >= 1 run on their own non-blocking
    // streams, so the scene must be in place before any of them is enqueued (a scene change is rare: a plain wait is enough)
    if (c->tables_dirty) build_tables(c);
    if (c->tlas_dirty) build_tlas(c);
    BRT_CUDA(cudaStreamSynchronize(c->stream));
    if (c->tlas.levels + max_blas_levels(c) > BRT_MAX_TREE_LEVELS) {
      c->built = false;
      throw LimitError("render_frame: BVH deeper than the traversal stack allows");
    }
  }
  if ((uint64_t)tiles_per_rank(o.width, o.height, c->tile_world) * 1024u > (1ull << BRT_SLOT_BITS))
    throw LimitError("render_frame: more than 2^26 pixels per rank (path ids carry the pixel slot in 26 bits)");
  const bool any_bounce = (o.flags & (BRT_RENDER_BOUNCE_REFLECT | BRT_RENDER_BOUNCE_REFRACT | BRT_RENDER_BOUNCE_DIFFUSE)) != 0;
  const uint32_t rounds = any_bounce ? u.depthMax : std::min(u.depthMax, 1u);
  ensure_frame_buffers(c, f, o, rounds);
  cudaStream_t s = f->stream;
  cudaStream_t s2 = (c->flags & BRT_CFG_NO_OVERLAP) ? f->stream : f->stream2;
  const TileMap map = make_tile_map(c, o);
  const uint32_t cap = f->cap;
  const size_t npx = (size_t)o.width * o.height;
  const uint32_t n_lights = (uint32_t)c->lights.size();
  const uint32_t n_slots = lbvh ? 1u : std::max(1u, n_lights);  // contribution / shadow-queue segments per path
  // buffers that only some frames need are sized before any stream work (nothing below may allocate: the frame may be captured)
  const uint32_t format = (o.flags & BRT_RENDER_FORMAT_MASK) >> BRT_RENDER_FORMAT_SHIFT;
  if (format > BRT_FORMAT_B8G8R8A8_SRGB) invalid("render_frame: unknown BRT_RENDER_FORMAT");
  if (format != BRT_FORMAT_R32G32B32A32_SFLOAT && c->tile_world > 1 && !to_peers)
    invalid("render_frame: with tile_world > 1 the 8-bit output formats need the fused exchange (brt_render_frame_peers*)");
  const bool denoise = (o.flags & BRT_RENDER_DENOISE) != 0u;
  if (denoise && c->tile_world > 1) invalid("render_frame: BRT_RENDER_DENOISE needs tile_world == 1");
  f->has_gbuffer = (o.flags & (BRT_RENDER_GBUFFER | BRT_RENDER_DENOISE)) != 0u;
  f->has_denoised = false;
  f->render_flags = o.flags;
  {
    const void* before[3] = {f->d_aov_pos.ptr(), f->d_aov_nrm.ptr(), f->d_image8.ptr()};
    if (f->has_gbuffer) {
      f->d_aov_pos.ensure(npx * 16);
      f->d_aov_nrm.ensure(npx * 16);
    }
    if (format != BRT_FORMAT_R32G32B32A32_SFLOAT && !to_peers) f->d_image8.ensure(npx * 4);
    if (before[0] != f->d_aov_pos.ptr() || before[1] != f->d_aov_nrm.ptr() || before[2] != f->d_image8.ptr()) f->generation++;
  }
  // per-frame constants: staged in pinned memory, copied by the first node of the frame
  f->d_consts.ensure(sizeof(FrameConsts));
  std::memcpy(f->h_consts->Vi, u.viewInverse, 64);
  std::memcpy(f->h_consts->Pi, u.projInverse, 64);
  f->h_consts->frame = u.frame;
  f->h_consts->sky = c->sky;
  for (int k = 0; k < 3; ++k) {
    const float ext = c->scene_hi[k] - c->scene_lo[k];
    f->h_consts->scene_lo[k] = c->scene_lo[k];
    f->h_consts->scene_inv[k] = ext > 0.0f ? (float)(1u << BRT_HITSORT_CELL_BITS) / ext : 0.0f;
  }
  // Frames in flight are staggered: this frame starts when the previous one has traced round 0 of its first sample batch, so
  // that its saturating head overlaps the latency-bound tail (bounce rounds, resolve, copy-out) of the previous frame
  // instead of running in lockstep with it, and two multi-batch frames run side by side (profiles/r2_configs.md).
  // where in this frame the NEXT one may start: frames of up to three rounds have no latency-bound tail worth waiting for and release their
  // successor right behind round 0's closest-hit trace, longer ones behind round 0's occlusion trace (measured: profiles/r2_configs.md)
  const uint32_t stagger = c->stagger != 4u ? c->stagger : (rounds <= 3u ? 1u : 3u);
  if (c->stagger && c->prev_head && c->prev_head != f->ev_head) BRT_CUDA(cudaStreamWaitEvent(s, c->prev_head, 0));

  // Full-frame clears that only the pixels this slot never writes depend on (AOVs of pixels outside the crop / of other ranks' tiles,
  // the local image outside this rank's tiles): every owned pixel inside the crop is rewritten by every frame, so they are done once per
  // buffer generation and traced window instead of once per frame (at 4K they were 233 MB of memsets per frame on every rank).
  {
    const uint64_t ck[3] = {((uint64_t)f->generation << 32) | c->tile_world, ((uint64_t)map.crop_x0 << 48) | ((uint64_t)map.crop_y0 << 32) | ((uint64_t)map.crop_x1 << 16) | map.crop_y1,
                            ((uint64_t)o.width << 32) | o.height};
    if (std::memcmp(ck, f->cleared_key, sizeof(ck)) != 0) {
      BRT_CUDA(cudaMemsetAsync(f->d_aov_prim.ptr(), 0xff, npx * 4, s));
      BRT_CUDA(cudaMemsetAsync(f->d_aov_inst.ptr(), 0xff, npx * 4, s));
      BRT_CUDA(cudaMemsetAsync(f->d_aov_t.ptr(), 0, npx * 4, s));
      if (c->tile_world > 1) BRT_CUDA(cudaMemsetAsync(f->d_image.ptr(), 0, npx * 16, s));
      std::memcpy(f->cleared_key, ck, sizeof(ck));
    }
  }

  // The stream work of the frame (about 45 dependent kernel launches, memsets and event records for C2) is captured into a CUDA
  // graph the first time a frame shape is seen on this slot and replayed afterwards: everything that changes from frame to frame
  // travels through FrameConsts, queue sizes already live on the device.
  const uint32_t gimg = to_peers ? (uint32_t)(f - c->slots) : 0u;  // gather image this frame's resolve stores into: the slot's own
  if (to_peers) {
    if (gimg >= BRT_GATHER_IMAGES) invalid("render_frame_peers: slot >= BRT_GATHER_IMAGES");
    if (!c->n_peers || c->gather_w != o.width || c->gather_h != o.height) bad_state("render_frame_peers: gather images not exported / opened for this frame size");
    if (!f->d_resolve_done.ptr()) {
      f->d_resolve_done.ensure(16);
      BRT_CUDA(cudaMemsetAsync(f->d_resolve_done.ptr(), 0, 16, c->stream));
      BRT_CUDA(cudaStreamSynchronize(c->stream));
    }
    f->h_consts->gather_seq = ++c->gather_seq[gimg];
  }
  auto enqueue = [&]() {
    BRT_CUDA(cudaMemcpyAsync(f->d_consts.ptr(), f->h_consts, sizeof(FrameConsts), cudaMemcpyHostToDevice, s));
    f->events_used = 0;
    FrameCounters* ctr = f->d_counters.as<FrameCounters>();
    ShadowCounters* sctr = reinterpret_cast<ShadowCounters*>(ctr + 1);
    FrameStats* fst = f->d_fstats.as<FrameStats>();
    EventPair& whole = next_events(f, CLS_COUNT);
    record_event(whole.a, s);
    BRT_CUDA(cudaMemsetAsync(f->d_accum.ptr(), 0, (size_t)cap * 16, s));
    BRT_CUDA(cudaMemsetAsync(fst, 0, sizeof(FrameStats), s));
    if (f->has_gbuffer) {
      BRT_CUDA(cudaMemsetAsync(f->d_aov_pos.ptr(), 0, npx * 16, s));
      BRT_CUDA(cudaMemsetAsync(f->d_aov_nrm.ptr(), 0, npx * 16, s));
    }
    uint32_t launches = 0, l_closest = 0, l_occl = 0;
    const Node8* tlas = c->tlas_count ? c->d_tlas_nodes.as<Node8>() : nullptr;
    const InstRec* insts = c->d_tlas_inst.as<InstRec>();
    bool acc_pending[2] = {false, false};  // ev_acc[p] has been recorded and not yet waited for by the main stream
    uint32_t global_round = 0;

    const uint32_t c_batch = f->batch;
    for (uint32_t sample = 0; sample < o.spp; sample += f->batch) {
      const uint32_t nb = std::min(f->batch, o.spp - sample);  // samples in this wavefront
      const bool head_batch = c->stagger_first ? sample == 0 : sample + c_batch >= o.spp;  // the batch whose round 0 releases the next frame in flight
      const uint32_t capw = cap * nb;                          // its path slots
      if (rounds) BRT_CUDA(cudaMemsetAsync(f->d_alive.ptr(), 0, (size_t)capw * 4, s));  // (the radiance terms themselves need no clearing)
      {
        RaygenParams rp;
        rp.count = capw;
        rp.count_ptr = nullptr;
        rp.map = map;
        rp.fc = f->d_consts.as<FrameConsts>();
        rp.sample0 = sample;
        rp.flags = o.flags;
        rp.cap = cap;
        rp.q = queue_of(f, 0);
        Timed t(f, CLS_RAYGEN, s);
        BRT_LAUNCH_1D(k_raygen, rp, grid_for(c, capw, 256, 8), 256, s);
        BRT_CHECK_LAUNCH();
        launches++;
      }
      int cur = 0;
      for (uint32_t round = 0; round < rounds; ++round, ++global_round) {
        const int par = (int)(global_round & 1u);
        // main-chain counters: n_paths[cur] is live (round 0 walks all `capw` slots, padding slots carry id == BRT_MISS)
        if (round == 0) {
          BRT_CUDA(cudaMemsetAsync(ctr, 0, sizeof(FrameCounters), s));
        } else {
          BRT_CUDA(cudaMemsetAsync(&ctr->n_paths[cur ^ 1], 0, 4, s));
          BRT_CUDA(cudaMemsetAsync(&ctr->work_closest, 0, 4, s));
        }
        const uint32_t* count_ptr = round == 0 ? nullptr : &ctr->n_paths[cur];
        const PathQueue qc = queue_of(f, cur), qn = queue_of(f, cur ^ 1);
        {
          TraceParams tp{};
          tp.count = capw;
          tp.count_ptr = count_ptr;
          tp.tlas = tlas;
          tp.insts = insts;
          tp.o = qc.o;
          tp.d = qc.d;
          tp.px = qc.px;
          tp.hit = f->d_hit.as<float4>();
          tp.hit_inst = f->d_hit_inst.as<uint32_t>();
          tp.work = &ctr->work_closest;
          tp.stats = fst;
          tp.refill_lanes = round == 0 ? c->refill_primary : c->refill_bounce;  // measured: refill pays for bounce rays only
          tp.defer_lanes = round == 0 ? c->defer_primary : c->defer_bounce;
          Timed t(f, CLS_CLOSEST, s);
          launch_trace<false>(c, tp, s, round == 0 ? ~0ull : (uint64_t)capw >> (round - 1));  // (a bounce round holds at most the previous round's hits)
          launches++;
          l_closest++;
          if (stagger == 1 && round == 0 && head_batch) record_event(f->ev_head, s);
        }
        const uint32_t* order = nullptr;
#ifndef BRT_EMU
        if (round > 0 && (c->flags & BRT_CFG_HIT_SORT)) {  // hit sort (render_kernels.cuh): shade this bounce round in the order of the hit positions
          HitSortParams hp{};
          hp.count = capw;
          hp.count_ptr = count_ptr;
          hp.o = qc.o;
          hp.d = qc.d;
          hp.px = qc.px;
          hp.hit = f->d_hit.as<float4>();
          hp.hit_inst = f->d_hit_inst.as<uint32_t>();
          hp.fc = f->d_consts.as<FrameConsts>();
          hp.key = f->d_sort_key.as<uint32_t>();
          hp.rank = f->d_sort_rank.as<uint32_t>();
          hp.bins = f->d_sort_bins.as<uint32_t>();
          hp.order = f->d_sort_order.as<uint32_t>();
          BRT_CUDA(cudaMemsetAsync(hp.bins, 0, (size_t)BRT_HITSORT_BINS * 4, s));
          Timed t(f, CLS_SHADE, s);
          k_hit_keys<<<grid_for(c, capw, 256, 8), 256, 0, s>>>(hp);
          k_bins_scan<<<1, 1024, 0, s>>>(hp.bins, BRT_HITSORT_BINS);
          BRT_LAUNCH_1D(k_hit_order, hp, grid_for(c, capw, 256, 8), 256, s);
          BRT_CHECK_LAUNCH();
          launches += 3;
          order = hp.order;
        }
#endif
        // the shadow-chain buffers of this parity were last used two rounds ago: wait for that accumulate
        if (acc_pending[par]) {
          if (s2 != s) BRT_CUDA(cudaStreamWaitEvent(s, f->ev_acc[par], 0));
          acc_pending[par] = false;
        }
        BRT_CUDA(cudaMemsetAsync(&sctr[par], 0, sizeof(ShadowCounters), s));
        {
          ShadeParams sp{};
          sp.count = capw;
          sp.count_ptr = count_ptr;
          sp.cur = qc;
          sp.next = qn;
          sp.hit = f->d_hit.as<float4>();
          sp.hit_inst = f->d_hit_inst.as<uint32_t>();
          sp.ctr = ctr;
          sp.sctr = &sctr[par];
          sp.aux = f->d_aux[par].as<float4>();
          sp.next_slot = (uint32_t)(cur ^ 1);
          sp.cap = capw;
          sp.inst = c->d_inst_shade.as<InstShade>();
          sp.materials = c->d_materials.as<float>();
          sp.mat_ext = c->d_mat_ext.as<float2>();
          sp.lights = c->d_lights.as<LightRec>();
          sp.light_bvh = c->d_light_bvh.as<float4>();
          sp.n_lights = n_lights;
          sp.contrib = f->d_contrib[par].as<float4>();
          sp.s_o = f->s_o[par].as<float4>();
          sp.s_d = f->s_d[par].as<float4>();
          sp.s_target = f->s_target[par].as<uint32_t>();
          sp.flags = o.flags;
          sp.last_round = round + 1 == rounds ? 1u : 0u;
          sp.write_aov = (sample == 0 && round == 0) ? 1u : 0u;
          sp.map = map;
          sp.aov_prim = f->d_aov_prim.as<uint32_t>();
          sp.aov_inst = f->d_aov_inst.as<uint32_t>();
          sp.aov_t = f->d_aov_t.as<float>();
          sp.aov_pos = f->has_gbuffer ? f->d_aov_pos.as<float4>() : nullptr;
          sp.aov_nrm = f->has_gbuffer ? f->d_aov_nrm.as<float4>() : nullptr;
          sp.fc = f->d_consts.as<FrameConsts>();
          sp.order = order;
          sp.ahead = c->shade_ahead;
          Timed t(f, CLS_SHADE, s);
  #ifdef BRT_EMU
          BRT_LAUNCH_1D(k_shade, sp, 1, 128, s);
  #else
          if (o.flags & BRT_RENDER_FAST_SHADING) launch_shade_fast(sp, grid_for(c, capw, 128, 16), round == 0, s);
          else if (round == 0) k_shade_primary<<<grid_for(c, capw, 128, 16), 128, 0, s>>>(sp);
          else k_shade<BRT_SHADE_WINDOW><<<grid_for(c, capw, 128, 16), 128, 0, s>>>(sp);
  #endif
          BRT_CHECK_LAUNCH();
          launches++;
        }
        if (stagger == 2 && round == 0 && head_batch) record_event(f->ev_head, s);
        if (s2 != s) {
          BRT_CUDA(cudaEventRecord(f->ev_shade, s));
          BRT_CUDA(cudaStreamWaitEvent(s2, f->ev_shade, 0));
        }
        if (n_lights) {
          TraceParams tp{};
          tp.count = 0;
          tp.seg_counts = sctr[par].n_shadow;
          tp.n_segs = n_slots;
          tp.seg_stride = capw;
          tp.tlas = tlas;
          tp.insts = insts;
          tp.o = f->s_o[par].as<float4>();
          tp.d = f->s_d[par].as<float4>();
          tp.target = f->s_target[par].as<uint32_t>();
          tp.contrib = f->d_contrib[par].as<float4>();
          tp.work = &sctr[par].work_occl;
          tp.stats = fst;
          tp.refill_lanes = round == 0 ? c->refill_primary : c->refill_bounce;
          tp.defer_lanes = round == 0 ? c->defer_primary : c->defer_bounce;
          Timed t(f, CLS_OCCL, s2);
          launch_trace<true>(c, tp, s2, round == 0 ? ~0ull : ((uint64_t)capw * n_slots) >> (round - 1));
          launches++;
          l_occl++;
        }
        if ((stagger == 3 || stagger == 0) && round == 0 && head_batch) record_event(f->ev_head, s2);
        {
          AccumParams ap{};
          ap.count = 0;
          ap.count_ptr = &sctr[par].n_items;
          ap.aux = f->d_aux[par].as<float4>();
          ap.contrib = f->d_contrib[par].as<float4>();
          ap.n_slots = n_slots;
          ap.cap = capw;
          ap.slots = cap;
          ap.rounds = rounds;
          ap.round = round;
          ap.rad = f->d_rad.as<float4>();
          ap.alive = f->d_alive.as<uint32_t>();
          Timed t(f, CLS_ACCUM, s2);
          BRT_LAUNCH_1D(k_accumulate, ap, grid_for(c, capw, 256, 8), 256, s2);
          BRT_CHECK_LAUNCH();
          launches++;
        }
        if (s2 != s) BRT_CUDA(cudaEventRecord(f->ev_acc[par], s2));
        acc_pending[par] = true;
        cur ^= 1;
      }
      // join the shadow chain, then add this batch's radiance terms to the per-slot sums in the shader's order
      for (int par = 0; par < 2; ++par) {
        if (acc_pending[par] && s2 != s) BRT_CUDA(cudaStreamWaitEvent(s, f->ev_acc[par], 0));
        acc_pending[par] = false;
      }
      if (rounds) {
        SumSamplesParams sp{};
        sp.count = cap;
        sp.count_ptr = nullptr;
        sp.samples = nb;
        sp.rounds = rounds;
        sp.rad = f->d_rad.as<float4>();
        sp.alive = f->d_alive.as<uint32_t>();
        sp.accum = f->d_accum.as<float4>();
        Timed t(f, CLS_ACCUM, s);
        BRT_LAUNCH_1D(k_sum_samples, sp, grid_for(c, cap, 256, 8), 256, s);
        BRT_CHECK_LAUNCH();
        launches++;
      }
    }
    {
      ResolveParams rp{};
      rp.count = cap;
      rp.count_ptr = nullptr;
      rp.map = map;
      rp.spp = (float)o.spp;
      rp.accum = f->d_accum.as<float4>();
      rp.image = f->d_image.as<float4>();
      rp.tiles = d_tiles_out ? static_cast<float4*>(d_tiles_out) : f->d_tiles.as<float4>();
      rp.n_peers = 0;
      rp.format = BRT_FORMAT_R32G32B32A32_SFLOAT;
#ifndef BRT_EMU
      if (to_peers) {
        // receivers: rank gather_root alone, or everybody (brt_gather_configure)
        const uint32_t first = c->gather_root_only ? c->gather_root : 0u, n_dst = c->gather_root_only ? 1u : c->n_peers;
        const size_t img_bytes = npx * 16;
        rp.tiles = nullptr;
        rp.n_peers = n_dst;
        for (uint32_t k = 0; k < n_dst; ++k) {
          rp.peers[k] = static_cast<char*>(c->peer_images[first + k]) + (size_t)gimg * img_bytes;
          rp.peer_flags[k] = reinterpret_cast<GatherFlags*>(static_cast<char*>(c->peer_images[first + k]) + (size_t)BRT_GATHER_IMAGES * img_bytes);
        }
        rp.format = format;
        rp.img = gimg;
        rp.rank = c->tile_rank;
        rp.done = f->d_resolve_done.as<uint32_t>();
        rp.fc = f->d_consts.as<FrameConsts>();
        GatherFlags* mine = reinterpret_cast<GatherFlags*>(static_cast<char*>(c->d_gather.ptr()) + (size_t)BRT_GATHER_IMAGES * img_bytes);
        Timed t(f, CLS_RESOLVE, s);
        if (c->peer_push) {
          ResolveParams local = rp;  // resolve into the local image and the packed tiles ...
          local.n_peers = 0;
          local.tiles = f->d_tiles.as<float4>();
          BRT_LAUNCH_1D(k_resolve, local, grid_for(c, cap, 256, 8), 256, s);
          rp.tiles = f->d_tiles.as<float4>();
          // ... the receivers must have released the previous frame of this image before its pixels are overwritten (device-side wait) ...
          k_gather_wait_consumed<<<1, 32, 0, s>>>(mine, gimg, rp.fc, first, n_dst);
          // ... and a small grid carries the tiles over NVLink while the SMs work on the other frames in flight
          k_push_peers<<<std::max(1u, std::min(c->peer_grid ? c->peer_grid : 64u * n_dst, grid_for(c, cap, 256, 8))), 256, 0, s>>>(rp);  // (64 blocks per receiver)
          BRT_CHECK_LAUNCH();
          launches += 3;
        } else {
          // the receivers must have released the previous frame of this image before its pixels are overwritten (device-side wait)
          k_gather_wait_consumed<<<1, 32, 0, s>>>(mine, gimg, rp.fc, first, n_dst);
          k_resolve_peers<<<c->peer_grid ? std::min(c->peer_grid, grid_for(c, cap, 256, 8)) : grid_for(c, cap, 256, 8), 256, 0, s>>>(rp);
          BRT_CHECK_LAUNCH();
          launches += 2;
        }
      } else
#endif
      {
        Timed t(f, CLS_RESOLVE, s);
        BRT_LAUNCH_1D(k_resolve, rp, grid_for(c, cap, 256, 8), 256, s);
        BRT_CHECK_LAUNCH();
        launches++;
      }
    }
    const uint32_t format = (o.flags & BRT_RENDER_FORMAT_MASK) >> BRT_RENDER_FORMAT_SHIFT;
    if (format != BRT_FORMAT_R32G32B32A32_SFLOAT && !denoise && !to_peers) {
      PresentParams pp{(uint32_t)npx, nullptr, format, f->d_image.as<float4>(), f->d_image8.as<uint32_t>()};
      Timed t(f, CLS_RESOLVE, s);
      BRT_LAUNCH_1D(k_present, pp, grid_for(c, (uint32_t)npx, 256, 8), 256, s);
      BRT_CHECK_LAUNCH();
      launches++;
    }
    record_event(whole.b, s);
    if (!rounds) record_event(f->ev_head, s);
    f->launches = launches;
    f->l_closest = l_closest;
    f->l_occl = l_occl;
    // device-side statistics of this frame travel back on its stream (read by finish_frame)
    BRT_CUDA(cudaMemcpyAsync(f->fs_host, f->d_fstats.ptr(), sizeof(FrameStats), cudaMemcpyDeviceToHost, s));
  };
#ifdef BRT_EMU
  enqueue();
#else
  // (the legacy default stream a caller may hand to brt_set_stream cannot be captured)
  const bool use_graph = !(c->flags & (BRT_CFG_NO_GRAPH | BRT_CFG_NO_OVERLAP | BRT_CFG_COUNTERS)) && (c->tile_world == 1 || to_peers) &&
                         s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread;
  const uint64_t key[18] = {((uint64_t)o.width << 32) | o.height, ((uint64_t)o.spp << 32) | o.flags, ((uint64_t)o.crop_x0 << 32) | o.crop_y0,
                            ((uint64_t)o.crop_w << 32) | o.crop_h, ((uint64_t)rounds << 32) | n_lights, ((uint64_t)f->generation << 32) | c->tlas_count,
                            (uint64_t)c->d_tlas_nodes.ptr(), (uint64_t)c->d_tlas_inst.ptr(), (uint64_t)c->d_inst_shade.ptr(), (uint64_t)c->d_materials.ptr(),
                            (uint64_t)c->d_mat_ext.ptr(), (uint64_t)c->d_lights.ptr(), (uint64_t)c->d_light_bvh.ptr(), (uint64_t)d_tiles_out, (uint64_t)s,
                            (uint64_t)f->d_consts.ptr(), ((uint64_t)to_peers << 32) | gimg | (c->gather_root_only << 8) | (c->gather_root << 16),
                            to_peers ? (uint64_t)c->peer_images[0] ^ ((uint64_t)c->peer_images[c->n_peers - 1] << 1) ^ ((uint64_t)f->d_resolve_done.ptr() << 2) : 0};
  const int gslot = to_peers ? 1 : 0;  // a slot keeps one graph for plain frames and one for frames of the fused exchange
  cudaGraphExec_t& gexec = f->graph_exec[gslot];
  if (use_graph && gexec && std::memcmp(key, f->graph_key[gslot], sizeof(key)) == 0) {
    BRT_CUDA(cudaGraphLaunch(gexec, s));
  } else if (use_graph) {
    if (gexec) {
      cudaGraphExecDestroy(gexec);
      gexec = nullptr;
    }
    BRT_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
    g_capturing = true;
    cudaGraph_t graph = nullptr;
    try {
      enqueue();
    } catch (...) {
      g_capturing = false;
      cudaStreamEndCapture(s, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    g_capturing = false;
    BRT_CUDA(cudaStreamEndCapture(s, &graph));
    const cudaError_t ie = cudaGraphInstantiate(&gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      gexec = nullptr;
      BRT_CUDA(ie);
    }
    std::memcpy(f->graph_key[gslot], key, sizeof(key));
    BRT_CUDA(cudaGraphLaunch(gexec, s));
  } else {
    enqueue();
  }
#endif
  if (to_peers) {
    f->gather_img = gimg;
    f->gather_seq = c->gather_seq[gimg];
    f->gather_format = format;
  }
  c->prev_head = f->ev_head;
  // Denoiser stages and the present conversion of a denoised frame follow the (replayed) frame as ordinary stream work: their
  // history is shared by all slots, so they wait for the previous frame's stages with stream semantics.
  if (denoise) {
    enqueue_denoise(c, f, u, c->dn_opts, s);
    if (format != BRT_FORMAT_R32G32B32A32_SFLOAT) {
      PresentParams pp{(uint32_t)npx, nullptr, format, f->d_denoised.as<float4>(), f->d_image8.as<uint32_t>()};
      BRT_LAUNCH_1D(k_present, pp, grid_for(c, (uint32_t)npx, 256, 8), 256, s);
      BRT_CHECK_LAUNCH();
    }
  }
  f->in_flight = true;
}

#ifndef BRT_EMU
void close_peers(brt_context* c) {
  for (uint32_t k = 0; k < BRT_MAX_PEERS; ++k) {
    if (c->peer_opened[k]) cudaIpcCloseMemHandle(c->peer_images[k]);
    c->peer_opened[k] = false;
    c->peer_images[k] = nullptr;
  }
  c->n_peers = 0;
}
#endif

// waits for the slot's frame and folds the device-side statistics / event timings into ctx->stats
void finish_frame(brt_context* c, FrameSlot* f) {
  if (!f->in_flight) return;
  BRT_CUDA(cudaStreamSynchronize(f->stream));
  f->in_flight = false;
  BRT_CUDA(cudaGetLastError());
  const FrameStats fs = *f->fs_host;
  float ms[CLS_COUNT + 1] = {0};
  for (size_t i = 0; i < f->events_used; ++i) {
    float t = 0.0f;
    cudaEventElapsedTime(&t, f->events[i].a, f->events[i].b);
    ms[f->events[i].cls] += t;
  }
  brt_stats& st = c->stats;
  st.launches_total = f->launches;
  st.launches_trace_closest = f->l_closest;
  st.launches_trace_occlusion = f->l_occl;
  st.rays_closest = fs.rays_closest;
  st.rays_occlusion = fs.rays_occlusion;
  st.nodes_visited_closest = fs.nodes_c;
  st.prims_tested_closest = fs.prims_c;
  st.spheres_tested_closest = fs.spheres_c;
  st.nodes_visited_occlusion = fs.nodes_o;
  st.prims_tested_occlusion = fs.prims_o;
  st.spheres_tested_occlusion = fs.spheres_o;
  st.ms_raygen = ms[CLS_RAYGEN];
  st.ms_trace_closest = ms[CLS_CLOSEST];
  st.ms_shade = ms[CLS_SHADE];
  st.ms_trace_occlusion = ms[CLS_OCCL];
  st.ms_accumulate = ms[CLS_ACCUM];
  st.ms_resolve = ms[CLS_RESOLVE];
  st.ms_total = ms[CLS_COUNT];
  if (f->has_denoised) {
    float t = 0.0f;
    cudaEventElapsedTime(&t, f->ev_dn[0], f->ev_dn[1]);
    st.ms_denoise = t;
    st.launches_denoise = f->dn_launches;
  }
  c->last_slot = (uint32_t)(f - c->slots);
}

// the frame's image to host memory in the format the caller asked for, on the frame's stream
void copy_out(FrameSlot* f, const brt_render_opts& o, void* host) {
  const size_t npx = (size_t)o.width * o.height;
  if (o.flags & BRT_RENDER_FORMAT_MASK) BRT_CUDA(cudaMemcpyAsync(host, f->d_image8.ptr(), npx * 4, cudaMemcpyDeviceToHost, f->stream));
  else BRT_CUDA(cudaMemcpyAsync(host, f->has_denoised ? f->d_denoised.ptr() : f->d_image.ptr(), npx * 16, cudaMemcpyDeviceToHost, f->stream));
}

// scene changes (builds, uploads, culling) and the synchronous entry points first drain every frame in flight
void wait_all_frames(brt_context* c) {
  for (FrameSlot& f : c->slots) finish_frame(c, &f);
}

// creates a slot's streams and events on first use (slot 0 shares the context's stream)
void ensure_slot(brt_context* c, FrameSlot* f) {
  if (f == &c->slots[0]) f->stream = c->stream;  // (brt_set_stream may have replaced it)
  if (f->ready) return;
  if (f != &c->slots[0]) {
    BRT_CUDA(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
    f->own_stream = true;
  }
  BRT_CUDA(cudaStreamCreateWithFlags(&f->stream2, cudaStreamNonBlocking));
  BRT_CUDA(cudaEventCreateWithFlags(&f->ev_shade, cudaEventDisableTiming));
  BRT_CUDA(cudaEventCreateWithFlags(&f->ev_acc[0], cudaEventDisableTiming));
  BRT_CUDA(cudaEventCreateWithFlags(&f->ev_acc[1], cudaEventDisableTiming));
  BRT_CUDA(cudaEventCreateWithFlags(&f->ev_head, cudaEventDisableTiming));
  BRT_CUDA(cudaEventCreate(&f->ev_dn[0]));
  BRT_CUDA(cudaEventCreate(&f->ev_dn[1]));
  BRT_CUDA(cudaMallocHost(reinterpret_cast<void**>(&f->fs_host), sizeof(FrameStats)));
  BRT_CUDA(cudaMallocHost(reinterpret_cast<void**>(&f->h_consts), sizeof(FrameConsts)));
  f->ready = true;
}

void destroy_slot(FrameSlot* f) {
  if (!f->ready) return;
  if (f->stream) cudaStreamSynchronize(f->stream);
  if (f->stream2) cudaStreamSynchronize(f->stream2);
  for (EventPair& e : f->events) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  if (f->own_stream && f->stream) cudaStreamDestroy(f->stream);
  if (f->stream2) cudaStreamDestroy(f->stream2);
  if (f->ev_shade) cudaEventDestroy(f->ev_shade);
  if (f->ev_head) cudaEventDestroy(f->ev_head);
  for (int k = 0; k < 2; ++k)
    if (f->ev_dn[k]) cudaEventDestroy(f->ev_dn[k]);
  for (int k = 0; k < 2; ++k)
    if (f->ev_acc[k]) cudaEventDestroy(f->ev_acc[k]);
  if (f->fs_host) cudaFreeHost(f->fs_host);
  if (f->h_consts) cudaFreeHost(f->h_consts);
#ifndef BRT_EMU
  for (int k = 0; k < 2; ++k)
    if (f->graph_exec[k]) cudaGraphExecDestroy(f->graph_exec[k]);
#endif
}

}  // namespace

// ========================================================================================================
extern "C" {

int brt_create(const brt_config* cfg, brt_context** out) {
  if (!out) return BRT_ERR_INVALID;
  *out = nullptr;
  brt_context* c = new brt_context();
  int rc = guarded(c, [&] {
    if (cfg) {
      c->device = cfg->device;
      c->tile_rank = cfg->tile_rank;
      c->tile_world = cfg->tile_world ? cfg->tile_world : 1;
      c->flags = cfg->flags;
      if (c->tile_rank >= c->tile_world) invalid("brt_create: tile_rank >= tile_world");
    }
    // more hardware work queues than the default 8 (read when CUDA initialises: only effective if this is the first CUDA call of the
    // process): two streams per frame in flight plus the exchange stream must not share queues
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = 0;
    BRT_CUDA(cudaGetDeviceCount(&n));
    if (n <= 0 || c->device < 0 || c->device >= n) throw CudaError("brt_create: no usable CUDA device (this library has no CPU path)");
    BRT_CUDA(cudaSetDevice(c->device));
    BRT_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
    BRT_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int k = 0; k < 3; ++k) BRT_CUDA(cudaEventCreate(&c->ev_t[k]));
    BRT_CUDA(cudaEventCreateWithFlags(&c->dn_event, cudaEventDisableTiming));
    c->own_stream = true;
    c->builder.reset(new Builder(c->sm_count, (c->flags & BRT_CFG_GREEDY_COLLAPSE) != 0, (c->flags & BRT_CFG_TREELET_PASSES_3) != 0));
    if (const char* e = getenv("BRT_WAVEFRONT_PATHS")) c->target_wavefront = (uint32_t)std::max(1L, atol(e));
    // measured slower on C2 / C3 / C5 at every distance tried (profiles/r2_ncu_summary.md §4): off unless asked for
    if (const char* e = getenv("BRT_SHADE_AHEAD")) c->shade_ahead = (uint32_t)std::max(0L, atol(e));
    if (const char* e = getenv("BRT_STAGGER_FIRST")) c->stagger_first = atoi(e) != 0;
    if (const char* e = getenv("BRT_STAGGER")) c->stagger = (uint32_t)std::max(0L, std::min(4L, atol(e)));
    if (const char* e = getenv("BRT_PEER_PUSH")) c->peer_push = atoi(e) != 0;
    if (const char* e = getenv("BRT_PEER_GRID")) c->peer_grid = (uint32_t)std::max(0L, atol(e));
    if (const char* e = getenv("BRT_RAYS_PER_WARP")) c->rays_per_warp = (uint32_t)std::max(0L, atol(e));
    if (const char* e = getenv("BRT_REFILL_PRIMARY")) c->refill_primary = (uint32_t)std::max(0L, std::min(32L, atol(e)));
    if (const char* e = getenv("BRT_REFILL_BOUNCE")) c->refill_bounce = (uint32_t)std::max(0L, std::min(32L, atol(e)));
    if (const char* e = getenv("BRT_DEFER_PRIMARY")) c->defer_primary = (uint32_t)std::max(0L, std::min(32L, atol(e)));  // tuning aids
    if (const char* e = getenv("BRT_DEFER_BOUNCE")) c->defer_bounce = (uint32_t)std::max(0L, std::min(32L, atol(e)));
  });
  if (rc != BRT_OK) {
    g_create_error = c->err;
    delete c;
    return rc;
  }
  *out = c;
  return BRT_OK;
}

void brt_destroy(brt_context* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (FrameSlot& f : c->slots) destroy_slot(&f);
#ifndef BRT_EMU
  close_peers(c);
#endif
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  for (int k = 0; k < 3; ++k)
    if (c->ev_t[k]) cudaEventDestroy(c->ev_t[k]);
  if (c->dn_event) cudaEventDestroy(c->dn_event);
  if (c->gather_stream) cudaStreamDestroy(c->gather_stream);
  delete c;
}

const char* brt_last_error(const brt_context* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int brt_set_stream(brt_context* c, void* stream) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    wait_all_frames(c);
    BRT_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream) cudaStreamDestroy(c->stream);
    c->own_stream = false;
    c->stream = static_cast<cudaStream_t>(stream);
  });
}

int brt_mesh_create(brt_context* c, const brt_vertex* v, uint32_t nv, const uint32_t* idx, uint32_t ni, uint32_t* mesh_id) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    wait_all_frames(c);
    if ((!v && nv) || (!idx && ni) || ni % 3) invalid("mesh_create: bad arguments");
    for (uint32_t i = 0; i < ni; ++i)
      if (idx[i] >= nv) invalid("mesh_create: index out of range");
    BRT_CUDA(cudaSetDevice(c->device));
    std::unique_ptr<MeshData> m(new MeshData());
    m->n_vertices = nv;
    m->n_indices = ni;
    upload(c->stream, m->vertices, v, (size_t)nv * sizeof(brt_vertex));
    upload(c->stream, m->indices, idx, (size_t)ni * 4);
    if (ni) {  // everything the first build of this mesh needs is allocated here, not inside the build (it was 128 ms cold)
      c->builder->reserve(ni / 3);
      m->nodes.ensure((size_t)Builder::node_capacity(ni / 3) * sizeof(Node8));
      m->tris.ensure((size_t)(ni / 3) * sizeof(TriRec));
    }
    BRT_CUDA(cudaStreamSynchronize(c->stream));  // "the library copies all input arrays during the call"
    c->meshes.push_back(std::move(m));
    if (mesh_id) *mesh_id = (uint32_t)c->meshes.size() - 1;
    c->built = false;
  });
}

int brt_mesh_update_vertices(brt_context* c, uint32_t mesh_id, const brt_vertex* v, uint32_t nv) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    wait_all_frames(c);
    if (mesh_id >= c->meshes.size() || c->meshes[mesh_id]->sphere || nv != c->meshes[mesh_id]->n_vertices || (!v && nv))
      invalid("mesh_update_vertices: bad mesh id or vertex count");
    BRT_CUDA(cudaSetDevice(c->device));
    MeshData& m = *c->meshes[mesh_id];
    BRT_CUDA(cudaMemcpyAsync(m.vertices.ptr(), v, (size_t)nv * sizeof(brt_vertex), cudaMemcpyHostToDevice, c->stream));
    BRT_CUDA(cudaStreamSynchronize(c->stream));
    m.dirty = true;
    m.dynamic = true;
    c->built = false;
  });
}

int brt_sphere_create(brt_context* c, const float center[3], float radius, uint32_t* mesh_id) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    wait_all_frames(c);
    if (!center || !(radius > 0.0f)) invalid("sphere_create: bad arguments");
    std::unique_ptr<MeshData> m(new MeshData());
    m->sphere = true;
    std::memcpy(m->center, center, 12);
    m->radius = radius;
    c->meshes.push_back(std::move(m));
    if (mesh_id) *mesh_id = (uint32_t)c->meshes.size() - 1;
    c->built = false;
  });
}

int brt_material_create(brt_context* c, const brt_material* m, uint32_t* id) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!m) invalid("material_create: null");
    c->materials.push_back(*m);
    c->mat_ext.push_back(0.0f);
    c->mat_ext.push_back(1.5f);
    if (id) *id = (uint32_t)c->materials.size() - 1;
    c->tables_dirty = true;
  });
}

int brt_material_set_transmission(brt_context* c, uint32_t id, float transmission, float ior) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (id >= c->materials.size()) invalid("material_set_transmission: bad id");
    c->mat_ext[2 * id] = transmission;
    c->mat_ext[2 * id + 1] = ior;
    c->tables_dirty = true;
  });
}

int brt_light_create(brt_context* c, const brt_light* l, uint32_t* id) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!l) invalid("light_create: null");
    c->lights.push_back(*l);
    if (id) *id = (uint32_t)c->lights.size() - 1;
    c->tables_dirty = true;
  });
}

int brt_sky_set(brt_context* c, const brt_sky* sky) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!sky) invalid("sky_set: null");
    c->sky = *sky;
    c->tables_dirty = true;  // (the device copy behind SceneBufferInfo.skyBuf)
  });
}

int brt_instance_create(brt_context* c, uint32_t mesh, uint32_t mat, const float x[12], uint32_t* id) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!x || mesh >= c->meshes.size() || mat >= c->materials.size()) invalid("instance_create: bad mesh/material id");
    InstanceData in;
    in.mesh = mesh;
    in.material = mat;
    std::memcpy(in.o2w, x, 48);
    invert3x4(in.o2w, in.w2o);
    c->instances.push_back(in);
    c->visible.push_back(1);
    if (id) *id = (uint32_t)c->instances.size() - 1;
    c->built = false;
    c->tables_dirty = c->tlas_dirty = true;
  });
}

int brt_instance_set_transform(brt_context* c, uint32_t id, const float x[12]) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!x || id >= c->instances.size()) invalid("instance_set_transform: bad id");
    std::memcpy(c->instances[id].o2w, x, 48);
    invert3x4(c->instances[id].o2w, c->instances[id].w2o);
    c->built = false;
    c->tables_dirty = c->tlas_dirty = true;
  });
}

int brt_instance_set_material(brt_context* c, uint32_t id, uint32_t mat) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (id >= c->instances.size() || mat >= c->materials.size()) invalid("instance_set_material: bad id");
    c->instances[id].material = mat;
    c->tables_dirty = true;
  });
}

int brt_instance_destroy(brt_context* c, uint32_t id) {  // RT/Scene.cpp:122-125: swap-remove
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (id >= c->instances.size()) invalid("instance_destroy: bad id");
    c->instances[id] = c->instances.back();
    c->instances.pop_back();
    c->visible[id] = c->visible.back();
    c->visible.pop_back();
    c->built = false;
    c->tables_dirty = c->tlas_dirty = true;
  });
}

int brt_scene_build(brt_context* c) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    wait_all_frames(c);
    BRT_CUDA(cudaSetDevice(c->device));
    scene_build(c);
  });
}

int brt_smart_cull(brt_context* c, const brt_uniform* u, uint32_t width, uint32_t height, float threshold_px2, float hysteresis,
                   uint32_t* visible_count) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    wait_all_frames(c);
    if (!u) invalid("smart_cull: null");
    (void)width;
    BRT_CUDA(cudaSetDevice(c->device));
    bool dirty_mesh = false;
    for (const auto& m : c->meshes) dirty_mesh |= m->dirty;
    if (dirty_mesh || !c->built) scene_build(c, false);  // BLASes of updated meshes; the TLAS follows below, once
    if (c->tables_dirty) build_tables(c);
    cudaStream_t s = c->stream;
    const uint32_t n = (uint32_t)c->instances.size();
    cudaEvent_t e0 = c->ev_t[0], e1 = c->ev_t[1], e2 = c->ev_t[2];
    BRT_CUDA(cudaEventRecord(e0, s));
    if (n) {
      upload(s, c->d_visible, c->visible.data(), n);
      CullParams p{};
      p.count = n;
      p.count_ptr = nullptr;
      p.inst = c->d_inst_shade.as<InstShade>();
      p.mesh_bounds = c->d_mesh_bounds.as<float4>();
      const float* Vi = u->viewInverse;
      p.eye[0] = Vi[3]; p.eye[1] = Vi[7]; p.eye[2] = Vi[11];
      p.fwd[0] = Vi[2]; p.fwd[1] = Vi[6]; p.fwd[2] = Vi[10];
      const float p11 = 1.0f / u->projInverse[5];
      p.k = p11 * ((float)height * 0.5f);
      p.threshold = threshold_px2;
      p.hysteresis = hysteresis;
      p.visible = c->d_visible.as<uint8_t>();
      BRT_LAUNCH_1D(k_cull, p, grid_for(c, n, 128, 8), 128, s);
      BRT_CHECK_LAUNCH();
      BRT_CUDA(cudaMemcpyAsync(c->visible.data(), c->d_visible.ptr(), n, cudaMemcpyDeviceToHost, s));
    }
    BRT_CUDA(cudaEventRecord(e1, s));
    BRT_CUDA(cudaStreamSynchronize(s));
    build_tlas(c);
    BRT_CUDA(cudaEventRecord(e2, s));
    BRT_CUDA(cudaStreamSynchronize(s));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    c->stats.ms_cull = ms;
    cudaEventElapsedTime(&ms, e1, e2);
    c->stats.ms_tlas_build = ms;
    if (c->tlas.levels + max_blas_levels(c) > BRT_MAX_TREE_LEVELS) {
      c->built = false;  // (scene_build above set it: a later render must not traverse this tree with the fixed-size stack)
      throw LimitError("smart_cull: BVH deeper than the traversal stack allows");
    }
    refresh_scene_stats(c);
    if (visible_count) {
      uint32_t k = 0;
      for (uint8_t v : c->visible) k += v ? 1u : 0u;
      *visible_count = k;
    }
  });
}

int brt_get_visibility(brt_context* c, uint8_t* out, uint32_t n) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!out || n != c->instances.size()) invalid("get_visibility: size mismatch");
    std::memcpy(out, c->visible.data(), n);
  });
}

int brt_get_scene_info_buffer(brt_context* c, brt_scene_buffer_info* out, uint64_t* d_copy) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!out) invalid("get_scene_info_buffer: null");
    if (!c->built) bad_state("get_scene_info_buffer: scene not built");
    BRT_CUDA(cudaSetDevice(c->device));
    if (c->tables_dirty) {
      wait_all_frames(c);
      build_tables(c);
      BRT_CUDA(cudaStreamSynchronize(c->stream));
    }
    *out = c->scene_info;
    if (d_copy) *d_copy = (uint64_t)c->d_scene_info.ptr();
  });
}

int brt_get_tlas(brt_context* c, brt_accel_info* out) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!out) invalid("get_tlas: null");
    if (!c->built) bad_state("get_tlas: scene not built");
    const uint64_t nodes = c->tlas_count ? (uint64_t)c->d_tlas_nodes.ptr() : 0;
    *out = brt_accel_info{nodes, nodes, c->tlas_count ? (uint64_t)c->d_tlas_inst.ptr() : 0, nodes, c->tlas.n_nodes, c->tlas_count};
  });
}

int brt_render_frame(brt_context* c, const brt_uniform* u, const brt_render_opts* o, float* rgba_host) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!u || !o) invalid("render_frame: null");
    BRT_CUDA(cudaSetDevice(c->device));
    wait_all_frames(c);
    FrameSlot* f = &c->slots[0];
    ensure_slot(c, f);
    render_frame_device(c, f, *u, *o, nullptr);
    if (rgba_host) copy_out(f, *o, rgba_host);
    finish_frame(c, f);
  });
}

int brt_render_frame_async(brt_context* c, const brt_uniform* u, const brt_render_opts* o, uint32_t slot, float* rgba_host) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!u || !o) invalid("render_frame_async: null");
    if (slot >= BRT_FRAMES_IN_FLIGHT) invalid("render_frame_async: slot >= BRT_FRAMES_IN_FLIGHT");
    BRT_CUDA(cudaSetDevice(c->device));
    FrameSlot* f = &c->slots[slot];
    finish_frame(c, f);  // the fence of this slot (VK/SwapChain.cpp:45-60): its previous frame must be complete
    if (c->tables_dirty || c->tlas_dirty) wait_all_frames(c);  // the scene tables are about to be rewritten
    ensure_slot(c, f);
    render_frame_device(c, f, *u, *o, nullptr);
    if (rgba_host) copy_out(f, *o, rgba_host);
  });
}

int brt_frame_wait(brt_context* c, uint32_t slot) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (slot >= BRT_FRAMES_IN_FLIGHT) invalid("frame_wait: slot >= BRT_FRAMES_IN_FLIGHT");
    BRT_CUDA(cudaSetDevice(c->device));
    finish_frame(c, &c->slots[slot]);
  });
}

void* brt_frame_stream(brt_context* c, uint32_t slot) {
  if (!c || slot >= BRT_FRAMES_IN_FLIGHT) return nullptr;
  if (guarded(c, [&] { BRT_CUDA(cudaSetDevice(c->device)); ensure_slot(c, &c->slots[slot]); }) != BRT_OK) return nullptr;
  return c->slots[slot].stream;
}

int brt_render_frame_tiles(brt_context* c, const brt_uniform* u, const brt_render_opts* o, void* d_tiles) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!u || !o) invalid("render_frame_tiles: null");
    BRT_CUDA(cudaSetDevice(c->device));
    wait_all_frames(c);
    FrameSlot* f = &c->slots[0];
    ensure_slot(c, f);
    render_frame_device(c, f, *u, *o, d_tiles);
    finish_frame(c, f);
  });
}

size_t brt_tile_buffer_bytes(uint32_t width, uint32_t height, uint32_t tile_world) {
  return (size_t)tiles_per_rank(width, height, tile_world) * 1024u * 16u;
}

int brt_untile(brt_context* c, const void* d_all, uint32_t width, uint32_t height, uint32_t tile_world, void* d_rgba) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!d_all || !d_rgba || !width || !height || !tile_world) invalid("untile: bad arguments");
    BRT_CUDA(cudaSetDevice(c->device));
    UntileParams p{};
    p.count = width * height;
    p.count_ptr = nullptr;
    p.width = width;
    p.height = height;
    p.tiles_x = div_up(width, BRT_TILE);
    p.tile_world = tile_world;
    p.slots_per_rank = tiles_per_rank(width, height, tile_world) * 1024u;
    p.all = static_cast<const float4*>(d_all);
    p.image = static_cast<float4*>(d_rgba);
    BRT_LAUNCH_1D(k_untile, p, grid_for(c, p.count, 256, 8), 256, c->stream);
    BRT_CHECK_LAUNCH();
  });
}

void* brt_device_image(brt_context* c) { return c ? c->slots[c->last_slot].d_image.ptr() : nullptr; }

int brt_gather_image_export(brt_context* c, uint32_t width, uint32_t height, void* handle_out) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!width || !height || !handle_out) invalid("gather_image_export: bad arguments");
#ifdef BRT_EMU
    bad_state("gather_image_export: peer memory needs CUDA devices");
#else
    static_assert(sizeof(cudaIpcMemHandle_t) == BRT_IPC_HANDLE_BYTES, "ipc handle size");
    BRT_CUDA(cudaSetDevice(c->device));
    wait_all_frames(c);
    close_peers(c);
    c->d_gather.release();  // a fresh allocation: an exported handle stays tied to its allocation
    const size_t bytes = (size_t)width * height * 16 * BRT_GATHER_IMAGES + sizeof(GatherFlags);
    c->d_gather.ensure(bytes);
    BRT_CUDA(cudaMemset(c->d_gather.ptr(), 0, bytes));
    for (uint32_t k = 0; k < BRT_GATHER_IMAGES; ++k) c->gather_seq[k] = 0;
    c->gather_next = 0;
    cudaIpcMemHandle_t h;
    BRT_CUDA(cudaIpcGetMemHandle(&h, c->d_gather.ptr()));
    std::memcpy(handle_out, &h, sizeof(h));
    c->gather_w = width;
    c->gather_h = height;
#endif
  });
}

int brt_gather_image_open(brt_context* c, const void* handles, uint32_t world) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!handles || world != c->tile_world || world > BRT_MAX_PEERS) invalid("gather_image_open: world must equal tile_world (<= 16)");
#ifdef BRT_EMU
    bad_state("gather_image_open: peer memory needs CUDA devices");
#else
    if (!c->gather_w) bad_state("gather_image_open: call brt_gather_image_export first");
    BRT_CUDA(cudaSetDevice(c->device));
    close_peers(c);
    for (uint32_t k = 0; k < world; ++k) {
      if (k == c->tile_rank) {
        c->peer_images[k] = c->d_gather.ptr();
      } else {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(handles) + (size_t)k * BRT_IPC_HANDLE_BYTES, sizeof(h));
        BRT_CUDA(cudaIpcOpenMemHandle(&c->peer_images[k], h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_opened[k] = true;
      }
    }
    c->n_peers = world;
#endif
  });
}

int brt_gather_configure(brt_context* c, uint32_t root_only, uint32_t root) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (root >= c->tile_world) invalid("gather_configure: root >= tile_world");
    for (uint32_t k = 0; k < BRT_GATHER_IMAGES; ++k)
      if (c->gather_seq[k]) bad_state("gather_configure: call it before the first frame of the exchange (after brt_gather_image_export)");
    c->gather_root_only = root_only ? 1u : 0u;
    c->gather_root = root;
  });
}

int brt_render_frame_peers(brt_context* c, const brt_uniform* u, const brt_render_opts* o) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!u || !o) invalid("render_frame_peers: null");
    BRT_CUDA(cudaSetDevice(c->device));
    FrameSlot* f = &c->slots[c->gather_next];
    c->gather_next = (c->gather_next + 1u) % BRT_GATHER_IMAGES;
    finish_frame(c, f);
    if (c->tables_dirty || c->tlas_dirty) wait_all_frames(c);
    ensure_slot(c, f);
    render_frame_device(c, f, *u, *o, nullptr, true);
    finish_frame(c, f);
  });
}

// Frames of the fused exchange in flight: slot k stores into gather image k. brt_frame_wait(slot) returns when this rank's stores of
// that frame have landed (local completion); whether EVERY rank's have is known on the device (brt_gather_wait).
int brt_render_frame_peers_async(brt_context* c, const brt_uniform* u, const brt_render_opts* o, uint32_t slot) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!u || !o) invalid("render_frame_peers_async: null");
    if (slot >= BRT_GATHER_IMAGES || slot >= BRT_FRAMES_IN_FLIGHT) invalid("render_frame_peers_async: slot >= BRT_GATHER_IMAGES");
    BRT_CUDA(cudaSetDevice(c->device));
    FrameSlot* f = &c->slots[slot];
    finish_frame(c, f);
    if (c->tables_dirty || c->tlas_dirty) wait_all_frames(c);
    ensure_slot(c, f);
    render_frame_device(c, f, *u, *o, nullptr, true);
  });
}

void* brt_gather_image(brt_context* c, uint32_t slot) {
  if (!c || !c->d_gather.ptr() || slot >= BRT_GATHER_IMAGES) return nullptr;
  return static_cast<char*>(c->d_gather.ptr()) + (size_t)slot * c->gather_w * c->gather_h * 16;
}

#ifndef BRT_EMU
namespace {
cudaStream_t gather_stream_of(brt_context* c, void* stream) {
  if (stream) return static_cast<cudaStream_t>(stream);
  if (!c->gather_stream) BRT_CUDA(cudaStreamCreateWithFlags(&c->gather_stream, cudaStreamNonBlocking));
  return c->gather_stream;
}
GatherFlags* my_gather_flags(brt_context* c) {
  return reinterpret_cast<GatherFlags*>(static_cast<char*>(c->d_gather.ptr()) + (size_t)BRT_GATHER_IMAGES * c->gather_w * c->gather_h * 16);
}
bool is_receiver(const brt_context* c) { return !c->gather_root_only || c->gather_root == c->tile_rank; }
void enqueue_gather_wait(brt_context* c, uint32_t slot, cudaStream_t s) {
  if (slot >= BRT_GATHER_IMAGES || !c->n_peers) invalid("gather_wait: bad slot / gather images not opened");
  if (!is_receiver(c)) bad_state("gather_wait: this rank does not receive the frame (brt_gather_configure)");
  k_gather_wait_arrive<<<1, 32, 0, s>>>(my_gather_flags(c), slot, c->gather_seq[slot], c->n_peers);
  BRT_CHECK_LAUNCH();
}
void enqueue_gather_release(brt_context* c, uint32_t slot, cudaStream_t s) {
  if (slot >= BRT_GATHER_IMAGES || !c->n_peers) invalid("gather_release: bad slot / gather images not opened");
  if (!is_receiver(c)) bad_state("gather_release: this rank does not receive the frame (brt_gather_configure)");
  GatherReleaseParams p{};
  p.n = c->n_peers;
  p.img = slot;
  p.seq = c->gather_seq[slot];
  p.me = c->tile_rank;
  const size_t off = (size_t)BRT_GATHER_IMAGES * c->gather_w * c->gather_h * 16;
  for (uint32_t k = 0; k < c->n_peers; ++k) p.peer_flags[k] = reinterpret_cast<GatherFlags*>(static_cast<char*>(c->peer_images[k]) + off);
  k_gather_release<<<1, 32, 0, s>>>(p);
  BRT_CHECK_LAUNCH();
}
}  // namespace
#endif

int brt_gather_wait(brt_context* c, uint32_t slot, void* stream) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
#ifdef BRT_EMU
    (void)slot; (void)stream;
    bad_state("gather_wait: peer memory needs CUDA devices");
#else
    BRT_CUDA(cudaSetDevice(c->device));
    enqueue_gather_wait(c, slot, gather_stream_of(c, stream));
#endif
  });
}

int brt_gather_release(brt_context* c, uint32_t slot, void* stream) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
#ifdef BRT_EMU
    (void)slot; (void)stream;
    bad_state("gather_release: peer memory needs CUDA devices");
#else
    BRT_CUDA(cudaSetDevice(c->device));
    enqueue_gather_release(c, slot, gather_stream_of(c, stream));
#endif
  });
}

int brt_gather_copy_to_host(brt_context* c, uint32_t slot, void* host, void* stream) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
#ifdef BRT_EMU
    (void)slot; (void)host; (void)stream;
    bad_state("gather_copy_to_host: peer memory needs CUDA devices");
#else
    if (!host) invalid("gather_copy_to_host: null");
    BRT_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = gather_stream_of(c, stream);
    enqueue_gather_wait(c, slot, s);
    const size_t npx = (size_t)c->gather_w * c->gather_h;
    const size_t bytes = c->slots[slot].gather_format == BRT_FORMAT_R32G32B32A32_SFLOAT ? npx * 16 : npx * 4;
    BRT_CUDA(cudaMemcpyAsync(host, brt_gather_image(c, slot), bytes, cudaMemcpyDeviceToHost, s));
    enqueue_gather_release(c, slot, s);
#endif
  });
}

int brt_gather_timed_out(brt_context* c) {
#ifdef BRT_EMU
  (void)c;
  return 0;
#else
  if (!c || !c->d_gather.ptr()) return 0;
  uint32_t v = 0;
  cudaSetDevice(c->device);
  if (cudaMemcpy(&v, &my_gather_flags(c)->timeout, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)v;
#endif
}

int brt_get_aov(brt_context* c, int kind, void* out) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!out) invalid("get_aov: null");
    wait_all_frames(c);
    FrameSlot* f = &c->slots[c->last_slot];
    if (!f->frame_w) bad_state("get_aov: no frame rendered");
    BRT_CUDA(cudaSetDevice(c->device));
    const size_t n = (size_t)f->frame_w * f->frame_h;
    const void* src = nullptr;
    size_t bytes = 4;
    switch (kind) {
      case BRT_AOV_PRIM_ID: src = f->d_aov_prim.ptr(); break;
      case BRT_AOV_INST_ID: src = f->d_aov_inst.ptr(); break;
      case BRT_AOV_HIT_T: src = f->d_aov_t.ptr(); break;
      case BRT_AOV_POSITION:
      case BRT_AOV_NORMAL:
        if (!f->has_gbuffer) bad_state("get_aov: the frame was not rendered with BRT_RENDER_GBUFFER");
        src = kind == BRT_AOV_POSITION ? f->d_aov_pos.ptr() : f->d_aov_nrm.ptr();
        bytes = 16;
        break;
      default: invalid("get_aov: bad kind");
    }
    BRT_CUDA(cudaMemcpyAsync(out, src, n * bytes, cudaMemcpyDeviceToHost, c->stream));
    BRT_CUDA(cudaStreamSynchronize(c->stream));
  });
}

int brt_get_stats(brt_context* c, brt_stats* out) {
  if (!c || !out) return BRT_ERR_INVALID;
  *out = c->stats;
  return BRT_OK;
}

int brt_trace_rays(brt_context* c, const float* rays, uint32_t n, int closest, uint32_t* out) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!rays || !out) invalid("trace_rays: null");
    if (!c->built) bad_state("trace_rays: scene not built");
    BRT_CUDA(cudaSetDevice(c->device));
    wait_all_frames(c);
    if (c->tlas_dirty) build_tlas(c);
    if (n == 0) return;
    cudaStream_t s = c->stream;
    FrameSlot* f = &c->slots[0];
    // split the interleaved rays into the two float4 arrays the kernels read
    std::vector<float> o4((size_t)n * 4), d4((size_t)n * 4);
    for (uint32_t i = 0; i < n; ++i) {
      std::memcpy(&o4[4 * (size_t)i], rays + 8 * (size_t)i, 16);
      std::memcpy(&d4[4 * (size_t)i], rays + 8 * (size_t)i + 4, 16);
    }
    c->d_rays.ensure((size_t)n * 32);
    c->d_ray_out.ensure((size_t)n * 16 + (size_t)n * 4 + (size_t)n * 4);
    f->d_counters.ensure(sizeof(FrameCounters) + 2 * sizeof(ShadowCounters));
    f->d_fstats.ensure(sizeof(FrameStats));
    float4* d_o = c->d_rays.as<float4>();
    float4* d_d = d_o + n;
    float4* d_hit = c->d_ray_out.as<float4>();
    uint32_t* d_inst = reinterpret_cast<uint32_t*>(d_hit + n);
    uint32_t* d_target = d_inst + n;
    BRT_CUDA(cudaMemcpyAsync(d_o, o4.data(), (size_t)n * 16, cudaMemcpyHostToDevice, s));
    BRT_CUDA(cudaMemcpyAsync(d_d, d4.data(), (size_t)n * 16, cudaMemcpyHostToDevice, s));
    BRT_CUDA(cudaMemsetAsync(f->d_counters.ptr(), 0, sizeof(FrameCounters) + 2 * sizeof(ShadowCounters), s));
    BRT_CUDA(cudaMemsetAsync(f->d_fstats.ptr(), 0, sizeof(FrameStats), s));
    FrameCounters* ctr = f->d_counters.as<FrameCounters>();
    TraceParams tp{};
    tp.count = n;
    tp.tlas = c->tlas_count ? c->d_tlas_nodes.as<Node8>() : nullptr;
    tp.insts = c->d_tlas_inst.as<InstRec>();
    tp.o = d_o;
    tp.d = d_d;
    tp.stats = f->d_fstats.as<FrameStats>();
    tp.refill_lanes = BRT_REFILL_LANES_INCOHERENT;  // arbitrary rays: the schedule of the bounce rounds
    tp.defer_lanes = c->defer_bounce;
    std::vector<float> hit((size_t)n * 4);
    std::vector<uint32_t> inst(n);
    if (closest) {
      tp.hit = d_hit;
      tp.hit_inst = d_inst;
      tp.work = &ctr->work_closest;
      launch_trace<false>(c, tp, s);
      BRT_CUDA(cudaMemcpyAsync(hit.data(), d_hit, (size_t)n * 16, cudaMemcpyDeviceToHost, s));
      BRT_CUDA(cudaMemcpyAsync(inst.data(), d_inst, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
      BRT_CUDA(cudaStreamSynchronize(s));
      for (uint32_t i = 0; i < n; ++i) {
        const bool h = inst[i] != BRT_MISS;
        uint32_t tb, pb;
        std::memcpy(&tb, &hit[4 * (size_t)i], 4);
        std::memcpy(&pb, &hit[4 * (size_t)i + 3], 4);
        out[4 * (size_t)i] = h ? tb : 0u;
        out[4 * (size_t)i + 1] = h ? pb : BRT_MISS;
        out[4 * (size_t)i + 2] = inst[i];
        out[4 * (size_t)i + 3] = h ? 1u : 0u;
      }
    } else {
      // occlusion: contrib[i] starts at 1 and is zeroed when ray i is blocked
      std::vector<float> ones((size_t)n * 4, 1.0f);
      std::vector<uint32_t> target(n);
      for (uint32_t i = 0; i < n; ++i) target[i] = i;
      BRT_CUDA(cudaMemcpyAsync(d_hit, ones.data(), (size_t)n * 16, cudaMemcpyHostToDevice, s));
      BRT_CUDA(cudaMemcpyAsync(d_target, target.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
      tp.target = d_target;
      tp.contrib = d_hit;
      tp.work = &reinterpret_cast<ShadowCounters*>(ctr + 1)->work_occl;
      launch_trace<true>(c, tp, s);
      BRT_CUDA(cudaMemcpyAsync(hit.data(), d_hit, (size_t)n * 16, cudaMemcpyDeviceToHost, s));
      BRT_CUDA(cudaStreamSynchronize(s));
      for (uint32_t i = 0; i < n; ++i) {
        out[4 * (size_t)i] = out[4 * (size_t)i + 1] = out[4 * (size_t)i + 2] = 0u;
        out[4 * (size_t)i + 3] = hit[4 * (size_t)i] == 0.0f ? 1u : 0u;
      }
    }
    FrameStats fs;
    BRT_CUDA(cudaMemcpy(&fs, f->d_fstats.ptr(), sizeof(fs), cudaMemcpyDeviceToHost));
    c->stats.rays_closest = fs.rays_closest;
    c->stats.rays_occlusion = fs.rays_occlusion;
    c->stats.nodes_visited_closest = fs.nodes_c;
    c->stats.prims_tested_closest = fs.prims_c;
    c->stats.spheres_tested_closest = fs.spheres_c;
    c->stats.nodes_visited_occlusion = fs.nodes_o;
    c->stats.prims_tested_occlusion = fs.prims_o;
    c->stats.spheres_tested_occlusion = fs.spheres_o;
  });
}

int brt_debug_sort_pairs(brt_context* c, uint32_t* keys, uint32_t* vals, uint32_t n, int bits) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if ((!keys || !vals) && n) invalid("debug_sort_pairs: null");
    if (bits < 1 || bits > 32) invalid("debug_sort_pairs: bits must be in 1..32");
    BRT_CUDA(cudaSetDevice(c->device));
    c->builder->debug_sort_pairs(c->stream, keys, vals, n, bits);
  });
}

int brt_debug_get_blas(brt_context* c, uint32_t mesh_id, void* nodes_out, uint32_t node_capacity, void* tris_out, uint32_t tri_capacity,
                       uint32_t* n_nodes, uint32_t* n_tris) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!n_nodes || !n_tris) invalid("debug_get_blas: null");
    if (mesh_id >= c->meshes.size()) invalid("debug_get_blas: no such mesh");
    const MeshData& m = *c->meshes[mesh_id];
    if (m.sphere) invalid("debug_get_blas: an analytic sphere has no BLAS");
    if (m.dirty) bad_state("debug_get_blas: mesh not built (call brt_scene_build)");
    wait_all_frames(c);
    BRT_CUDA(cudaSetDevice(c->device));
    const uint32_t nt = m.n_indices / 3;
    *n_nodes = nt ? m.blas.n_nodes : 0u;
    *n_tris = nt;
    const size_t kn = std::min<size_t>(node_capacity, *n_nodes), kt = std::min<size_t>(tri_capacity, nt);
    if (nodes_out && kn) BRT_CUDA(cudaMemcpyAsync(nodes_out, m.nodes.ptr(), kn * sizeof(Node8), cudaMemcpyDeviceToHost, c->stream));
    if (tris_out && kt) BRT_CUDA(cudaMemcpyAsync(tris_out, m.tris.ptr(), kt * sizeof(TriRec), cudaMemcpyDeviceToHost, c->stream));
    BRT_CUDA(cudaStreamSynchronize(c->stream));
  });
}

// 4x4 inverse in double (cofactor expansion along 2x2 minors), row-major in/out

// Measurement aid for the roofline context of bench.py: GB/s of a read-only pass with 16-byte loads over `bytes` of device memory
// (64 MiB fits the 126 MB L2: after the warm-up pass the loads are L2 hits), best of `iters`, CUDA events on the context's stream.
int brt_debug_l2_read_gbs(brt_context* c, size_t bytes, uint32_t iters, float* gbs_out) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
#ifdef BRT_EMU
    (void)bytes; (void)iters; (void)gbs_out;
    bad_state("debug_l2_read_gbs: needs a CUDA device");
#else
    if (!gbs_out || bytes < 4096 || !iters) invalid("debug_l2_read_gbs: bad arguments");
    BRT_CUDA(cudaSetDevice(c->device));
    wait_all_frames(c);
    DevBuf buf, out;
    buf.ensure(bytes);
    out.ensure(16);
    BRT_CUDA(cudaMemsetAsync(buf.ptr(), 1, bytes, c->stream));
    const uint32_t n16 = (uint32_t)(bytes / 16);
    const uint32_t grid = (uint32_t)c->sm_count * 8u;
    float best = 0.0f;
    for (uint32_t k = 0; k <= iters; ++k) {  // pass 0 warms the L2
      BRT_CUDA(cudaEventRecord(c->ev_t[0], c->stream));
      k_l2_read<<<grid, 256, 0, c->stream>>>(buf.as<uint4>(), n16, out.as<uint32_t>());
      BRT_CUDA(cudaEventRecord(c->ev_t[1], c->stream));
      BRT_CUDA(cudaStreamSynchronize(c->stream));
      float ms = 0.0f;
      cudaEventElapsedTime(&ms, c->ev_t[0], c->ev_t[1]);
      if (k && ms > 0.0f) best = std::max(best, (float)((double)n16 * 16.0 / (ms * 1e-3) / 1e9));
    }
    *gbs_out = best;
#endif
  });
}

void brt_camera_uniform(const float pos[3], const float rot[3], float fovy, float aspect, float znear, float zfar, uint32_t frame,
                        uint32_t depth_max, brt_uniform* out) {
  // Camera::updateView (Graphics/Camera.cpp:71-95): Tait-Bryan Y(yaw = rot.y) - X(rot.x) - Z(rot.z), rows u, v, w
  const float c3 = std::cos(rot[2]), s3 = std::sin(rot[2]);
  const float c2 = std::cos(rot[0]), s2 = std::sin(rot[0]);
  const float c1 = std::cos(rot[1]), s1 = std::sin(rot[1]);
  const float ux = c1 * c3 + s1 * s2 * s3, uy = c2 * s3, uz = c1 * s2 * s3 - c3 * s1;
  const float vx = c3 * s1 * s2 - c1 * s3, vy = c2 * c3, vz = c1 * c3 * s2 + s1 * s3;
  const float wx = c2 * s1, wy = -s2, wz = c1 * c2;
  const float tu = -(ux * pos[0] + uy * pos[1] + uz * pos[2]);
  const float tv = -(vx * pos[0] + vy * pos[1] + vz * pos[2]);
  const float tw = -(wx * pos[0] + wy * pos[1] + wz * pos[2]);
  const double view[16] = {ux, uy, uz, tu, vx, vy, vz, tv, wx, wy, wz, tw, 0, 0, 0, 1};
  // Camera::setPerspectiveProjection (Graphics/Camera.cpp:8-17)
  const float t = std::tan(fovy / 2.0f);
  double proj[16] = {0};
  proj[0] = 1.0f / (aspect * t);
  proj[5] = 1.0f / t;
  proj[10] = zfar / (zfar - znear);
  proj[14] = 1.0f;
  proj[11] = -(zfar * znear) / (zfar - znear);
  // RTApp::run (RT/RTApp.cpp:44-49): glm::inverse(glm::transpose(M)) in column-major memory is M^-1 in row-major memory
  double vi[16], pi[16];
  invert4x4(view, vi);
  invert4x4(proj, pi);
  for (int i = 0; i < 16; ++i) {
    out->viewInverse[i] = (float)(vi[i] + 0.0);  // + 0.0: no negative zeros
    out->projInverse[i] = (float)(pi[i] + 0.0);
  }
  out->frame = frame;
  out->depthMax = depth_max;
  out->lightThreshold = 0.0001f;
}

// Camera::handleInputs (Graphics/Camera.cpp:26-61) with the key state passed in instead of polled from GLFW.
void brt_camera_handle_inputs(uint32_t keys, float dt, float position[3], float rotation[3]) {
  auto key = [&](uint32_t k) { return (keys & k) != 0u; };
  const float eps = 1.1920929e-07f;  // std::numeric_limits<float>::epsilon()
  float rx = 0.0f, ry = 0.0f;
  if (key(BRT_KEY_LOOK_RIGHT)) ry += 1.0f;
  if (key(BRT_KEY_LOOK_LEFT)) ry -= 1.0f;
  if (key(BRT_KEY_LOOK_UP)) rx += 1.0f;
  if (key(BRT_KEY_LOOK_DOWN)) rx -= 1.0f;
  const float rr = rx * rx + ry * ry;
  if (rr > eps) {
    const float inv = 1.0f / std::sqrt(rr);
    rotation[0] += 1.5f * dt * (rx * inv);
    rotation[1] += 1.5f * dt * (ry * inv);
  }
  rotation[0] = std::fmin(std::fmax(rotation[0], -1.5f), 1.5f);
  const float two_pi = 6.28318530717958647692f;
  rotation[1] = rotation[1] - two_pi * std::floor(rotation[1] / two_pi);  // glm::mod
  const float yaw = rotation[1];
  const float fwd[3] = {std::sin(yaw), 0.0f, std::cos(yaw)};
  const float right[3] = {fwd[2], 0.0f, -fwd[0]};
  const float up[3] = {0.0f, -1.0f, 0.0f};
  float mv[3] = {0.0f, 0.0f, 0.0f};
  auto add = [&](const float* d, float sgn) { for (int k = 0; k < 3; ++k) mv[k] += sgn * d[k]; };
  if (key(BRT_KEY_MOVE_FORWARD)) add(fwd, 1.0f);
  if (key(BRT_KEY_MOVE_BACKWARD)) add(fwd, -1.0f);
  if (key(BRT_KEY_MOVE_RIGHT)) add(right, 1.0f);
  if (key(BRT_KEY_MOVE_LEFT)) add(right, -1.0f);
  if (key(BRT_KEY_MOVE_UP)) add(up, 1.0f);
  if (key(BRT_KEY_MOVE_DOWN)) add(up, -1.0f);
  const float mm = mv[0] * mv[0] + mv[1] * mv[1] + mv[2] * mv[2];
  if (mm > eps) {
    const float inv = 1.0f / std::sqrt(mm);
    for (int k = 0; k < 3; ++k) position[k] += 3.0f * dt * (mv[k] * inv);
  }
}

// Extensions::Denoiser::denoise (Graphics/Denoiser/Denoiser.h:5-20) on the frame rendered last; DESIGN.md §12.
int brt_denoise(brt_context* c, const brt_uniform* u, const brt_denoise_opts* d, float* rgba_host) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!u || !d) invalid("denoise: null");
    check_denoise_opts(*d);
    BRT_CUDA(cudaSetDevice(c->device));
    wait_all_frames(c);
    FrameSlot* f = &c->slots[c->last_slot];
    if (!f->frame_w || !f->has_gbuffer) bad_state("denoise: render the frame with BRT_RENDER_GBUFFER first");
    if (c->tile_world > 1) bad_state("denoise: single-GPU contexts only (tile_world == 1)");
    ensure_slot(c, f);
    cudaStream_t s = c->stream;
    enqueue_denoise(c, f, *u, *d, s);
    const size_t npx = (size_t)f->frame_w * f->frame_h;
    if (rgba_host) BRT_CUDA(cudaMemcpyAsync(rgba_host, f->d_denoised.ptr(), npx * 16, cudaMemcpyDeviceToHost, s));
    BRT_CUDA(cudaStreamSynchronize(s));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, f->ev_dn[0], f->ev_dn[1]);
    c->stats.ms_denoise = ms;
    c->stats.launches_denoise = f->dn_launches;
  });
}

/* options used by frames rendered with BRT_RENDER_DENOISE */
int brt_denoise_configure(brt_context* c, const brt_denoise_opts* d) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!d) invalid("denoise_configure: null");
    check_denoise_opts(*d);
    wait_all_frames(c);
    c->dn_opts = *d;
  });
}

int brt_get_light_bvh(brt_context* c, brt_light_bvh_node* out, uint32_t max_nodes, uint32_t* n_nodes) {
  if (!c) return BRT_ERR_INVALID;
  return guarded(c, [&] {
    if (!n_nodes) invalid("get_light_bvh: null");
    if (!c->built) bad_state("get_light_bvh: scene not built");
    if (c->tables_dirty) { wait_all_frames(c); build_tables(c); }
    *n_nodes = (uint32_t)c->light_bvh.size();
    if (out) std::memcpy(out, c->light_bvh.data(), std::min<size_t>(max_nodes, c->light_bvh.size()) * sizeof(brt_light_bvh_node));
  });
}

void* brt_denoised_image(brt_context* c) { return c && c->dn_w ? c->slots[c->last_slot].d_denoised.ptr() : nullptr; }

}  // extern "C"
