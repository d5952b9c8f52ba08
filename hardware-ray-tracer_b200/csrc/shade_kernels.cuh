// The shade kernels (rchitMain + calculateColor + rmissMain over a wavefront), shared by two translation units: brt_api.cu compiles them
// under the arithmetic contract (-fmad=false, IEEE division and square root: bit-identical to the oracle), shade_fast.cu compiles the same
// source with FMA contraction and approximate division / reciprocal square root for the opt-in BRT_RENDER_FAST_SHADING. The including file
// may set BRT_SHADE_PRIMARY / BRT_SHADE_BOUNCE to rename the kernels.
#pragma once
#include "render_kernels.cuh"

#ifndef BRT_SHADE_PRIMARY
#define BRT_SHADE_PRIMARY k_shade_primary
#define BRT_SHADE_BOUNCE k_shade
#endif

namespace brt {

#ifndef BRT_SHADE_MIN_BLOCKS
#define BRT_SHADE_MIN_BLOCKS 5  // 96 registers, no spills: measured 1-1.5 % faster than 4 (118) on C3 / C5; 6 (80, spills) is slower
#endif
#ifndef BRT_SHADE_PRIMARY_MIN_BLOCKS
#define BRT_SHADE_PRIMARY_MIN_BLOCKS BRT_SHADE_MIN_BLOCKS
#endif
#ifndef BRT_SHADE_HEADS
#define BRT_SHADE_HEADS 1  // path heads of a window requested together (2: spills at 96 registers)
#endif
#ifndef BRT_SHADE_WINDOW
#define BRT_SHADE_WINDOW 4  // x 128 path slots are classified before their hits are shaded together
#endif
#ifdef BRT_EMU
BRT_KERNEL_1D_LB(k_shade, ShadeParams, shade_body, 128, BRT_SHADE_MIN_BLOCKS)
#else
// primary round: the wavefront is in pixel order, hits and misses come in large coherent runs — one path per thread, no compaction
__global__ void __launch_bounds__(128, BRT_SHADE_PRIMARY_MIN_BLOCKS) BRT_SHADE_PRIMARY(const ShadeParams p) {
  const uint32_t n = p.count_ptr ? *p.count_ptr : p.count;
  const uint32_t stride = gridDim.x * 128u;
  for (uint32_t i = blockIdx.x * 128u + threadIdx.x; i < n; i += stride) {
    if (p.ahead && i + p.ahead < n) shade_prefetch(p, i + p.ahead);
    shade_body(p, i);
  }
}
// Shade with block-level hit compaction. After the first bounce the hits and misses of a wavefront are interleaved at random, and the
// hit shader (geometry fetch, BRDF per light, shadow-ray emission, bounce sampling: ~95 % of the kernel's instructions) ran with ~6 of
// 32 lanes active (ncu, profiles/). Each block therefore first runs the cheap prologue for its 128 path slots (bookkeeping, AOVs, the
// whole miss shader) — K = BRT_SHADE_WINDOW times, so that the window is 512 slots and the dependent-load latency of the hit shader is
// paid once per window —, compacts the indices of the hits into shared memory (ballot + per-warp prefix), and then shades the compacted list with
// full warps. Which thread shades which path is irrelevant: every output is addressed by the path slot.
template <uint32_t K>
__global__ void __launch_bounds__(128, BRT_SHADE_MIN_BLOCKS) BRT_SHADE_BOUNCE(const ShadeParams p) {
  __shared__ uint32_t s_idx[128u * K];
  __shared__ uint32_t s_warp[4 * K];
  const uint32_t n = p.count_ptr ? *p.count_ptr : p.count;
  // a short queue (fewer than K chunks per block) is latency-bound on the number of blocks in flight: window of one chunk then
  const uint32_t kk = n >= K * 128u * gridDim.x ? K : 1u;
  const uint32_t slots = 128u * kk;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t base = blockIdx.x * slots; base < n; base += gridDim.x * slots) {  // block-uniform trip count
    bool is_hit[K];
    uint32_t slot[K];
    unsigned m[K];
    constexpr uint32_t G = BRT_SHADE_HEADS;  // path heads requested together (all K at once spills at 96 registers)
#pragma unroll
    for (uint32_t k = 0; k < K; ++k) {
      bool live[G];
      PathHead head[G];
      if (k % G == 0u) {
#pragma unroll
        for (uint32_t g = 0; g < G; ++g) {
          const uint32_t w = base + (k + g) * 128u + threadIdx.x;
          live[g] = k + g < kk && w < n;
          if (p.ahead && !p.order && live[g] && w + p.ahead < n) shade_prefetch(p, w + p.ahead);
          slot[k + g] = live[g] && p.order ? p.order[w] : w;  // hit-sorted order of a bounce round (render_kernels.cuh), else queue order
          if (live[g]) head[g] = shade_load(p, slot[k + g]);
        }
#pragma unroll
        for (uint32_t g = 0; g < G; ++g) is_hit[k + g] = live[g] && shade_prologue(p, slot[k + g], head[g]);
      }
      m[k] = __ballot_sync(0xffffffffu, is_hit[k]);
      if (lane == 0) s_warp[k * 4 + warp] = (uint32_t)__popc(m[k]);
    }
    __syncthreads();
    uint32_t total = 0, off[K];
#pragma unroll
    for (uint32_t j = 0; j < 4 * K; ++j) {
      const uint32_t cnt = s_warp[j];
#pragma unroll
      for (uint32_t k = 0; k < K; ++k)
        if (j == k * 4 + warp) off[k] = total;
      total += cnt;
    }
#pragma unroll
    for (uint32_t k = 0; k < K; ++k)
      if (is_hit[k]) s_idx[off[k] + (uint32_t)__popc(m[k] & ((1u << lane) - 1u))] = slot[k];
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < total; t += 128u) {
      const uint32_t i = s_idx[t];
      shade_hit(p, i, p.cur.px[i], p.hit_inst[i], p.hit[i]);
    }
    __syncthreads();  // s_idx / s_warp are rewritten by the next window
  }
}
#endif

}  // namespace brt
