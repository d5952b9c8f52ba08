// Stable LSD radix sort of (uint32 key, uint32 value) pairs for the Morton codes of the LBVH builder:
// 8-bit digits, three launches per pass (per-tile digit histogram -> exclusive scan of the digit-major table ->
// stable scatter). Ranks inside a tile come from warp-level __match_any_sync groups plus per-warp digit counters in
// shared memory, with items laid out (warp, round, lane) so that equal digits keep their input order.
#pragma once
#include "host_util.h"

#ifdef BRT_EMU
#include <algorithm>
#include <numeric>
#include <vector>
#endif

namespace brt {

#define BRT_SORT_THREADS 256
// keys per thread: 16 for large inputs; 4 below BRT_SORT_SMALL_N keys, where the tiles of 4096 keys would occupy a handful of
// SMs and every pass would be bound by the 16 dependent ranking rounds of its few blocks (the re-built dynamic BLAS of C4:
// 20,480 keys were 5 blocks, 23 us per scatter)
#define BRT_SORT_ITEMS 16
#define BRT_SORT_ITEMS_SMALL 4
#define BRT_SORT_SMALL_N 262144u
inline uint32_t radix_sort_items(uint32_t n) { return n <= BRT_SORT_SMALL_N ? BRT_SORT_ITEMS_SMALL : BRT_SORT_ITEMS; }

inline uint32_t radix_sort_tiles(uint32_t n) {
  const uint32_t tile = BRT_SORT_THREADS * radix_sort_items(n);
  return n ? (n + tile - 1) / tile : 1;
}
// scratch: the digit-major histogram table, 256 x tiles counters
inline size_t radix_sort_temp_bytes(uint32_t n) { return (size_t)radix_sort_tiles(n) * 256 * 4 + 16; }

#ifndef BRT_EMU
// digit histogram of tile `tile` (block-cooperative, BRT_SORT_THREADS threads; ends with a barrier so that it can be called in a loop)
template <int ITEMS>
__device__ __forceinline__ void radix_hist_tile(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t dmask,
                                                uint32_t* __restrict__ hist, uint32_t tiles, uint32_t tile) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = tile * (BRT_SORT_THREADS * ITEMS);
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const uint32_t i = base + k * BRT_SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & dmask], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * tiles + tile] = h[threadIdx.x];
  __syncthreads();
}
template <int ITEMS>
__global__ void __launch_bounds__(BRT_SORT_THREADS) k_radix_hist(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t dmask,
                                                                 uint32_t* __restrict__ hist, uint32_t tiles) {
  radix_hist_tile<ITEMS>(keys, n, shift, dmask, hist, tiles, blockIdx.x);
}

// exclusive scan of `count` counters in place, one block of 1024 threads (the table is at most a few 100k entries)
__global__ void __launch_bounds__(1024) k_radix_scan(uint32_t* __restrict__ hist, uint32_t count) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  const uint32_t per = (count + 1023u) / 1024u;
  const uint32_t begin = threadIdx.x * per, end = min(begin + per, count);
  uint32_t sum = 0;
  for (uint32_t i = begin; i < end; ++i) sum += hist[i];
  // block-wide exclusive scan of the per-thread sums
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t incl = sum;
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
    if ((int)lane >= off) incl += v;
  }
  if (lane == 31u) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = warp_sums[lane];
    uint32_t wi = w;
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, wi, off);
      if ((int)lane >= off) wi += v;
    }
    warp_sums[lane] = wi - w;
    if (lane == 31u) carry = wi;
  }
  __syncthreads();
  uint32_t run = warp_sums[warp] + incl - sum;
  for (uint32_t i = begin; i < end; ++i) {
    const uint32_t v = hist[i];
    hist[i] = run;
    run += v;
  }
  (void)carry;
}

template <int ITEMS>
__device__ __forceinline__ void radix_scatter_tile(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                   uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                   uint32_t dmask, const uint32_t* __restrict__ hist, uint32_t tiles, uint32_t tile) {
  constexpr int WARPS = BRT_SORT_THREADS / 32;
  __shared__ uint32_t warp_count[WARPS][256];
  for (int k = threadIdx.x; k < WARPS * 256; k += BRT_SORT_THREADS) (&warp_count[0][0])[k] = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t warp_base = tile * (BRT_SORT_THREADS * ITEMS) + warp * (ITEMS * 32);
  uint32_t key[ITEMS], val[ITEMS], rank[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const uint32_t i = warp_base + k * 32 + lane;
    const bool valid = i < n;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    rank[k] = 0;
    if (valid) {
      key[k] = keys_in[i];
      val[k] = vals_in[i];
      const uint32_t d = (key[k] >> shift) & dmask;
      const unsigned peers = __match_any_sync(vmask, d);
      const int leader = __ffs(peers) - 1;
      uint32_t before = 0;
      if ((int)lane == leader) {
        before = warp_count[warp][d];
        warp_count[warp][d] = before + __popc(peers);
      }
      before = __shfl_sync(peers, before, leader);
      rank[k] = before + __popc(peers & lt);
    }
    __syncwarp();
  }
  __syncthreads();
  {  // thread d turns the per-warp counts of digit d into start offsets (global base of this tile + earlier warps)
    const uint32_t d = threadIdx.x;
    uint32_t run = hist[d * tiles + tile];
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      const uint32_t c = warp_count[w][d];
      warp_count[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const uint32_t i = warp_base + k * 32 + lane;
    if (i < n) {
      const uint32_t pos = warp_count[warp][(key[k] >> shift) & dmask] + rank[k];
      keys_out[pos] = key[k];
      vals_out[pos] = val[k];
    }
  }
  __syncthreads();  // (callable in a loop: warp_count is rewritten by the next tile)
}
template <int ITEMS>
__global__ void __launch_bounds__(BRT_SORT_THREADS) k_radix_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n,
                                                                    int shift, uint32_t dmask, const uint32_t* __restrict__ hist, uint32_t tiles) {
  radix_scatter_tile<ITEMS>(keys_in, vals_in, keys_out, vals_out, n, shift, dmask, hist, tiles, blockIdx.x);
}
// in-place exclusive scan of `count` counters by ONE block of any size (<= 1024 threads)
__device__ __forceinline__ void radix_scan_block(uint32_t* __restrict__ hist, uint32_t count) {
  __shared__ uint32_t warp_sums[32];
  const uint32_t nt = blockDim.x;
  const uint32_t per = (count + nt - 1u) / nt;
  const uint32_t begin = min(threadIdx.x * per, count), end = min(begin + per, count);
  uint32_t sum = 0;
  for (uint32_t i = begin; i < end; ++i) sum += hist[i];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t incl = sum;
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
    if ((int)lane >= off) incl += v;
  }
  if (threadIdx.x < 32u) warp_sums[threadIdx.x] = 0u;
  __syncthreads();
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
    const uint32_t v = hist[i];
    hist[i] = run;
    run += v;
  }
  __syncthreads();
}
#endif  // !BRT_EMU

// Sorts n pairs by the low `bits` bits of the key. Returns which of the two buffers (0/1) holds the result.
inline int radix_sort_pairs(cudaStream_t stream, uint32_t* keys0, uint32_t* keys1, uint32_t* vals0, uint32_t* vals1, uint32_t n, int bits,
                            void* temp, size_t temp_bytes, int sm_count) {
  (void)sm_count;
#ifdef BRT_EMU
  (void)stream; (void)temp; (void)temp_bytes;
  const uint32_t mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return (keys0[a] & mask) < (keys0[b] & mask); });
  for (uint32_t i = 0; i < n; ++i) {
    keys1[i] = keys0[order[i]];
    vals1[i] = vals0[order[i]];
  }
  return 1;
#else
  const uint32_t tiles = radix_sort_tiles(n);
  if (temp_bytes < radix_sort_temp_bytes(n)) throw LimitError("radix sort: scratch too small");
  uint32_t* hist = static_cast<uint32_t*>(temp);
  uint32_t* k[2] = {keys0, keys1};
  uint32_t* v[2] = {vals0, vals1};
  int cur = 0;
  for (int shift = 0; shift < bits; shift += 8) {
    const uint32_t dmask = bits - shift >= 8 ? 0xffu : ((1u << (bits - shift)) - 1u);  // the last digit may be narrower
    if (radix_sort_items(n) == BRT_SORT_ITEMS_SMALL) {
      k_radix_hist<BRT_SORT_ITEMS_SMALL><<<tiles, BRT_SORT_THREADS, 0, stream>>>(k[cur], n, shift, dmask, hist, tiles);
      k_radix_scan<<<1, 1024, 0, stream>>>(hist, tiles * 256u);
      k_radix_scatter<BRT_SORT_ITEMS_SMALL><<<tiles, BRT_SORT_THREADS, 0, stream>>>(k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], n, shift, dmask, hist, tiles);
    } else {
      k_radix_hist<BRT_SORT_ITEMS><<<tiles, BRT_SORT_THREADS, 0, stream>>>(k[cur], n, shift, dmask, hist, tiles);
      k_radix_scan<<<1, 1024, 0, stream>>>(hist, tiles * 256u);
      k_radix_scatter<BRT_SORT_ITEMS><<<tiles, BRT_SORT_THREADS, 0, stream>>>(k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], n, shift, dmask, hist, tiles);
    }
    BRT_CHECK_LAUNCH();
    cur ^= 1;
  }
  return cur;
#endif
}

}  // namespace brt
