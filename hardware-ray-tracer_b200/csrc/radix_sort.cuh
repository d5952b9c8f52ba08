// LSD radix sort of (uint32 key, uint32 value) pairs for the Morton codes of the LBVH builder.
#pragma once
#include "host_util.h"

#ifdef BRT_EMU
#include <algorithm>
#include <numeric>
#include <vector>
#else
#include <cub/device/device_radix_sort.cuh>
#endif

namespace brt {

inline size_t radix_sort_temp_bytes(uint32_t n) {
#ifdef BRT_EMU
  (void)n;
  return 16;
#else
  size_t bytes = 0;
  cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (int)n, 0, 30);
  return bytes + 16;
#endif
}

// Sorts n pairs by the low `bits` bits of the key. Returns which of the two buffers (0/1) holds the result.
inline int radix_sort_pairs(cudaStream_t stream, uint32_t* keys0, uint32_t* keys1, uint32_t* vals0, uint32_t* vals1, uint32_t n, int bits,
                            void* temp, size_t temp_bytes, int sm_count) {
  (void)sm_count;
#ifdef BRT_EMU
  (void)stream; (void)temp; (void)temp_bytes; (void)bits;
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys0[a] < keys0[b]; });
  for (uint32_t i = 0; i < n; ++i) {
    keys1[i] = keys0[order[i]];
    vals1[i] = vals0[order[i]];
  }
  return 1;
#else
  cub::DoubleBuffer<uint32_t> k(keys0, keys1), v(vals0, vals1);
  BRT_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k, v, (int)n, 0, bits, stream));
  return k.Current() == keys0 ? 0 : 1;
#endif
}

}  // namespace brt
