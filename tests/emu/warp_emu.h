// TEST INFRASTRUCTURE ONLY — a lock-step stand-in for the warp intrinsics used by the builder's warp-cooperative collapse
// (collapse_warp, build_kernels.cuh), so that this device-only code can be checked on the CPU against the scalar collapse_body
// item by item (builder.cu, k_collapse under BRT_EMU), and the warp-cooperative treelet kernel (k_treelet_warp, treelet.cuh) can run on a copy of the
// binary tree next to the scalar passes (tests/test_emu_parity.py::test_warp_kernels_equal_scalar).
// A "warp" is 32 host threads; every intrinsic is an exchange through a 32-slot array between two barriers. Only full-mask,
// fully converged use is supported — which is all these two kernels do.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#define __device__
#define __forceinline__ inline
#define __global__ static
#define __shared__ static  /* one warp runs at a time */
#define __launch_bounds__(...)

namespace brt_warp_emu {
struct Warp {
  pthread_barrier_t bar;
  uint32_t slot[32];
};
inline thread_local Warp* t_warp = nullptr;
inline thread_local unsigned t_lane = 0;
struct Dim { unsigned x; };

inline void exchange(uint32_t v, uint32_t* all) {
  Warp* w = t_warp;
  w->slot[t_lane] = v;
  pthread_barrier_wait(&w->bar);
  memcpy(all, w->slot, sizeof(w->slot));
  pthread_barrier_wait(&w->bar);
}
template <class T>
inline uint32_t bits(T v) {
  static_assert(sizeof(T) == 4, "32-bit values only");
  uint32_t u;
  memcpy(&u, &v, 4);
  return u;
}
template <class T>
inline T from_bits(uint32_t u) {
  T v;
  memcpy(&v, &u, 4);
  return v;
}
}  // namespace brt_warp_emu
// launch geometry of the kernel being emulated: one warp of one block at a time (run_kernel_warps)
inline thread_local brt_warp_emu::Dim threadIdx{0}, blockIdx{0};
inline brt_warp_emu::Dim blockDim{32}, gridDim{1};
namespace brt_warp_emu {
// runs fn(lane) on 32 lock-stepped host threads
template <class F>
inline void run_warp(F fn) {
  Warp w;
  pthread_barrier_init(&w.bar, nullptr, 32);
  std::vector<std::thread> th;
  th.reserve(32);
  for (unsigned l = 0; l < 32; ++l)
    th.emplace_back([&w, l, &fn] {
      t_warp = &w;
      t_lane = l;
      fn(l);
    });
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&w.bar);
}
// a kernel of `blocks` blocks of `threads` threads that only synchronises inside warps: its warps run one after the other
template <class F>
inline void run_kernel_warps(unsigned blocks, unsigned threads, F kernel) {
  blockDim.x = threads;
  gridDim.x = blocks;
  for (unsigned b = 0; b < blocks; ++b)
    for (unsigned w = 0; w < threads / 32; ++w)
      run_warp([&](unsigned lane) {
        blockIdx.x = b;
        threadIdx.x = w * 32 + lane;
        kernel();
      });
}
}  // namespace brt_warp_emu

template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
  uint32_t all[32];
  brt_warp_emu::exchange(brt_warp_emu::bits(v), all);
  return brt_warp_emu::from_bits<T>(all[src & 31]);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  uint32_t all[32];
  brt_warp_emu::exchange(brt_warp_emu::bits(v), all);
  return brt_warp_emu::from_bits<T>(all[(brt_warp_emu::t_lane ^ (unsigned)lane_mask) & 31]);
}
inline void __syncwarp() { pthread_barrier_wait(&brt_warp_emu::t_warp->bar); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline bool __any_sync(unsigned, bool pred) {
  uint32_t all[32];
  brt_warp_emu::exchange(pred ? 1u : 0u, all);
  for (int l = 0; l < 32; ++l)
    if (all[l]) return true;
  return false;
}
inline uint32_t __ballot_sync(unsigned, bool pred) {
  uint32_t all[32], m = 0;
  brt_warp_emu::exchange(pred ? 1u : 0u, all);
  for (int l = 0; l < 32; ++l) m |= (all[l] & 1u) << l;
  return m;
}
inline uint32_t __reduce_min_sync(unsigned, uint32_t v) {
  uint32_t all[32];
  brt_warp_emu::exchange(v, all);
  uint32_t r = all[0];
  for (int l = 1; l < 32; ++l) r = all[l] < r ? all[l] : r;
  return r;
}
inline uint32_t __reduce_max_sync(unsigned, uint32_t v) {
  uint32_t all[32];
  brt_warp_emu::exchange(v, all);
  uint32_t r = all[0];
  for (int l = 1; l < 32; ++l) r = all[l] > r ? all[l] : r;
  return r;
}
inline uint32_t __reduce_add_sync(unsigned, uint32_t v) {
  uint32_t all[32], r = 0;
  brt_warp_emu::exchange(v, all);
  for (int l = 0; l < 32; ++l) r += all[l];
  return r;
}
inline uint32_t __reduce_or_sync(unsigned, uint32_t v) {
  uint32_t all[32], r = 0;
  brt_warp_emu::exchange(v, all);
  for (int l = 0; l < 32; ++l) r |= all[l];
  return r;
}
inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
inline int __popc(uint32_t x) { return __builtin_popcount(x); }
