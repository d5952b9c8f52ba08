// Function qualifiers, launch helpers and the handful of intrinsics the kernels use.
//
// The product is built by nvcc for sm_100a (csrc/Makefile); that is the only thing libbrt.so ever
// contains. The same kernel bodies can also be compiled by g++ with -DBRT_EMU (tests/emu/Makefile)
// into a *test-only* library that runs every kernel body as a sequential host loop: it exists so the
// kernel logic (quantisation, traversal order, queue handling) can be debugged in the GPU-less build
// container, is never loaded by the package and is not a fallback — see DESIGN.md §9.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef BRT_EMU
// ---- host emulation: plain C++ ------------------------------------------------------------------
#include "cuda_emu.h"  // tests/emu/: vector types + a stand-in for the few runtime calls used
#define BRT_HD static inline
#define BRT_HDM inline  /* member functions */
#define BRT_HDH static inline
#define BRT_DEVICE_ONLY 0
#else
#include <cuda_runtime.h>
#define BRT_HD __device__ __forceinline__
#define BRT_HDM __device__ __forceinline__
#define BRT_HDH __host__ __device__ __forceinline__
#define BRT_DEVICE_ONLY 1
#endif

namespace brt {

// bit casts -------------------------------------------------------------------------------------
BRT_HDH uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
BRT_HDH float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
#ifndef BRT_EMU
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif
// explicit fused multiply-add (the build uses -fmad=false, so this is the only way to get an FFMA)
BRT_HD float fma_rn(float a, float b, float c) {
#ifdef BRT_EMU
  return fmaf(a, b, c);
#else
  return __fmaf_rn(a, b, c);
#endif
}
BRT_HD int popc(uint32_t x) {
#ifdef BRT_EMU
  return __builtin_popcount(x);
#else
  return __popc(x);
#endif
}
BRT_HD int clz32(uint32_t x) {
#ifdef BRT_EMU
  return x ? __builtin_clz(x) : 32;
#else
  return __clz((int)x);
#endif
}
BRT_HD int ffs32(uint32_t x) {  // 1-based index of the lowest set bit, 0 if none
#ifdef BRT_EMU
  return __builtin_ffs((int)x);
#else
  return __ffs((int)x);
#endif
}
BRT_HD float fast_rcp(float x) {
#ifdef BRT_EMU
  return 1.0f / x;
#else
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}
// directed roundings used by the builder's conservative quantisation
BRT_HD float add_rd(float a, float b) {
#ifdef BRT_EMU
  float r = a + b;
  return ((double)r > (double)a + (double)b) ? nextafterf(r, -INFINITY) : r;
#else
  return __fadd_rd(a, b);
#endif
}
BRT_HD float add_ru(float a, float b) {
#ifdef BRT_EMU
  float r = a + b;
  return ((double)r < (double)a + (double)b) ? nextafterf(r, INFINITY) : r;
#else
  return __fadd_ru(a, b);
#endif
}

// atomics (sequential stand-ins under emulation) ------------------------------------------------
BRT_HD uint32_t atomic_add(uint32_t* p, uint32_t v) {
#ifdef BRT_EMU
  uint32_t o = *p;
  *p = o + v;
  return o;
#else
  return atomicAdd(p, v);
#endif
}
BRT_HD unsigned long long atomic_add64(unsigned long long* p, unsigned long long v) {
#ifdef BRT_EMU
  unsigned long long o = *p;
  *p = o + v;
  return o;
#else
  return atomicAdd(p, v);
#endif
}
BRT_HD uint32_t atomic_min(uint32_t* p, uint32_t v) {
#ifdef BRT_EMU
  uint32_t o = *p;
  if (v < o) *p = v;
  return o;
#else
  return atomicMin(p, v);
#endif
}
BRT_HD uint32_t atomic_max(uint32_t* p, uint32_t v) {
#ifdef BRT_EMU
  uint32_t o = *p;
  if (v > o) *p = v;
  return o;
#else
  return atomicMax(p, v);
#endif
}
BRT_HD void fence() {
#ifndef BRT_EMU
  __threadfence();
#endif
}
// fetch-and-add with acquire-release semantics at device scope: the arrival counters of the bottom-up builder passes. One instruction
// orders this thread's earlier stores before the add and its later loads after it — what `fence(); atomic_add(); fence();` spelled with
// two full membars (each ~1 us on the dependent chain of a refit walk).
BRT_HD uint32_t atomic_add_acq_rel(uint32_t* p, uint32_t v) {
#ifdef BRT_EMU
  uint32_t o = *p;
  *p = o + v;
  return o;
#else
  uint32_t o;
  asm volatile("atom.acq_rel.gpu.add.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
  return o;
#endif
}

// float <-> order-preserving uint (for atomic min/max on floats)
BRT_HDH uint32_t float_to_ordered(float f) {
  uint32_t u = f2u(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
BRT_HDH float ordered_to_float(uint32_t u) { return u2f((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// read-only 16-byte loads ----------------------------------------------------------------------
BRT_HD float4 ldg4(const float4* p) {
#ifdef BRT_EMU
  return *p;
#else
  return __ldg(p);
#endif
}
BRT_HD uint4 ldg4(const uint4* p) {
#ifdef BRT_EMU
  return *p;
#else
  return __ldg(p);
#endif
}

}  // namespace brt

// ---- kernel definition / launch helpers ---------------------------------------------------------
// A "1-D kernel" is a Params struct with members `count` (and optionally a device-side `count_ptr`
// that overrides it) plus a body `void body(const Params&, uint32_t i)`. BRT_KERNEL_1D generates a
// grid-stride __global__ function (or, under emulation, a host loop); BRT_LAUNCH_1D launches it on
// `stream` with a grid that is a multiple of the SM count (persistent-style for device-side counts).
#ifdef BRT_EMU
#define BRT_KERNEL_1D(name, Params, body)                               \
  static void name(const Params p) {                                    \
    const uint32_t n = p.count_ptr ? *p.count_ptr : p.count;            \
    for (uint32_t i = 0; i < n; ++i) body(p, i);                        \
  }
#define BRT_LAUNCH_1D(name, params, grid, block, stream) name(params)
#else
#define BRT_KERNEL_1D(name, Params, body)                                                               \
  __global__ void __launch_bounds__(256) name(const Params p) {                                         \
    const uint32_t n = p.count_ptr ? *p.count_ptr : p.count;                                            \
    const uint32_t stride = gridDim.x * blockDim.x;                                                     \
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) body(p, i);            \
  }
#define BRT_LAUNCH_1D(name, params, grid, block, stream) name<<<(grid), (block), 0, (stream)>>>(params)
#endif
// same, with an explicit (threads per block, resident blocks per SM) occupancy target for register-heavy kernels
#ifdef BRT_EMU
#define BRT_KERNEL_1D_LB(name, Params, body, threads, min_blocks) BRT_KERNEL_1D(name, Params, body)
#else
#define BRT_KERNEL_1D_LB(name, Params, body, threads, min_blocks)                                       \
  __global__ void __launch_bounds__(threads, min_blocks) name(const Params p) {                         \
    const uint32_t n = p.count_ptr ? *p.count_ptr : p.count;                                            \
    const uint32_t stride = gridDim.x * blockDim.x;                                                     \
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) body(p, i);            \
  }
#endif
