// Host orchestration of the GPU BVH builder (kernels in build_kernels.cuh, radix_sort.cuh, treelet.cuh).
#include "builder.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "build_kernels.cuh"
#include "radix_sort.cuh"
#include "treelet.cuh"

// One primitive per leaf slot: a node addresses its primitives as prim_base + rank of the slot among its leaf slots.
// (Leaves of up to 3 triangles were measured 3-4 % slower on B200: a primitive test runs with ~5 active lanes, a node
// test with ~27, so trading node tests for primitive tests loses.)
#define BRT_TRI_LEAF 1

namespace brt {

BRT_KERNEL_1D(k_tri_bounds, TriBoundsParams, tri_bounds_body)
BRT_KERNEL_1D(k_inst_bounds, InstBoundsParams, inst_bounds_body)
BRT_KERNEL_1D(k_morton, MortonParams, morton_body)
BRT_KERNEL_1D(k_hierarchy, HierarchyParams, hierarchy_body)
BRT_KERNEL_1D(k_refit, RefitParams, refit_body)
BRT_KERNEL_1D(k_wide_cost, WideCostParams, wide_cost_body)
#if defined(BRT_EMU) && defined(BRT_EMU_WARP)
// Host emulation, test only: every item is collapsed by the scalar collapse_body and then again — with the builder's counters put
// back — by the device's warp-cooperative collapse_warp on 32 lock-stepped host threads (tests/emu/warp_emu.h); the node, the queue
// entries and the primitive records the two wrote are compared byte for byte. brt_emu_warp_collapse_stats reports the totals.
static unsigned long long g_warp_items = 0, g_warp_mismatches = 0;
extern "C" __attribute__((visibility("default"))) void brt_emu_warp_collapse_stats(unsigned long long* items, unsigned long long* mismatches) {
  *items = g_warp_items;
  *mismatches = g_warp_mismatches;
}
// Same idea for the treelet passes: the device's k_treelet_warp runs on a copy of the binary tree (warps one after the other, 32
// lock-stepped host threads each); its result must be a valid tree (parents, boxes, counts and costs consistent, every leaf once) whose
// SAH cost is close to the scalar passes' (ties between equally good treelet topologies are broken differently, so the trees — and
// with them the later treelets — can differ: the largest relative deviation is reported).
static unsigned long long g_treelet_checks = 0, g_treelet_bad = 0;
static float g_treelet_max_dev = 0.0f;
extern "C" __attribute__((visibility("default"))) void brt_emu_warp_treelet_stats(unsigned long long* checks, unsigned long long* invalid,
                                                                                  float* max_cost_deviation) {
  *checks = g_treelet_checks;
  *invalid = g_treelet_bad;
  *max_cost_deviation = g_treelet_max_dev;
}
static bool emu_valid_binary_tree(const std::vector<BNode>& nodes, const std::vector<uint32_t>& parent, const std::vector<uint32_t>& sub_count,
                                  const std::vector<float>& cost, uint32_t n) {
  const uint32_t n_int = n - 1;
  std::vector<uint8_t> seen(2 * n - 1, 0);
  std::vector<uint32_t> stack{0u};
  uint32_t leaves = 0;
  if (parent[0] != BRT_MISS) return false;
  while (!stack.empty()) {
    const uint32_t id = stack.back();
    stack.pop_back();
    if (id >= 2 * n - 1 || seen[id]) return false;
    seen[id] = 1;
    if (id >= n_int) {
      leaves++;
      continue;
    }
    const uint32_t c[2] = {f2u(nodes[id].lo.w), f2u(nodes[id].hi.w)};
    if (c[0] >= 2 * n - 1 || c[1] >= 2 * n - 1 || parent[c[0]] != id || parent[c[1]] != id) return false;
    const BNode &a = nodes[c[0]], &b = nodes[c[1]], &m = nodes[id];
    if (m.lo.x != fminf(a.lo.x, b.lo.x) || m.lo.y != fminf(a.lo.y, b.lo.y) || m.lo.z != fminf(a.lo.z, b.lo.z)) return false;
    if (m.hi.x != fmaxf(a.hi.x, b.hi.x) || m.hi.y != fmaxf(a.hi.y, b.hi.y) || m.hi.z != fmaxf(a.hi.z, b.hi.z)) return false;
    if (sub_count[id] != sub_count[c[0]] + sub_count[c[1]]) return false;
    const float lo[3] = {m.lo.x, m.lo.y, m.lo.z}, hi[3] = {m.hi.x, m.hi.y, m.hi.z};
    const float want = BRT_SAH_CI * area_of(lo, hi) + cost[c[0]] + cost[c[1]];
    if (fabsf(cost[id] - want) > 1e-4f * fabsf(want)) return false;
    stack.push_back(c[0]);
    stack.push_back(c[1]);
  }
  return leaves == n;
}
static void k_collapse(const CollapseParams p) {
  const uint32_t n = p.count_ptr ? *p.count_ptr : p.count;
  const bool check = getenv("BRT_EMU_WARP_CHECK") != nullptr;
  for (uint32_t i = 0; i < n; ++i) {
    if (!check) {
      collapse_body(p, i);
      continue;
    }
    const BuildGlobals before = *p.g;
    collapse_body(p, i);
    const BuildGlobals after = *p.g;
    const uint2 work = p.queue_in[i];
    const uint32_t q0 = before.level_count[p.level + 1], q1 = after.level_count[p.level + 1];
    const uint32_t r0 = before.prim_count, r1 = after.prim_count;
    const size_t rec = p.out_tris ? sizeof(TriRec) : sizeof(InstRec);
    char* recs = p.out_tris ? reinterpret_cast<char*>(p.out_tris) : reinterpret_cast<char*>(p.out_inst);
    const Node8 node_a = p.out_nodes[work.y];
    std::vector<uint2> queue_a(p.queue_out + q0, p.queue_out + q1);
    std::vector<char> recs_a(recs + r0 * rec, recs + r1 * rec);
    memset(&p.out_nodes[work.y], 0xcd, sizeof(Node8));
    if (q1 > q0) memset(p.queue_out + q0, 0xcd, (q1 - q0) * sizeof(uint2));
    if (r1 > r0) memset(recs + r0 * rec, 0xcd, (r1 - r0) * rec);
    *p.g = before;
    brt_warp_emu::run_warp([&](unsigned lane) { collapse_warp(p, i, lane); });
    const BuildGlobals redo = *p.g;
    bool same = memcmp(&redo, &after, sizeof(BuildGlobals)) == 0 && memcmp(&p.out_nodes[work.y], &node_a, sizeof(Node8)) == 0;
    same = same && (q1 == q0 || memcmp(p.queue_out + q0, queue_a.data(), (q1 - q0) * sizeof(uint2)) == 0);
    same = same && (r1 == r0 || memcmp(recs + r0 * rec, recs_a.data(), (r1 - r0) * rec) == 0);
    g_warp_items++;
    if (!same) g_warp_mismatches++;
  }
}
#elif defined(BRT_EMU)
BRT_KERNEL_1D(k_collapse, CollapseParams, collapse_body)
#else
// One level of the collapse. Levels of up to BRT_COLLAPSE_WARP_MAX items (the top of every tree, all of a small one) are
// latency-bound: one warp per item (collapse_warp); wide levels are throughput-bound: one thread per item (collapse_body).
#define BRT_COLLAPSE_WARP_MAX 8192u
__global__ void __launch_bounds__(64) k_collapse(const CollapseParams p) {
  const uint32_t n = p.count_ptr ? *p.count_ptr : p.count;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  if (n <= BRT_COLLAPSE_WARP_MAX) {
    for (uint32_t i = tid >> 5; i < n; i += nt >> 5) collapse_warp(p, i, threadIdx.x & 31u);
  } else {
    for (uint32_t i = tid; i < n; i += nt) collapse_body(p, i);
  }
}
#endif
#ifdef BRT_EMU
BRT_KERNEL_1D(k_treelet, TreeletParams, treelet_body)  // the device runs the warp-cooperative k_treelet_warp (treelet.cuh)
#endif

struct InitGlobalsParams {
  uint32_t count;
  const uint32_t* count_ptr;
  BuildGlobals* g;
  uint32_t root;  // binary root id pushed as the first collapse work item
  uint2* queue0;
};
BRT_HD void init_globals_body(const InitGlobalsParams& p, uint32_t) {
  BuildGlobals& g = *p.g;
  for (int k = 0; k < 3; ++k) {
    g.bounds[k] = 0xffffffffu; g.bounds[3 + k] = 0u;
    g.exact[k] = 0xffffffffu; g.exact[3 + k] = 0u;
  }
  g.node_count = 1u;  // the root
  g.prim_count = 0u;
  g.levels = 0u;
  g.overflow = 0u;
  for (int k = 0; k < 64; ++k) g.level_count[k] = 0u;
  g.level_count[0] = 1u;
  p.queue0[0] = make_uint2(p.root, 0u);
}
BRT_KERNEL_1D(k_init_globals, InitGlobalsParams, init_globals_body)

struct MeshBoundsParams {
  uint32_t count;
  const uint32_t* count_ptr;
  const BuildGlobals* g;
  float4* out;  // lo, hi
};
BRT_HD void mesh_bounds_body(const MeshBoundsParams& p, uint32_t) {
  p.out[0] = make_float4(ordered_to_float(p.g->exact[0]), ordered_to_float(p.g->exact[1]), ordered_to_float(p.g->exact[2]), 0.0f);
  p.out[1] = make_float4(ordered_to_float(p.g->exact[3]), ordered_to_float(p.g->exact[4]), ordered_to_float(p.g->exact[5]), 0.0f);
}
BRT_KERNEL_1D(k_mesh_bounds, MeshBoundsParams, mesh_bounds_body)

#ifndef BRT_EMU
#define BRT_SMALL_BUILD_MAX 4096u
#define BRT_SMALL_BUILD_THREADS 1024
struct SmallBuildParams {
  uint32_t n;
  MortonParams morton;     // writes keys / vals (unsorted)
  uint32_t* keys_sorted;
  uint32_t* vals_sorted;
  HierarchyParams hier;
  RefitParams refit;  // also fills the cost table of the collapse
  CollapseParams collapse; // level / queues / count_ptr are set per level by the kernel
  uint2* queue[2];
  float4* mesh_bounds;     // may be null
  uint32_t in_smem;        // the binary tree and every intermediate array live in dynamic shared memory (n <= BRT_SMEM_BUILD_MAX)
  uint32_t fused_prologue; // instance builds: the kernel also resets the globals and computes the instance boxes (two launches less)
  InitGlobalsParams init;
  InstBoundsParams inst;
};
// Up to BRT_SMEM_BUILD_MAX primitives (the per-frame TLAS of C4: 513 instances) the binary nodes, parent links, arrival counters, subtree
// counts, the collapse's cost table and plan, the sorted keys and both collapse queues fit the SM's shared memory (140 bytes per
// primitive next to the 32 KB sort buffer): the dependent chains of the bottom-up refit and of the collapse then wait for shared memory
// (~30 ns) instead of the L2 (~700 ns) at every hop. The kernel bodies take plain pointers, so they do not know the difference.
#define BRT_SMEM_BUILD_MAX 1280u
inline size_t small_build_smem_bytes(uint32_t n) { return (size_t)n * 144 + 512; }
__global__ void __launch_bounds__(BRT_SMALL_BUILD_THREADS) k_build_small(const SmallBuildParams p_in) {
  __shared__ unsigned long long sk[BRT_SMALL_BUILD_MAX];  // (Morton key << 32 | primitive): unique, so the sort is stable by construction
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  SmallBuildParams p = p_in;
#ifdef BRT_BUILD_SMALL_TIMING
  unsigned long long ts[10];
  int nts = 0;
#define BRT_STAMP() do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); ts[nts++] = t_; } } while (0)
#else
#define BRT_STAMP() do {} while (0)
#endif
  BRT_STAMP();
  if (p.fused_prologue) {
    if (threadIdx.x == 0) init_globals_body(p.init, 0);
    __threadfence();
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < p.n; i += blockDim.x) inst_bounds_body(p.inst, i);
    __threadfence();
    __syncthreads();
  }
  if (p.in_smem) {
    const uint32_t n = p.n;
    unsigned char* at = dyn_smem;
    auto take = [&](size_t bytes) { unsigned char* r = at; at += (bytes + 15) & ~(size_t)15; return r; };
    BNode* nodes = reinterpret_cast<BNode*>(take((size_t)2 * n * sizeof(BNode)));
    float* wcost = reinterpret_cast<float*>(take((size_t)n * BRT_WCOST_STRIDE * 4));
    unsigned long long* wplan = reinterpret_cast<unsigned long long*>(take((size_t)n * 8));
    uint32_t* parent = reinterpret_cast<uint32_t*>(take((size_t)2 * n * 4));
    uint32_t* sub_count = reinterpret_cast<uint32_t*>(take((size_t)2 * n * 4));
    uint32_t* arrive = reinterpret_cast<uint32_t*>(take((size_t)n * 4));
    uint32_t* keys_sorted = reinterpret_cast<uint32_t*>(take((size_t)n * 4));
    uint32_t* vals_sorted = reinterpret_cast<uint32_t*>(take((size_t)n * 4));
    uint2* q0 = reinterpret_cast<uint2*>(take((size_t)(n / 2 + 8) * 8));
    uint2* q1 = reinterpret_cast<uint2*>(take((size_t)(n / 2 + 8) * 8));
    if (threadIdx.x == 0) q0[0] = p.queue[0][0];  // the root work item written by init_globals_body
    p.keys_sorted = keys_sorted;
    p.vals_sorted = vals_sorted;
    p.hier.keys = keys_sorted;
    p.hier.nodes = nodes;
    p.hier.parent = parent;
    p.refit.vals = vals_sorted;
    p.refit.nodes = nodes;
    p.refit.parent = parent;
    p.refit.arrive = arrive;
    p.refit.sub_count = sub_count;
    if (p.refit.wcost) p.refit.wcost = wcost;
    p.refit.wplan = wplan;
    p.collapse.nodes = nodes;
    p.collapse.sub_count = sub_count;
    if (p.collapse.wcost) p.collapse.wcost = wcost;
    p.collapse.wplan = wplan;
    p.queue[0] = q0;
    p.queue[1] = q1;
  }
  const uint32_t tid = threadIdx.x, nt = blockDim.x, n = p.n;
  BRT_STAMP();  // 1: prologue + smem carve-up
  uint32_t np2 = 2;
  while (np2 < n) np2 <<= 1;
  for (uint32_t i = tid; i < np2; i += nt) {
    if (i < n) {
      morton_body(p.morton, i);
      sk[i] = ((unsigned long long)p.morton.keys[i] << 32) | i;
    } else {
      sk[i] = ~0ull;
    }
  }
  for (uint32_t i = tid; i < 2 * n; i += nt) p.hier.parent[i] = 0xffffffffu;
  for (uint32_t i = tid; i < n; i += nt) p.refit.arrive[i] = 0u;
  __syncthreads();
  for (uint32_t k = 2; k <= np2; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = tid; i < np2; i += nt) {
        const uint32_t l = i ^ j;
        if (l > i) {
          const unsigned long long a = sk[i], b = sk[l];
          if ((a > b) == ((i & k) == 0u)) { sk[i] = b; sk[l] = a; }
        }
      }
      __syncthreads();
    }
  BRT_STAMP();  // 2: morton + sort
  for (uint32_t i = tid; i < n; i += nt) {
    p.keys_sorted[i] = (uint32_t)(sk[i] >> 32);
    p.vals_sorted[i] = (uint32_t)sk[i];
  }
  __threadfence();
  __syncthreads();
  for (uint32_t i = tid; i + 1 < n; i += nt) hierarchy_body(p.hier, i);
  __threadfence();
  __syncthreads();
  BRT_STAMP();  // 3: hierarchy
  for (uint32_t i = tid; i < n; i += nt) refit_body(p.refit, i);
  __threadfence();
  __syncthreads();
  BRT_STAMP();  // 4: refit + cost table
  CollapseParams cp = p.collapse;
  for (uint32_t level = 0; level < 62; ++level) {
    const uint32_t count = *reinterpret_cast<volatile uint32_t*>(&cp.g->level_count[level]);
    if (count == 0u) break;  // block-uniform
    cp.level = level;
    cp.queue_in = p.queue[level & 1];
    cp.queue_out = p.queue[(level + 1) & 1];
    for (uint32_t i = tid >> 5; i < count; i += nt >> 5) collapse_warp(cp, i, tid & 31u);
    __threadfence();
    __syncthreads();
  }
  BRT_STAMP();  // 5: collapse
  if (p.mesh_bounds && tid == 0) {
    MeshBoundsParams mb{1, nullptr, cp.g, p.mesh_bounds};
    mesh_bounds_body(mb, 0);
  }
#ifdef BRT_BUILD_SMALL_TIMING
  if (threadIdx.x == 0)
    printf("k_build_small n=%u smem=%u: prologue %llu sort %llu hierarchy %llu refit %llu collapse %llu ns (levels %u)\n", p.n, p.in_smem, ts[1] - ts[0], ts[2] - ts[1],
           ts[3] - ts[2], ts[4] - ts[3], ts[5] - ts[4], cp.g->levels);
#endif
}
#endif

void Builder::ensure_scratch(uint32_t n) {
  const size_t N = n;
  globals_.ensure(sizeof(BuildGlobals));
  prim_lo_.ensure(N * 16);
  prim_hi_.ensure(N * 16);
  for (int k = 0; k < 2; ++k) {
    keys_[k].ensure(N * 4);
    vals_[k].ensure(N * 4);
    queue_[k].ensure((N / 2 + 8) * 8);
  }
  sort_tmp_.ensure(radix_sort_temp_bytes(n));
  nodes_.ensure((2 * N) * sizeof(BNode));
  parent_.ensure(2 * N * 4);
  arrive_.ensure(N * 4);
  sub_count_.ensure(2 * N * 4);
  treelet_.ensure(2 * N * 4 + 16);  // SAH cost per binary node
  wcost_.ensure(N * BRT_WCOST_STRIDE * 4);  // collapse cost table per internal binary node
  wplan_.ensure(N * 8);                      // ... and the choices behind it
}

void Builder::run(cudaStream_t stream, uint32_t n, uint32_t max_leaf, bool treelets, Node8* out_nodes, const float* d_vertices,
                  const uint32_t* d_indices, TriRec* out_tris, const InstRec* d_src, InstRec* out_inst, float4* d_mesh_bounds,
                  BuildResult* res) {
  // prim_lo_/prim_hi_ and globals_->bounds have been filled by the caller
  BuildGlobals* g = globals_.as<BuildGlobals>();
  const uint32_t grid_n = std::max(1u, std::min(div_up(n, 256u), (uint32_t)sm_count_ * 8u));
  float cost_before = 0.0f, cost_after = 0.0f;
  BNode root_before{}, root_after{};
  const bool want_sah = out_tris != nullptr && n > 1;
  BuildGlobals hg;
  uint32_t level = 0;
  bool small = false;
#ifndef BRT_EMU
  // Small builds (the per-frame TLAS over a few hundred instances): the ~30 dependent launches of the general path cost far more
  // in launch latency than in work, so ONE block runs the whole pipeline — Morton codes, a bitonic sort of (key, index) pairs in
  // shared memory, Karras hierarchy, bottom-up refit, level-by-level collapse — with __syncthreads() between the phases.
  if (out_tris == nullptr && n >= 2 && n <= BRT_SMALL_BUILD_MAX) {
    SmallBuildParams sp{};
    sp.n = n;
    sp.morton = MortonParams{n, nullptr, prim_lo_.as<float4>(), prim_hi_.as<float4>(), g, keys_[0].as<uint32_t>(), vals_[0].as<uint32_t>()};
    sp.keys_sorted = keys_[1].as<uint32_t>();
    sp.vals_sorted = vals_[1].as<uint32_t>();
    sp.hier = HierarchyParams{n - 1, nullptr, n, sp.keys_sorted, nodes_.as<BNode>(), parent_.as<uint32_t>()};
    sp.refit = RefitParams{n, nullptr, sp.vals_sorted, prim_lo_.as<float4>(), prim_hi_.as<float4>(), nodes_.as<BNode>(), parent_.as<uint32_t>(),
                           arrive_.as<uint32_t>(), sub_count_.as<uint32_t>(), nullptr, greedy_collapse_ ? nullptr : wcost_.as<float>(),
                           wplan_.as<unsigned long long>()};
    CollapseParams& cp = sp.collapse;
    cp.n = n;
    cp.max_leaf = max_leaf;
    cp.nodes = nodes_.as<BNode>();
    cp.sub_count = sub_count_.as<uint32_t>();
    cp.wcost = greedy_collapse_ ? nullptr : wcost_.as<float>();
    cp.wplan = wplan_.as<unsigned long long>();
    cp.queue_cap = n / 2 + 8;
    cp.g = g;
    cp.out_nodes = out_nodes;
    cp.node_cap = node_capacity(n);
    cp.vertices = d_vertices;
    cp.indices = d_indices;
    cp.out_tris = out_tris;
    cp.src_inst = d_src;
    cp.out_inst = out_inst;
    sp.queue[0] = queue_[0].as<uint2>();
    sp.queue[1] = queue_[1].as<uint2>();
    sp.mesh_bounds = d_mesh_bounds;
    sp.in_smem = n <= BRT_SMEM_BUILD_MAX && !getenv("BRT_NO_SMEM_BUILD") ? 1u : 0u;
    sp.fused_prologue = fused_small_ ? 1u : 0u;
    if (fused_small_) {
      std::memcpy(&sp.init, fused_params_, sizeof(sp.init));
      std::memcpy(&sp.inst, fused_params_ + sizeof(sp.init), sizeof(sp.inst));
    }
    const size_t dyn = sp.in_smem ? small_build_smem_bytes(n) : 0;
    if (dyn > small_build_smem_set_) {
      BRT_CUDA(cudaFuncSetAttribute(k_build_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_build_smem_bytes(BRT_SMEM_BUILD_MAX)));
      small_build_smem_set_ = small_build_smem_bytes(BRT_SMEM_BUILD_MAX);
    }
    k_build_small<<<1, BRT_SMALL_BUILD_THREADS, dyn, stream>>>(sp);
    BRT_CHECK_LAUNCH();
    BRT_CUDA(cudaMemcpyAsync(&hg, g, sizeof(hg), cudaMemcpyDeviceToHost, stream));
    BRT_CUDA(cudaStreamSynchronize(stream));
    for (level = 0; level < 62 && hg.level_count[level] != 0; ++level) {}
    small = true;
  }
#endif
  if (!small) {
    {
      MortonParams p{n, nullptr, prim_lo_.as<float4>(), prim_hi_.as<float4>(), g, keys_[0].as<uint32_t>(), vals_[0].as<uint32_t>()};
      BRT_LAUNCH_1D(k_morton, p, grid_n, 256, stream);
      BRT_CHECK_LAUNCH();
    }
    uint32_t* keys = keys_[0].as<uint32_t>();
    uint32_t* vals = vals_[0].as<uint32_t>();
    if (n > 1) {
      int out = radix_sort_pairs(stream, keys_[0].as<uint32_t>(), keys_[1].as<uint32_t>(), vals_[0].as<uint32_t>(), vals_[1].as<uint32_t>(), n,
                                 BRT_MORTON_BITS, sort_tmp_.ptr(), sort_tmp_.capacity(), sm_count_);
      keys = keys_[out].as<uint32_t>();
      vals = vals_[out].as<uint32_t>();
    }
    BNode* nodes = nodes_.as<BNode>();
    uint32_t* parent = parent_.as<uint32_t>();
    uint32_t* sub_count = sub_count_.as<uint32_t>();
    BRT_CUDA(cudaMemsetAsync(parent, 0xff, (size_t)(2 * n) * 4, stream));
    if (n > 1) {
      BRT_CUDA(cudaMemsetAsync(arrive_.ptr(), 0, (size_t)n * 4, stream));
      HierarchyParams p{n - 1, nullptr, n, keys, nodes, parent};
      BRT_LAUNCH_1D(k_hierarchy, p, grid_n, 256, stream);
      BRT_CHECK_LAUNCH();
    }
    // The refit also leaves the SAH cost of the LBVH (statistic) and, when no treelet pass is going to change the topology, the
    // cost table of the collapse: a fast build walks the tree bottom-up once.
    const bool do_treelets = treelets && want_sah && n >= 2 * BRT_TREELET_LEAVES;
    const bool want_wcost = n > 1 && !greedy_collapse_;
    float* cost = treelet_.as<float>();
    {
      RefitParams p{n, nullptr, vals, prim_lo_.as<float4>(), prim_hi_.as<float4>(), nodes, parent, arrive_.as<uint32_t>(), sub_count,
                    want_sah ? cost : nullptr, want_wcost && !do_treelets ? wcost_.as<float>() : nullptr, wplan_.as<unsigned long long>()};
      BRT_LAUNCH_1D(k_refit, p, grid_n, 256, stream);
      BRT_CHECK_LAUNCH();
    }
    if (want_sah) {
      BRT_CUDA(cudaMemcpyAsync(&cost_before, cost, 4, cudaMemcpyDeviceToHost, stream));
      BRT_CUDA(cudaMemcpyAsync(&root_before, nodes, sizeof(BNode), cudaMemcpyDeviceToHost, stream));
    }
    // SAH treelet restructuring (first builds of a triangle BLAS), then the cost table of the collapse on the final topology
    if (do_treelets) {
#if defined(BRT_EMU) && defined(BRT_EMU_WARP)
      std::vector<BNode> w_nodes;
      std::vector<uint32_t> w_parent, w_sub, w_arrive;
      std::vector<float> w_cost;
      const bool warp_check = getenv("BRT_EMU_WARP_CHECK") != nullptr;
      if (warp_check) {
        w_nodes.assign(nodes, nodes + (2 * n - 1));
        w_parent.assign(parent, parent + 2 * n);
        w_sub.assign(sub_count, sub_count + 2 * n);
        w_cost.assign(cost, cost + 2 * n);
      }
#endif
      for (int pass = 1; pass <= treelet_passes_; ++pass) {
        BRT_CUDA(cudaMemsetAsync(arrive_.ptr(), 0, (size_t)n * 4, stream));
        TreeletParams tp{n, nullptr, nodes, parent, sub_count, arrive_.as<uint32_t>(), cost, 1u};
  #ifdef BRT_EMU
        BRT_LAUNCH_1D(k_treelet, tp, 1, 64, stream);
  #else
        k_treelet_warp<<<std::max(1u, std::min(div_up(n, 128u), (uint32_t)sm_count_ * 8u)), 128, 0, stream>>>(tp);
  #endif
        BRT_CHECK_LAUNCH();
      }
#if defined(BRT_EMU) && defined(BRT_EMU_WARP)
      if (warp_check) {
        for (int pass = 1; pass <= treelet_passes_; ++pass) {
          w_arrive.assign(n, 0u);
          const TreeletParams tp{n, nullptr, w_nodes.data(), w_parent.data(), w_sub.data(), w_arrive.data(), w_cost.data(), 1u};
          brt_warp_emu::run_kernel_warps(std::max(1u, div_up(n, 128u)), 128u, [&] { k_treelet_warp(tp); });
        }
        const float cs = cost[0], cw = w_cost[0];
        g_treelet_checks++;
        g_treelet_max_dev = std::max(g_treelet_max_dev, fabsf(cw - cs) / std::max(fabsf(cs), 1e-30f));
        if (!emu_valid_binary_tree(w_nodes, w_parent, w_sub, w_cost, n) || !(cw == cw)) g_treelet_bad++;
      }
#endif
      if (want_wcost) {
        BRT_CUDA(cudaMemsetAsync(arrive_.ptr(), 0, (size_t)n * 4, stream));
        WideCostParams p{n, nullptr, nodes, parent, arrive_.as<uint32_t>(), wcost_.as<float>(), wplan_.as<unsigned long long>()};
        BRT_LAUNCH_1D(k_wide_cost, p, grid_n, 256, stream);
        BRT_CHECK_LAUNCH();
      }
    }
    if (want_sah) {
      BRT_CUDA(cudaMemcpyAsync(&cost_after, cost, 4, cudaMemcpyDeviceToHost, stream));
      BRT_CUDA(cudaMemcpyAsync(&root_after, nodes, sizeof(BNode), cudaMemcpyDeviceToHost, stream));
    }
    // collapse, level by level; the per-level work count lives on the device
    const uint32_t node_cap = node_capacity(n);
    const uint32_t queue_cap = n / 2 + 8;
    CollapseParams cp{};
    cp.n = n;
    cp.max_leaf = max_leaf;
    cp.nodes = nodes;
    cp.sub_count = sub_count;
    cp.wcost = n > 1 && !greedy_collapse_ ? wcost_.as<float>() : nullptr;
    cp.wplan = wplan_.as<unsigned long long>();
    cp.queue_cap = queue_cap;
    cp.g = g;
    cp.out_nodes = out_nodes;
    cp.node_cap = node_cap;
    cp.vertices = d_vertices;
    cp.indices = d_indices;
    cp.out_tris = out_tris;
    cp.src_inst = d_src;
    cp.out_inst = out_inst;
    for (;;) {
      const uint32_t chunk_end = level + 8;
      for (; level < chunk_end && level < 62; ++level) {
        cp.level = level;
        cp.count = 0;
        cp.count_ptr = &g->level_count[level];
        cp.queue_in = queue_[level & 1].as<uint2>();
        cp.queue_out = queue_[(level + 1) & 1].as<uint2>();
        // level L has at most min(8^L, queue_cap) items
        uint64_t max_items = 1;
        for (uint32_t k = 0; k < level && max_items < queue_cap; ++k) max_items *= 8;
        max_items = std::min<uint64_t>(max_items, queue_cap);
        uint32_t grid = std::max(1u, std::min(div_up((uint32_t)max_items, 64u), (uint32_t)sm_count_ * 16u));
#ifndef BRT_EMU
        grid = std::max(grid, div_up((uint32_t)std::min<uint64_t>(max_items, BRT_COLLAPSE_WARP_MAX), 2u));  // two warps per block
#endif
        BRT_LAUNCH_1D(k_collapse, cp, grid, 64, stream);
        BRT_CHECK_LAUNCH();
      }
      BRT_CUDA(cudaMemcpyAsync(&hg, g, sizeof(hg), cudaMemcpyDeviceToHost, stream));
      BRT_CUDA(cudaStreamSynchronize(stream));
      if (level >= 62 || hg.level_count[level] == 0) break;
    }
  }
  if (d_mesh_bounds && !small) {
    MeshBoundsParams p{1, nullptr, g, d_mesh_bounds};
    BRT_LAUNCH_1D(k_mesh_bounds, p, 1, 32, stream);
    BRT_CHECK_LAUNCH();
  }
  if (hg.overflow) throw LimitError("BVH build: node or queue capacity exceeded");
  if (hg.level_count[level] != 0) throw LimitError("BVH build: hierarchy deeper than 62 levels");
  res->n_prims = n;
  res->n_nodes = hg.node_count;
  res->levels = hg.levels;
  auto host_area = [](const BNode& b) {
    const float ex = b.hi.x - b.lo.x, ey = b.hi.y - b.lo.y, ez = b.hi.z - b.lo.z;
    return 2.0f * ((ex * ey + ey * ez) + ez * ex);
  };
  res->sah_lbvh = want_sah && host_area(root_before) > 0.0f ? cost_before / host_area(root_before) : 0.0f;
  res->sah_final = want_sah && host_area(root_after) > 0.0f ? cost_after / host_area(root_after) : 0.0f;
  for (int k = 0; k < 3; ++k) {
    res->lo[k] = ordered_to_float(hg.exact[k]);
    res->hi[k] = ordered_to_float(hg.exact[3 + k]);
  }
}

void Builder::debug_sort_pairs(cudaStream_t stream, uint32_t* keys_host, uint32_t* vals_host, uint32_t n, int bits) {
  if (n == 0) return;
  ensure_scratch(n);
  BRT_CUDA(cudaMemcpyAsync(keys_[0].ptr(), keys_host, (size_t)n * 4, cudaMemcpyHostToDevice, stream));
  BRT_CUDA(cudaMemcpyAsync(vals_[0].ptr(), vals_host, (size_t)n * 4, cudaMemcpyHostToDevice, stream));
  const int out = radix_sort_pairs(stream, keys_[0].as<uint32_t>(), keys_[1].as<uint32_t>(), vals_[0].as<uint32_t>(), vals_[1].as<uint32_t>(), n, bits,
                                   sort_tmp_.ptr(), sort_tmp_.capacity(), sm_count_);
  BRT_CUDA(cudaMemcpyAsync(keys_host, keys_[out].ptr(), (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
  BRT_CUDA(cudaMemcpyAsync(vals_host, vals_[out].ptr(), (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
  BRT_CUDA(cudaStreamSynchronize(stream));
}

void Builder::build_triangles(cudaStream_t stream, const float* d_vertices, const uint32_t* d_indices, uint32_t n_tris, Node8* out_nodes,
                              TriRec* out_tris, float4* d_mesh_bounds, bool treelets, BuildResult* res) {
  ensure_scratch(n_tris);
  BuildGlobals* g = globals_.as<BuildGlobals>();
  const uint32_t root = n_tris > 1 ? 0u : 0u;  // internal node 0, or leaf id 0 when n == 1 (n_internal == 0)
  InitGlobalsParams ip{1, nullptr, g, root, queue_[0].as<uint2>()};
  BRT_LAUNCH_1D(k_init_globals, ip, 1, 32, stream);
  BRT_CHECK_LAUNCH();
  const uint32_t grid_n = std::max(1u, std::min(div_up(n_tris, 256u), (uint32_t)sm_count_ * 8u));
  TriBoundsParams p{n_tris, nullptr, d_vertices, d_indices, prim_lo_.as<float4>(), prim_hi_.as<float4>(), g};
  BRT_LAUNCH_1D(k_tri_bounds, p, grid_n, 256, stream);
  BRT_CHECK_LAUNCH();
  run(stream, n_tris, BRT_TRI_LEAF, treelets, out_nodes, d_vertices, d_indices, out_tris, nullptr, nullptr, d_mesh_bounds, res);
}

void Builder::build_instances(cudaStream_t stream, const InstShade* d_shade, const uint32_t* d_inst_ids, const InstRec* d_src, uint32_t n,
                              const float4* d_mesh_bounds, Node8* out_nodes, InstRec* out_inst, BuildResult* res) {
  ensure_scratch(n);
  BuildGlobals* g = globals_.as<BuildGlobals>();
  InitGlobalsParams ip{1, nullptr, g, 0u, queue_[0].as<uint2>()};
  InstBoundsParams p{n, nullptr, d_shade, d_inst_ids, d_mesh_bounds, prim_lo_.as<float4>(), prim_hi_.as<float4>(), g};
#ifndef BRT_EMU
  if (n >= 2 && n <= BRT_SMALL_BUILD_MAX) {  // the one-block kernel resets the globals and computes the instance boxes itself
    static_assert(sizeof(InitGlobalsParams) + sizeof(InstBoundsParams) <= sizeof(fused_params_), "fused_params_ too small");
    fused_small_ = true;
    std::memcpy(fused_params_, &ip, sizeof(ip));
    std::memcpy(fused_params_ + sizeof(ip), &p, sizeof(p));
    try {
      run(stream, n, 1, false, out_nodes, nullptr, nullptr, nullptr, d_src, out_inst, nullptr, res);
    } catch (...) {
      fused_small_ = false;
      throw;
    }
    fused_small_ = false;
    return;
  }
#endif
  BRT_LAUNCH_1D(k_init_globals, ip, 1, 32, stream);
  BRT_CHECK_LAUNCH();
  const uint32_t grid_n = std::max(1u, std::min(div_up(n, 256u), (uint32_t)sm_count_ * 8u));
  BRT_LAUNCH_1D(k_inst_bounds, p, grid_n, 256, stream);
  BRT_CHECK_LAUNCH();
  run(stream, n, 1, false, out_nodes, nullptr, nullptr, nullptr, d_src, out_inst, nullptr, res);
}

}  // namespace brt
