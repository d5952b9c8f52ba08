// SAH treelet restructuring of the binary LBVH (after Karras & Aila, "Fast Parallel Construction of
// High-Quality Bounding Volume Hierarchies", HPG 2013) and the SAH cost statistic.
//
// One bottom-up pass: every leaf thread walks towards the root; the second thread to arrive at an internal
// node owns it (atomic arrival counter), so all work below a node is finished before the node is touched.
// At every node with at least 7 primitives below it a treelet of 7 leaves is formed by repeatedly opening the
// treelet leaf with the largest surface area, the optimal binary topology over those 7 leaves is found by
// dynamic programming over all 127 subsets, and — if it is cheaper — the 6 internal nodes of the treelet are
// rewritten in place. The pass also leaves the SAH cost of every subtree in `cost[]`, so cost[root] / area(root)
// is the statistic reported by brt_get_stats (sah_cost_lbvh before, sah_cost after restructuring).
#pragma once
#include "build_kernels.cuh"

namespace brt {

#define BRT_TREELET_LEAVES 7
// (BRT_SAH_CI / BRT_SAH_CT: build_kernels.cuh)

struct TreeletParams {
  uint32_t count;  // n leaves
  const uint32_t* count_ptr;
  BNode* nodes;
  uint32_t* parent;
  uint32_t* sub_count;
  uint32_t* arrive;  // n-1, zeroed before the pass
  float* cost;       // 2n-1
  uint32_t restructure;  // 0: only compute cost[]
};

BRT_HD float area_of(const float* lo, const float* hi) {
  const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
  return 2.0f * ((ex * ey + ey * ez) + ez * ex);
}

// Optimises the treelet rooted at internal node `root`. Returns the SAH cost of the (possibly rewritten) subtree.
BRT_HD float optimize_treelet(const TreeletParams& p, uint32_t root, uint32_t n_int) {
  volatile BNode* nodes = p.nodes;
  volatile float* cost = p.cost;
  volatile uint32_t* sub_count = p.sub_count;
  const int NL = BRT_TREELET_LEAVES;
  uint32_t leaf[NL];      // treelet leaves (node ids)
  uint32_t internal[NL];  // treelet internal nodes, internal[0] == root
  float larea[NL];
  int nl = 2, ni = 1;
  internal[0] = root;
  leaf[0] = f2u(nodes[root].lo.w);
  leaf[1] = f2u(nodes[root].hi.w);
  for (int k = 0; k < 2; ++k) {
    const float lo[3] = {nodes[leaf[k]].lo.x, nodes[leaf[k]].lo.y, nodes[leaf[k]].lo.z};
    const float hi[3] = {nodes[leaf[k]].hi.x, nodes[leaf[k]].hi.y, nodes[leaf[k]].hi.z};
    larea[k] = area_of(lo, hi);
  }
  while (nl < NL) {
    int best = -1;
    float ba = -1.0f;
    for (int k = 0; k < nl; ++k)
      if (leaf[k] < n_int && larea[k] > ba) { best = k; ba = larea[k]; }
    if (best < 0) break;
    const uint32_t id = leaf[best];
    internal[ni++] = id;
    const uint32_t c0 = f2u(nodes[id].lo.w), c1 = f2u(nodes[id].hi.w);
    leaf[best] = c0;
    leaf[nl] = c1;
    const int slots[2] = {best, nl};
    for (int k = 0; k < 2; ++k) {
      const uint32_t c = leaf[slots[k]];
      const float lo[3] = {nodes[c].lo.x, nodes[c].lo.y, nodes[c].lo.z};
      const float hi[3] = {nodes[c].hi.x, nodes[c].hi.y, nodes[c].hi.z};
      larea[slots[k]] = area_of(lo, hi);
    }
    nl++;
  }
  if (nl < NL) return cost[root];  // fewer than 7 leaves reachable (cannot happen for sub_count >= 7, kept for safety)

  float llo[NL][3], lhi[NL][3], lcost[NL];
  uint32_t lcount[NL];
  for (int k = 0; k < NL; ++k) {
    llo[k][0] = nodes[leaf[k]].lo.x; llo[k][1] = nodes[leaf[k]].lo.y; llo[k][2] = nodes[leaf[k]].lo.z;
    lhi[k][0] = nodes[leaf[k]].hi.x; lhi[k][1] = nodes[leaf[k]].hi.y; lhi[k][2] = nodes[leaf[k]].hi.z;
    lcost[k] = cost[leaf[k]];
    lcount[k] = sub_count[leaf[k]];
  }
  // dynamic programming over the subsets of the 7 leaves
  float copt[128];
  uint8_t popt[128];
  for (uint32_t s = 1; s < 128u; ++s) {
    if ((s & (s - 1u)) == 0u) {  // singleton
      copt[s] = lcost[ffs32(s) - 1];
      popt[s] = 0;
      continue;
    }
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = 0; k < NL; ++k)
      if (s & (1u << k))
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], llo[k][a]); hi[a] = fmaxf(hi[a], lhi[k][a]); }
    const float a = area_of(lo, hi);
    // enumerate the partitions of s: subsets q of (s without its lowest bit), the other side keeps the lowest bit
    const uint32_t low = s & (0u - s);
    const uint32_t rest = s ^ low;
    float best = INFINITY;
    uint32_t bq = 0;
    for (uint32_t q = rest; q != 0u; q = (q - 1u) & rest) {
      const float c = copt[q] + copt[s ^ q];
      if (c < best) { best = c; bq = q; }
    }
    copt[s] = BRT_SAH_CI * a + best;
    popt[s] = (uint8_t)bq;
  }
  const float old_cost = cost[root];
  if (!(copt[127] < old_cost * 0.9999f)) return old_cost;

  // rewrite the treelet: internal[0] (the root) gets the full set, the other 5 ids are handed out top-down
  uint32_t stack_set[NL], stack_node[NL];
  int sp = 0, next_internal = 1;
  stack_set[sp] = 127u;
  stack_node[sp] = root;
  sp++;
  // children must be complete before a parent's bounds can be formed: first assign topology top-down, remembering
  // the order, then fill bounds / counts / costs bottom-up in reverse order
  uint32_t order_node[NL], order_set[NL];
  int n_order = 0;
  while (sp) {
    --sp;
    const uint32_t s = stack_set[sp], id = stack_node[sp];
    order_node[n_order] = id;
    order_set[n_order] = s;
    n_order++;
    const uint32_t q = popt[s], r = s ^ q;
    uint32_t child[2];
    const uint32_t sub[2] = {q, r};
    for (int k = 0; k < 2; ++k) {
      if ((sub[k] & (sub[k] - 1u)) == 0u) {
        child[k] = leaf[ffs32(sub[k]) - 1];
      } else {
        child[k] = internal[next_internal++];
        stack_set[sp] = sub[k];
        stack_node[sp] = child[k];
        sp++;
      }
      p.parent[child[k]] = id;
    }
    nodes[id].lo.w = u2f(child[0]);
    nodes[id].hi.w = u2f(child[1]);
  }
  for (int k = n_order - 1; k >= 0; --k) {
    const uint32_t id = order_node[k], s = order_set[k];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    uint32_t cnt = 0;
    for (int j = 0; j < NL; ++j)
      if (s & (1u << j)) {
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], llo[j][a]); hi[a] = fmaxf(hi[a], lhi[j][a]); }
        cnt += lcount[j];
      }
    nodes[id].lo.x = lo[0]; nodes[id].lo.y = lo[1]; nodes[id].lo.z = lo[2];
    nodes[id].hi.x = hi[0]; nodes[id].hi.y = hi[1]; nodes[id].hi.z = hi[2];
    sub_count[id] = cnt;
    cost[id] = copt[s];
  }
  return copt[127];
}

BRT_HD void treelet_body(const TreeletParams& p, uint32_t i) {
  const uint32_t n_int = p.count - 1;
  volatile BNode* nodes = p.nodes;
  volatile float* cost = p.cost;
  uint32_t cur = n_int + i;
  {
    const float lo[3] = {nodes[cur].lo.x, nodes[cur].lo.y, nodes[cur].lo.z};
    const float hi[3] = {nodes[cur].hi.x, nodes[cur].hi.y, nodes[cur].hi.z};
    cost[cur] = BRT_SAH_CT * area_of(lo, hi);
  }
  for (;;) {
    const uint32_t par = p.parent[cur];
    if (par == BRT_MISS) break;
    if (atomic_add_acq_rel(&p.arrive[par], 1u) == 0u) break;
    const uint32_t l = f2u(nodes[par].lo.w), r = f2u(nodes[par].hi.w);
    const float lo[3] = {nodes[par].lo.x, nodes[par].lo.y, nodes[par].lo.z};
    const float hi[3] = {nodes[par].hi.x, nodes[par].hi.y, nodes[par].hi.z};
    cost[par] = BRT_SAH_CI * area_of(lo, hi) + cost[l] + cost[r];
    if (p.restructure && ((volatile uint32_t*)p.sub_count)[par] >= BRT_TREELET_LEAVES) optimize_treelet(p, par, n_int);
    cur = par;
  }
}


#if !defined(BRT_EMU) || defined(BRT_EMU_WARP)  // (BRT_EMU_WARP: the test-only lock-step warp of tests/emu/warp_emu.h)
// ---- warp-cooperative version (device only) ------------------------------------------------------------------
// Same algorithm, but the 32 lanes of a warp share the dynamic programme of ONE treelet at a time (areas of the
// 127 subsets, then optimal costs by subset size with the partitions of the large subsets spread over the lanes).
// The bottom-up walk is unchanged (one leaf per thread); lanes that reach a treelet root in the same step have
// their treelets processed one after the other by the whole warp. This shortens the serial chain towards the
// root — the critical path of the pass — by more than an order of magnitude compared with one thread per treelet.
struct TreeletShared {  // per warp
  float area[128];
  float copt[128];
  uint8_t popt[128];
  float lo[BRT_TREELET_LEAVES][3], hi[BRT_TREELET_LEAVES][3];
  float lcost[BRT_TREELET_LEAVES];
  uint32_t lcount[BRT_TREELET_LEAVES];
  uint32_t leaf[BRT_TREELET_LEAVES], internal[BRT_TREELET_LEAVES];
};

// k-th (1-based, k < 2^popc(mask)) non-empty subset of `mask`: deposits the bits of k at the set bits of mask
__device__ __forceinline__ uint32_t deposit_bits(uint32_t k, uint32_t mask) {
  uint32_t out = 0;
  while (k) {
    const uint32_t low = mask & (0u - mask);
    if (k & 1u) out |= low;
    mask ^= low;
    k >>= 1;
  }
  return out;
}

__device__ __forceinline__ void warp_optimize_treelet(const TreeletParams& p, uint32_t root, uint32_t n_int, TreeletShared& sh, unsigned lane) {
  volatile BNode* nodes = p.nodes;
  volatile float* cost = p.cost;
  volatile uint32_t* sub_count = p.sub_count;
  const int NL = BRT_TREELET_LEAVES;
  // 1. treelet formation (lane 0), published through shared memory
  if (lane == 0) {
    float larea[NL];
    int nl = 2, ni = 1;
    sh.internal[0] = root;
    sh.leaf[0] = f2u(nodes[root].lo.w);
    sh.leaf[1] = f2u(nodes[root].hi.w);
    for (int k = 0; k < 2; ++k) {
      const uint32_t c = sh.leaf[k];
      const float lo[3] = {nodes[c].lo.x, nodes[c].lo.y, nodes[c].lo.z};
      const float hi[3] = {nodes[c].hi.x, nodes[c].hi.y, nodes[c].hi.z};
      larea[k] = area_of(lo, hi);
    }
    while (nl < NL) {
      int best = -1;
      float ba = -1.0f;
      for (int k = 0; k < nl; ++k)
        if (sh.leaf[k] < n_int && larea[k] > ba) { best = k; ba = larea[k]; }
      if (best < 0) break;
      const uint32_t id = sh.leaf[best];
      sh.internal[ni++] = id;
      sh.leaf[best] = f2u(nodes[id].lo.w);
      sh.leaf[nl] = f2u(nodes[id].hi.w);
      const int slots[2] = {best, nl};
      for (int k = 0; k < 2; ++k) {
        const uint32_t c = sh.leaf[slots[k]];
        const float lo[3] = {nodes[c].lo.x, nodes[c].lo.y, nodes[c].lo.z};
        const float hi[3] = {nodes[c].hi.x, nodes[c].hi.y, nodes[c].hi.z};
        larea[slots[k]] = area_of(lo, hi);
      }
      nl++;
    }
    if (nl < NL) sh.leaf[0] = BRT_MISS;
  }
  __syncwarp();
  if (sh.leaf[0] == BRT_MISS) return;
  // 2. leaf data
  if (lane < (unsigned)NL) {
    const uint32_t c = sh.leaf[lane];
    sh.lo[lane][0] = nodes[c].lo.x; sh.lo[lane][1] = nodes[c].lo.y; sh.lo[lane][2] = nodes[c].lo.z;
    sh.hi[lane][0] = nodes[c].hi.x; sh.hi[lane][1] = nodes[c].hi.y; sh.hi[lane][2] = nodes[c].hi.z;
    sh.lcost[lane] = cost[c];
    sh.lcount[lane] = sub_count[c];
  }
  __syncwarp();
  // 3. surface area of every subset; singletons seed the cost table
  for (uint32_t s = lane; s < 128u; s += 32u) {
    if (s == 0u) continue;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = 0; k < NL; ++k)
      if (s & (1u << k))
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], sh.lo[k][a]); hi[a] = fmaxf(hi[a], sh.hi[k][a]); }
    sh.area[s] = area_of(lo, hi);
    if ((s & (s - 1u)) == 0u) { sh.copt[s] = sh.lcost[__ffs(s) - 1]; sh.popt[s] = 0; }
  }
  __syncwarp();
  // 4. subsets of 2..5 leaves: one subset per lane, partitions enumerated serially (at most 15)
  for (int size = 2; size <= 5; ++size) {
    for (uint32_t s = lane; s < 128u; s += 32u) {
      if (__popc(s) != size) continue;
      const uint32_t low = s & (0u - s), rest = s ^ low;
      float best = INFINITY;
      uint32_t bq = 0;
      for (uint32_t q = rest; q != 0u; q = (q - 1u) & rest) {
        const float c = sh.copt[q] + sh.copt[s ^ q];
        if (c < best) { best = c; bq = q; }
      }
      sh.copt[s] = BRT_SAH_CI * sh.area[s] + best;
      sh.popt[s] = (uint8_t)bq;
    }
    __syncwarp();
  }
  // 5. subsets of 6 and 7 leaves (127 without one leaf, then 127): the 31 / 63 partitions of each are spread over the lanes
  for (int idx = 0; idx < 8; ++idx) {
    const uint32_t s = idx < 7 ? (127u ^ (1u << idx)) : 127u;
    const int size = idx < 7 ? 6 : 7;
    const uint32_t low = s & (0u - s), rest = s ^ low;
    const uint32_t n_part = (1u << (size - 1)) - 1u;
    float best = INFINITY;
    uint32_t bq = 0;
    for (uint32_t k = lane + 1u; k <= n_part; k += 32u) {
      const uint32_t q = deposit_bits(k, rest);
      const float c = sh.copt[q] + sh.copt[s ^ q];
      if (c < best) { best = c; bq = q; }
    }
    for (int off = 16; off > 0; off >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, off);
      const uint32_t oq = __shfl_xor_sync(0xffffffffu, bq, off);
      if (ob < best || (ob == best && oq < bq)) { best = ob; bq = oq; }
    }
    if (lane == 0) { sh.copt[s] = BRT_SAH_CI * sh.area[s] + best; sh.popt[s] = (uint8_t)bq; }
    if (idx >= 6) __syncwarp();  // the full set reads all 6-subsets
  }
  __syncwarp();
  // 6. rewrite (lane 0)
  if (lane == 0) {
    const float old_cost = cost[root];
    if (sh.copt[127] < old_cost * 0.9999f) {
      uint32_t stack_set[NL], stack_node[NL], order_node[NL], order_set[NL];
      int sp = 0, next_internal = 1, n_order = 0;
      stack_set[sp] = 127u;
      stack_node[sp] = root;
      sp++;
      while (sp) {
        --sp;
        const uint32_t s = stack_set[sp], id = stack_node[sp];
        order_node[n_order] = id;
        order_set[n_order] = s;
        n_order++;
        const uint32_t q = sh.popt[s], r = s ^ q;
        const uint32_t sub[2] = {q, r};
        uint32_t child[2];
        for (int k = 0; k < 2; ++k) {
          if ((sub[k] & (sub[k] - 1u)) == 0u) {
            child[k] = sh.leaf[__ffs(sub[k]) - 1];
          } else {
            child[k] = sh.internal[next_internal++];
            stack_set[sp] = sub[k];
            stack_node[sp] = child[k];
            sp++;
          }
          p.parent[child[k]] = id;
        }
        nodes[id].lo.w = u2f(child[0]);
        nodes[id].hi.w = u2f(child[1]);
      }
      for (int k = n_order - 1; k >= 0; --k) {
        const uint32_t id = order_node[k], s = order_set[k];
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        uint32_t cnt = 0;
        for (int j = 0; j < NL; ++j)
          if (s & (1u << j)) {
            for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], sh.lo[j][a]); hi[a] = fmaxf(hi[a], sh.hi[j][a]); }
            cnt += sh.lcount[j];
          }
        nodes[id].lo.x = lo[0]; nodes[id].lo.y = lo[1]; nodes[id].lo.z = lo[2];
        nodes[id].hi.x = hi[0]; nodes[id].hi.y = hi[1]; nodes[id].hi.z = hi[2];
        sub_count[id] = cnt;
        cost[id] = sh.copt[s];
      }
    }
    __threadfence();
  }
  __syncwarp();
}

__global__ void __launch_bounds__(128) k_treelet_warp(const TreeletParams p) {
  __shared__ TreeletShared shared[4];
  const unsigned lane = threadIdx.x & 31u;
  TreeletShared& sh = shared[threadIdx.x >> 5];
  const uint32_t n = p.count, n_int = p.count - 1;
  volatile BNode* nodes = p.nodes;
  volatile float* cost = p.cost;
  // whole warps iterate together (the cooperative part needs all 32 lanes), lanes beyond n idle
  for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += gridDim.x * blockDim.x) {
    const uint32_t i = base + lane;
    bool walking = i < n;
    uint32_t cur = n_int + i;
    if (walking) {
      const float lo[3] = {nodes[cur].lo.x, nodes[cur].lo.y, nodes[cur].lo.z};
      const float hi[3] = {nodes[cur].hi.x, nodes[cur].hi.y, nodes[cur].hi.z};
      cost[cur] = BRT_SAH_CT * area_of(lo, hi);
    }
    while (__any_sync(0xffffffffu, walking)) {
      uint32_t root = BRT_MISS;
      if (walking) {
        const uint32_t par = p.parent[cur];
        if (par == BRT_MISS) {
          walking = false;
        } else {
          __threadfence();
          if (atomicAdd(&p.arrive[par], 1u) == 0u) {
            walking = false;
          } else {
            __threadfence();
            const uint32_t l = f2u(nodes[par].lo.w), r = f2u(nodes[par].hi.w);
            const float lo[3] = {nodes[par].lo.x, nodes[par].lo.y, nodes[par].lo.z};
            const float hi[3] = {nodes[par].hi.x, nodes[par].hi.y, nodes[par].hi.z};
            cost[par] = BRT_SAH_CI * area_of(lo, hi) + cost[l] + cost[r];
            if (p.restructure && ((volatile uint32_t*)p.sub_count)[par] >= BRT_TREELET_LEAVES) root = par;
            cur = par;
          }
        }
      }
      unsigned todo = __ballot_sync(0xffffffffu, root != BRT_MISS);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1u;
        const uint32_t r = __shfl_sync(0xffffffffu, root, src);
        warp_optimize_treelet(p, r, n_int, sh, lane);
      }
    }
  }
}
#endif  // !BRT_EMU

}  // namespace brt
