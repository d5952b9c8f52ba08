// SAH statistics and SAH treelet restructuring of the binary LBVH (Karras & Aila, HPG 2013).
#pragma once
#include "build_kernels.cuh"

namespace brt {

// SAH cost of the binary tree as the collapse will see it: subtrees with <= max_leaf primitives are
// leaves (cost = area * count), larger internal nodes cost their area once; normalised by the root area.
struct SahParams {
  uint32_t count;  // internal nodes
  const uint32_t* count_ptr;
  uint32_t n;
  const BNode* nodes;
  const uint32_t* sub_count;
  uint32_t max_leaf;
  float* out;
};
BRT_HD void sah_cost_body(const SahParams& p, uint32_t i) {
  const uint32_t n_int = p.n - 1;
  const BNode root = p.nodes[0];
  const float ra = box_area(xyz(root.lo), xyz(root.hi));
  if (!(ra > 0.0f)) return;
  if (p.sub_count[i] <= p.max_leaf) return;  // interior of a leaf
  const BNode nd = p.nodes[i];
  float c = box_area(xyz(nd.lo), xyz(nd.hi));
  const uint32_t ch[2] = {f2u(nd.lo.w), f2u(nd.hi.w)};
  for (int k = 0; k < 2; ++k) {
    const uint32_t cnt = ch[k] >= n_int ? 1u : p.sub_count[ch[k]];
    if (cnt <= p.max_leaf) {
      const BNode cn = p.nodes[ch[k]];
      c += box_area(xyz(cn.lo), xyz(cn.hi)) * (float)cnt;
    }
  }
#ifdef BRT_EMU
  *p.out += c / ra;
#else
  atomicAdd(p.out, c / ra);
#endif
}

struct TreeletParams {
  uint32_t count;
  const uint32_t* count_ptr;
};
BRT_HD void treelet_body(const TreeletParams&, uint32_t) {}

inline void run_treelet_passes(cudaStream_t, uint32_t, BNode*, uint32_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t, int) {}

}  // namespace brt
