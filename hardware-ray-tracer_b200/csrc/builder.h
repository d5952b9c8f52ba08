// Host interface of the GPU BVH builder (builder.cu).
#pragma once
#include "device_types.cuh"
#include <algorithm>
#include <cstdlib>

#include "host_util.h"

namespace brt {

struct BuildResult {
  uint32_t n_prims = 0;
  uint32_t n_nodes = 0;    // 8-wide nodes written
  uint32_t levels = 0;     // depth of the 8-wide tree
  float sah_lbvh = 0.0f;   // SAH cost of the binary tree before / after treelet restructuring
  float sah_final = 0.0f;
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};  // exact bounds of the primitives (triangle builds)
};

class Builder {
 public:
  explicit Builder(int sm_count, bool greedy_collapse = false, bool max_quality = false)
      : sm_count_(sm_count), greedy_collapse_(greedy_collapse) {
    if (max_quality) treelet_passes_ = 3;
    if (const char* e = getenv("BRT_TREELET_PASSES")) treelet_passes_ = std::max(1, std::min(8, atoi(e)));  // tuning aid
  }
  // BLAS over an indexed triangle mesh (brt_vertex stride). out_nodes must hold node_capacity(n_tris)
  // nodes, out_tris n_tris records. d_mesh_bounds (2 x float4, may be null) receives the exact mesh box.
  // Synchronises `stream` once at the end to read the result back.
  void build_triangles(cudaStream_t stream, const float* d_vertices, const uint32_t* d_indices, uint32_t n_tris, Node8* out_nodes,
                       TriRec* out_tris, float4* d_mesh_bounds, bool treelets, BuildResult* res);
  // TLAS over instances: primitive k is instance d_inst_ids[k], its traversal record is d_src[k].
  void build_instances(cudaStream_t stream, const InstShade* d_shade, const uint32_t* d_inst_ids, const InstRec* d_src, uint32_t n,
                       const float4* d_mesh_bounds, Node8* out_nodes, InstRec* out_inst, BuildResult* res);
  // test hook: the builder's radix sort on host arrays
  void debug_sort_pairs(cudaStream_t stream, uint32_t* keys_host, uint32_t* vals_host, uint32_t n, int bits);
  static uint32_t node_capacity(uint32_t n_prims) { return n_prims < 8 ? 8 : n_prims; }
  // allocates the scratch of a build of n primitives ahead of time (mesh upload), so that the first build does not pay for it
  void reserve(uint32_t n) { ensure_scratch(n); }

 private:
  void run(cudaStream_t stream, uint32_t n, uint32_t max_leaf, bool treelets, Node8* out_nodes, const float* d_vertices,
           const uint32_t* d_indices, TriRec* out_tris, const InstRec* d_src, InstRec* out_inst, float4* d_mesh_bounds, BuildResult* res);
  void ensure_scratch(uint32_t n);
  bool fused_small_ = false;  // build_instances -> run(): the one-block kernel runs the prologue (globals, instance boxes) too
  alignas(16) unsigned char fused_params_[192] = {0};  // InitGlobalsParams + InstBoundsParams of that call (their types live in builder.cu)
  size_t small_build_smem_set_ = 0;  // dynamic shared memory the one-block build kernel has been allowed so far
  int treelet_passes_ = 1;  // SAH treelet passes of a first ("fast trace") build: 1 (4.7 ms for 1M triangles, SAH 35.56) or 3 (9.5 ms, SAH 35.02)
  int sm_count_;
  bool greedy_collapse_;  // BRT_CFG_GREEDY_COLLAPSE
  DevBuf globals_, prim_lo_, prim_hi_, keys_[2], vals_[2], sort_tmp_, nodes_, parent_, arrive_, sub_count_, queue_[2], treelet_, wcost_, wplan_;
  friend struct BuilderAccess;
};

}  // namespace brt
