/*
 * brt.h — C ABI of the B200-native Bloon RT hot path (libbrt.so).
 *
 * This is the drop-in boundary for the reference's ray-gen / closest-hit / miss shader pipeline:
 * everything the reference does through `RayTracing::Scene`, `RayTracing::Pipeline` and the four
 * descriptor bindings (TLAS, outImage, Uniform, SceneBufferInfo) is reachable through the plain-C
 * entry points below. Citations use the shorthand
 *   RT/ = /root/reference/Hardware Ray Tracer/Graphics/RayTracing/
 *   SH/ = /root/reference/Hardware Ray Tracer/shaders/
 *
 * Conventions
 *   - every function returns an int status (BRT_OK == 0); brt_last_error(ctx) gives the message.
 *     The reference throws std::runtime_error (Graphics/Definitions.h:5); the C++ facade in
 *     include/bloon/ re-raises a non-zero status as std::runtime_error.
 *   - handles are opaque; input arrays are copied during the call; one context is not thread-safe.
 *   - the library has NO CPU path: brt_create fails when no CUDA device is usable.
 *   - POD layouts are byte-identical to the reference's host structs / shader structs.
 */
#ifndef BRT_H_
#define BRT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* every entry point below is exported from libbrt.so; everything else in the library is hidden */
#if defined(__GNUC__)
#define BRT_API __attribute__((visibility("default")))
#else
#define BRT_API
#endif

#define BRT_OK 0
#define BRT_ERR_INVALID 1  /* bad argument / bad id */
#define BRT_ERR_CUDA 2     /* CUDA runtime error (message in brt_last_error) */
#define BRT_ERR_STATE 3    /* call order violated (e.g. render before build) */
#define BRT_ERR_LIMIT 4    /* internal capacity exceeded (e.g. BVH deeper than the traversal stack) */

/* ---- POD layouts shared with the reference ------------------------------------------------- */

/* RT/Scene.h:28-31, read by SH/objects.slang:49-51 at offsets 0/12/24, stride 32 */
typedef struct brt_vertex {
  float pos[3];
  float normal[3];
  float uv[2];
} brt_vertex;

/* RT/Scene.h:50-62 == SH/material.slang:3-15, 13 floats, stride 52 */
typedef struct brt_material {
  float color[3];
  float subsurface;
  float metallic;
  float roughness;
  float specular;
  float specularTint;
  float anisotropic;
  float sheen;
  float sheenTint;
  float clearCoat;
  float clearCoatGloss;
} brt_material;

/* RT/Scene.h:64-75, read by SH/light.slang:25-28 at offsets 0/12/24/28, stride 32 */
enum { BRT_LIGHT_POINT = 0, BRT_LIGHT_SPOT = 1, BRT_LIGHT_DIRECTIONAL = 2 };
typedef struct brt_light {
  float pos[3];
  float color[3];
  float intensity;
  uint8_t type;
  uint8_t pad_[3];
} brt_light;

/* RT/RTPipeline.h:24-30 == SH/raytracing.slang:9-15 (std140 offsets 0/64/128/132/136), 140 bytes.
 * viewInverse / projInverse hold glm::inverse(glm::transpose(M)) in glm column-major memory
 * (RT/RTApp.cpp:44-49); the kernels read them exactly the way the shader's mul(rowVec, M) does. */
typedef struct brt_uniform {
  float viewInverse[16];
  float projInverse[16];
  uint32_t frame;
  uint32_t depthMax;
  float lightThreshold; /* uploaded but never read by the shader (SH/raytracing.slang:79 uses the macro) */
} brt_uniform;

/* RT/Scene.h:90-104, 88 bytes. The reference uploads it (RT/Scene.cpp:333-355) and never reads it;
 * here it feeds the optional BRT_RENDER_SKY miss colour (extension). */
typedef struct brt_sky {
  float skyColor[3];
  float horizonColor[3];
  float groundColor[3];
  float sunDirection[3];
  float upDirection[3];
  float brightness;
  float horizonSize;
  float angularSize;
  float glowIntensity;
  float glowSharpness;
  float glowSize;
  float lightRadiance;
} brt_sky;

/* RT/Scene.h:84-88, 24 bytes, one per instance: what SH/objects.slang:15-28 reads through instanceBuffer. Addresses are CUDA device
 * pointers of the mesh's vertex (stride 32) and index buffers. */
typedef struct brt_instance_info {
  uint64_t vertexAddress;
  uint64_t indexAddress;
  uint32_t materialId;
  uint32_t pad;
} brt_instance_info;

/* RT/Scene.h:106-121 = SceneInfo of SH/raytracing.slang:17-32 (descriptor binding 3), 80 bytes: device addresses and strides of the
 * scene tables. The tables behind the addresses have the reference's byte layouts (brt_material 52 B, brt_light 32 B, brt_instance_info
 * 24 B, brt_sky 88 B), so device code written against the reference's SceneInfo reads them unchanged. */
typedef struct brt_scene_buffer_info {
  uint64_t mBuf, mStride;         /* materials */
  uint64_t lBuf, lStride, lCount; /* lights */
  uint64_t vStride;               /* vertex stride (32) */
  uint64_t sBuf, sStride;         /* brt_instance_info per instance */
  uint64_t skyBuf, skyStride;
} brt_scene_buffer_info;

/* RT/Scene.h:77-82 AccelerationStructure {handle, buffer, memory, address} of Scene::getTlas(): here the device arrays of the software
 * TLAS (compressed 8-wide nodes, 80 B each; instance records, 96 B each). */
typedef struct brt_accel_info {
  uint64_t handle;   /* = address */
  uint64_t buffer;   /* device pointer of the node array */
  uint64_t memory;   /* device pointer of the instance-record array */
  uint64_t address;  /* device pointer of the root node */
  uint32_t n_nodes, n_instances; /* extra */
} brt_accel_info;

/* ---- library-specific PODs ------------------------------------------------------------------ */

typedef struct brt_config {
  uint32_t struct_size; /* sizeof(brt_config) */
  int32_t device;       /* CUDA device ordinal */
  uint32_t tile_rank;   /* this context renders the 32x32 image tiles t (row-major ids) with (t % W + W - (t / W) % W) % W == tile_rank, W = tile_world:
                         * groups of W consecutive tiles, ranks rotated by the group index (diagonals, not columns, when the tile columns are a multiple of W) */
  uint32_t tile_world;  /* 1 = whole image */
  uint32_t flags;       /* BRT_CFG_* */
} brt_config;

#define BRT_CFG_COUNTERS 1u /* run the instrumented traversal kernels (node/primitive visit counters) */
#define BRT_CFG_NO_TREELET 2u /* skip the SAH treelet refinement pass of the LBVH builder */
/* Build policy (the Vulkan build flags of RT/Scene.cpp:169 restated): the first build of a mesh is
 * PREFER_FAST_TRACE (LBVH + SAH treelet restructuring); rebuilds after brt_mesh_update_vertices are
 * PREFER_FAST_BUILD (plain LBVH, what the reference's prepareRendering placeholder names) unless this flag is set. */
#define BRT_CFG_TREELET_ON_REBUILD 4u
/* run every kernel of a frame on one stream (default: the shadow chain of a round overlaps the next round's
 * closest-hit traversal on a second stream). Same results; used when per-kernel event times must not overlap. */
#define BRT_CFG_NO_OVERLAP 8u
/* launch every kernel of every frame individually instead of replaying the CUDA graph captured for the frame shape (single-GPU
 * contexts capture one graph per frame slot and shape). A replayed frame reports ms_total only: the per-class kernel times of
 * brt_stats need this flag (BRT_CFG_NO_OVERLAP and BRT_CFG_COUNTERS imply it). */
#define BRT_CFG_NO_GRAPH 16u
/* measurement only: collapse the binary hierarchy into 8-wide nodes by opening the child with the largest area until the 8 slots
 * are full (the builder's first rule) instead of the area-optimal choice of wide nodes (build_kernels.cuh, wide_cost_body).
 * Same frames (results never depend on the BVH's shape), more and emptier nodes. */
#define BRT_CFG_GREEDY_COLLAPSE 32u
/* Opt-in: bounce rounds shade their paths in the order of the hit positions (counting sort by the cell of the hit on a 64^3 grid
 * over the scene), so that the shadow and bounce rays they emit start at neighbouring origins. Frames are bit-identical either way.
 * Measured on B200 (profiles/r2_traversal.md): traversal -0..4 %, shade + sort +2 ms on C5 — a loss, hence off by default. */
#define BRT_CFG_HIT_SORT 64u
/* First ("fast trace") builds run ONE SAH treelet pass (1M triangles: 4.7 ms, SAH 35.56); with this flag three (9.5 ms, SAH 35.02,
 * -0.7 % frame time on C5): for scenes built once and rendered for a long time. */
#define BRT_CFG_TREELET_PASSES_3 128u

/* render mode flags. With none of the BOUNCE flags set the behaviour is the reference's live path:
 * direct light + hard shadows, weight = 0 after the first hit (SH/raytracing.slang:168). */
#define BRT_RENDER_BOUNCE_REFLECT 1u /* the reference's dormant bounce, SH/raytracing.slang:166-167 */
#define BRT_RENDER_BOUNCE_REFRACT 2u /* extension: dielectric transmission for materials with transmission > 0 */
#define BRT_RENDER_BOUNCE_DIFFUSE 4u /* extension: cosine-hemisphere GI bounce (SH/sampler.slang:53-65 + toWorld) */
#define BRT_RENDER_JITTER 8u         /* use the sub-pixel jitter the shader computes but drops (SH/raytracing.slang:96-98) */
#define BRT_RENDER_SKY 16u           /* miss returns a sky gradient instead of black (SH/raytracing.slang:173-176) */
/* Light BVH (the reference declares LightBVHNode, RT/Scene.h:123-130, and announces it at SH/raytracing.slang:76: "This will
 * later be replaced with a light bounding volume hierarchy system"): instead of looping over all lights, every hit samples ONE
 * light by a stochastic descent of the light BVH (child chosen with probability ~ totalFlux / distance^2) and divides its
 * contribution by the probability — unbiased, one shadow ray per hit whatever the light count, no BRT_MAX_LIGHTS limit.
 * All lights must be POINT lights (the only kind Scene::createLight creates, RT/Scene.cpp:88-97). */
#define BRT_RENDER_LIGHT_BVH 64u
/* run the denoiser stages (brt_denoise_configure) right after the resolve, on the frame's own stream: the image handed back (and
 * converted to the present format) is the denoised one. Implies BRT_RENDER_GBUFFER. Frames must be submitted in display order. */
#define BRT_RENDER_DENOISE 128u
#define BRT_RENDER_FAST_SHADING 2048u /* opt-in: shade kernels with FMA contraction and approximate division / rsqrt (not bit-identical to the
                                       * oracle any more; primary ids unchanged, radiance within the 1e-3 relative RMSE bar) */
#define BRT_RENDER_GBUFFER 32u       /* also keep world position + shading normal of the primary hit (input of brt_denoise) */

/* Output format of the image handed back by the render entry points (bits 8..10 of brt_render_opts.flags): the format
 * Pipeline::rebuildRenderOutput(format, extent) creates the storage image in — the swapchain's (RT/RTPipeline.cpp:49-55,
 * VK/SwapChain.cpp:384-392 prefers R32G32B32A32_SFLOAT and falls back to what the surface offers, in practice 8-bit BGRA) —
 * so that copyImageToSwapchain (RT/RTApp.cpp:87-152) is a plain copy. 8-bit formats: 4 bytes per pixel; UNORM = clamp to
 * [0, 1], x255, round to nearest even; SRGB additionally applies the sRGB transfer function to R, G, B. The conversion runs
 * on the GPU after the resolve; the library keeps the linear RGBA32F image as well (brt_device_image, AOVs). */
enum { BRT_FORMAT_R32G32B32A32_SFLOAT = 0, BRT_FORMAT_R8G8B8A8_UNORM = 1, BRT_FORMAT_B8G8R8A8_UNORM = 2,
       BRT_FORMAT_R8G8B8A8_SRGB = 3, BRT_FORMAT_B8G8R8A8_SRGB = 4 };
#define BRT_RENDER_FORMAT_SHIFT 8
#define BRT_RENDER_FORMAT_MASK 0x700u
#define BRT_RENDER_FORMAT(fmt) ((uint32_t)(fmt) << BRT_RENDER_FORMAT_SHIFT)

typedef struct brt_render_opts {
  uint32_t width, height;   /* vkCmdTraceRaysKHR(w, h, 1), RT/RTPipeline.cpp:41-43 */
  uint32_t spp;             /* SAMPLES (SH/constants.slang:23-25); sample s uses frame + s as RNG frame */
  uint32_t flags;           /* BRT_RENDER_* | BRT_RENDER_FORMAT(BRT_FORMAT_*) */
  uint32_t crop_x0, crop_y0, crop_w, crop_h; /* crop_w == 0: whole image; else only this window is traced */
} brt_render_opts;

enum { BRT_AOV_PRIM_ID = 0, BRT_AOV_INST_ID = 1, BRT_AOV_HIT_T = 2,
       BRT_AOV_POSITION = 3, BRT_AOV_NORMAL = 4 /* 4 floats per pixel, frames rendered with BRT_RENDER_GBUFFER */ };
#define BRT_AOV_MISS 0xffffffffu

typedef struct brt_stats {
  /* last frame */
  uint64_t rays_closest;   /* closest-hit queries (primary + bounce) */
  uint64_t rays_occlusion; /* occlusion queries (shadow rays) */
  /* counters, only with BRT_CFG_COUNTERS */
  uint64_t nodes_visited_closest, prims_tested_closest, spheres_tested_closest;
  uint64_t nodes_visited_occlusion, prims_tested_occlusion, spheres_tested_occlusion;
  /* device time of the last frame per kernel class, CUDA events on the library's stream (ms); frames replayed from a CUDA graph
   * fill ms_total only (see BRT_CFG_NO_GRAPH) */
  float ms_raygen, ms_trace_closest, ms_shade, ms_trace_occlusion, ms_accumulate, ms_resolve, ms_total;
  uint32_t launches_trace_closest, launches_trace_occlusion, launches_total;
  /* last brt_scene_build / brt_smart_cull */
  float ms_blas_build, ms_tlas_build, ms_cull;
  uint32_t blas_built; /* number of BLAS rebuilt by the last brt_scene_build */
  /* scene */
  uint64_t total_triangles; /* sum over instances */
  uint64_t bvh_nodes;       /* all BLAS + TLAS 8-wide nodes */
  uint64_t bvh_bytes;       /* nodes + primitive records resident in HBM */
  float sah_cost;           /* SAH cost of the largest BLAS (binary tree, after refinement) */
  float sah_cost_lbvh;      /* same before refinement */
  uint32_t instances_visible, instances_total;
  /* last brt_denoise: device time of all its kernels (ms) and how many were launched */
  float ms_denoise;
  uint32_t launches_denoise;
} brt_stats;

/* RT/Scene.h:123-130, 48 bytes. Binary tree: childIndex >= 0: index of the left child (the right one follows it);
 * childIndex < 0: leaf holding light -1 - childIndex. Point lights emit in every direction: coneAngle = pi. */
typedef struct brt_light_bvh_node {
  float bBoxMin[3];
  float bBoxMax[3];
  float totalFlux; /* sum of intensity * luminance(color) below this node */
  float coneAxis[3];
  float coneAngle;
  int32_t childIndex;
} brt_light_bvh_node;

/* ---- context ------------------------------------------------------------------------------- */
typedef struct brt_context brt_context;

/* Device::Device + Pipeline ctor (vulkan_core/Device.cpp:45-53, RT/RTPipeline.cpp:4-26): picks the
 * CUDA device, creates the stream and the per-frame buffers. Fails (BRT_ERR_CUDA) without a GPU. */
BRT_API int brt_create(const brt_config* cfg, brt_context** out);
BRT_API void brt_destroy(brt_context* ctx);
BRT_API const char* brt_last_error(const brt_context* ctx);
/* Launch all work of this context on an externally owned cudaStream_t (e.g. torch's current stream). */
BRT_API int brt_set_stream(brt_context* ctx, void* cuda_stream);

/* ---- scene upload (RayTracing::Scene, RT/Scene.h:134-151) -------------------------------- */
/* Mesh::Mesh (RT/Scene.cpp:419-458): vertices (stride 32) + uint32 indices, 3 per triangle. */
BRT_API int brt_mesh_create(brt_context* ctx, const brt_vertex* vertices, uint32_t n_vertices,
                    const uint32_t* indices, uint32_t n_indices, uint32_t* mesh_id);
/* Scene::prepareRendering placeholder (RT/Scene.cpp:135-138, "LBVH not implemented!"): replace the
 * vertex array of a mesh; the next brt_scene_build rebuilds that BLAS with the GPU LBVH builder. */
BRT_API int brt_mesh_update_vertices(brt_context* ctx, uint32_t mesh_id, const brt_vertex* vertices, uint32_t n_vertices);
/* extension (no reference counterpart): an analytic sphere usable as a mesh id in brt_instance_create */
BRT_API int brt_sphere_create(brt_context* ctx, const float center[3], float radius, uint32_t* mesh_id);
/* Scene::createMaterial (RT/Scene.cpp:80-86) */
BRT_API int brt_material_create(brt_context* ctx, const brt_material* m, uint32_t* material_id);
/* extension: dielectric parameters kept in a parallel array so the 52-byte material stays intact */
BRT_API int brt_material_set_transmission(brt_context* ctx, uint32_t material_id, float transmission, float ior);
/* Scene::createLight (RT/Scene.cpp:88-97) */
BRT_API int brt_light_create(brt_context* ctx, const brt_light* l, uint32_t* light_id);
/* Scene::createSky (RT/Scene.cpp:333-355) */
BRT_API int brt_sky_set(brt_context* ctx, const brt_sky* sky);
/* Scene::createInstance (RT/Scene.cpp:76-78) + MeshInstance::calculateTransformation
 * (RT/MeshInstance.h:82-85): xform is the row-major 3x4 object->world matrix of
 * VkAccelerationStructureInstanceKHR (RT/Scene.cpp:183-190). */
BRT_API int brt_instance_create(brt_context* ctx, uint32_t mesh_id, uint32_t material_id, const float xform3x4[12], uint32_t* instance_id);
BRT_API int brt_instance_set_transform(brt_context* ctx, uint32_t instance_id, const float xform3x4[12]);
BRT_API int brt_instance_set_material(brt_context* ctx, uint32_t instance_id, uint32_t material_id);
/* Scene::destroyInstance (RT/Scene.cpp:122-125): swap-remove, the last instance takes this id */
BRT_API int brt_instance_destroy(brt_context* ctx, uint32_t instance_id);
/* Scene::build (RT/Scene.cpp:100-120): BLAS for new/dirty meshes (GPU LBVH + SAH treelets + BVH8),
 * TLAS over the visible instances, material/light/instance tables. */
BRT_API int brt_scene_build(brt_context* ctx);
/* README.md:15-18 "Smart Culling": screen-space footprint per instance, hysteresis, then TLAS rebuild
 * over the survivors. threshold_px2 <= 0 makes every instance visible again.
 * This is also the per-frame entry (Scene::prepareRendering, RT/Scene.cpp:135-138): meshes updated since the last build
 * (brt_mesh_update_vertices) get their BLAS rebuilt here, and the TLAS is built once, after the visibility is known —
 * no brt_scene_build is needed in between. */
BRT_API int brt_smart_cull(brt_context* ctx, const brt_uniform* u, uint32_t width, uint32_t height,
                   float threshold_px2, float hysteresis, uint32_t* visible_count);
/* copy the per-instance visibility flags (1 byte each) of the last brt_smart_cull to the host */
BRT_API int brt_get_visibility(brt_context* ctx, uint8_t* out, uint32_t n);
/* Scene::getSceneInfoBuffer() (RT/Scene.h:151; built by createSceneInformation / createSceneInfoBuffer, RT/Scene.cpp:357-403): fills
 * *out with the 80-byte table and, if d_copy is not NULL, stores the device address of its copy in device memory there (what
 * binding 3 points at). The scene must be built. */
BRT_API int brt_get_scene_info_buffer(brt_context* ctx, brt_scene_buffer_info* out, uint64_t* d_copy);
/* Scene::getTlas() (RT/Scene.h:150) */
BRT_API int brt_get_tlas(brt_context* ctx, brt_accel_info* out);

/* ---- render-frame entry (Pipeline::writeToUniformBuffer + traceRays, RT/RTPipeline.cpp:41-47) - */
/* Traces the frame and, when rgba_host != NULL, copies the image back to host memory — the reference's outImage
 * (SH/raytracing.slang:132): linear RGBA32F (w*h*16 bytes, row major, alpha = 1) by default, or w*h*4 bytes in the 8-bit
 * format selected with BRT_RENDER_FORMAT (tile_world == 1 only).
 * With tile_world > 1 only the pixels of this rank's tiles are written, the rest stays 0. */
BRT_API int brt_render_frame(brt_context* ctx, const brt_uniform* u, const brt_render_opts* opts, float* rgba_host);
/* Frames in flight. The reference records frame k+1 while frame k is still on the GPU: MAX_FRAMES_IN_FLIGHT = 2
 * (VK/SwapChain.h:8), one fence per frame slot waited in acquireNextImage (VK/SwapChain.cpp:45-60) and signalled by
 * submitCommandBuffers (VK/SwapChain.cpp:92-131), one uniform buffer and descriptor set per slot
 * (RT/RTPipeline.cpp:44-47, bindDescriptorSets(cmd, frameIndex)). brt_render_frame_async is that submit: it first waits
 * for the slot's previous frame (the fence), enqueues the whole frame on the slot's own streams and buffers — including
 * the copy to rgba_host (pinned memory for a truly asynchronous copy; NULL = leave it on the device) — and returns.
 * brt_frame_wait(slot) returns when that frame and its copy are complete and folds its timings into brt_get_stats.
 * The latency-bound tail of one frame's wavefronts then overlaps the head of the next one. Scene-changing calls
 * (build, mesh update, Smart Culling) drain all slots first. With tile_world > 1 the image holds this rank's tiles only. */
#ifndef BRT_FRAMES_IN_FLIGHT
#define BRT_FRAMES_IN_FLIGHT 4 /* slots available; the reference uses 2; more let the latency-bound tail of small frames (one rank's share of a tiled frame) overlap */
#endif
BRT_API int brt_render_frame_async(brt_context* ctx, const brt_uniform* u, const brt_render_opts* opts, uint32_t slot, float* rgba_host);
BRT_API int brt_frame_wait(brt_context* ctx, uint32_t slot);
/* the cudaStream_t a slot's frame is enqueued on (slot 0: the context's stream), e.g. to order caller work after it */
BRT_API void* brt_frame_stream(brt_context* ctx, uint32_t slot);
/* Same, but leaves the result on the device: d_tiles (device pointer, may be NULL) receives this
 * rank's tiles packed tile-major (brt_tile_buffer_bytes bytes) for the NCCL gather. */
BRT_API int brt_render_frame_tiles(brt_context* ctx, const brt_uniform* u, const brt_render_opts* opts, void* d_tiles);
/* bytes of one rank's packed tile buffer (identical on every rank: the tile count is padded) */
BRT_API size_t brt_tile_buffer_bytes(uint32_t width, uint32_t height, uint32_t tile_world);
/* after the gather: d_all = tile_world packed buffers back to back (device) -> row-major RGBA32F
 * image at d_rgba (device). */
BRT_API int brt_untile(brt_context* ctx, const void* d_all, uint32_t width, uint32_t height, uint32_t tile_world, void* d_rgba);
/* Fused framebuffer exchange for the ranks of one box (the alternative to brt_render_frame_tiles + NCCL all-gather + brt_untile):
 * every rank exports one allocation holding BRT_GATHER_IMAGES full-frame "gather images" and a block of completion flags (a cudaIpc
 * handle), opens everybody else's, and the resolve kernel of brt_render_frame_peers* stores each pixel this rank owns straight into
 * the RECEIVERS' gather images — row-major, the final layout, RGBA32F or the 8-bit present format of opts.flags — with st.global on
 * the mapped peer pointers over NVLink. Receivers are all ranks (default) or one root rank (brt_gather_configure; the north star only
 * asks for the frame on one rank). Completion needs no collective and no host barrier: the last block of a rank's resolve kernel
 * publishes a sequence number in every receiver's flag block (release at system scope); a receiver enqueues brt_gather_wait on the
 * stream that reads the image, reads it, and enqueues brt_gather_release, which lets the producers overwrite that image BRT_GATHER_IMAGES
 * frames later (their resolve kernels wait for it on the device). Slot k of brt_render_frame_peers_async uses gather image k, so up to
 * BRT_GATHER_IMAGES frames are in flight per rank. Every rank must submit the same frames on the same slots in the same order; a
 * receiver must release every frame it receives. A wait gives up after 20 s (brt_gather_timed_out) instead of hanging the GPU.
 * Replaces nothing in the reference (single GPU, VK/Device.cpp:388-389); it is the tile split the north star names. */
#ifndef BRT_GATHER_IMAGES
#define BRT_GATHER_IMAGES 4
#endif
#define BRT_IPC_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t) */
BRT_API int brt_gather_image_export(brt_context* ctx, uint32_t width, uint32_t height, void* handle_out);
/* handles: tile_world handles of BRT_IPC_HANDLE_BYTES bytes each, in rank order (the own one is ignored) */
BRT_API int brt_gather_image_open(brt_context* ctx, const void* handles, uint32_t world);
/* root_only != 0: only rank `root` receives the pixels; before the first frame of the exchange */
BRT_API int brt_gather_configure(brt_context* ctx, uint32_t root_only, uint32_t root);
/* synchronous: next slot in turn, returns when this rank's stores have landed (the receivers' view: brt_gather_wait) */
BRT_API int brt_render_frame_peers(brt_context* ctx, const brt_uniform* u, const brt_render_opts* opts);
BRT_API int brt_render_frame_peers_async(brt_context* ctx, const brt_uniform* u, const brt_render_opts* opts, uint32_t slot);
/* receiver side, for the frame submitted last on `slot`; stream = a cudaStream_t, NULL = the library's own gather stream */
BRT_API int brt_gather_wait(brt_context* ctx, uint32_t slot, void* stream);
BRT_API int brt_gather_release(brt_context* ctx, uint32_t slot, void* stream);
/* wait + copy of the gather image (16 or 4 bytes per pixel) to host memory + release, all on `stream` */
BRT_API int brt_gather_copy_to_host(brt_context* ctx, uint32_t slot, void* host, void* stream);
BRT_API int brt_gather_timed_out(brt_context* ctx);
/* device pointer of gather image `slot` (complete once brt_gather_wait has passed on the reading stream) */
BRT_API void* brt_gather_image(brt_context* ctx, uint32_t slot);
/* device pointer of the context's own full-frame RGBA32F image of the last frame (the slot last waited for) */
BRT_API void* brt_device_image(brt_context* ctx);

/* ---- denoiser slot (Graphics/Denoiser/Denoiser.h:5-20) -------------------------------------------------------
 * The reference declares Extensions::Denoiser::denoise() without a body; the comment above it lists the stages: temporal
 * accumulation (with reprojection), history clamping (to prevent ghosting), variance estimation, a-trous wavelet denoiser,
 * bilateral pass. (The README's DLSS Ray Reconstruction is proprietary: out of scope.) brt_denoise runs those stages on the
 * frame rendered last with BRT_RENDER_GBUFFER — `u` is that frame's uniform — keeps the history (accumulated colour, luminance
 * moments, G-buffer, camera) in the context for the next call, and copies the filtered linear RGBA32F image to rgba_host
 * (may be NULL; brt_denoised_image() is the device copy). The arithmetic is specified in DESIGN.md §12. */
typedef struct brt_denoise_opts {
  uint32_t struct_size;   /* sizeof(brt_denoise_opts) */
  uint32_t flags;         /* BRT_DENOISE_* */
  uint32_t iterations;    /* a-trous iterations, hole size 2^i, 0..6 */
  uint32_t sigma_n_log2;  /* normal edge-stopping weight max(0, N.N')^(2^k), 0..8 */
  float sigma_z;          /* depth edge-stopping, relative to hit distance x hole size */
  float sigma_l;          /* luminance edge-stopping, in standard deviations */
  float clamp_gamma;      /* history clamped to mean +- gamma sigma of the 3x3 neighbourhood; <= 0: no clamping */
  float max_history;      /* frames a pixel accumulates at most */
} brt_denoise_opts;
#define BRT_DENOISE_RESET 1u      /* drop the history first (camera cut, scene change) */
#define BRT_DENOISE_BILATERAL 2u  /* final 3x3 joint-bilateral pass */
BRT_API int brt_denoise(brt_context* ctx, const brt_uniform* u, const brt_denoise_opts* opts, float* rgba_host);
/* options of the denoiser stages that frames rendered with BRT_RENDER_DENOISE run (default: 4 a-trous iterations + bilateral pass,
 * sigma_n 2^5, sigma_z 0.05, sigma_l 4, no history clamping, 32 frames of history); BRT_DENOISE_RESET in flags is honoured by every
 * such frame until it is configured away */
BRT_API int brt_denoise_configure(brt_context* ctx, const brt_denoise_opts* opts);
BRT_API void* brt_denoised_image(brt_context* ctx);

/* copies the light BVH built by the last brt_scene_build to the host: up to max_nodes nodes, returns the node count in *n_nodes
 * (2 * lights - 1; 0 without lights). Built on the host (lights are few), deterministic. */
BRT_API int brt_get_light_bvh(brt_context* ctx, brt_light_bvh_node* out, uint32_t max_nodes, uint32_t* n_nodes);

/* ---- test / measurement only ---------------------------------------------------------------- */
/* primary-hit AOVs of sample 0 of the last frame: uint32 prim id, uint32 instance id
 * (BRT_AOV_MISS on a miss), float hit distance; w*h elements each */
BRT_API int brt_get_aov(brt_context* ctx, int kind, void* out_host);
BRT_API int brt_get_stats(brt_context* ctx, brt_stats* out);
/* Trace caller-supplied rays (8 floats each: origin xyz, tmin, direction xyz, tmax) against the
 * built scene; closest != 0: out = 4 x uint32 per ray {float bits of t, prim, instance, hit?};
 * closest == 0: out[4*i+3] = occluded ? 1 : 0. Used by the brute-force-vs-BVH equivalence tests. */
BRT_API int brt_trace_rays(brt_context* ctx, const float* rays_host, uint32_t n_rays, int closest, uint32_t* out_host);

/* Sorts n (key, value) pairs in place (host arrays) by the low `bits` bits of the key with the LBVH builder's own GPU
 * radix sort; stable. Exposed so that the sort can be tested on its own. */
BRT_API int brt_debug_sort_pairs(brt_context* ctx, uint32_t* keys_host, uint32_t* vals_host, uint32_t n, int bits);
/* Copies the BLAS of a triangle mesh to the host: compressed 8-wide nodes (80 bytes each, layout in DESIGN.md §4) and triangle records
 * (48 bytes each: three float4 = the vertices, primitive index in the bits of the first w), up to the given capacities; the counts
 * come back in *n_nodes / *n_tris (null pointers: counts only). Used by the structural tests of the builder (every primitive once, every
 * quantised child box contains what is below it); the driver's acceleration structures the reference uses are opaque. */
BRT_API int brt_debug_get_blas(brt_context* ctx, uint32_t mesh_id, void* nodes_out, uint32_t node_capacity, void* tris_out, uint32_t tri_capacity,
                               uint32_t* n_nodes, uint32_t* n_tris);

/* ---- host helper: Core::Camera + the uniform block of RTApp::run ---------------------------- */
/* measurement aid (bench.py's roofline context): GB/s of a hand-written read-only pass (16-byte loads) over `bytes` of device memory,
 * best of `iters` after a warm-up pass; 64 MiB stays in the 126 MB L2 */
BRT_API int brt_debug_l2_read_gbs(brt_context* ctx, size_t bytes, uint32_t iters, float* gbs_out);
/* Camera::setView/updateView (Graphics/Camera.cpp:19-24,71-95), Camera::setPerspectiveProjection
 * (Graphics/Camera.cpp:8-17) and Uniform{inverse(transpose(view)), inverse(transpose(proj)), frame,
 * depthMax} (RT/RTApp.cpp:44-49). rot = (pitch x, yaw y, roll z), Tait-Bryan Y-X-Z. Pure host code. */
BRT_API void brt_camera_uniform(const float pos[3], const float rot[3], float fovy, float aspect, float znear, float zfar,
                        uint32_t frame, uint32_t depth_max, brt_uniform* out);

/* Camera::handleInputs (Graphics/Camera.cpp:26-61) with the key state as a bit mask instead of glfwGetKey: arrow keys turn the
 * camera at 1.5 rad/s (pitch clamped to +-1.5, yaw wrapped to [0, 2 pi)), W/S/D/A/E/Q move it at 3 units/s along the yaw-forward,
 * right and up (0,-1,0) axes. Updates position and rotation in place. Pure host code. */
#define BRT_KEY_MOVE_LEFT 1u      /* GLFW_KEY_A, Graphics/Camera.h:24-35 */
#define BRT_KEY_MOVE_RIGHT 2u     /* D */
#define BRT_KEY_MOVE_FORWARD 4u   /* W */
#define BRT_KEY_MOVE_BACKWARD 8u  /* S */
#define BRT_KEY_MOVE_UP 16u       /* E */
#define BRT_KEY_MOVE_DOWN 32u     /* Q */
#define BRT_KEY_LOOK_RIGHT 64u    /* arrow keys */
#define BRT_KEY_LOOK_LEFT 128u
#define BRT_KEY_LOOK_UP 256u
#define BRT_KEY_LOOK_DOWN 512u
BRT_API void brt_camera_handle_inputs(uint32_t keys, float dt, float position[3], float rotation[3]);

#ifdef __cplusplus
}
#endif
#endif /* BRT_H_ */
