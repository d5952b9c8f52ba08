/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Kind "port", PINNED to the reference itself: it reproduces frames computed by
 * executing the reference's shipped shader binary (tests/golden/spv_kat.json, tests/test_spv_kat.py), its host maths agree with the
 * reference's own Camera.cpp / MeshInstance.h / glm compiled into oracle/_ref (tests/test_ref_pin.py), its RNG with the integer
 * KATs of SH/random.slang (tests/golden/rng_kat.json). Not pinnable by any reference artefact, because the reference has no source
 * for them: BVH build, traversal, the ray/triangle test (brute force + analytic cases instead) and the extensions of DESIGN.md §7.
 *
 * C API of the CPU oracle (liboracle.so): a scalar, multithreaded C++ restatement of the
 * reference's shaders/ logic plus the driver pieces that have no source (BVH, ray/triangle test).
 * It mirrors include/brt.h one-to-one (orc_* instead of brt_*) so the parity tests drive both
 * through the same Python code. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libbrt.so) never does.
 */
#ifndef ORACLE_H_
#define ORACLE_H_
#include "../include/brt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_context orc_context;

/* cfg->device is ignored; cfg->flags bit 31 (ORC_CFG_BRUTE_FORCE) disables the BVH (tests only) */
#define ORC_CFG_BRUTE_FORCE 0x80000000u
int orc_create(const brt_config* cfg, orc_context** out);
void orc_destroy(orc_context* ctx);
const char* orc_last_error(const orc_context* ctx);
int orc_set_threads(orc_context* ctx, uint32_t n_threads); /* 0 = hardware_concurrency */
uint32_t orc_get_threads(const orc_context* ctx);

int orc_mesh_create(orc_context* ctx, const brt_vertex* v, uint32_t nv, const uint32_t* idx, uint32_t ni, uint32_t* mesh_id);
int orc_mesh_update_vertices(orc_context* ctx, uint32_t mesh_id, const brt_vertex* v, uint32_t nv);
int orc_sphere_create(orc_context* ctx, const float center[3], float radius, uint32_t* mesh_id);
int orc_material_create(orc_context* ctx, const brt_material* m, uint32_t* id);
int orc_material_set_transmission(orc_context* ctx, uint32_t id, float transmission, float ior);
int orc_light_create(orc_context* ctx, const brt_light* l, uint32_t* id);
int orc_sky_set(orc_context* ctx, const brt_sky* sky);
int orc_instance_create(orc_context* ctx, uint32_t mesh_id, uint32_t material_id, const float xform3x4[12], uint32_t* id);
int orc_instance_set_transform(orc_context* ctx, uint32_t id, const float xform3x4[12]);
int orc_instance_set_material(orc_context* ctx, uint32_t id, uint32_t material_id);
int orc_instance_destroy(orc_context* ctx, uint32_t id);
int orc_scene_build(orc_context* ctx);
int orc_smart_cull(orc_context* ctx, const brt_uniform* u, uint32_t w, uint32_t h, float threshold_px2, float hysteresis, uint32_t* visible);
int orc_get_visibility(orc_context* ctx, uint8_t* out, uint32_t n);
int orc_render_frame(orc_context* ctx, const brt_uniform* u, const brt_render_opts* opts, float* rgba_host);
int orc_get_aov(orc_context* ctx, int kind, void* out_host);
int orc_get_stats(orc_context* ctx, brt_stats* out);
int orc_debug_sort_pairs(orc_context* ctx, uint32_t* keys, uint32_t* vals, uint32_t n, int bits);
int orc_trace_rays(orc_context* ctx, const float* rays, uint32_t n, int closest, uint32_t* out);

/* known-answer / unit-test hooks into the transcribed shader functions */
uint32_t orc_kat_hash(uint32_t x, uint32_t y, uint32_t z);              /* SH/random.slang:2-12 */
uint32_t orc_kat_pcg(uint32_t* state);                                  /* SH/random.slang:14-19 */
float orc_kat_rand(uint32_t* state);                                    /* SH/random.slang:21-24 */
void orc_kat_brdf(const brt_material* m, const float N[3], const float V[3], const float L[3], float out[3]); /* SH/disney.slang:95-116 */
void orc_kat_sample_vndf(const brt_material* m, const float V[3], const float N[3], float r1, float r2, float out_dir[3], float* pdf); /* SH/sampler.slang:67-93 */
void orc_kat_sample_cosine(float r1, float r2, float out_dir[3], float* pdf); /* SH/sampler.slang:53-65 */
void orc_kat_sincos(float x, float* s, float* c);
float orc_kat_log2(float x);
/* one watertight ray/triangle test; returns 1 on hit and writes t,u,v */
int orc_kat_intersect_tri(const float o[3], const float d[3], float tmin, float tmax, const float v0[3], const float v1[3], const float v2[3], float tuv[3]);
/* Camera::setPerspectiveProjection + setView (Graphics/Camera.cpp:8-17,71-95) and the uniform
 * block of RTApp::run (RT/RTApp.cpp:44-49) */
float orc_kat_exp2(float x);
float orc_kat_srgb(float x);
uint32_t orc_kat_unorm8(float x);
int orc_kat_light_pdfs(orc_context* c, const float P[3], float* out, uint32_t n);
uint32_t orc_kat_light_sample(orc_context* c, const float P[3], float r, float* inv_pdf);
int orc_get_light_bvh(orc_context* c, brt_light_bvh_node* out, uint32_t max_nodes, uint32_t* n_nodes);
int orc_denoise(orc_context* c, const brt_uniform* u, const brt_denoise_opts* opts, float* rgba_host);
void orc_camera_handle_inputs(uint32_t keys, float dt, float position[3], float rotation[3]);
void orc_camera_uniform(const float pos[3], const float rot[3], float fovy, float aspect, float znear, float zfar,
                        uint32_t frame, uint32_t depth_max, brt_uniform* out);

#ifdef __cplusplus
}
#endif
#endif
