// bloon.hpp — C++ host facade with the reference's class names over the C ABI of libbrt.so.
//
// A user of the reference's App/Graphics host API (RayTracing::Scene, RayTracing::Pipeline, Core::Camera,
// RayTracing::MeshInstance, the Vertex / Material / Light / Uniform structs) finds the same names, argument
// meaning and error behaviour (std::runtime_error, Graphics/Definitions.h:5) here; Vulkan-typed parameters
// are replaced by neutral ones. Header-only; link with -lbrt. Citations: RT/ = reference
// Graphics/RayTracing/, GFX/ = reference Graphics/.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../brt.h"

namespace bloon {
struct vec3 {  // stands in for glm::vec3 in the signatures
  float x = 0.0f, y = 0.0f, z = 0.0f;
  vec3() = default;
  vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
inline void check(int status, brt_context* ctx, const char* what) {
  if (status != BRT_OK) throw std::runtime_error(std::string(what) + ": " + brt_last_error(ctx));
}
}  // namespace bloon

// ---- OBJ ingestion: what Scene::loadModel (RT/Scene.cpp:29-74) does with tinyobjloader 1.0.6 (libs/tinyobj/tiny_obj_loader.h) ----
// Host-only (no GPU needed). The behaviour below is tinyobj 1.0.6's published behaviour, restated: LF / CR / CRLF line ends; `v`,
// `vn`, `vt`, `f`, `g`, `o`, `usemtl`, `mtllib` records (everything else ignored); missing numbers default to 0; its own decimal
// parser (digits accumulated in double, value = mantissa x 5^e x 2^e, then narrowed to float — not correctly rounded, and results
// must match it bit for bit); 1-based, 0 and negative (relative) indices; polygons fan-triangulated; faces collected per shape, a
// `usemtl` that changes the material moves the pending faces into the current shape, `g` / `o` close the shape — and drop it when
// no face followed the last material change (a quirk of that version, kept); material names come from the `mtllib` files, looked
// up relative to the working directory as the reference's call (no base directory) does. tests/test_ref_pin.py compares the
// result with the reference's own loadModel + tinyobj (oracle/_ref/libref_obj.so) on fuzzed files, byte for byte.
namespace bloon { namespace obj {

struct Mesh {
  std::vector<brt_vertex> vertices;
  std::vector<uint32_t> indices;
};
struct Corner { int v, vt, vn; };  // resolved zero-based indices, -1 = absent

inline bool isSpace(char c) { return c == ' ' || c == '\t'; }
inline bool isDigit(char c) { return (unsigned)(c - '0') < 10u; }
inline bool isLineEnd(char c) { return c == '\r' || c == '\n' || c == '\0'; }

// One line without its terminator; accepts "\n", "\r\n" and a lone "\r". False once the stream is exhausted.
inline bool getLine(std::istream& in, std::string& t) {
  t.clear();
  std::streambuf* sb = in.rdbuf();
  if (sb->sgetc() == std::char_traits<char>::eof()) return false;
  for (;;) {
    const int c = sb->sbumpc();
    if (c == '\n' || c == std::char_traits<char>::eof()) return true;
    if (c == '\r') { if (sb->sgetc() == '\n') sb->sbumpc(); return true; }
    t += (char)c;
  }
}

// [sign] digits ["." digits] [("e"|"E") [sign] digits]; greedy; false on a malformed number (the caller then keeps its default).
inline bool parseDouble(const char* s, const char* end, double* result) {
  if (s >= end) return false;
  double mantissa = 0.0;
  int exponent = 0, read = 0;
  char sign = '+', expSign = '+';
  const char* c = s;
  if (*c == '+' || *c == '-') sign = *c++;
  else if (!isDigit(*c)) return false;
  while (c != end && isDigit(*c)) { mantissa *= 10; mantissa += (int)(*c - '0'); ++c; ++read; }
  if (read == 0) return false;
  bool more = c != end;
  if (more && *c == '.') {
    static const double lut[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
    ++c;
    read = 1;
    while (c != end && isDigit(*c)) {
      mantissa += (int)(*c - '0') * (read < 8 ? lut[read] : std::pow(10.0, -read));
      ++read; ++c;
    }
    more = c != end;
  } else if (more && !(*c == 'e' || *c == 'E')) {
    more = false;  // trailing garbage ends the number
  }
  if (more && (*c == 'e' || *c == 'E')) {
    ++c;
    if (c != end && (*c == '+' || *c == '-')) expSign = *c++;
    else if (!isDigit(*c)) return false;  // "1e" alone is malformed
    read = 0;
    while (c != end && isDigit(*c)) { exponent *= 10; exponent += (int)(*c - '0'); ++c; ++read; }
    if (expSign == '-') exponent = -exponent;
    if (read == 0) return false;
  }
  *result = (sign == '+' ? 1 : -1) * (exponent ? std::ldexp(mantissa * std::pow(5.0, exponent), exponent) : mantissa);
  return true;
}

inline float parseReal(const char** tok, double dflt = 0.0) {
  *tok += std::strspn(*tok, " \t");
  const char* end = *tok + std::strcspn(*tok, " \t\r");
  double v = dflt;
  parseDouble(*tok, end, &v);
  *tok = end;
  return (float)v;
}

inline int fixIndex(int idx, int n) { return idx > 0 ? idx - 1 : (idx == 0 ? 0 : n + idx); }

// i, i/j, i//k, i/j/k
inline Corner parseCorner(const char** tok, int nv, int nvn, int nvt) {
  Corner c{-1, -1, -1};
  c.v = fixIndex(std::atoi(*tok), nv);
  *tok += std::strcspn(*tok, "/ \t\r");
  if (**tok != '/') return c;
  ++*tok;
  if (**tok == '/') {
    ++*tok;
    c.vn = fixIndex(std::atoi(*tok), nvn);
    *tok += std::strcspn(*tok, "/ \t\r");
    return c;
  }
  c.vt = fixIndex(std::atoi(*tok), nvt);
  *tok += std::strcspn(*tok, "/ \t\r");
  if (**tok != '/') return c;
  ++*tok;
  c.vn = fixIndex(std::atoi(*tok), nvn);
  *tok += std::strcspn(*tok, "/ \t\r");
  return c;
}

inline std::string firstWord(const char* s) {  // what sscanf("%s") reads
  s += std::strspn(s, " \t\n\v\f\r");
  return std::string(s, std::strcspn(s, " \t\n\v\f\r"));
}

// Material names of one .mtl file, numbered as tinyobj numbers them (a named material is registered when the next `newmtl` or the
// end of the file closes it; the last one is registered even when it has no name).
inline bool readMaterialNames(const std::string& file, std::map<std::string, int>& ids, int& count) {
  std::ifstream in(file.c_str());
  if (!in) return false;
  std::string line, name;
  while (getLine(in, line)) {
    const size_t last = line.find_last_not_of(" \t");
    line = last == std::string::npos ? std::string() : line.substr(0, last + 1);
    const char* tok = line.c_str();
    tok += std::strspn(tok, " \t");
    if (*tok == '\0' || *tok == '#') continue;
    if (std::strncmp(tok, "newmtl", 6) == 0 && isSpace(tok[6])) {
      if (!name.empty()) { ids.insert(std::make_pair(name, count)); ++count; }
      name = firstWord(tok + 7);
    }
  }
  ids.insert(std::make_pair(name, count));
  ++count;
  return true;
}

// The corner list of every shape of the file, in tinyobj's shape order (triangulated).
inline std::vector<Corner> parse(std::istream& in, std::vector<float>& v, std::vector<float>& vn, std::vector<float>& vt) {
  std::vector<Corner> out;              // shapes already closed, concatenated (loadModel walks them in order)
  std::vector<Corner> shape;            // the open shape
  std::vector<std::vector<Corner>> faceGroup;  // faces not yet moved into the open shape
  std::map<std::string, int> materialIds;
  int materialCount = 0, material = -1;
  auto flush = [&]() {  // faces -> open shape, fan-triangulated; false when there was nothing to move
    if (faceGroup.empty()) return false;
    for (const std::vector<Corner>& face : faceGroup)
      for (size_t k = 2; k < face.size(); ++k) { shape.push_back(face[0]); shape.push_back(face[k - 1]); shape.push_back(face[k]); }
    return true;
  };
  std::string line;
  while (getLine(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    const char* tok = line.c_str();
    tok += std::strspn(tok, " \t");
    if (*tok == '\0' || *tok == '#') continue;
    if (tok[0] == 'v' && isSpace(tok[1])) {
      tok += 2;
      for (int k = 0; k < 3; ++k) v.push_back(parseReal(&tok));
    } else if (tok[0] == 'v' && tok[1] == 'n' && isSpace(tok[2])) {
      tok += 3;
      for (int k = 0; k < 3; ++k) vn.push_back(parseReal(&tok));
    } else if (tok[0] == 'v' && tok[1] == 't' && isSpace(tok[2])) {
      tok += 3;
      for (int k = 0; k < 2; ++k) vt.push_back(parseReal(&tok));
    } else if (tok[0] == 'f' && isSpace(tok[1])) {
      tok += 2;
      tok += std::strspn(tok, " \t");
      std::vector<Corner> face;
      while (!isLineEnd(*tok)) {
        face.push_back(parseCorner(&tok, (int)(v.size() / 3), (int)(vn.size() / 3), (int)(vt.size() / 2)));
        tok += std::strspn(tok, " \t\r");
      }
      faceGroup.push_back(std::move(face));
    } else if (std::strncmp(tok, "usemtl", 6) == 0 && isSpace(tok[6])) {
      const auto it = materialIds.find(firstWord(tok + 7));
      const int id = it == materialIds.end() ? -1 : it->second;
      if (id != material) { flush(); faceGroup.clear(); material = id; }
    } else if (std::strncmp(tok, "mtllib", 6) == 0 && isSpace(tok[6])) {
      std::stringstream names(std::string(tok + 7));
      std::string file;
      while (std::getline(names, file, ' '))
        if (readMaterialNames(file, materialIds, materialCount)) break;
    } else if ((tok[0] == 'g' || tok[0] == 'o') && isSpace(tok[1])) {
      if (flush()) out.insert(out.end(), shape.begin(), shape.end());
      shape.clear();  // a shape whose last faces were already moved by a material change is dropped here, as in tinyobj 1.0.6
      faceGroup.clear();
    }
  }
  if (flush() || !shape.empty()) out.insert(out.end(), shape.begin(), shape.end());
  return out;
}

inline Mesh load(std::istream& in) {
  std::vector<float> v, vn, vt;
  const std::vector<Corner> corners = parse(in, v, vn, vt);
  Mesh m;
  // key: the 8 floats with -0 folded onto +0 (the reference compares with ==, RT/Scene.h:33-37); the first vertex seen is kept
  std::unordered_map<std::string, uint32_t> unique;
  auto inRange = [](int i, size_t n, const char* what) {
    if ((size_t)i >= n) throw std::runtime_error(std::string("OBJ: ") + what + " index out of range");
  };
  for (const Corner& c : corners) {
    brt_vertex x{};
    if (c.v >= 0) { inRange(c.v, v.size() / 3, "vertex"); x.pos[0] = v[3 * c.v]; x.pos[1] = -v[3 * c.v + 1]; x.pos[2] = v[3 * c.v + 2]; }                 // :47-51
    if (c.vn >= 0) { inRange(c.vn, vn.size() / 3, "normal"); x.normal[0] = vn[3 * c.vn]; x.normal[1] = -vn[3 * c.vn + 1]; x.normal[2] = vn[3 * c.vn + 2]; }  // :53-57
    if (c.vt >= 0) { inRange(c.vt, vt.size() / 2, "texcoord"); x.uv[0] = vt[2 * c.vt]; x.uv[1] = vt[2 * c.vt + 1]; }                                       // :59-62
    float k[8];
    std::memcpy(k, &x, sizeof(k));
    for (float& f : k) if (f == 0.0f) f = 0.0f;
    const std::string key(reinterpret_cast<const char*>(k), sizeof(k));
    auto it = unique.find(key);
    if (it == unique.end()) { it = unique.emplace(key, (uint32_t)m.vertices.size()).first; m.vertices.push_back(x); }  // :64-67
    m.indices.push_back(it->second);                                                                                  // :69
  }
  return m;
}

inline Mesh load(const std::string& path) {
  std::ifstream in(path.c_str());
  if (!in) throw std::runtime_error("Cannot open file [" + path + "]\n");  // tinyobj's message, thrown at RT/Scene.cpp:38-41
  return load(in);
}

}}  // namespace bloon::obj

namespace Core {

// Core::Device (vulkan_core/Device.h): owns the connection to the GPU — here one brt_context.
class Device {
 public:
  explicit Device(int cudaDevice = 0, uint32_t tileRank = 0, uint32_t tileWorld = 1, uint32_t flags = 0) {
    brt_config cfg{};
    cfg.struct_size = sizeof(cfg);
    cfg.device = cudaDevice;
    cfg.tile_rank = tileRank;
    cfg.tile_world = tileWorld;
    cfg.flags = flags;
    int rc = brt_create(&cfg, &ctx_);
    if (rc != BRT_OK) throw std::runtime_error(std::string("failed to create the ray tracing device: ") + brt_last_error(nullptr));
  }
  ~Device() { brt_destroy(ctx_); }
  Device(const Device&) = delete;
  Device& operator=(const Device&) = delete;
  brt_context* getDevice() const { return ctx_; }

 private:
  brt_context* ctx_ = nullptr;
};

// Core::Camera (GFX/Camera.h:8-17): setView / setPerspectiveProjection keep the reference's conventions
// (Y down, +Z forward, Tait-Bryan Y-X-Z, Vulkan 0..1 depth). handleInputs takes the key state as a BRT_KEY_* mask instead of a
// GLFWwindow to poll.
class Camera {
 public:
  void setPerspectiveProjection(float fovy, float aspectRatio, float near_, float far_) {
    fovy_ = fovy; aspect_ = aspectRatio; near_z_ = near_; far_z_ = far_;
  }
  void setView(bloon::vec3 position, bloon::vec3 rotation) { position_ = position; rotation_ = rotation; }
  // Camera::handleInputs(GLFWwindow*, float dt), Graphics/Camera.cpp:26-61 — keys = BRT_KEY_* bits of the pressed keys
  void handleInputs(uint32_t keys, float dt) {
    float p[3] = {position_.x, position_.y, position_.z}, r[3] = {rotation_.x, rotation_.y, rotation_.z};
    brt_camera_handle_inputs(keys, dt, p, r);
    position_ = bloon::vec3(p[0], p[1], p[2]);
    rotation_ = bloon::vec3(r[0], r[1], r[2]);
  }
  bloon::vec3 getPosition() const { return position_; }
  bloon::vec3 getRotation() const { return rotation_; }
  // glm-style column-major matrices (GFX/Camera.cpp:8-17, 71-95)
  std::array<float, 16> getProjection() const {
    std::array<float, 16> m{};
    const float t = std::tan(fovy_ / 2.0f);
    m[0] = 1.0f / (aspect_ * t);
    m[5] = 1.0f / t;
    m[10] = far_z_ / (far_z_ - near_z_);
    m[11] = 1.0f;
    m[14] = -(far_z_ * near_z_) / (far_z_ - near_z_);
    return m;
  }
  std::array<float, 16> getView() const {
    const float c3 = std::cos(rotation_.z), s3 = std::sin(rotation_.z);
    const float c2 = std::cos(rotation_.x), s2 = std::sin(rotation_.x);
    const float c1 = std::cos(rotation_.y), s1 = std::sin(rotation_.y);
    const float u[3] = {c1 * c3 + s1 * s2 * s3, c2 * s3, c1 * s2 * s3 - c3 * s1};
    const float v[3] = {c3 * s1 * s2 - c1 * s3, c2 * c3, c1 * c3 * s2 + s1 * s3};
    const float w[3] = {c2 * s1, -s2, c1 * c2};
    const float p[3] = {position_.x, position_.y, position_.z};
    std::array<float, 16> m{};
    for (int k = 0; k < 3; ++k) { m[4 * k + 0] = u[k]; m[4 * k + 1] = v[k]; m[4 * k + 2] = w[k]; }
    m[12] = -(u[0] * p[0] + u[1] * p[1] + u[2] * p[2]);
    m[13] = -(v[0] * p[0] + v[1] * p[1] + v[2] * p[2]);
    m[14] = -(w[0] * p[0] + w[1] * p[1] + w[2] * p[2]);
    m[15] = 1.0f;
    return m;
  }
  // Uniform{inverse(transpose(view)), inverse(transpose(proj)), frame, depthMax} of RTApp::run (RT/RTApp.cpp:44-49)
  brt_uniform uniform(uint32_t frame, uint32_t depthMax) const {
    brt_uniform u;
    const float pos[3] = {position_.x, position_.y, position_.z}, rot[3] = {rotation_.x, rotation_.y, rotation_.z};
    brt_camera_uniform(pos, rot, fovy_, aspect_, near_z_, far_z_, frame, depthMax, &u);
    return u;
  }

 private:
  bloon::vec3 position_, rotation_;
  float fovy_ = 1.0471975512f, aspect_ = 1.0f, near_z_ = 0.001f, far_z_ = 100000.0f;
};

}  // namespace Core

namespace RayTracing {

using Vertex = brt_vertex;      // RT/Scene.h:28-38
using Material = brt_material;  // RT/Scene.h:50-62
using Light = brt_light;        // RT/Scene.h:70-75
using Uniform = brt_uniform;    // RT/RTPipeline.h:24-30
struct Extent2D { uint32_t width, height; };

// RT/MeshInstance.h: mesh id, material id, position / rotation / scale -> 3x4 transform. As in the
// reference only scale + translate is live (RT/MeshInstance.h:82-85; the rotation code is commented out).
class MeshInstance {
 public:
  MeshInstance(uint32_t meshId, uint32_t materialId, bloon::vec3 position = {}, bloon::vec3 rotation = {}, bloon::vec3 scale = {1, 1, 1})
      : meshId(meshId), materialId(materialId), position(position), rotation(rotation), scale(scale) { calculateTransformation(); }
  void setPosition(bloon::vec3 p) { position = p; calculateTransformation(); }
  void setRotation(bloon::vec3 r) { rotation = r; calculateTransformation(); }
  void setScale(bloon::vec3 s) { scale = s; calculateTransformation(); }
  void setMeshId(uint32_t id) { meshId = id; }
  void setMaterialId(uint32_t id) { materialId = id; }
  bloon::vec3 getPosition() const { return position; }
  bloon::vec3 getRotation() const { return rotation; }
  const float* getTransformation() const { return transform; }
  uint32_t getMeshId() const { return meshId; }
  uint32_t getMaterialId() const { return materialId; }

 private:
  void calculateTransformation() {
    const float t[12] = {scale.x, 0.0f, 0.0f, position.x, 0.0f, scale.y, 0.0f, position.y, 0.0f, 0.0f, scale.z, position.z};
    std::memcpy(transform, t, sizeof(t));
  }
  uint32_t meshId, materialId;
  bloon::vec3 position, rotation, scale;
  float transform[12];
};

// RT/Scene.h:132-192
class Scene {
 public:
  explicit Scene(Core::Device& device) : device(device) {}
  Scene(const Scene&) = delete;
  Scene& operator=(const Scene&) = delete;

  // RT/Scene.cpp:29-74: OBJ -> flip Y of positions and normals -> de-duplicate identical vertices -> ONE indexed mesh over all
  // shapes of the file. Parsing follows tinyobjloader 1.0.6, the reference's parser (bloon::obj below); returns the mesh id
  // (the reference returns void and the id is the running mesh count — the same number).
  uint32_t loadModel(const std::string& path) {
    const bloon::obj::Mesh m = bloon::obj::load(path);
    return createMesh(m.vertices, m.indices);
  }
  // Mesh::Mesh (RT/Scene.cpp:419-458) for procedural geometry
  uint32_t createMesh(const std::vector<Vertex>& vertices, const std::vector<uint32_t>& indices) {
    uint32_t id = 0;
    bloon::check(brt_mesh_create(ctx(), vertices.data(), (uint32_t)vertices.size(), indices.data(), (uint32_t)indices.size(), &id), ctx(), "createMesh");
    return id;
  }
  void updateMeshVertices(uint32_t meshId, const std::vector<Vertex>& vertices) {
    bloon::check(brt_mesh_update_vertices(ctx(), meshId, vertices.data(), (uint32_t)vertices.size()), ctx(), "updateMeshVertices");
  }
  uint32_t createSphere(bloon::vec3 center, float radius) {  // extension
    const float c[3] = {center.x, center.y, center.z};
    uint32_t id = 0;
    bloon::check(brt_sphere_create(ctx(), c, radius, &id), ctx(), "createSphere");
    return id;
  }
  void createInstance(uint32_t meshId, uint32_t materialId, bloon::vec3 position = {}, bloon::vec3 rotation = {}, bloon::vec3 scale = {1, 1, 1}) {
    MeshInstance inst(meshId, materialId, position, rotation, scale);  // RT/Scene.cpp:76-78
    uint32_t id = 0;
    bloon::check(brt_instance_create(ctx(), meshId, materialId, inst.getTransformation(), &id), ctx(), "createInstance");
    instances.push_back(inst);
  }
  // RT/Scene.cpp:80-86: emissive arguments are accepted and ignored, specular defaults to 0.5, the rest to 0
  uint32_t createMaterial(bloon::vec3 color, float metallic = 0.f, float roughness = 1.f, bloon::vec3 /*emissiveColor*/ = {}, float /*emissionStrength*/ = 0.f) {
    Material m{};
    m.color[0] = color.x; m.color[1] = color.y; m.color[2] = color.z;
    m.metallic = metallic;
    m.roughness = roughness;
    m.specular = 0.5f;
    uint32_t id = 0;
    bloon::check(brt_material_create(ctx(), &m, &id), ctx(), "createMaterial");
    return id;
  }
  void createLight(bloon::vec3 position, bloon::vec3 color, float intensity) {  // RT/Scene.cpp:88-97: always POINT
    Light l{};
    l.pos[0] = position.x; l.pos[1] = position.y; l.pos[2] = position.z;
    l.color[0] = color.x; l.color[1] = color.y; l.color[2] = color.z;
    l.intensity = intensity;
    l.type = BRT_LIGHT_POINT;
    bloon::check(brt_light_create(ctx(), &l, nullptr), ctx(), "createLight");
  }
  void build() { bloon::check(brt_scene_build(ctx()), ctx(), "Scene::build"); }  // RT/Scene.cpp:100-120
  void destroyInstance(uint32_t instanceID) {                                    // RT/Scene.cpp:122-125 (swap-remove)
    bloon::check(brt_instance_destroy(ctx(), instanceID), ctx(), "destroyInstance");
    instances[instanceID] = instances.back();
    instances.pop_back();
  }
  void unloadModel(uint32_t) {}      // empty in the reference too (RT/Scene.cpp:126-133)
  void destroyLight(uint32_t) {}
  void destroyMaterial(uint32_t) {}
  MeshInstance& getInstance(uint32_t id) { return instances.at(id); }
  void commitInstance(uint32_t id) {  // push setPosition / setScale / setMaterialId changes of getInstance(id) to the GPU
    bloon::check(brt_instance_set_transform(ctx(), id, instances.at(id).getTransformation()), ctx(), "commitInstance");
    bloon::check(brt_instance_set_material(ctx(), id, instances.at(id).getMaterialId()), ctx(), "commitInstance");
  }
  // The reference's prepareRendering() prints "Not implemented!" and throws "LBVH not implemented!"
  // (RT/Scene.cpp:135-138). Here it is the per-frame entry: rebuild the BLAS of updated meshes with the GPU LBVH
  // builder, and — when a threshold is given — run Smart Culling (README.md:15-18) and rebuild the TLAS.
  uint32_t prepareRendering(const Uniform* uniform = nullptr, Extent2D extent = {0, 0}, float thresholdPx2 = 0.0f, float hysteresis = 0.25f) {
    uint32_t visible = (uint32_t)instances.size();
    // brt_smart_cull rebuilds the BLAS of updated meshes itself and builds the TLAS once, after the visibility is known
    if (uniform) bloon::check(brt_smart_cull(ctx(), uniform, extent.width, extent.height, thresholdPx2, hysteresis, &visible), ctx(), "prepareRendering");
    else build();
    return visible;
  }
  // Scene::getTlas() / getSceneInfoBuffer() (RT/Scene.h:150-151): the reference hands out its Vulkan acceleration structure and the buffer
  // behind descriptor binding 3; here the device arrays of the software TLAS and the 80-byte SceneBufferInfo (host copy + the
  // address of its device copy), whose tables keep the reference's byte layouts.
  using AccelerationStructure = brt_accel_info;
  using SceneBufferInfo = brt_scene_buffer_info;
  using InstanceInfo = brt_instance_info;
  AccelerationStructure getTlas() {
    AccelerationStructure a{};
    bloon::check(brt_get_tlas(ctx(), &a), ctx(), "getTlas");
    return a;
  }
  SceneBufferInfo getSceneInfoBuffer(uint64_t* deviceAddress = nullptr) {
    SceneBufferInfo info{};
    bloon::check(brt_get_scene_info_buffer(ctx(), &info, deviceAddress), ctx(), "getSceneInfoBuffer");
    return info;
  }
  Core::Device& getDeviceRef() { return device; }

 private:
  brt_context* ctx() { return device.getDevice(); }
  Core::Device& device;
  std::vector<MeshInstance> instances;
};

// RT/RTPipeline.h:34-51 — the render-frame entry. bind / bindDescriptorSets have no CUDA analogue and are no-ops.
class Pipeline {
 public:
  Pipeline(Core::Device& device, Extent2D extent, Scene& scene) : device(device), extent(extent), scene(scene) { rebuildRenderOutput(extent); }
  static std::unique_ptr<Pipeline> createPipeline(Core::Device& device, Extent2D extent, Scene& scene) {
    return std::unique_ptr<Pipeline>(new Pipeline(device, extent, scene));
  }
  void bind() {}
  void bindDescriptorSets(uint32_t) {}
  void writeToUniformBuffer(const void* data, uint32_t index) { std::memcpy(&uniforms[index % 2], data, sizeof(Uniform)); current = index % 2; }  // RT/RTPipeline.cpp:44-47
  // vkCmdTraceRaysKHR(width, height, 1) (RT/RTPipeline.cpp:41-43) + the image read-back of copyImageToSwapchain
  void traceRays(uint32_t width, uint32_t height, uint32_t /*depth*/ = 1, uint32_t spp = 1, uint32_t renderFlags = 0) {
    if (width != extent.width || height != extent.height) rebuildRenderOutput({width, height});
    brt_render_opts o{};
    o.width = width; o.height = height; o.spp = spp; o.flags = renderFlags | BRT_RENDER_FORMAT(format);
    void* dst = format == BRT_FORMAT_R32G32B32A32_SFLOAT ? (void*)storageImage.data() : (void*)storageImage8.data();
    bloon::check(brt_render_frame(device.getDevice(), &uniforms[current], &o, static_cast<float*>(dst)), device.getDevice(), "traceRays");
  }
  // Two frames in flight, as RTApp::beginFrame / endFrame over SwapChain::MAX_FRAMES_IN_FLIGHT = 2 (RT/RTApp.cpp:171-212,
  // vulkan_core/SwapChain.h:8): submitFrame(frameIndex) = bindDescriptorSets(cmd, frameIndex) + traceRays + submitCommandBuffers
  // with the uniform written for that index; waitFrame(frameIndex) = the slot's fence, returns that slot's image.
  void submitFrame(uint32_t frameIndex, uint32_t width, uint32_t height, uint32_t spp = 1, uint32_t renderFlags = 0) {
    if (width != extent.width || height != extent.height) rebuildRenderOutput({width, height});
    const uint32_t k = frameIndex % 2;
    bloon::check(brt_frame_wait(device.getDevice(), k), device.getDevice(), "submitFrame");  // the image below is about to be rewritten
    inFlight[k].assign((size_t)width * height * 4, 0.0f);
    brt_render_opts o{};
    o.width = width; o.height = height; o.spp = spp; o.flags = renderFlags;
    bloon::check(brt_render_frame_async(device.getDevice(), &uniforms[k], &o, k, inFlight[k].data()), device.getDevice(), "submitFrame");
  }
  std::vector<float>& waitFrame(uint32_t frameIndex) {
    bloon::check(brt_frame_wait(device.getDevice(), frameIndex % 2), device.getDevice(), "waitFrame");
    return inFlight[frameIndex % 2];
  }
  void rebuildRenderOutput(Extent2D e) { rebuildRenderOutput(format, e); }
  // Pipeline::rebuildRenderOutput(VkFormat format, VkExtent2D extent), RT/RTPipeline.cpp:49-55: the storage image is created in the
  // swapchain's format (BRT_FORMAT_*: R32G32B32A32_SFLOAT, or an 8-bit RGBA / BGRA UNORM / SRGB format)
  void rebuildRenderOutput(uint32_t fmt, Extent2D e) {
    brt_frame_wait(device.getDevice(), 0);
    brt_frame_wait(device.getDevice(), 1);
    extent = e;
    format = fmt;
    storageImage.assign(fmt == BRT_FORMAT_R32G32B32A32_SFLOAT ? (size_t)e.width * e.height * 4 : 0, 0.0f);
    storageImage8.assign(fmt == BRT_FORMAT_R32G32B32A32_SFLOAT ? 0 : (size_t)e.width * e.height * 4, 0);
  }
  std::vector<uint8_t>& getRenderOutput8() { return storageImage8; }  // the 8-bit formats: 4 bytes per texel, row-major
  void updateTopLevelAS() {}                                                                                       // empty in the reference (RT/RTPipeline.cpp:57-59)
  std::vector<float>& getRenderOutput() { return storageImage; }  // linear RGBA32F, row-major (outImage, SH/raytracing.slang:132)
  brt_stats getStats() { brt_stats s{}; brt_get_stats(device.getDevice(), &s); return s; }

 private:
  Core::Device& device;
  Extent2D extent;
  Scene& scene;
  Uniform uniforms[2]{};  // MAX_FRAMES_IN_FLIGHT = 2 (vulkan_core/SwapChain.h:8)
  uint32_t current = 0;
  uint32_t format = BRT_FORMAT_R32G32B32A32_SFLOAT;
  std::vector<float> storageImage;
  std::vector<uint8_t> storageImage8;
  std::vector<float> inFlight[2];
};

}  // namespace RayTracing

namespace Extensions {

// Extensions::Denoiser (Graphics/Denoiser/Denoiser.h:14-20 declares the class without a body; the stages listed at :5-12 are what
// brt_denoise runs). denoise() filters the frame the pipeline traced last — trace it with BRT_RENDER_GBUFFER — and keeps the
// temporal history in the device context.
class Denoiser {
 public:
  explicit Denoiser(Core::Device& device) : device(device) {
    opts.struct_size = sizeof(opts);
    opts.flags = BRT_DENOISE_BILATERAL;
    opts.iterations = 4;
    opts.sigma_n_log2 = 5;
    opts.sigma_z = 0.05f;
    opts.sigma_l = 4.0f;
    opts.clamp_gamma = 0.0f;
    opts.max_history = 32.0f;
  }
  ~Denoiser() = default;
  Denoiser(const Denoiser&) = delete;
  Denoiser& operator=(const Denoiser&) = delete;
  brt_denoise_opts& options() { return opts; }
  void reset() { resetNext = true; }
  // returns the filtered linear RGBA32F image
  const std::vector<float>& denoise(const RayTracing::Uniform& uniform, uint32_t width, uint32_t height) {
    out.assign((size_t)width * height * 4, 0.0f);
    brt_denoise_opts o = opts;
    if (resetNext) o.flags |= BRT_DENOISE_RESET;
    resetNext = false;
    bloon::check(brt_denoise(device.getDevice(), &uniform, &o, out.data()), device.getDevice(), "Denoiser::denoise");
    return out;
  }

 private:
  Core::Device& device;
  brt_denoise_opts opts{};
  bool resetNext = false;
  std::vector<float> out;
};

}  // namespace Extensions

