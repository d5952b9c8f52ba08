// C entry points over the reference's OWN host code, compiled where it lies under /root/reference (oracle/ref/Makefile):
//   Graphics/Camera.cpp (whole file, unmodified): setPerspectiveProjection :8-17, setView :19-24, handleInputs :26-61, updateView :71-95
//   Graphics/RayTracing/MeshInstance.h: calculateTransformation :38-86
//   libs/glm-1.0.1: glm::inverse, glm::transpose as called at Graphics/RayTracing/RTApp.cpp:45-46
// TEST INFRASTRUCTURE: the result (oracle/_ref/libref_host.so) checks the oracle's and the product's host maths; nothing in the
// product links or loads it. GLFW itself is a Win64 binary in the reference: glfwGetKey is answered from a key table here.
#include <cstdint>
#include <cstring>
#include <map>
#define private public  // Core::Camera keeps position / rotation private; the shim reads them back after handleInputs
#include "Graphics/Camera.h"
#undef private
#include "Graphics/RayTracing/MeshInstance.h"
#include <glm/glm.hpp>
#include <glm/gtc/type_ptr.hpp>

static std::map<int, int> g_keys;
extern "C" int glfwGetKey(GLFWwindow*, int key) {
	auto it = g_keys.find(key);
	return it == g_keys.end() ? GLFW_RELEASE : it->second;
}

// bit k of `mask` = the key the product's BRT_KEY_* bit k stands for (include/brt.h), in the order of Camera.h:24-35
static const int kKeyOfBit[10] = { GLFW_KEY_A, GLFW_KEY_D, GLFW_KEY_W, GLFW_KEY_S, GLFW_KEY_E, GLFW_KEY_Q,
                                   GLFW_KEY_RIGHT, GLFW_KEY_LEFT, GLFW_KEY_UP, GLFW_KEY_DOWN };

extern "C" {

void* ref_camera_new() { return new Core::Camera(); }
void ref_camera_delete(void* c) { delete static_cast<Core::Camera*>(c); }
void ref_camera_set_view(void* c, const float pos[3], const float rot[3]) {
	static_cast<Core::Camera*>(c)->setView(glm::make_vec3(pos), glm::make_vec3(rot));
}
void ref_camera_set_perspective(void* c, float fovy, float aspect, float znear, float zfar) {
	static_cast<Core::Camera*>(c)->setPerspectiveProjection(fovy, aspect, znear, zfar);
}
void ref_camera_handle_inputs(void* c, uint32_t mask, float dt) {
	g_keys.clear();
	for (int b = 0; b < 10; b++) g_keys[kKeyOfBit[b]] = (mask >> b) & 1u ? GLFW_PRESS : GLFW_RELEASE;
	static_cast<Core::Camera*>(c)->handleInputs(nullptr, dt);
}
void ref_camera_state(void* c, float pos[3], float rot[3]) {
	auto* cam = static_cast<Core::Camera*>(c);
	std::memcpy(pos, glm::value_ptr(cam->position), 12);
	std::memcpy(rot, glm::value_ptr(cam->rotation), 12);
}
// view / projection in glm memory order (column-major, 16 floats each)
void ref_camera_matrices(void* c, float view[16], float proj[16]) {
	auto* cam = static_cast<Core::Camera*>(c);
	glm::mat4 v = cam->getView(), p = cam->getProjection();
	std::memcpy(view, glm::value_ptr(v), 64);
	std::memcpy(proj, glm::value_ptr(p), 64);
}
// The 140-byte block RTApp::run writes every frame (RTApp.cpp:44-49; struct Uniform, RTPipeline.h:24-30):
// the two glm calls are the reference's, the byte layout is the struct's (two mat4, frame, depthMax, LIGHT_TRESHOLD = .0001f).
void ref_uniform(void* c, uint32_t frame, uint32_t depthMax, unsigned char out[140]) {
	auto* cam = static_cast<Core::Camera*>(c);
	glm::mat4 vi = glm::inverse(glm::transpose(cam->getView()));
	glm::mat4 pi = glm::inverse(glm::transpose(cam->getProjection()));
	float thr = .0001f;
	std::memcpy(out, glm::value_ptr(vi), 64);
	std::memcpy(out + 64, glm::value_ptr(pi), 64);
	std::memcpy(out + 128, &frame, 4);
	std::memcpy(out + 132, &depthMax, 4);
	std::memcpy(out + 136, &thr, 4);
}
void ref_inverse_transpose(const float m[16], float out[16]) {
	glm::mat4 r = glm::inverse(glm::transpose(glm::make_mat4(m)));
	std::memcpy(out, glm::value_ptr(r), 64);
}
// MeshInstance's 3x4 row-major VkTransformMatrixKHR (MeshInstance.h:82-85: scale + translate, rotation ignored)
void ref_instance_transform(const float pos[3], const float rot[3], const float scale[3], float out[12]) {
	RayTracing::MeshInstance inst(0, 0, glm::make_vec3(pos), glm::make_vec3(rot), glm::make_vec3(scale));
	VkTransformMatrixKHR t = inst.getTransformation();
	std::memcpy(out, &t, 48);
}

}
