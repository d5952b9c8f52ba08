// The reference's demo application (RT/RTApp.cpp:3-59) on the B200 path, through the C++ facade that keeps the
// reference's class names: scene set-up of RTApp::RTApp, the per-frame uniform block and rayTraceScene() of
// RTApp::run — minus window, swapchain and present (no display on a compute box): the frame is written to disk.
//
//   g++ -std=c++17 -O2 -Iinclude examples/rtapp_demo.cpp -Lhardware-ray-tracer_b200/lib -lbrt -Wl,-rpath,... -o rtapp_demo
//   ./rtapp_demo out_prefix [frames] [async]      (async: two frames in flight, RTApp::beginFrame / endFrame)
//   ./rtapp_demo out_prefix [frames] post         (camera walk with Camera::handleInputs, denoised frames, BGRA8 present image)
#include <cstdio>
#include <fstream>
#include <iostream>

#include "bloon/bloon.hpp"

using bloon::vec3;

static void writePlaneObj(const std::string& path) {  // stands in for the missing models/Plane.obj (a 2 x 2 quad in the XZ plane)
  std::ofstream f(path);
  f << "v -1 0 -1\nv 1 0 -1\nv 1 0 1\nv -1 0 1\n"
       "vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n"
       "vn 0 1 0\n"
       "f 1/1/1 2/2/1 3/3/1 4/4/1\n";
}

int main(int argc, char** argv) {
  const std::string prefix = argc > 1 ? argv[1] : "rtapp_demo";
  const int frames = argc > 2 ? std::atoi(argv[2]) : 1;
  const bool async = argc > 3 && std::string(argv[3]) == "async";
  const bool post = argc > 3 && std::string(argv[3]) == "post";
  try {
    const uint32_t width = 800, height = 600;  // Window({800, 600, ...}), RT/RTApp.cpp:3
    Core::Device device(0);
    RayTracing::Scene scene(device);
    writePlaneObj(prefix + "_Plane.obj");
    scene.loadModel(prefix + "_Plane.obj");                                    // RT/RTApp.cpp:4
    scene.createMaterial(vec3(1.f, 1.f, 1.f), 1.0f);                           // :6
    scene.createMaterial(vec3(1.f, 1.f, 1.f), 1.0f, 0.0f);                     // :7
    scene.createLight(vec3(1.f, 0.f, 0.f), vec3(0.0f, 0.0f, 1.0f), 2.0f);      // :9-11
    scene.createLight(vec3(-1.f, 0.0f, 0.0f), vec3(0.0f, 1.0f, 0.0f), 2.0f);
    scene.createLight(vec3(0.0f, 0.0f, -1.0f), vec3(1.0f, 0.0f, 0.0f), 2.0f);
    scene.createInstance(0, 1, vec3(0.f, -1.f, 0.f), vec3(), vec3(1.0f, 1.0f, 1.0f));  // :13-14
    scene.createInstance(0, 0, vec3(0.f, 1.f, 0.f), vec3(), vec3(4.0f, 1.0f, 4.0f));
    scene.build();                                                             // :16
    auto rtPipeline = RayTracing::Pipeline::createPipeline(device, {width, height}, scene);  // :19
    Core::Camera camera;
    camera.setView(vec3(0.0f, 0.0f, -2.0f), vec3());                           // :25

    if (post) {
      // the parts of RTApp::run the plain path above leaves out: keyboard camera (Camera::handleInputs, :39), the denoiser slot
      // (Graphics/Denoiser/Denoiser.h) and a storage image in the swapchain's 8-bit format (rebuildRenderOutput(format, extent))
      Extensions::Denoiser denoiser(device);
      std::vector<float> denoised;
      for (int frame = 0; frame < frames; ++frame) {
        camera.handleInputs(BRT_KEY_MOVE_FORWARD | BRT_KEY_LOOK_RIGHT, 1.0f / 60.0f);
        camera.setPerspectiveProjection(1.0471975512f, (float)width / (float)height, 0.001f, 100000.f);
        RayTracing::Uniform uniform = camera.uniform((uint32_t)frame, 2);
        rtPipeline->writeToUniformBuffer(&uniform, (uint32_t)frame % 2);
        rtPipeline->traceRays(width, height, 1, 1, BRT_RENDER_GBUFFER);
        denoised = denoiser.denoise(uniform, width, height);
      }
      rtPipeline->rebuildRenderOutput(BRT_FORMAT_B8G8R8A8_UNORM, {width, height});
      rtPipeline->traceRays(width, height, 1);
      const std::vector<uint8_t>& bgra = rtPipeline->getRenderOutput8();
      std::ofstream raw(prefix + ".denoised.rgba32f", std::ios::binary);
      raw.write(reinterpret_cast<const char*>(denoised.data()), (std::streamsize)(denoised.size() * sizeof(float)));
      std::ofstream raw8(prefix + ".bgra8", std::ios::binary);
      raw8.write(reinterpret_cast<const char*>(bgra.data()), (std::streamsize)bgra.size());
      const bloon::vec3 p = camera.getPosition(), r = camera.getRotation();
      std::printf("rtapp_demo post: camera %.6f %.6f %.6f rot %.6f %.6f %.6f, %zu denoised floats, %zu present bytes, denoise %.3f ms\n", p.x, p.y, p.z,
                  r.x, r.y, r.z, denoised.size(), bgra.size(), rtPipeline->getStats().ms_denoise);
      return EXIT_SUCCESS;
    }
    for (int frame = 0; frame < frames; ++frame) {                             // RTApp::run, :29-59
      camera.setPerspectiveProjection(1.0471975512f /* glm::radians(60.f) */, (float)width / (float)height, 0.001f, 100000.f);  // :41
      RayTracing::Uniform uniform = camera.uniform(/*frame = imageIndex*/ (uint32_t)frame % 2, /*depthMax*/ 2);  // :44-49
      rtPipeline->writeToUniformBuffer(&uniform, (uint32_t)frame % 2);          // :51
      if (async) rtPipeline->submitFrame((uint32_t)frame % 2, width, height);   // beginFrame (fence of the slot) ... endFrame (submit), :171-212
      else rtPipeline->traceRays(width, height, 1);                             // rayTraceScene -> traceRays, :159-160
    }
    const std::vector<float>& img = async ? rtPipeline->waitFrame((uint32_t)(frames - 1) % 2) : rtPipeline->getRenderOutput();
    std::ofstream raw(prefix + ".rgba32f", std::ios::binary);
    raw.write(reinterpret_cast<const char*>(img.data()), (std::streamsize)(img.size() * sizeof(float)));
    std::ofstream ppm(prefix + ".ppm", std::ios::binary);
    ppm << "P6\n" << width << " " << height << "\n255\n";
    for (size_t i = 0; i < (size_t)width * height; ++i)
      for (int c = 0; c < 3; ++c) {
        float v = std::pow(std::fmin(std::fmax(img[4 * i + c], 0.0f), 1.0f), 1.0f / 2.2f);
        ppm.put((char)(unsigned char)(v * 255.0f + 0.5f));
      }
    brt_stats st = rtPipeline->getStats();
    std::printf("rtapp_demo: %ux%u, %llu closest + %llu occlusion rays, %.3f ms/frame on the GPU\n", width, height,
                (unsigned long long)st.rays_closest, (unsigned long long)st.rays_occlusion, st.ms_total);
  } catch (const std::runtime_error& e) {  // HRT/main.cpp:9-13
    std::cerr << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
