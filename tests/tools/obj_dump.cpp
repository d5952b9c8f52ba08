// Test helper: runs the facade's OBJ ingestion (bloon::obj::load, what RayTracing::Scene::loadModel uploads) on every path
// given and writes <path>.bin = { u32 n_vertices, u32 n_indices, vertices (32 B each), indices } or <path>.err = message.
#include <cstdio>
#include "bloon/bloon.hpp"

int main(int argc, char** argv) {
  for (int a = 1; a < argc; ++a) {
    const std::string path = argv[a];
    try {
      const bloon::obj::Mesh m = bloon::obj::load(path);
      FILE* f = std::fopen((path + ".bin").c_str(), "wb");
      const uint32_t n[2] = {(uint32_t)m.vertices.size(), (uint32_t)m.indices.size()};
      std::fwrite(n, 4, 2, f);
      std::fwrite(m.vertices.data(), sizeof(brt_vertex), m.vertices.size(), f);
      std::fwrite(m.indices.data(), 4, m.indices.size(), f);
      std::fclose(f);
    } catch (const std::exception& e) {
      FILE* f = std::fopen((path + ".err").c_str(), "w");
      std::fputs(e.what(), f);
      std::fclose(f);
    }
  }
  return 0;
}
