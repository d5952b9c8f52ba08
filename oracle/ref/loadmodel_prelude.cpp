// Head of the generated translation unit oracle/_ref/gen/loadmodel_tu.cpp (see oracle/ref/Makefile). The Makefile appends, by line
// range and straight from /root/reference (never copied into this repository):
//   Graphics/RayTracing/Scene.h:20-38    hashCombine, struct Vertex (+ the closing brace of the namespace)
//   [loadmodel_mid.cpp]                  stand-ins for Core::Device, Mesh and the three Scene members loadModel touches
//   Graphics/RayTracing/Scene.cpp:5-14   std::hash<Vertex>
//   Graphics/RayTracing/Scene.cpp:29-74  Scene::loadModel
//   [loadmodel_tail.cpp]                 the C entry point
// so the OBJ parser (libs/tinyobj/tiny_obj_loader.h 1.0.6), the Y flip and the de-duplication that run are the reference's.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#define TINYOBJLOADER_IMPLEMENTATION
#include "libs/tinyobj/tiny_obj_loader.h"
