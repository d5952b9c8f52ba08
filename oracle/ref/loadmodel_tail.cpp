// out_v: 8 floats per vertex (pos, normal, uv = RayTracing::Vertex, 32 bytes); buffers are malloc'd, release with ref_free.
// Returns 0, or 1 with the exception's text in err when the reference throws (tinyobj error, RT/Scene.cpp:38-41).
extern "C" int ref_load_model(const char* path, float** out_v, uint32_t* nv, uint32_t** out_i, uint32_t* ni, char* err, uint32_t errcap) {
	static_assert(sizeof(RayTracing::Vertex) == 32, "Vertex layout");
	Core::Device dev;
	RayTracing::Scene scene(dev);
	std::streambuf* keep = std::cout.rdbuf(nullptr);  // loadModel prints its error before throwing
	try {
		scene.loadModel(path);
	} catch (const std::exception& e) {
		std::cout.rdbuf(keep);
		if (err && errcap) { std::strncpy(err, e.what(), errcap - 1); err[errcap - 1] = 0; }
		return 1;
	}
	std::cout.rdbuf(keep);
	const RayTracing::Mesh& m = scene.meshes.back();
	*nv = (uint32_t)m.vertices.size();
	*ni = (uint32_t)m.indices.size();
	*out_v = (float*)std::malloc(m.vertices.size() * 32 + 1);
	*out_i = (uint32_t*)std::malloc(m.indices.size() * 4 + 1);
	std::memcpy(*out_v, m.vertices.data(), m.vertices.size() * 32);
	std::memcpy(*out_i, m.indices.data(), m.indices.size() * 4);
	return 0;
}
extern "C" void ref_free(void* p) { std::free(p); }
