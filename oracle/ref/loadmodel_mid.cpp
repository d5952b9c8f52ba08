namespace Core { struct Device {}; }
namespace RayTracing {
	struct Mesh {  // Scene.h:40-48 without the Vulkan buffers
		Core::Device& device;
		std::vector<Vertex> vertices;
		std::vector<uint32_t> indices;
	};
	class Scene {
	public:
		explicit Scene(Core::Device& d) : device(d) {}
		void loadModel(std::string path);
		Core::Device& device;
		std::vector<Mesh> meshes;
	};
}
