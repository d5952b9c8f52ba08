/* Stand-in for <vulkan/vulkan.h> (the Vulkan SDK is not in this image): only the names that
 * Graphics/Window.h, Graphics/RayTracing/MeshInstance.h and GLFW/glfw3.h (GLFW_INCLUDE_VULKAN) mention.
 * Written for oracle/ref/Makefile; declares types, calls nothing. */
#pragma once
#include <stdint.h>
typedef struct VkInstance_T* VkInstance;
typedef struct VkPhysicalDevice_T* VkPhysicalDevice;
typedef struct VkSurfaceKHR_T* VkSurfaceKHR;
typedef struct VkAllocationCallbacks VkAllocationCallbacks;
typedef int VkResult;
typedef struct VkExtent2D { uint32_t width, height; } VkExtent2D;
typedef struct VkTransformMatrixKHR { float matrix[3][4]; } VkTransformMatrixKHR;
typedef void (*PFN_vkVoidFunction)(void);
typedef PFN_vkVoidFunction (*PFN_vkGetInstanceProcAddr)(VkInstance, const char*);
