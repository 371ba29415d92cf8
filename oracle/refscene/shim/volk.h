/* Shim so that the reference's src/scene.cpp (which includes common.h -> <volk.h>) compiles
 * without a Vulkan SDK.  scene.cpp itself uses no Vulkan entry points; textures.h only
 * *declares* a function taking these handle types.  Test infrastructure only. */
#pragma once
typedef struct VkDevice_T* VkDevice;
typedef struct VkCommandPool_T* VkCommandPool;
typedef struct VkCommandBuffer_T* VkCommandBuffer;
typedef struct VkQueue_T* VkQueue;
typedef int VkResult;
#define VK_SUCCESS 0
#define VK_SUBOPTIMAL_KHR 1
#define VK_ERROR_OUT_OF_DATE_KHR 2
#define VK_NOT_READY 3
struct VkPhysicalDeviceMemoryProperties;
