// TEST INFRASTRUCTURE: produces golden vectors for the host-side helpers (niagara_b200/csrc/nvc_host.cpp) with the
// reference's own math library (glm, compiled from /root/reference/extern/glm with the reference's defines).
// niagara.cpp's main() cannot be compiled here (Vulkan/GLFW), so the few lines involved are restated verbatim in
// spirit: PCG32 + random scene (niagara.cpp:449-481, 969-998) and the CullData fill (niagara.cpp:424-437, 1487-1516).
//
// Output "NVCH" v1: u32 magic, u32 version, u32 drawCountA, u32 meshCountA, u32 drawCountB, u32 meshCountB, u32 cameraCount, u32 pad
//   MeshDraw[drawCountA], MeshDraw[drawCountB], then per camera: {float pos[3], quat xyzw[4], fovY, znear, u32 w, u32 h, u32 drawCount, u32 lodStep} + CullData(144 B)
#include "math.h"

#include <glm/gtc/matrix_transform.hpp>

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

struct MeshDraw
{
	vec3 position;
	float scale;
	quat orientation;
	uint32_t meshIndex, meshletVisibilityOffset, postPass, materialIndex;
};

struct alignas(16) CullData
{
	mat4 view;
	float P00, P11, znear, zfar;
	float frustum[4];
	float lodTarget;
	float pyramidWidth, pyramidHeight;
	uint32_t drawCount;
	int cullingEnabled, lodEnabled, occlusionEnabled, clusterOcclusionEnabled, clusterBackfaceEnabled;
	uint32_t postPass;
};
static_assert(sizeof(CullData) == 144, "CullData");
static_assert(sizeof(MeshDraw) == 48, "MeshDraw");

struct pcg32_random_t
{
	uint64_t state, inc;
};
static pcg32_random_t rngstate = { 0x853c49e6748fea9bULL, 0xda3e39cb94b95bdbULL };

static uint32_t pcg32_random_r(pcg32_random_t* rng)
{
	uint64_t oldstate = rng->state;
	rng->state = oldstate * 6364136223846793005ULL + (rng->inc | 1);
	uint32_t xorshifted = uint32_t(((oldstate >> 18u) ^ oldstate) >> 27u);
	uint32_t rot = oldstate >> 59u;
	return (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
}
static double rand01() { return pcg32_random_r(&rngstate) / double(1ull << 32); }
static uint32_t rand32() { return pcg32_random_r(&rngstate); }

static std::vector<MeshDraw> randomScene(uint32_t drawCount, size_t meshCount)
{
	rngstate.state = 0x42;
	std::vector<MeshDraw> draws(drawCount);
	float sceneRadius = 300;
	for (uint32_t i = 0; i < drawCount; ++i)
	{
		MeshDraw& draw = draws[i];
		memset(&draw, 0, sizeof(draw));
		size_t meshIndex = rand32() % meshCount;
		draw.position[0] = float(rand01()) * sceneRadius * 2 - sceneRadius;
		draw.position[1] = float(rand01()) * sceneRadius * 2 - sceneRadius;
		draw.position[2] = float(rand01()) * sceneRadius * 2 - sceneRadius;
		draw.scale = float(rand01()) + 1;
		draw.scale *= 2;
		vec3 axis = normalize(vec3(float(rand01()) * 2 - 1, float(rand01()) * 2 - 1, float(rand01()) * 2 - 1));
		float angle = glm::radians(float(rand01()) * 90.f);
		draw.orientation = quat(cosf(angle * 0.5f), axis * sinf(angle * 0.5f));
		draw.meshIndex = uint32_t(meshIndex);
	}
	return draws;
}

static mat4 perspectiveProjection(float fovY, float aspectWbyH, float zNear)
{
	float f = 1.0f / tanf(fovY / 2.0f);
	return mat4(f / aspectWbyH, 0.0f, 0.0f, 0.0f, 0.0f, f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, zNear, 0.0f);
}
static vec4 normalizePlane(vec4 p) { return p / length(vec3(p)); }
static uint32_t previousPow2(uint32_t v)
{
	uint32_t r = 1;
	while (r * 2 < v)
		r *= 2;
	return r;
}

struct CameraCase
{
	float pos[3];
	float q[4]; // xyzw
	float fovY, znear;
	uint32_t width, height, drawCount, lodStep;
};

static CullData fillCullData(const CameraCase& c)
{
	quat orientation(c.q[3], c.q[0], c.q[1], c.q[2]); // (w, x, y, z) scalar-first constructor is unaffected by GLM_FORCE_QUAT_CTOR_XYZW? no: use named init below
	orientation.x = c.q[0], orientation.y = c.q[1], orientation.z = c.q[2], orientation.w = c.q[3];
	mat4 view = glm::mat4_cast(orientation);
	view[3] = vec4(vec3(c.pos[0], c.pos[1], c.pos[2]), 1.0f);
	view = inverse(view);
	view = glm::scale(glm::identity<glm::mat4>(), vec3(1, 1, -1)) * view;

	mat4 projection = perspectiveProjection(c.fovY, float(c.width) / float(c.height), c.znear);
	mat4 projectionT = transpose(projection);
	vec4 frustumX = normalizePlane(projectionT[3] + projectionT[0]);
	vec4 frustumY = normalizePlane(projectionT[3] + projectionT[1]);

	CullData cullData = {};
	cullData.view = view;
	cullData.P00 = projection[0][0];
	cullData.P11 = projection[1][1];
	cullData.znear = c.znear;
	cullData.zfar = 200;
	cullData.frustum[0] = frustumX.x;
	cullData.frustum[1] = frustumX.z;
	cullData.frustum[2] = frustumY.y;
	cullData.frustum[3] = frustumY.z;
	cullData.drawCount = c.drawCount;
	cullData.cullingEnabled = 1;
	cullData.lodEnabled = 1;
	cullData.occlusionEnabled = 1;
	cullData.lodTarget = (2 / cullData.P11) * (1.f / float(c.height)) * (1 << c.lodStep);
	cullData.pyramidWidth = float(previousPow2(c.width));
	cullData.pyramidHeight = float(previousPow2(c.height));
	cullData.clusterOcclusionEnabled = 1;
	return cullData;
}

int main(int argc, char** argv)
{
	if (argc < 2)
		return 2;
	std::vector<MeshDraw> a = randomScene(4096, 1), b = randomScene(2048, 7);
	CameraCase cams[] = {
		{ { 0, 0, 0 }, { 0, 0, 0, 1 }, glm::radians(70.f), 0.1f, 1024, 768, 4096, 0 },
		{ { 10.5f, -3.25f, 42.f }, { 0.1825742f, 0.3651484f, 0.5477226f, 0.7302967f }, glm::radians(70.f), 0.1f, 1920, 1080, 1000000, 0 },
		{ { -120.f, 60.f, -250.f }, { -0.3f, 0.1f, 0.2f, 0.9273618f }, glm::radians(50.f), 0.5f, 4096, 4096, 1000000, 2 },
		{ { 1.f, 2.f, 3.f }, { 0.f, 0.7071068f, 0.f, 0.7071068f }, glm::radians(90.f), 1.f, 2560, 1440, 12345, 1 },
	};
	uint32_t ncam = sizeof(cams) / sizeof(cams[0]);
	FILE* f = fopen(argv[1], "wb");
	if (!f)
		return 1;
	uint32_t header[8] = { 0x4843564eu, 1u, uint32_t(a.size()), 1u, uint32_t(b.size()), 7u, ncam, 0 };
	fwrite(header, sizeof(header), 1, f);
	fwrite(a.data(), sizeof(MeshDraw), a.size(), f);
	fwrite(b.data(), sizeof(MeshDraw), b.size(), f);
	for (uint32_t i = 0; i < ncam; ++i)
	{
		CullData cd = fillCullData(cams[i]);
		fwrite(&cams[i], sizeof(CameraCase), 1, f);
		fwrite(&cd, sizeof(cd), 1, f);
	}
	fclose(f);
	printf("%s: %zu + %zu draws, %u cameras\n", argv[1], a.size(), b.size(), ncam);
	return 0;
}
