// TEST INFRASTRUCTURE: the glm-based golden generators shared by host_golden.cpp (committed fixture) and anim_ref.cpp (property
// tests): what niagara.cpp's frame loop computes on the host for the visibility path, around the reference's own math library.
#pragma once
#include "math.h"

#include <glm/gtc/matrix_transform.hpp>

#include <math.h>
#include <stdint.h>
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

// PCG32 (pcg-random.org minimal generator) with the stream constant niagara.cpp:449-466 seeds it with
struct Pcg32
{
	uint64_t state = 0x853c49e6748fea9bULL;
	uint64_t stream = 0xda3e39cb94b95bdbULL;

	uint32_t next()
	{
		uint64_t s0 = state;
		state = s0 * 6364136223846793005ULL + (stream | 1);
		uint32_t x = uint32_t(((s0 >> 18u) ^ s0) >> 27u), r = uint32_t(s0 >> 59u);
		return (x >> r) | (x << ((32 - r) & 31));
	}
	double unit() { return next() / double(1ull << 32); }
};
static Pcg32 g_rng;
static double rand01() { return g_rng.unit(); }

// The random scene of niagara.cpp:969-998: per draw one mesh pick, a position in a 600^3 cube, scale in [2, 4), a rotation
// of up to 90 degrees about a random axis.  The three-argument vec3(...) below must stay ONE expression: the reference
// relies on the compiler's argument evaluation order there (GCC: right to left), and so does the golden.
static std::vector<MeshDraw> randomScene(uint32_t count, size_t meshes)
{
	g_rng.state = 0x42;
	const float extent = 300;
	std::vector<MeshDraw> out(count);
	for (MeshDraw& d : out)
	{
		memset(&d, 0, sizeof(d));
		d.meshIndex = uint32_t(g_rng.next() % meshes);
		for (int c = 0; c < 3; ++c)
			d.position[c] = float(rand01()) * extent * 2 - extent;
		d.scale = float(rand01()) + 1;
		d.scale *= 2;
		vec3 axis = normalize(vec3(float(rand01()) * 2 - 1, float(rand01()) * 2 - 1, float(rand01()) * 2 - 1));
		float half = glm::radians(float(rand01()) * 90.f) * 0.5f;
		d.orientation = quat(cosf(half), axis * sinf(half));
	}
	return out;
}

struct CameraCase
{
	float pos[3];
	float q[4]; // xyzw
	float fovY, znear;
	uint32_t width, height, drawCount, lodStep;
};

// CullData as the frame loop fills it (niagara.cpp:1487-1516) from: the view matrix (camera transform inverted, Z flipped),
// the infinite reverse-Z projection (niagara.cpp:424-432) and its two symmetric frustum planes — all through glm.
static CullData fillCullData(const CameraCase& c)
{
	quat rotation;
	rotation.x = c.q[0], rotation.y = c.q[1], rotation.z = c.q[2], rotation.w = c.q[3];
	mat4 camera = glm::mat4_cast(rotation);
	camera[3] = vec4(vec3(c.pos[0], c.pos[1], c.pos[2]), 1.0f);
	mat4 view = glm::scale(glm::identity<glm::mat4>(), vec3(1, 1, -1)) * inverse(camera);

	float f = 1.0f / tanf(c.fovY / 2.0f), aspect = float(c.width) / float(c.height);
	mat4 projection(f / aspect, 0.0f, 0.0f, 0.0f, 0.0f, f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, c.znear, 0.0f);
	mat4 rows = transpose(projection);
	vec4 planeX = rows[3] + rows[0], planeY = rows[3] + rows[1];
	planeX = planeX / length(vec3(planeX));
	planeY = planeY / length(vec3(planeY));

	auto pow2_below = [](uint32_t v) {
		uint32_t r = 1;
		while (r * 2 < v)
			r *= 2;
		return r;
	};

	CullData cd = {};
	cd.view = view;
	cd.P00 = projection[0][0];
	cd.P11 = projection[1][1];
	cd.znear = c.znear;
	cd.zfar = 200;
	cd.frustum[0] = planeX.x, cd.frustum[1] = planeX.z;
	cd.frustum[2] = planeY.y, cd.frustum[3] = planeY.z;
	cd.drawCount = c.drawCount;
	cd.cullingEnabled = cd.lodEnabled = cd.occlusionEnabled = cd.clusterOcclusionEnabled = 1;
	cd.lodTarget = (2 / cd.P11) * (1.f / float(c.height)) * (1 << c.lodStep);
	cd.pyramidWidth = float(pow2_below(c.width));
	cd.pyramidHeight = float(pow2_below(c.height));
	return cd;
}

