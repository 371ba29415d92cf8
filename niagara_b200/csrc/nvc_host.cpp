// nvc_host.cpp — host-side mirror of the reference's CPU code around the visibility path (pure C++, no CUDA):
//   previousPow2 / getImageMipLevels / pyramid layout   niagara.cpp:439-447,1339-1342  resources.cpp:280-292
//   PCG32 + the built-in random scene                   niagara.cpp:449-481,969-998
//   meshletVisibilityOffset prefix sum                  niagara.cpp:1002-1020
//   view / projection / frustum / lodTarget -> CullData niagara.cpp:424-437,1487-1516
// The reference does this math with glm; the few glm routines it touches are restated here (same operation
// order, so the results are bit-identical — checked against glm-generated fixtures in tests/golden/host_*.bin).
#include "../../include/niagara_cull.h"

#include <math.h>
#include <string.h>

#include <algorithm>

namespace
{

struct Pcg32
{
	uint64_t state = 0x853c49e6748fea9bULL; // PCG32_INITIALIZER niagara.cpp:455-458
	uint64_t inc = 0xda3e39cb94b95bdbULL;

	uint32_t next() // niagara.cpp:460-469 (XSH RR)
	{
		uint64_t old = state;
		state = old * 6364136223846793005ULL + (inc | 1);
		uint32_t xorshifted = uint32_t(((old >> 18u) ^ old) >> 27u);
		uint32_t rot = uint32_t(old >> 59u);
		return (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
	}

	double rand01() { return next() / double(1ull << 32); } // niagara.cpp:473-476
};

typedef float Mat4[4][4]; // [column][row], like glm

void mat4_from_quat(const float q[4], Mat4 m) // glm::mat4_cast (gtc/quaternion.inl mat3_cast)
{
	float x = q[0], y = q[1], z = q[2], w = q[3];
	float xx = x * x, yy = y * y, zz = z * z, xz = x * z, xy = x * y, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
	memset(m, 0, sizeof(Mat4));
	m[0][0] = 1.f - 2.f * (yy + zz);
	m[0][1] = 2.f * (xy + wz);
	m[0][2] = 2.f * (xz - wy);
	m[1][0] = 2.f * (xy - wz);
	m[1][1] = 1.f - 2.f * (xx + zz);
	m[1][2] = 2.f * (yz + wx);
	m[2][0] = 2.f * (xz + wy);
	m[2][1] = 2.f * (yz - wx);
	m[2][2] = 1.f - 2.f * (xx + yy);
	m[3][3] = 1.f;
}

// general 4x4 inverse by cofactors, operation order of glm::inverse (detail/func_matrix.inl compute_inverse<4,4>)
void mat4_inverse(const Mat4 m, Mat4 out)
{
	// 2x2 sub-determinants of rows {2,3}/{1,3}/{1,2} for each column pair
	float sub[6][3];
	const int colpairs[6][2] = { { 2, 3 }, { 2, 3 }, { 2, 3 }, { 2, 3 }, { 2, 3 }, { 2, 3 } };
	(void)colpairs;
	// c[k] for (rowA,rowB) pairs in the order glm names Coef00..23
	const int rows[6][2] = { { 2, 3 }, { 1, 3 }, { 1, 2 }, { 0, 3 }, { 0, 2 }, { 0, 1 } };
	for (int k = 0; k < 6; ++k)
	{
		int ra = rows[k][0], rb = rows[k][1];
		sub[k][0] = m[2][ra] * m[3][rb] - m[3][ra] * m[2][rb];
		sub[k][1] = m[1][ra] * m[3][rb] - m[3][ra] * m[1][rb];
		sub[k][2] = m[1][ra] * m[2][rb] - m[2][ra] * m[1][rb];
	}
	// Fac_k = (sub[k][0], sub[k][0], sub[k][1], sub[k][2])
	float fac[6][4];
	for (int k = 0; k < 6; ++k)
	{
		fac[k][0] = sub[k][0];
		fac[k][1] = sub[k][0];
		fac[k][2] = sub[k][1];
		fac[k][3] = sub[k][2];
	}
	float vec[4][4]; // Vec_r = (m[1][r], m[0][r], m[0][r], m[0][r])
	for (int r = 0; r < 4; ++r)
	{
		vec[r][0] = m[1][r];
		vec[r][1] = vec[r][2] = vec[r][3] = m[0][r];
	}
	float inv[4][4];
	for (int i = 0; i < 4; ++i)
	{
		inv[0][i] = (vec[1][i] * fac[0][i] - vec[2][i] * fac[1][i]) + vec[3][i] * fac[2][i];
		inv[1][i] = (vec[0][i] * fac[0][i] - vec[2][i] * fac[3][i]) + vec[3][i] * fac[4][i];
		inv[2][i] = (vec[0][i] * fac[1][i] - vec[1][i] * fac[3][i]) + vec[3][i] * fac[5][i];
		inv[3][i] = (vec[0][i] * fac[2][i] - vec[1][i] * fac[4][i]) + vec[2][i] * fac[5][i];
	}
	const float signA[4] = { 1.f, -1.f, 1.f, -1.f };
	for (int c = 0; c < 4; ++c)
		for (int i = 0; i < 4; ++i)
			inv[c][i] = inv[c][i] * ((c & 1) ? -signA[i] : signA[i]);

	float d0 = m[0][0] * inv[0][0], d1 = m[0][1] * inv[1][0], d2 = m[0][2] * inv[2][0], d3 = m[0][3] * inv[3][0];
	float det = (d0 + d1) + (d2 + d3);
	float ood = 1.f / det;
	for (int c = 0; c < 4; ++c)
		for (int i = 0; i < 4; ++i)
			out[c][i] = inv[c][i] * ood;
}

void mat4_mul(const Mat4 a, const Mat4 b, Mat4 out) // glm operator*(mat4, mat4): column c = ((a0*b[c][0] + a1*b[c][1]) + a2*b[c][2]) + a3*b[c][3]
{
	Mat4 r;
	for (int c = 0; c < 4; ++c)
		for (int i = 0; i < 4; ++i)
			r[c][i] = ((a[0][i] * b[c][0] + a[1][i] * b[c][1]) + a[2][i] * b[c][2]) + a[3][i] * b[c][3];
	memcpy(out, r, sizeof(Mat4));
}

void normalize_plane(const float p[4], float out[4]) // niagara.cpp:434-437
{
	float len = sqrtf((p[0] * p[0] + p[1] * p[1]) + p[2] * p[2]);
	for (int i = 0; i < 4; ++i)
		out[i] = p[i] / len;
}

} // namespace

extern "C"
{

NVC_API uint32_t nvc_previous_pow2(uint32_t v) // niagara.cpp:439-447
{
	uint32_t r = 1;
	while (uint64_t(r) * 2 < v)
		r *= 2;
	return r;
}

NVC_API uint32_t nvc_image_mip_levels(uint32_t width, uint32_t height) // resources.cpp:280-292
{
	uint32_t result = 1;
	while (width > 1 || height > 1)
	{
		result++;
		width /= 2;
		height /= 2;
	}
	return result;
}

NVC_API int nvc_hiz_layout(uint32_t depth_width, uint32_t depth_height, NvcHiZ* out) // niagara.cpp:1339-1342
{
	if (!out || depth_width == 0 || depth_height == 0)
		return NVC_ERROR_INVALID_ARGUMENT;
	memset(out, 0, sizeof(*out));
	out->width = nvc_previous_pow2(depth_width);
	out->height = nvc_previous_pow2(depth_height);
	out->levels = nvc_image_mip_levels(out->width, out->height);
	if (out->levels > NVC_MAX_HIZ_LEVELS)
		return NVC_ERROR_INVALID_ARGUMENT;
	uint64_t offset = 0;
	for (uint32_t l = 0; l < out->levels; ++l)
	{
		out->level_offset[l] = uint32_t(offset);
		offset += uint64_t(std::max(1u, out->width >> l)) * std::max(1u, out->height >> l);
	}
	if (offset > 0xffffffffull)
		return NVC_ERROR_INVALID_ARGUMENT;
	out->total_texels = uint32_t(offset);
	return NVC_OK;
}

NVC_API void nvc_host_random_draws(NvcMeshDraw* draws, uint32_t draw_count, uint32_t mesh_count, float scene_radius) // niagara.cpp:969-998
{
	Pcg32 rng;
	rng.state = 0x42;

	for (uint32_t i = 0; i < draw_count; ++i)
	{
		NvcMeshDraw& draw = draws[i];
		memset(&draw, 0, sizeof(draw));

		uint32_t meshIndex = rng.next() % mesh_count;

		draw.position[0] = float(rng.rand01()) * scene_radius * 2 - scene_radius;
		draw.position[1] = float(rng.rand01()) * scene_radius * 2 - scene_radius;
		draw.position[2] = float(rng.rand01()) * scene_radius * 2 - scene_radius;
		draw.scale = float(rng.rand01()) + 1;
		draw.scale *= 2;

		// vec3(rand, rand, rand): C++ leaves the evaluation order of the three arguments unspecified; GCC (the
		// toolchain the reference's CI and this image use) evaluates them right to left, so z is drawn first.
		// Verified against the glm/GCC-built fixture tests/golden/host_golden.nvch.
		float az = float(rng.rand01()) * 2 - 1;
		float ay = float(rng.rand01()) * 2 - 1;
		float ax = float(rng.rand01()) * 2 - 1;
		// glm::normalize = v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x)
		float inv = 1.f / sqrtf((ax * ax + ay * ay) + az * az);
		ax *= inv, ay *= inv, az *= inv;
		float angle = (float(rng.rand01()) * 90.f) * 0.01745329251994329576923690768489f; // glm::radians

		float s = sinf(angle * 0.5f);
		draw.orientation[0] = ax * s;
		draw.orientation[1] = ay * s;
		draw.orientation[2] = az * s;
		draw.orientation[3] = cosf(angle * 0.5f);

		draw.meshIndex = meshIndex;
	}
}

NVC_API uint32_t nvc_host_visibility_offsets(NvcMeshDraw* draws, uint32_t draw_count, const NvcMesh* meshes, uint32_t* post_pass_mask) // niagara.cpp:1002-1020
{
	uint32_t meshletVisibilityCount = 0;
	uint32_t meshPostPasses = 0;
	for (uint32_t i = 0; i < draw_count; ++i)
	{
		NvcMeshDraw& draw = draws[i];
		const NvcMesh& mesh = meshes[draw.meshIndex];

		draw.meshletVisibilityOffset = meshletVisibilityCount;

		uint32_t meshletCount = 0;
		for (uint32_t l = 0; l < mesh.lodCount; ++l)
			meshletCount = std::max(meshletCount, mesh.lods[l].meshletCount);

		meshletVisibilityCount += meshletCount;
		meshPostPasses |= 1u << draw.postPass;
	}
	if (post_pass_mask)
		*post_pass_mask = meshPostPasses;
	return meshletVisibilityCount;
}

NVC_API void nvc_host_cull_data(const NvcCamera* camera, uint32_t screen_width, uint32_t screen_height,
    uint32_t draw_count, const NvcCullOptions* options, NvcCullData* out, float* out_projection16)
{
	// niagara.cpp:1487-1490
	Mat4 cam, view, flip;
	mat4_from_quat(camera->orientation, cam);
	cam[3][0] = camera->position[0];
	cam[3][1] = camera->position[1];
	cam[3][2] = camera->position[2];
	cam[3][3] = 1.f;
	mat4_inverse(cam, view);
	memset(flip, 0, sizeof(flip));
	// glm::scale(identity, vec3(1, 1, -1)): column 2 is (0, 0, 1, 0) * -1, i.e. its zeros are negative zeros
	flip[0][0] = 1.f, flip[1][1] = 1.f, flip[3][3] = 1.f;
	flip[2][0] = -0.f, flip[2][1] = -0.f, flip[2][2] = -1.f, flip[2][3] = -0.f;
	mat4_mul(flip, view, view);

	// niagara.cpp:424-432 perspectiveProjection(fovY, aspect, znear): infinite far plane, reverse Z
	float aspect = float(screen_width) / float(screen_height);
	float f = 1.0f / tanf(camera->fovY / 2.0f);
	Mat4 proj;
	memset(proj, 0, sizeof(proj));
	proj[0][0] = f / aspect;
	proj[1][1] = f;
	proj[2][3] = 1.0f;
	proj[3][2] = camera->znear;
	if (out_projection16)
		memcpy(out_projection16, proj, sizeof(proj));

	// niagara.cpp:1494-1497: rows of the projection (columns of its transpose)
	float row0[4] = { proj[0][0], proj[1][0], proj[2][0], proj[3][0] };
	float row1[4] = { proj[0][1], proj[1][1], proj[2][1], proj[3][1] };
	float row3[4] = { proj[0][3], proj[1][3], proj[2][3], proj[3][3] };
	float px[4], py[4], frustumX[4], frustumY[4];
	for (int i = 0; i < 4; ++i)
	{
		px[i] = row3[i] + row0[i];
		py[i] = row3[i] + row1[i];
	}
	normalize_plane(px, frustumX);
	normalize_plane(py, frustumY);

	NvcCullData cd;
	memset(&cd, 0, sizeof(cd));
	memcpy(cd.view, view, sizeof(view));
	cd.P00 = proj[0][0];
	cd.P11 = proj[1][1];
	cd.znear = camera->znear;
	cd.zfar = options->draw_distance;
	cd.frustum[0] = frustumX[0];
	cd.frustum[1] = frustumX[2];
	cd.frustum[2] = frustumY[1];
	cd.frustum[3] = frustumY[2];
	cd.drawCount = draw_count;
	cd.cullingEnabled = options->culling;
	cd.lodEnabled = options->lod;
	cd.occlusionEnabled = options->occlusion;
	cd.lodTarget = (2 / cd.P11) * (1.f / float(screen_height)) * (1 << options->debug_lod_step); // 1px
	cd.pyramidWidth = float(nvc_previous_pow2(screen_width));
	cd.pyramidHeight = float(nvc_previous_pow2(screen_height));
	cd.clusterOcclusionEnabled = options->occlusion && options->cluster_occlusion && options->mesh_shading;
	*out = cd;
}

NVC_API void nvc_host_pass_data(const NvcCullData* frame, int for_drawcull, uint32_t post_pass, NvcCullData* out)
{
	NvcCullData pass = *frame;
	if (for_drawcull)
		pass.clusterBackfaceEnabled = post_pass == 0; // niagara.cpp:1549 (only the drawcull pass data gets it — SURVEY F8)
	pass.postPass = post_pass;                          // niagara.cpp:1550,1596
	*out = pass;
}

// ---- N3: keyframe evaluation of the frame loop, niagara.cpp:1362-1390 (draw branch) ---------------------------------
// glm::mix(x, y, a) = x * (1 - a) + y * a (detail/func_common.inl:104-112); glm::slerp (ext/quaternion_common.inl:41-73):
// dot = (w*w' + x*x') + (y*y' + z*z'), negate y when dot < 0, lerp when dot > 1 - FLT_EPSILON, else
// (sin((1 - a) * angle) * x + sin(a * angle) * z) / sin(angle) with angle = acos(dot) — same libm calls as glm makes.
static inline float mix1(float x, float y, float a) { return x * (1.0f - a) + y * a; }

NVC_API int nvc_host_animate(const NvcAnimation* animations, uint32_t animation_count, const NvcKeyframe* keyframes, uint32_t keyframe_count,
    double animation_time, NvcMeshDraw* draws, uint32_t draw_count, uint32_t* update_indices, NvcMeshDraw* update_values, uint32_t max_updates)
{
	if ((animation_count && (!animations || !keyframes)) || !draws)
		return NVC_ERROR_INVALID_ARGUMENT;
	uint32_t updates = 0;
	for (uint32_t ai = 0; ai < animation_count; ++ai)
	{
		const NvcAnimation& animation = animations[ai];
		if (animation.keyframeCount == 0 || uint64_t(animation.keyframeOffset) + animation.keyframeCount > keyframe_count)
			return NVC_ERROR_INVALID_ARGUMENT;

		double index = (animation_time - animation.startTime) / animation.period; // :1368
		if (index < 0)
			continue;
		index = fmod(index, double(animation.keyframeCount)); // :1373

		uint32_t index0 = uint32_t(int(index)) % animation.keyframeCount; // :1375-1376
		uint32_t index1 = (index0 + 1) % animation.keyframeCount;
		float t = float(index - floor(index)); // :1378, used as float(t)

		const NvcKeyframe& k0 = keyframes[animation.keyframeOffset + index0];
		const NvcKeyframe& k1 = keyframes[animation.keyframeOffset + index1];

		if (animation.drawIndex < 0)
			continue; // light animations (:1403-1410) are not part of the visibility path
		if (uint32_t(animation.drawIndex) >= draw_count)
			return NVC_ERROR_INVALID_ARGUMENT;

		NvcMeshDraw& draw = draws[animation.drawIndex];
		for (int c = 0; c < 3; ++c)
			draw.position[c] = mix1(k0.translation[c], k1.translation[c], t); // :1386
		draw.scale = mix1(k0.scale, k1.scale, t);                             // :1387

		// :1388 glm::slerp(keyframe0.rotation, keyframe1.rotation, float(t)); rotation[] = x, y, z, w
		const float* x = k0.rotation;
		float z[4] = { k1.rotation[0], k1.rotation[1], k1.rotation[2], k1.rotation[3] };
		float cosTheta = (x[3] * z[3] + x[0] * z[0]) + (x[1] * z[1] + x[2] * z[2]);
		if (cosTheta < 0.0f)
		{
			for (int c = 0; c < 4; ++c)
				z[c] = -z[c];
			cosTheta = -cosTheta;
		}
		if (cosTheta > 1.0f - 1.1920928955078125e-07f)
		{
			for (int c = 0; c < 4; ++c)
				draw.orientation[c] = mix1(x[c], z[c], t);
		}
		else
		{
			float angle = acosf(cosTheta);
			float s0 = sinf((1.0f - t) * angle), s1 = sinf(t * angle), sd = sinf(angle);
			for (int c = 0; c < 4; ++c)
				draw.orientation[c] = (x[c] * s0 + z[c] * s1) / sd;
		}

		if (update_indices && update_values)
		{
			if (updates >= max_updates)
				return NVC_ERROR_INVALID_ARGUMENT;
			update_indices[updates] = uint32_t(animation.drawIndex);
			update_values[updates] = draw;
		}
		++updates;
	}
	return int(updates);
}

} // extern "C"
