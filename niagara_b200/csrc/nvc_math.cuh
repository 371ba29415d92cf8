// nvc_math.cuh — device-side restatement of src/shaders/math.h:1-49 and the fixed-function pieces the GLSL
// relies on (fp16/s8 decode, MIN-reduction sampler), written so that every result is BIT-IDENTICAL to the
// strict-IEEE interpretation documented in DESIGN.md ("arithmetic contract"):
//   * no FMA contraction: this translation unit is compiled with -fmad=false, and every multiply/add below is a
//     separate correctly-rounded binary32 operation in GLSL source order;
//   * division and square root are the correctly rounded __fdiv_rn / __fsqrt_rn;
//   * x / 127.0 (s8 decode) is evaluated with an FMA residual correction that is proven bit-identical to the
//     true quotient for all 256 inputs (tests/test_layout_and_math.py::test_div127_exact);
//   * ceil(log2(x)) comes from the float's exponent/mantissa, exp2(int) is an exact power of two.
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/niagara_cull.h"

namespace nvc
{

struct f3
{
	float x, y, z;
};

__device__ __forceinline__ float half_bits_to_float(uint32_t h16)
{
	return __half2float(__ushort_as_half((unsigned short)h16)); // exact
}

// true quotient float(i) / 127.0f for i in [-128, 127]; q0 = i * fl(1/127) is within 1 ulp, one FMA residual step
// yields the correctly rounded quotient (Markstein); explicit __fmaf_rn is not affected by -fmad=false.
__device__ __forceinline__ float s8_div127(int i)
{
	const float r = 0.00787401574803149606f; // fl(1/127)
	float a = (float)i;
	float q0 = __fmul_rn(a, r);
	float rem = __fmaf_rn(-q0, 127.0f, a);
	return __fmaf_rn(rem, r, q0);
}

// GLSL cross(): [x1*y2 - y1*x2, x2*y0 - y2*x0, x0*y1 - y0*x1]
__device__ __forceinline__ f3 cross3(f3 a, f3 b)
{
	f3 r;
	r.x = __fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(b.y, a.z));
	r.y = __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(b.z, a.x));
	r.z = __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(b.x, a.y));
	return r;
}

// src/shaders/math.h:46-49
__device__ __forceinline__ f3 rotate_quat(f3 v, float4 q)
{
	f3 qv = { q.x, q.y, q.z };
	f3 c1 = cross3(qv, v);
	f3 t = { __fadd_rn(c1.x, __fmul_rn(q.w, v.x)), __fadd_rn(c1.y, __fmul_rn(q.w, v.y)), __fadd_rn(c1.z, __fmul_rn(q.w, v.z)) };
	f3 c2 = cross3(qv, t);
	f3 r = { __fadd_rn(v.x, __fmul_rn(2.0f, c2.x)), __fadd_rn(v.y, __fmul_rn(2.0f, c2.y)), __fadd_rn(v.z, __fmul_rn(2.0f, c2.z)) };
	return r;
}

// (view * vec4(p, 1)).xyz, column-major, left-to-right
__device__ __forceinline__ f3 transform_point(const float* __restrict__ m, f3 p)
{
	f3 r;
	r.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], p.x), __fmul_rn(m[4], p.y)), __fmul_rn(m[8], p.z)), m[12]);
	r.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[1], p.x), __fmul_rn(m[5], p.y)), __fmul_rn(m[9], p.z)), m[13]);
	r.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[2], p.x), __fmul_rn(m[6], p.y)), __fmul_rn(m[10], p.z)), m[14]);
	return r;
}

// mat3(view) * v
__device__ __forceinline__ f3 transform_vector(const float* __restrict__ m, f3 v)
{
	f3 r;
	r.x = __fadd_rn(__fadd_rn(__fmul_rn(m[0], v.x), __fmul_rn(m[4], v.y)), __fmul_rn(m[8], v.z));
	r.y = __fadd_rn(__fadd_rn(__fmul_rn(m[1], v.x), __fmul_rn(m[5], v.y)), __fmul_rn(m[9], v.z));
	r.z = __fadd_rn(__fadd_rn(__fmul_rn(m[2], v.x), __fmul_rn(m[6], v.y)), __fmul_rn(m[10], v.z));
	return r;
}

__device__ __forceinline__ float dot3(f3 a, f3 b)
{
	return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}

__device__ __forceinline__ float length3(f3 a)
{
	return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)), __fmul_rn(a.z, a.z)));
}

// drawcull.comp.glsl:77-83 == clustercull.comp.glsl:104-108
template <typename CD>
__device__ __forceinline__ bool frustum_visible(const CD& cd, f3 c, float radius)
{
	bool visible = __fsub_rn(__fmul_rn(c.z, cd.frustum[1]), __fmul_rn(fabsf(c.x), cd.frustum[0])) > -radius;
	visible = visible && __fsub_rn(__fmul_rn(c.z, cd.frustum[3]), __fmul_rn(fabsf(c.y), cd.frustum[2])) > -radius;
	visible = visible && __fadd_rn(c.z, radius) > cd.znear && __fsub_rn(c.z, radius) < cd.zfar;
	return visible;
}

// src/shaders/math.h:1-22.  Branch-free: the aabb is always computed (garbage, possibly NaN, when the sphere crosses
// the near plane) and the early `return false` becomes the returned flag, so that a warp runs one straight path.
__device__ __forceinline__ bool project_sphere(f3 c, float r, float znear, float P00, float P11, float4& aabb)
{
	bool ok = !(c.z < __fadd_rn(r, znear));

	float crx = __fmul_rn(c.x, r), cry = __fmul_rn(c.y, r), crz = __fmul_rn(c.z, r);
	float czr2 = __fsub_rn(__fmul_rn(c.z, c.z), __fmul_rn(r, r));

	float vx = __fsqrt_rn(__fadd_rn(__fmul_rn(c.x, c.x), czr2));
	float vxx = __fmul_rn(vx, c.x), vxz = __fmul_rn(vx, c.z);
	float minx = __fdiv_rn(__fsub_rn(vxx, crz), __fadd_rn(vxz, crx));
	float maxx = __fdiv_rn(__fadd_rn(vxx, crz), __fsub_rn(vxz, crx));

	float vy = __fsqrt_rn(__fadd_rn(__fmul_rn(c.y, c.y), czr2));
	float vyy = __fmul_rn(vy, c.y), vyz = __fmul_rn(vy, c.z);
	float miny = __fdiv_rn(__fsub_rn(vyy, crz), __fadd_rn(vyz, cry));
	float maxy = __fdiv_rn(__fadd_rn(vyy, crz), __fsub_rn(vyz, cry));

	aabb.x = __fadd_rn(__fmul_rn(__fmul_rn(minx, P00), 0.5f), 0.5f);
	aabb.y = __fadd_rn(__fmul_rn(__fmul_rn(maxy, P11), -0.5f), 0.5f);
	aabb.z = __fadd_rn(__fmul_rn(__fmul_rn(maxx, P00), 0.5f), 0.5f);
	aabb.w = __fadd_rn(__fmul_rn(__fmul_rn(miny, P11), -0.5f), 0.5f);
	return ok;
}

// src/shaders/math.h:24-39; returns an integer level already clamped to [0, max_level].
// level = ceil(log2(m)) with m = max(size.x * pw, size.y * ph), minus one if the finer mip still fits 2x2, max(., 0).
// ceil(log2 m) <= 0 for every m <= 1 (and log2 of a non-positive / NaN m is -inf / NaN, max(level, 0) = 0), so
// anything but m > 1 yields level 0; for m > 1 the float is normal and L = exponent + (mantissa != 0).
__device__ __forceinline__ int occlusion_mip(float4 aabb, float pw, float ph, int max_level)
{
	float sizex = __fsub_rn(aabb.z, aabb.x);
	float sizey = __fsub_rn(aabb.w, aabb.y);
	// GLSL max(x, y) = (x < y) ? y : x; fmaxf agrees for ordered operands and keeps the defined operand when exactly
	// one is NaN (the oracle's rule)
	float m = fmaxf(__fmul_rn(sizex, pw), __fmul_rn(sizey, ph));
	if (!(m > 1.0f))
		return 0;
	uint32_t bits = __float_as_uint(m);
	if (bits == 0x7f800000u)
		return max_level;
	int L = int(bits >> 23) - 127 + ((bits & 0x7fffffu) ? 1 : 0); // 1..128

	// exp2(1 - L): exact power of two built from exponent bits (subnormal 2^-127 when L == 128)
	float scale = (L <= 127) ? __uint_as_float(uint32_t(128 - L) << 23) : __uint_as_float(0x00400000u);
	float fmx = __fmul_rn(pw, scale);
	float fmy = __fmul_rn(ph, scale);
	float px = __fmul_rn(aabb.x, fmx);
	float py = __fmul_rn(aabb.y, fmy);
	float fx = __fsub_rn(px, floorf(px));
	float fy = __fsub_rn(py, floorf(py));
	bool fits = (__fadd_rn(fx, __fmul_rn(sizex, fmx)) <= 2.0f) && (__fadd_rn(fy, __fmul_rn(sizey, fmy)) <= 2.0f);
	L -= fits ? 1 : 0;
	return min(L, max_level);
}

// MIN-reduction bilinear footprint on one level (resources.cpp:294-325 sampler): indices + which texels count
struct Footprint
{
	uint32_t x0, x1, y0, y1;
	bool usex1, usey1;
};

// x = u * w - 0.5; texels floor(x) and floor(x) + 1, clamped to the edge; the second one only counts when its weight
// fract(x) is non-zero.  Non-finite coordinates need no special case: a NaN clamps both indices to 0 and +-inf clamps
// both to the same edge texel, so whether the second texel "counts" cannot change the minimum.
__device__ __forceinline__ Footprint min_footprint(uint32_t w, uint32_t h, float u, float v)
{
	float x = __fsub_rn(__fmul_rn(u, (float)w), 0.5f);
	float y = __fsub_rn(__fmul_rn(v, (float)h), 0.5f);
	float fx0 = floorf(x), fy0 = floorf(y);
	float wmax = (float)(w - 1), hmax = (float)(h - 1);
	Footprint f;
	f.x0 = (uint32_t)fminf(fmaxf(fx0, 0.f), wmax);
	f.y0 = (uint32_t)fminf(fmaxf(fy0, 0.f), hmax);
	f.x1 = (uint32_t)fminf(fmaxf(__fadd_rn(fx0, 1.f), 0.f), wmax);
	f.y1 = (uint32_t)fminf(fmaxf(__fadd_rn(fy0, 1.f), 0.f), hmax);
	f.usex1 = __fsub_rn(x, fx0) != 0.f;
	f.usey1 = __fsub_rn(y, fy0) != 0.f;
	return f;
}

template <typename Load>
__device__ __forceinline__ float sample_min(Load load, uint32_t w, uint32_t h, float u, float v)
{
	Footprint f = min_footprint(w, h, u, v);
	// all four loads are issued unconditionally (clamped addresses are always valid) so they are in flight together;
	// texels whose weight is zero are replaced by +inf before the min (VK_SAMPLER_REDUCTION_MODE_MIN ignores them)
	const float inf = __int_as_float(0x7f800000);
	uint32_t r0 = f.y0 * w, r1 = f.y1 * w;
	float t00 = load(r0 + f.x0);
	float t01 = load(r0 + f.x1);
	float t10 = load(r1 + f.x0);
	float t11 = load(r1 + f.x1);
	t01 = f.usex1 ? t01 : inf;
	t10 = f.usey1 ? t10 : inf;
	t11 = (f.usex1 && f.usey1) ? t11 : inf;
	return fminf(fminf(t00, t01), fminf(t10, t11));
}

} // namespace nvc
