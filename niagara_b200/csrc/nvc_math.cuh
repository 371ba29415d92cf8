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
__device__ __forceinline__ bool frustum_visible(const NvcCullData& cd, f3 c, float radius)
{
	bool visible = __fsub_rn(__fmul_rn(c.z, cd.frustum[1]), __fmul_rn(fabsf(c.x), cd.frustum[0])) > -radius;
	visible = visible && __fsub_rn(__fmul_rn(c.z, cd.frustum[3]), __fmul_rn(fabsf(c.y), cd.frustum[2])) > -radius;
	visible = visible && __fadd_rn(c.z, radius) > cd.znear && __fsub_rn(c.z, radius) < cd.zfar;
	return visible;
}

// src/shaders/math.h:1-22
__device__ __forceinline__ bool project_sphere(f3 c, float r, float znear, float P00, float P11, float4& aabb)
{
	if (c.z < __fadd_rn(r, znear))
		return false;

	float crx = __fmul_rn(c.x, r), cry = __fmul_rn(c.y, r), crz = __fmul_rn(c.z, r);
	float czr2 = __fsub_rn(__fmul_rn(c.z, c.z), __fmul_rn(r, r));

	float vx = __fsqrt_rn(__fadd_rn(__fmul_rn(c.x, c.x), czr2));
	float minx = __fdiv_rn(__fsub_rn(__fmul_rn(vx, c.x), crz), __fadd_rn(__fmul_rn(vx, c.z), crx));
	float maxx = __fdiv_rn(__fadd_rn(__fmul_rn(vx, c.x), crz), __fsub_rn(__fmul_rn(vx, c.z), crx));

	float vy = __fsqrt_rn(__fadd_rn(__fmul_rn(c.y, c.y), czr2));
	float miny = __fdiv_rn(__fsub_rn(__fmul_rn(vy, c.y), crz), __fadd_rn(__fmul_rn(vy, c.z), cry));
	float maxy = __fdiv_rn(__fadd_rn(__fmul_rn(vy, c.y), crz), __fsub_rn(__fmul_rn(vy, c.z), cry));

	aabb.x = __fadd_rn(__fmul_rn(__fmul_rn(minx, P00), 0.5f), 0.5f);
	aabb.y = __fadd_rn(__fmul_rn(__fmul_rn(maxy, P11), -0.5f), 0.5f);
	aabb.z = __fadd_rn(__fmul_rn(__fmul_rn(maxx, P00), 0.5f), 0.5f);
	aabb.w = __fadd_rn(__fmul_rn(__fmul_rn(miny, P11), -0.5f), 0.5f);
	return true;
}

// smallest integer L with 2^L >= x for finite x > 0
__device__ __forceinline__ int ceil_log2_exact(float x)
{
	uint32_t bits = __float_as_uint(x);
	int e = int((bits >> 23) & 0xff);
	uint32_t man = bits & 0x7fffffu;
	if (e == 0)
	{
		int top = 31 - __clz(man);
		bool pow2 = (man & (man - 1)) == 0;
		return top - 149 + (pow2 ? 0 : 1);
	}
	return e - 127 + (man ? 1 : 0);
}

// src/shaders/math.h:24-39; returns an integer level already clamped to [0, max_level]
__device__ __forceinline__ int occlusion_mip(float4 aabb, float pw, float ph, int max_level)
{
	float sizex = __fsub_rn(aabb.z, aabb.x);
	float sizey = __fsub_rn(aabb.w, aabb.y);
	float a = __fmul_rn(sizex, pw);
	float b = __fmul_rn(sizey, ph);
	float m = (a > b) ? a : b;
	if (b != b && !(a != a))
		m = a;

	if (!(m > 0.f))
		return 0;
	if (m == __int_as_float(0x7f800000))
		return max_level;

	int L = ceil_log2_exact(m);
	if (L <= 0)
		return 0;

	// exp2(1 - L): exact power of two; 1 - L is in [-127, 0] here (L <= 128), build it from the exponent bits,
	// falling back to the subnormal 2^-127 when L == 128
	float scale = (L <= 127) ? __uint_as_float(uint32_t(127 + 1 - L) << 23) : __uint_as_float(0x00400000u);
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

__device__ __forceinline__ Footprint min_footprint(uint32_t w, uint32_t h, float u, float v)
{
	float x = __fsub_rn(__fmul_rn(u, (float)w), 0.5f);
	float y = __fsub_rn(__fmul_rn(v, (float)h), 0.5f);
	float fx0 = floorf(x), fy0 = floorf(y);
	float fx = __fsub_rn(x, fx0), fy = __fsub_rn(y, fy0);
	float wmax = (float)(w - 1), hmax = (float)(h - 1);
	float fx1 = __fadd_rn(fx0, 1.f), fy1 = __fadd_rn(fy0, 1.f);
	float cx0 = fx0 < 0.f ? 0.f : (fx0 > wmax ? wmax : fx0);
	float cy0 = fy0 < 0.f ? 0.f : (fy0 > hmax ? hmax : fy0);
	float cx1 = fx1 < 0.f ? 0.f : (fx1 > wmax ? wmax : fx1);
	float cy1 = fy1 < 0.f ? 0.f : (fy1 > hmax ? hmax : fy1);
	if (!(x == x))
	{
		cx0 = cx1 = 0.f;
		fx = 0.f;
	}
	if (!(y == y))
	{
		cy0 = cy1 = 0.f;
		fy = 0.f;
	}
	Footprint f;
	f.x0 = (uint32_t)cx0;
	f.x1 = (uint32_t)cx1;
	f.y0 = (uint32_t)cy0;
	f.y1 = (uint32_t)cy1;
	f.usex1 = fx != 0.f;
	f.usey1 = fy != 0.f;
	return f;
}

template <typename Load>
__device__ __forceinline__ float sample_min(Load load, uint32_t w, uint32_t h, float u, float v)
{
	Footprint f = min_footprint(w, h, u, v);
	// issue all four loads unconditionally (clamped addresses are always valid), select afterwards: keeps the
	// loads independent and in flight together
	float t00 = load(f.y0 * w + f.x0);
	float t01 = load(f.y0 * w + f.x1);
	float t10 = load(f.y1 * w + f.x0);
	float t11 = load(f.y1 * w + f.x1);
	float r = t00;
	if (f.usex1)
		r = fminf(r, t01);
	if (f.usey1)
	{
		r = fminf(r, t10);
		if (f.usex1)
			r = fminf(r, t11);
	}
	return r;
}

} // namespace nvc
