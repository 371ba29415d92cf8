// nvc_math2.cuh — the same arithmetic as nvc_math.cuh on TWO items per lane with Blackwell's packed FP32x2
// instructions (PTX mul/fma .f32x2 -> SASS FMUL2 / FFMA2, new in sm_100).  Each half is an independent, correctly
// rounded binary32 operation, so results are bit-identical to the scalar path while every multiply / add / subtract
// costs ONE issue slot for two meshlets — the cluster kernels are issue-bound, not FMA-pipe-bound.
//
// ptxas contracts `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 (it does not honour the .rn no-contraction rule for the
// packed forms, with or without -fmad=false) and even rewrites fma(x, 1.0, y) with a literal 1.0 back into a
// contractible add.  Additions are therefore issued as fma(a, ONE, b) with ONE = 1.0f read from a kernel parameter:
// the compiler cannot fold a run-time multiplier, a product cannot be fused INTO a fused multiply-add, and a * 1.0 + b
// rounds exactly like a + b.  Subtraction a - b is fma(b, -1.0 (run-time), a).  Checked by the bit-exact parity tests.
#pragma once

#include "nvc_math.cuh"

namespace nvc
{

typedef unsigned long long f2; // {lo = item A, hi = item B}

struct Pk // run-time constants that keep ptxas from contracting (see above)
{
	f2 one, neg_one;
};

#if defined(NVC_EMU) // CPU emulation of the kernels (tests/cuda_emu): each half is one IEEE operation
__device__ __forceinline__ f2 pk(float a, float b)
{
	return f2(__float_as_uint(a)) | (f2(__float_as_uint(b)) << 32);
}
#else
__device__ __forceinline__ f2 pk(float a, float b)
{
	f2 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
	return r;
}
#endif

__device__ __forceinline__ float lo(f2 v)
{
	return __uint_as_float(static_cast<uint32_t>(v));
}

__device__ __forceinline__ float hi(f2 v)
{
	return __uint_as_float(static_cast<uint32_t>(v >> 32));
}

__device__ __forceinline__ f2 bc(float s) // broadcast a scalar to both halves
{
	return pk(s, s);
}

#if defined(NVC_EMU)
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return pk(__fmul_rn(lo(a), lo(b)), __fmul_rn(hi(a), hi(b))); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return pk(__fmaf_rn(lo(a), lo(b), lo(c)), __fmaf_rn(hi(a), hi(b), hi(c))); }
#else
__device__ __forceinline__ f2 mul2(f2 a, f2 b)
{
	f2 d;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}

__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) // a genuine fused multiply-add (used where the scalar code fuses too)
{
	f2 d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}
#endif

__device__ __forceinline__ f2 add2(const Pk& k, f2 a, f2 b)
{
	return fma2(a, k.one, b); // a * 1 + b == a + b, one rounding
}

__device__ __forceinline__ f2 sub2(const Pk& k, f2 a, f2 b)
{
	return fma2(b, k.neg_one, a); // b * -1 + a == a - b, one rounding
}

struct f3x2
{
	f2 x, y, z;
};

// GLSL cross()
__device__ __forceinline__ f3x2 cross3_2(const Pk& k, f3x2 a, f3x2 b)
{
	f3x2 r;
	r.x = sub2(k, mul2(a.y, b.z), mul2(b.y, a.z));
	r.y = sub2(k, mul2(a.z, b.x), mul2(b.z, a.x));
	r.z = sub2(k, mul2(a.x, b.y), mul2(b.x, a.y));
	return r;
}

// src/shaders/math.h:46-49; q = (qx, qy, qz, qw) per item
__device__ __forceinline__ f3x2 rotate_quat2(const Pk& k, f3x2 v, f3x2 qv, f2 qw)
{
	f3x2 c1 = cross3_2(k, qv, v);
	f3x2 t = { add2(k, c1.x, mul2(qw, v.x)), add2(k, c1.y, mul2(qw, v.y)), add2(k, c1.z, mul2(qw, v.z)) };
	f3x2 c2 = cross3_2(k, qv, t);
	const f2 two = bc(2.0f);
	f3x2 r = { add2(k, v.x, mul2(two, c2.x)), add2(k, v.y, mul2(two, c2.y)), add2(k, v.z, mul2(two, c2.z)) };
	return r;
}

// (view * vec4(p, 1)).xyz, column-major, left-to-right
__device__ __forceinline__ f3x2 transform_point2(const Pk& k, const float* __restrict__ m, f3x2 p)
{
	f3x2 r;
	r.x = add2(k, add2(k, add2(k, mul2(bc(m[0]), p.x), mul2(bc(m[4]), p.y)), mul2(bc(m[8]), p.z)), bc(m[12]));
	r.y = add2(k, add2(k, add2(k, mul2(bc(m[1]), p.x), mul2(bc(m[5]), p.y)), mul2(bc(m[9]), p.z)), bc(m[13]));
	r.z = add2(k, add2(k, add2(k, mul2(bc(m[2]), p.x), mul2(bc(m[6]), p.y)), mul2(bc(m[10]), p.z)), bc(m[14]));
	return r;
}

__device__ __forceinline__ f3x2 transform_vector2(const Pk& k, const float* __restrict__ m, f3x2 v)
{
	f3x2 r;
	r.x = add2(k, add2(k, mul2(bc(m[0]), v.x), mul2(bc(m[4]), v.y)), mul2(bc(m[8]), v.z));
	r.y = add2(k, add2(k, mul2(bc(m[1]), v.x), mul2(bc(m[5]), v.y)), mul2(bc(m[9]), v.z));
	r.z = add2(k, add2(k, mul2(bc(m[2]), v.x), mul2(bc(m[6]), v.y)), mul2(bc(m[10]), v.z));
	return r;
}

__device__ __forceinline__ f2 dot3_2(const Pk& k, f3x2 a, f3x2 b)
{
	return add2(k, add2(k, mul2(a.x, b.x), mul2(a.y, b.y)), mul2(a.z, b.z));
}

__device__ __forceinline__ f2 sqrt2(f2 v) // correctly rounded, per half
{
	return pk(__fsqrt_rn(lo(v)), __fsqrt_rn(hi(v)));
}

__device__ __forceinline__ f2 div2(f2 a, f2 b) // correctly rounded, per half
{
	return pk(__fdiv_rn(lo(a), lo(b)), __fdiv_rn(hi(a), hi(b)));
}

__device__ __forceinline__ f2 length3_2(const Pk& k, f3x2 a)
{
	return sqrt2(add2(k, add2(k, mul2(a.x, a.x), mul2(a.y, a.y)), mul2(a.z, a.z)));
}

// true quotient float(i) / 127.0f (see s8_div127): the two FMAs are genuinely fused in the scalar code as well
__device__ __forceinline__ f2 s8_div127_2(int ia, int ib)
{
	const f2 r = bc(0.00787401574803149606f);
	f2 a = pk((float)ia, (float)ib);
	f2 q0 = mul2(a, r);
	f2 rem = fma2(q0, bc(-127.0f), a); // -q0 * 127 + a
	return fma2(rem, r, q0);
}

// drawcull.comp.glsl:77-83 == clustercull.comp.glsl:104-108; returns the two verdicts
template <typename CD>
__device__ __forceinline__ void frustum_visible2(const Pk& k, const CD& cd, f3x2 c, f2 radius, bool& va, bool& vb)
{
	// |c.x|, |c.y| are free operand modifiers in scalar FMUL: keep these two products scalar
	f2 ax = pk(__fmul_rn(fabsf(lo(c.x)), cd.frustum[0]), __fmul_rn(fabsf(hi(c.x)), cd.frustum[0]));
	f2 ay = pk(__fmul_rn(fabsf(lo(c.y)), cd.frustum[2]), __fmul_rn(fabsf(hi(c.y)), cd.frustum[2]));
	f2 tx = sub2(k, mul2(c.z, bc(cd.frustum[1])), ax);
	f2 ty = sub2(k, mul2(c.z, bc(cd.frustum[3])), ay);
	f2 zn = add2(k, c.z, radius);
	f2 zf = sub2(k, c.z, radius);
	float ra = lo(radius), rb = hi(radius);
	va = lo(tx) > -ra && lo(ty) > -ra && lo(zn) > cd.znear && lo(zf) < cd.zfar;
	vb = hi(tx) > -rb && hi(ty) > -rb && hi(zn) > cd.znear && hi(zf) < cd.zfar;
}

// src/shaders/math.h:1-22, branch-free like project_sphere(); aabb components as packed pairs
struct aabb2
{
	f2 x, y, z, w;
};

__device__ __forceinline__ void project_sphere2(const Pk& k, f3x2 c, f2 r, float znear, float P00, float P11, aabb2& aabb, bool& oka, bool& okb)
{
	f2 rz = add2(k, r, bc(znear));
	oka = !(lo(c.z) < lo(rz));
	okb = !(hi(c.z) < hi(rz));

	f2 crx = mul2(c.x, r), cry = mul2(c.y, r), crz = mul2(c.z, r);
	f2 czr2 = sub2(k, mul2(c.z, c.z), mul2(r, r));

	f2 vx = sqrt2(add2(k, mul2(c.x, c.x), czr2));
	f2 vxx = mul2(vx, c.x), vxz = mul2(vx, c.z);
	f2 minx = div2(sub2(k, vxx, crz), add2(k, vxz, crx));
	f2 maxx = div2(add2(k, vxx, crz), sub2(k, vxz, crx));

	f2 vy = sqrt2(add2(k, mul2(c.y, c.y), czr2));
	f2 vyy = mul2(vy, c.y), vyz = mul2(vy, c.z);
	f2 miny = div2(sub2(k, vyy, crz), add2(k, vyz, cry));
	f2 maxy = div2(add2(k, vyy, crz), sub2(k, vyz, cry));

	const f2 half = bc(0.5f), nhalf = bc(-0.5f);
	aabb.x = add2(k, mul2(mul2(minx, bc(P00)), half), half);
	aabb.y = add2(k, mul2(mul2(maxy, bc(P11)), nhalf), half);
	aabb.z = add2(k, mul2(mul2(maxx, bc(P00)), half), half);
	aabb.w = add2(k, mul2(mul2(miny, bc(P11)), nhalf), half);
}

// src/shaders/math.h:24-39 for two items; float arithmetic packed, exponent logic per item (see occlusion_mip)
__device__ __forceinline__ void occlusion_mip2(const Pk& k, const aabb2& aabb, float pw, float ph, int max_level, int& la, int& lb)
{
	f2 sizex = sub2(k, aabb.z, aabb.x);
	f2 sizey = sub2(k, aabb.w, aabb.y);
	f2 mx = mul2(sizex, bc(pw)), my = mul2(sizey, bc(ph));
	float ma = fmaxf(lo(mx), lo(my)), mb = fmaxf(hi(mx), hi(my));

	// per item: L = ceil(log2 m) for m > 1 (else level 0), exp2(1 - L) as a float; items at level 0 / max get scale 0
	uint32_t ba = __float_as_uint(ma), bb = __float_as_uint(mb);
	bool pa = ma > 1.0f, pb = mb > 1.0f;
	bool ia = ba == 0x7f800000u, ib = bb == 0x7f800000u;
	int La = int(ba >> 23) - 127 + ((ba & 0x7fffffu) ? 1 : 0);
	int Lb = int(bb >> 23) - 127 + ((bb & 0x7fffffu) ? 1 : 0);
	La = pa ? La : 1; // any valid exponent: the result is overridden below
	Lb = pb ? Lb : 1;
	La = min(max(La, 1), 128);
	Lb = min(max(Lb, 1), 128);
	float sa = (La <= 127) ? __uint_as_float(uint32_t(128 - La) << 23) : __uint_as_float(0x00400000u);
	float sb = (Lb <= 127) ? __uint_as_float(uint32_t(128 - Lb) << 23) : __uint_as_float(0x00400000u);
	f2 scale = pk(sa, sb);
	f2 fmx = mul2(bc(pw), scale), fmy = mul2(bc(ph), scale);
	f2 px = mul2(aabb.x, fmx), py = mul2(aabb.y, fmy);
	f2 fx = sub2(k, px, pk(floorf(lo(px)), floorf(hi(px))));
	f2 fy = sub2(k, py, pk(floorf(lo(py)), floorf(hi(py))));
	f2 ex = add2(k, fx, mul2(sizex, fmx));
	f2 ey = add2(k, fy, mul2(sizey, fmy));
	bool fa = lo(ex) <= 2.0f && lo(ey) <= 2.0f;
	bool fb = hi(ex) <= 2.0f && hi(ey) <= 2.0f;
	La -= fa ? 1 : 0;
	Lb -= fb ? 1 : 0;
	la = !pa ? 0 : (ia ? max_level : min(La, max_level));
	lb = !pb ? 0 : (ib ? max_level : min(Lb, max_level));
}

} // namespace nvc
