// =====================================================================================================
// glsl_shim.h — a CPU stand-in for the GLSL compute environment.  TEST INFRASTRUCTURE ONLY (oracle/).
//
// Purpose: let g++ compile the reference's OWN shader text (src/shaders/{drawcull,tasksubmit,clustercull,
// clustersubmit,depthreduce}.comp.glsl, meshlet.task.glsl with mesh.h / math.h / ../config.h) so that the shaders
// themselves — not a restatement — can be executed on the host and compared with oracle/oracle.cpp and with the
// CUDA path.  gen.py rewrites only the declarations C++ cannot parse (layout(...) blocks, `out` parameters,
// `shared`, specialisation constants, unsuffixed float literals -> `f`); every statement of every function body
// is the reference's, compiled from where it lies under /root/reference into oracle/_ref/.
//
// What in here is OURS (and therefore an interpretation) rather than the reference's:
//   * vector / matrix arithmetic: component-wise IEEE binary32, one rounding per operation, no contraction
//     (the TU is built -ffp-contract=off); dot and length summed left to right; mat * vec as the column sum
//     ((m0*x + m1*y) + m2*z) + m3*w.
//   * min / max follow IEEE minNum / maxNum (the GPU's FMNMX): a NaN operand loses.
//   * log2 is rounded toward +inf so that ceil(log2(x)) is the exact "smallest L with 2^L >= x" (a GPU's
//     approximate log2 and libm's round-to-nearest one both misplace x within 1 ulp of a power of two).
//   * the MIN-reduction sampler (Vulkan texel filtering: unnormalised x = u*w - 0.5, floor / fract, texels with
//     a zero weight excluded, clamp-to-edge, nearest mip level clamped to the view's range).
//   * out-of-bounds storage-buffer reads return memory the driver pads with zeros (robust-buffer behaviour),
//     out-of-bounds imageStore is discarded.
// =====================================================================================================
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__FAST_MATH__)
#error "the shader shim must not be built with -ffast-math"
#endif

namespace glsl
{

typedef unsigned int uint;

// ---- storage-only 16-bit float (GL_EXT_shader_16bit_storage): converts exactly to float ------------------------
struct float16_t
{
	uint16_t bits;

	operator float() const
	{
		uint32_t sign = uint32_t(bits & 0x8000u) << 16;
		uint32_t e = (bits >> 10) & 0x1fu, m = bits & 0x3ffu, out;
		if (e == 0x1fu)
			out = sign | 0x7f800000u | (m << 13);
		else if (e != 0)
			out = sign | ((e + 112u) << 23) | (m << 13);
		else if (m == 0)
			out = sign;
		else
		{
			int shift = __builtin_clz(m) - 21; // normalise the subnormal: leading one to bit 10
			m = (m << shift) & 0x3ffu;
			out = sign | (uint32_t(113 - shift) << 23) | (m << 13);
		}
		float f;
		memcpy(&f, &out, 4);
		return f;
	}
};

// ---- swizzle proxy: lives in a union with the parent's components, P = parent component count -------------------
template <typename T, typename V, int P, int... I>
struct Swz
{
	T v[P];

	operator V() const { return V(v[I]...); }

	Swz& operator=(const V& r)
	{
		int k = 0;
		((v[I] = r[k++]), ...);
		return *this;
	}
	Swz& operator+=(const V& r) { return *this = V(*this) + r; }
	Swz& operator-=(const V& r) { return *this = V(*this) - r; }
	Swz& operator*=(const V& r) { return *this = V(*this) * r; }
};

struct vec2;
struct vec3;
struct vec4;
struct uvec2;
struct ivec2;
struct ivec3;

struct vec2
{
	union
	{
		struct
		{
			float x, y;
		};
		Swz<float, vec2, 2, 0, 1> xy;
		Swz<float, vec2, 2, 1, 0> yx;
	};

	vec2() : x(0), y(0) {}
	explicit vec2(float s) : x(s), y(s) {}
	vec2(float x_, float y_) : x(x_), y(y_) {}
	explicit vec2(const uvec2& u);
	explicit vec2(const ivec2& u);
	float operator[](int i) const { return (&x)[i]; }
	float& operator[](int i) { return (&x)[i]; }
};

struct vec3
{
	union
	{
		struct
		{
			float x, y, z;
		};
		Swz<float, vec2, 3, 0, 1> xy;
		Swz<float, vec2, 3, 1, 0> yx;
		Swz<float, vec3, 3, 0, 1, 2> xyz;
	};

	vec3() : x(0), y(0), z(0) {}
	explicit vec3(float s) : x(s), y(s), z(s) {}
	vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
	vec3(const vec2& a, float z_) : x(a.x), y(a.y), z(z_) {}
	explicit vec3(const ivec3& u);
	float operator[](int i) const { return (&x)[i]; }
	float& operator[](int i) { return (&x)[i]; }
};

struct vec4
{
	union
	{
		struct
		{
			float x, y, z, w;
		};
		Swz<float, vec2, 4, 0, 1> xy;
		Swz<float, vec2, 4, 2, 3> zw;
		Swz<float, vec3, 4, 0, 1, 2> xyz;
		Swz<float, vec4, 4, 0, 3, 2, 1> xwzy;
	};

	vec4() : x(0), y(0), z(0), w(0) {}
	explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
	vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
	vec4(const vec3& a, float w_) : x(a.x), y(a.y), z(a.z), w(w_) {}
	float operator[](int i) const { return (&x)[i]; }
	float& operator[](int i) { return (&x)[i]; }
};

#define GLSL_VEC_OPS(V, N) \
	inline V operator+(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r; } \
	inline V operator-(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - b[i]; return r; } \
	inline V operator*(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * b[i]; return r; } \
	inline V operator/(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / b[i]; return r; } \
	inline V operator+(const V& a, float b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + b; return r; } \
	inline V operator-(const V& a, float b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - b; return r; } \
	inline V operator*(const V& a, float b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * b; return r; } \
	inline V operator/(const V& a, float b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / b; return r; } \
	inline V operator+(float a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a + b[i]; return r; } \
	inline V operator-(float a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a - b[i]; return r; } \
	inline V operator*(float a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a * b[i]; return r; } \
	inline V operator/(float a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a / b[i]; return r; } \
	inline V operator-(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r; } \
	inline V& operator+=(V& a, const V& b) { return a = a + b; } \
	inline V& operator-=(V& a, const V& b) { return a = a - b; } \
	inline V& operator*=(V& a, const V& b) { return a = a * b; } \
	inline V& operator*=(V& a, float b) { return a = a * b; } \
	inline V abs(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = fabsf(a[i]); return r; } \
	inline V floor(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = floorf(a[i]); return r; } \
	inline V fract(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - floorf(a[i]); return r; } \
	inline V sqrt(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = sqrtf(a[i]); return r; } \
	inline V pow(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = powf(a[i], b[i]); return r; } \
	inline V max(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = fmaxf(a[i], b[i]); return r; } \
	inline V min(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = fminf(a[i], b[i]); return r; }

GLSL_VEC_OPS(vec2, 2)
GLSL_VEC_OPS(vec3, 3)
GLSL_VEC_OPS(vec4, 4)

// ---- scalar built-ins (names must hide ::abs, ::min ... inside namespace glsl) ----------------------------------
inline float abs(float a) { return fabsf(a); }
inline int abs(int a) { return a < 0 ? -a : a; }
inline float sqrt(float a) { return sqrtf(a); }
inline float floor(float a) { return floorf(a); }
inline float ceil(float a) { return ceilf(a); }
inline float fract(float a) { return a - floorf(a); }
inline float exp2(float a) { return exp2f(a); } // exact for integral arguments (glibc), which is all the path uses
inline float pow(float a, float b) { return powf(a, b); }

// log2 rounded toward +inf (see the header comment): ceil(log2(x)) == smallest L with 2^L >= x, for every x > 0.
inline float log2(float a)
{
	if (!(a > 0.0f) || isinf(a))
		return log2f(a); // -inf, NaN, +inf as IEEE
	double d = ::log2(double(a));
	float f = float(d);
	if (double(f) < d)
		f = nextafterf(f, INFINITY);
	// double log2 is accurate to < 1 ulp(double); an exact integer result can only come from a power of two
	return f;
}

inline float max(float a, float b) { return fmaxf(a, b); }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, int b) { return fmaxf(a, float(b)); }
inline float max(int a, float b) { return fmaxf(float(a), b); }
inline float min(float a, int b) { return fminf(a, float(b)); }
inline float min(int a, float b) { return fminf(float(a), b); }
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline uint max(uint a, uint b) { return a > b ? a : b; }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline uint max(uint a, int b) { return max(a, uint(b)); } // GLSL implicit int -> uint
inline uint max(int a, uint b) { return max(uint(a), b); }
inline uint min(uint a, int b) { return min(a, uint(b)); }
inline uint min(int a, uint b) { return min(uint(a), b); }

inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline float length(const vec2& a) { return sqrtf(dot(a, a)); }
inline float length(const vec3& a) { return sqrtf(dot(a, a)); }
inline vec3 normalize(const vec3& a) { return a / length(a); }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }

// ---- bool / int vectors (only what the shaders use) -------------------------------------------------------------
struct bvec2
{
	bool x, y;
};
inline bvec2 lessThanEqual(const vec2& a, const vec2& b) { return bvec2{ a.x <= b.x, a.y <= b.y }; }
inline bool all(const bvec2& b) { return b.x && b.y; }
inline bool any(const bvec2& b) { return b.x || b.y; }

struct uvec2
{
	uint x, y;
	uvec2() : x(0), y(0) {}
	uvec2(uint x_, uint y_) : x(x_), y(y_) {}
	uint operator[](int i) const { return (&x)[i]; }
};

struct ivec2
{
	int x, y;
	ivec2() : x(0), y(0) {}
	explicit ivec2(int s) : x(s), y(s) {}
	ivec2(int x_, int y_) : x(x_), y(y_) {}
	explicit ivec2(const uvec2& u) : x(int(u.x)), y(int(u.y)) {}
	explicit ivec2(uint s) : x(int(s)), y(int(s)) {}
};

struct ivec3
{
	int x, y, z;
	ivec3() : x(0), y(0), z(0) {}
	explicit ivec3(int s) : x(s), y(s), z(s) {}
	explicit ivec3(uint s) : x(int(s)), y(int(s)), z(int(s)) {}
	ivec3(int x_, int y_, int z_) : x(x_), y(y_), z(z_) {}
};

inline ivec2 operator>>(const ivec2& a, const ivec2& b) { return ivec2(a.x >> b.x, a.y >> b.y); }
inline ivec2 operator&(const ivec2& a, const ivec2& b) { return ivec2(a.x & b.x, a.y & b.y); }
inline ivec3 operator>>(const ivec3& a, const ivec3& b) { return ivec3(a.x >> b.x, a.y >> b.y, a.z >> b.z); }
inline ivec3 operator&(const ivec3& a, const ivec3& b) { return ivec3(a.x & b.x, a.y & b.y, a.z & b.z); }
inline vec2::vec2(const uvec2& u) : x(float(u.x)), y(float(u.y)) {}
inline vec2::vec2(const ivec2& u) : x(float(u.x)), y(float(u.y)) {}
inline vec3::vec3(const ivec3& u) : x(float(u.x)), y(float(u.y)), z(float(u.z)) {}
inline vec2 operator/(const ivec2& a, float b) { return vec2(a) / b; } // GLSL implicit int -> float
inline vec3 operator/(const ivec3& a, float b) { return vec3(a) / b; }

struct uvec3
{
	union
	{
		struct
		{
			uint x, y, z;
		};
		Swz<uint, uvec2, 3, 0, 1> xy;
	};
	uvec3() : x(0), y(0), z(0) {}
	uvec3(uint x_, uint y_, uint z_) : x(x_), y(y_), z(z_) {}
};

// ---- matrices: column-major ------------------------------------------------------------------------------------
struct alignas(16) mat4 // std430: a mat4 (and every struct holding one, e.g. CullData = 144 B) is 16-byte aligned
{
	vec4 col[4];
	const vec4& operator[](int i) const { return col[i]; }
	vec4& operator[](int i) { return col[i]; }
};

struct mat3
{
	vec3 col[3];
	mat3() {}
	explicit mat3(const mat4& m)
	{
		for (int i = 0; i < 3; ++i)
			col[i] = vec3(m.col[i].x, m.col[i].y, m.col[i].z);
	}
};

inline vec4 operator*(const mat4& m, const vec4& v) { return ((m.col[0] * v.x + m.col[1] * v.y) + m.col[2] * v.z) + m.col[3] * v.w; }
inline vec3 operator*(const mat3& m, const vec3& v) { return (m.col[0] * v.x + m.col[1] * v.y) + m.col[2] * v.z; }

// ---- atomics on storage / shared memory (dispatches may run workgroups on several host threads) ------------------
inline uint atomicAdd(uint& mem, uint data) { return __atomic_fetch_add(&mem, data, __ATOMIC_RELAXED); }
inline int atomicAdd(int& mem, int data) { return __atomic_fetch_add(&mem, data, __ATOMIC_RELAXED); }
inline uint atomicOr(uint& mem, uint data) { return __atomic_fetch_or(&mem, data, __ATOMIC_RELAXED); }
inline uint atomicAnd(uint& mem, uint data) { return __atomic_fetch_and(&mem, data, __ATOMIC_RELAXED); }

// ---- images and the MIN-reduction sampler ------------------------------------------------------------------------
// A texture2D is a view of `levels` consecutive mips; mip l is row-major, tightly packed at texels[l].
struct texture2D
{
	const float* texels[16];
	uint width[16], height[16];
	uint levels;
};

struct image2D
{
	float* texels;
	uint width, height;
};

struct sampler
{
	int unused;
};

struct sampler2D
{
	const texture2D* tex;
	sampler2D(const texture2D& t, const sampler&) : tex(&t) {}
};

inline float sampleMinLevel(const texture2D& t, uint l, const vec2& uv)
{
	uint w = t.width[l], h = t.height[l];
	const float* img = t.texels[l];
	float x = uv.x * float(w) - 0.5f, y = uv.y * float(h) - 0.5f;
	float x0 = floorf(x), y0 = floorf(y);
	float fx = x - x0, fy = y - y0;
	// CLAMP_TO_EDGE; a NaN coordinate (undefined in Vulkan) lands on texel 0, as fmax(NaN, 0) does in the oracle / CUDA path
	auto clampi = [](float v, uint n) -> uint { return !(v > 0.0f) ? 0u : (v >= float(n - 1) ? n - 1 : uint(v)); };
	uint i0 = clampi(x0, w), i1 = clampi(x0 + 1.0f, w), j0 = clampi(y0, h), j1 = clampi(y0 + 1.0f, h);
	float r = img[size_t(j0) * w + i0];
	if (fx != 0.0f)
		r = fminf(r, img[size_t(j0) * w + i1]);
	if (fy != 0.0f)
		r = fminf(r, img[size_t(j1) * w + i0]);
	if (fx != 0.0f && fy != 0.0f)
		r = fminf(r, img[size_t(j1) * w + i1]);
	return r;
}

// mipmapMode NEAREST (niagara.cpp:629), lod clamped to the view's mip range (maxLod 16 >= any pyramid, resources.cpp:304)
inline vec4 textureLod(const sampler2D& s, const vec2& uv, float lod)
{
	const texture2D& t = *s.tex;
	float top = float(t.levels - 1);
	float d = lod <= 0.0f ? 0.0f : (lod >= top ? top : lod);
	if (d != d)
		d = 0.0f;
	uint l = uint(ceilf(d + 0.5f) - 1.0f); // Vulkan's nearest-level rule; d is integral on this path
	float v = sampleMinLevel(t, l, uv);
	return vec4(v, 0.0f, 0.0f, 1.0f);
}

inline vec4 texture(const sampler2D& s, const vec2& uv) { return textureLod(s, uv, 0.0f); } // compute: implicit lod 0

inline void imageStore(image2D& img, const ivec2& p, const vec4& v)
{
	if (p.x < 0 || p.y < 0 || uint(p.x) >= img.width || uint(p.y) >= img.height)
		return; // out-of-bounds stores are discarded
	img.texels[size_t(p.y) * img.width + uint(p.x)] = v.x;
}

// ---- storage-buffer arrays with robustBufferAccess behaviour: out-of-range reads give zero, writes are discarded ----
template <typename T>
struct RobustArray
{
	T* p = nullptr;
	size_t n = 0;

	void bind(void* ptr, size_t bytes)
	{
		p = static_cast<T*>(ptr);
		n = bytes / sizeof(T);
	}
	T& operator[](size_t i) const
	{
		if (i < n)
			return p[i];
		static thread_local T outside;
		memset(static_cast<void*>(&outside), 0, sizeof(T));
		return outside;
	}
};

// ---- invocation state --------------------------------------------------------------------------------------------
extern thread_local uvec3 gl_GlobalInvocationID, gl_LocalInvocationID, gl_WorkGroupID;
extern thread_local uint gl_LocalInvocationIndex;

// ---- mesh stage outputs (GL_EXT_mesh_shader) of the workgroup running on this host thread ------------------------------
struct MeshPerVertex
{
	vec4 gl_Position;
};
struct MeshPerPrimitive
{
	bool gl_CullPrimitiveEXT;
};
extern thread_local MeshPerVertex gl_MeshVerticesEXT[256];
extern thread_local uvec3 gl_PrimitiveTriangleIndicesEXT[256];
extern thread_local MeshPerPrimitive gl_MeshPrimitivesEXT[256];
void SetMeshOutputsEXT(uint vertexCount, uint primitiveCount);

// ---- vertex stage (mesh.vert.glsl): one invocation per vertex of an indexed indirect draw ------------------------------
extern thread_local int gl_VertexIndex, gl_DrawIDARB; // GL_ARB_shader_draw_parameters
extern thread_local vec4 gl_Position;
inline float round(float a) { return roundf(a); }

void barrier();                              // yields to the other invocations of the workgroup (fibers), see driver.cpp
void EmitMeshTasksEXT(uint x, uint y, uint z); // records the workgroup's emit count

} // namespace glsl
