// =====================================================================================================
// oracle.cpp — CPU restatement of niagara's GPU-driven visibility path.  TEST INFRASTRUCTURE ONLY.
//
// This file is the parity oracle and the reported CPU baseline.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it; the product (niagara_b200/) never does.
//
// PARITY: the reference (zeux/niagara @ 95f289b) has no tests, golden vectors or CPU cull path
// and there is no Vulkan driver or glslang here, so parity against a LIVE VULKAN RUN stays unpinned.  What this
// restatement IS pinned to: the reference's own shader text executed on the host — oracle/refshader compiles
// src/shaders/{drawcull,tasksubmit,clustercull,clustersubmit,depthreduce}.comp.glsl and meshlet.task.glsl (with
// mesh.h / math.h / config.h) through a GLSL shim into oracle/_ref/librefshader.so, and tests/test_refshader.py
// requires every pass of this file to equal those shaders bit for bit.  What remains an interpretation (shared by
// this file, the shim and the CUDA path) is what GLSL leaves to the implementation:
//   * IEEE-754 binary32, round-to-nearest-even, NO fused multiply-add (build with -ffp-contract=off),
//     operations in GLSL source order and associativity; true division; correctly rounded sqrt.
//   * ceil(log2(x)) evaluated exactly from the float's exponent/mantissa; exp2(int) exact.
//   * MIN-reduction sampler: bilinear 2x2 footprint at x = u*w - 0.5, zero-weight texels excluded,
//     clamp-to-edge, integer mip level clamped to [0, levels-1].
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// Input geometry for the tests comes from the reference's OWN scene.cpp (oracle/refscene, tests/golden/*.nvcg).
// =====================================================================================================
#include "../include/niagara_cull.h"

#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <sched.h>
#include <pthread.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#if defined(__FAST_MATH__)
#error "the oracle must not be built with -ffast-math"
#endif

namespace
{

struct vec3
{
	float x, y, z;
};

struct vec4
{
	float x, y, z, w;
};

// binary16 -> binary32, exact (GLSL float16_t -> float conversion; clustercull.comp.glsl:72,76)
inline float halfToFloat(uint16_t h)
{
	uint32_t sign = uint32_t(h & 0x8000u) << 16;
	uint32_t exp = (h >> 10) & 0x1f;
	uint32_t man = h & 0x3ffu;
	uint32_t bits;
	if (exp == 0)
	{
		if (man == 0)
			bits = sign;
		else
		{
			// subnormal half: value = man * 2^-24, renormalise
			int e = -1;
			do
			{
				e++;
				man <<= 1;
			} while ((man & 0x400u) == 0);
			bits = sign | uint32_t(127 - 15 - e) << 23 | (man & 0x3ffu) << 13;
		}
	}
	else if (exp == 31)
		bits = sign | 0x7f800000u | man << 13;
	else
		bits = sign | (exp + 127 - 15) << 23 | man << 13;
	float f;
	memcpy(&f, &bits, 4);
	return f;
}

// GLSL cross(): [x1*y2 - y1*x2, x2*y0 - y2*x0, x0*y1 - y0*x1]
inline vec3 cross(vec3 a, vec3 b)
{
	vec3 r;
	r.x = a.y * b.z - b.y * a.z;
	r.y = a.z * b.x - b.z * a.x;
	r.z = a.x * b.y - b.x * a.y;
	return r;
}

// src/shaders/math.h:46-49   v + 2.0 * cross(q.xyz, cross(q.xyz, v) + q.w * v)
inline vec3 rotateQuat(vec3 v, const float q[4])
{
	vec3 qv = { q[0], q[1], q[2] };
	vec3 c1 = cross(qv, v);
	vec3 t = { c1.x + q[3] * v.x, c1.y + q[3] * v.y, c1.z + q[3] * v.z };
	vec3 c2 = cross(qv, t);
	vec3 r = { v.x + 2.0f * c2.x, v.y + 2.0f * c2.y, v.z + 2.0f * c2.z };
	return r;
}

// (cullData.view * vec4(p, 1)).xyz — column-major, summed left to right (Appendix C.1); m3 * 1 is exact
inline vec3 transformPoint(const float m[16], vec3 p)
{
	vec3 r;
	r.x = ((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12];
	r.y = ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13];
	r.z = ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14];
	return r;
}

// mat3(cullData.view) * v
inline vec3 transformVector(const float m[16], vec3 v)
{
	vec3 r;
	r.x = (m[0] * v.x + m[4] * v.y) + m[8] * v.z;
	r.y = (m[1] * v.x + m[5] * v.y) + m[9] * v.z;
	r.z = (m[2] * v.x + m[6] * v.y) + m[10] * v.z;
	return r;
}

inline float dot3(vec3 a, vec3 b)
{
	return (a.x * b.x + a.y * b.y) + a.z * b.z;
}

inline float length3(vec3 a)
{
	return sqrtf((a.x * a.x + a.y * a.y) + a.z * a.z);
}

// src/shaders/math.h:1-22
inline bool projectSphere(vec3 c, float r, float znear, float P00, float P11, vec4& aabb)
{
	if (c.z < r + znear)
		return false;

	vec3 cr = { c.x * r, c.y * r, c.z * r };
	float czr2 = c.z * c.z - r * r;

	float vx = sqrtf(c.x * c.x + czr2);
	float minx = (vx * c.x - cr.z) / (vx * c.z + cr.x);
	float maxx = (vx * c.x + cr.z) / (vx * c.z - cr.x);

	float vy = sqrtf(c.y * c.y + czr2);
	float miny = (vy * c.y - cr.z) / (vy * c.z + cr.y);
	float maxy = (vy * c.y + cr.z) / (vy * c.z - cr.y);

	// aabb = vec4(minx * P00, miny * P11, maxx * P00, maxy * P11).xwzy * vec4(0.5, -0.5, 0.5, -0.5) + 0.5
	aabb.x = (minx * P00) * 0.5f + 0.5f;
	aabb.y = (maxy * P11) * -0.5f + 0.5f;
	aabb.z = (maxx * P00) * 0.5f + 0.5f;
	aabb.w = (miny * P11) * -0.5f + 0.5f;
	return true;
}

// Appendix C.2: smallest integer L with 2^L >= x, for finite x > 0 (exact, from the float's bits)
inline int ceilLog2Exact(float x)
{
	uint32_t bits;
	memcpy(&bits, &x, 4);
	int e = int((bits >> 23) & 0xff);
	uint32_t man = bits & 0x7fffffu;
	if (e == 0)
	{
		// subnormal: x = man * 2^-149
		int top = 31 - __builtin_clz(man);
		bool pow2 = (man & (man - 1)) == 0;
		return top - 149 + (pow2 ? 0 : 1);
	}
	return e - 127 + (man ? 1 : 0);
}

// src/shaders/math.h:24-39.  Returns the (integer-valued) level; +inf is represented by a large value that
// the sampler clamps to the last mip.
inline float getOcclusionMip(vec4 aabb, float pyramidWidth, float pyramidHeight)
{
	float sizex = aabb.z - aabb.x;
	float sizey = aabb.w - aabb.y;

	float a = sizex * pyramidWidth;
	float b = sizey * pyramidHeight;
	float m = (a > b) ? a : b; // GLSL max(x,y) = y if x < y else x
	if (b != b && !(a != a))
		m = a; // keep the defined operand if exactly one is NaN (driver-dependent; never hit by finite inputs)

	if (!(m > 0.f))
		return 0.f; // log2 -> -inf or NaN; max(level, 0) = 0
	if (m == INFINITY)
		return 1e9f;

	int L = ceilLog2Exact(m);
	if (L <= 0)
		return 0.f; // level or level-1 <= 0 -> max(.,0) = 0

	float level = float(L);
	float scale = ldexpf(1.f, 1 - L); // exp2(1 - level), exact
	float fmx = pyramidWidth * scale;
	float fmy = pyramidHeight * scale;

	float px = aabb.x * fmx;
	float py = aabb.y * fmy;
	float fx = px - floorf(px); // fract
	float fy = py - floorf(py);
	bool fits = (fx + sizex * fmx <= 2.0f) && (fy + sizey * fmy <= 2.0f);
	level -= fits ? 1.f : 0.f;

	return level > 0.f ? level : 0.f;
}

struct HiZView
{
	const float* texels;
	uint32_t width, height, levels;
	const uint32_t* level_offset;
};

// Appendix C.3: VK_SAMPLER_REDUCTION_MODE_MIN, LINEAR filter, CLAMP_TO_EDGE on one image level
// (resources.cpp:294-325, niagara.cpp:629).
inline float sampleMin(const float* img, uint32_t w, uint32_t h, float u, float v)
{
	float x = u * float(w) - 0.5f;
	float y = v * float(h) - 0.5f;
	float fx0 = floorf(x), fy0 = floorf(y);
	float fx = x - fx0, fy = y - fy0;

	// clamp in float first so that huge / non-finite coordinates are defined
	float wmax = float(w - 1), hmax = float(h - 1);
	float cx0 = fx0 < 0.f ? 0.f : (fx0 > wmax ? wmax : fx0);
	float cy0 = fy0 < 0.f ? 0.f : (fy0 > hmax ? hmax : fy0);
	float cx1 = fx0 + 1.f < 0.f ? 0.f : (fx0 + 1.f > wmax ? wmax : fx0 + 1.f);
	float cy1 = fy0 + 1.f < 0.f ? 0.f : (fy0 + 1.f > hmax ? hmax : fy0 + 1.f);
	if (!(x == x))
		cx0 = cx1 = 0.f, fx = 0.f;
	if (!(y == y))
		cy0 = cy1 = 0.f, fy = 0.f;

	uint32_t ix0 = uint32_t(cx0), iy0 = uint32_t(cy0), ix1 = uint32_t(cx1), iy1 = uint32_t(cy1);
	bool usex1 = fx != 0.f, usey1 = fy != 0.f;

	float r = img[size_t(iy0) * w + ix0];
	if (usex1)
		r = fminf(r, img[size_t(iy0) * w + ix1]);
	if (usey1)
	{
		r = fminf(r, img[size_t(iy1) * w + ix0]);
		if (usex1)
			r = fminf(r, img[size_t(iy1) * w + ix1]);
	}
	return r;
}

inline float sampleHiZ(const HiZView& hiz, float u, float v, float level)
{
	float lmax = float(hiz.levels - 1);
	float lc = level < 0.f ? 0.f : (level > lmax ? lmax : level);
	uint32_t l = uint32_t(lc);
	uint32_t w = std::max(1u, hiz.width >> l), h = std::max(1u, hiz.height >> l);
	return sampleMin(hiz.texels + hiz.level_offset[l], w, h, u, v);
}

// shared by drawcull.comp.glsl:88-103 and clustercull.comp.glsl:110-124
inline bool occlusionVisible(const NvcCullData& cd, const HiZView& hiz, vec3 center, float radius)
{
	vec4 aabb;
	if (projectSphere(center, radius, cd.znear, cd.P00, cd.P11, aabb))
	{
		float level = getOcclusionMip(aabb, cd.pyramidWidth, cd.pyramidHeight);
		float depth = sampleHiZ(hiz, (aabb.x + aabb.z) * 0.5f, (aabb.y + aabb.w) * 0.5f, level);
		float depthSphere = cd.znear / (center.z - radius);
		return depthSphere > depth;
	}
	return true;
}

inline bool frustumVisible(const NvcCullData& cd, vec3 center, float radius)
{
	bool visible = true;
	visible = visible && center.z * cd.frustum[1] - fabsf(center.x) * cd.frustum[0] > -radius;
	visible = visible && center.z * cd.frustum[3] - fabsf(center.y) * cd.frustum[2] > -radius;
	visible = visible && center.z + radius > cd.znear && center.z - radius < cd.zfar;
	return visible;
}

struct DrawResult
{
	bool emit;
	bool visible;
	uint32_t lodIndex;
};

// drawcull.comp.glsl:54-127 — the per-draw decision, without the command write
inline bool drawDecision(const NvcCullData& cd, bool late, const NvcMeshDraw& draw, const NvcMesh& mesh, uint32_t dv, const HiZView* hiz, DrawResult& out)
{
	out.emit = false;
	out.visible = false;
	out.lodIndex = 0;

	if (draw.postPass != cd.postPass) // :63
		return false;
	if (!late && dv == 0) // :67
		return false;

	vec3 mc = { mesh.center[0], mesh.center[1], mesh.center[2] };
	vec3 rc = rotateQuat(mc, draw.orientation); // :73
	vec3 center = { rc.x * draw.scale + draw.position[0], rc.y * draw.scale + draw.position[1], rc.z * draw.scale + draw.position[2] };
	center = transformPoint(cd.view, center); // :74
	float radius = mesh.radius * draw.scale;   // :75

	bool visible = frustumVisible(cd, center, radius); // :77-83
	visible = visible || cd.cullingEnabled == 0;       // :85

	if (late && visible && cd.occlusionEnabled == 1) // :87-103
		visible = visible && occlusionVisible(cd, *hiz, center, radius);

	out.visible = visible;

	// :108  TASK_CULL == 1 (config.h:8)
	if (visible && (!late || cd.clusterOcclusionEnabled == 1 || dv == 0 || cd.postPass != 0))
	{
		uint32_t lodIndex = 0;
		if (cd.lodEnabled == 1) // :112-120
		{
			float d = length3(center) - radius;
			float distance = d > 0.f ? d : 0.f; // max(x, 0)
			float threshold = distance * cd.lodTarget / draw.scale;
			for (uint32_t i = 1; i < mesh.lodCount; ++i)
				if (mesh.lods[i].error < threshold)
					lodIndex = i;
		}
		out.lodIndex = lodIndex;
		out.emit = true;
	}
	return true;
}

struct Emit
{
	uint32_t di, lod, dv;
	uint32_t groups; // task groups of the selected LOD (drawcull.comp.glsl:122)
};

// drawcull.comp.glsl:54-156 over draws [begin, end) in ascending di (= one legal order of the GLSL atomics).
// Decisions only; the command write happens in orc_drawcull once the global order is known.
void drawcullRange(const NvcCullData& cd, bool late, const NvcMeshDraw* draws, const NvcMesh* meshes, uint32_t* dvb,
    const HiZView* hiz, uint32_t begin, uint32_t end, std::vector<Emit>& emitted, uint8_t* lod_out)
{
	for (uint32_t di = begin; di < end; ++di)
	{
		const NvcMeshDraw& draw = draws[di];
		const NvcMesh& mesh = meshes[draw.meshIndex];
		uint32_t dv = dvb[di];
		DrawResult r;
		bool reached = drawDecision(cd, late, draw, mesh, dv, hiz, r);
		if (lod_out)
			lod_out[di] = r.emit ? uint8_t(r.lodIndex) : 0xff;
		if (!reached)
			continue;
		if (r.emit)
		{
			const NvcMeshLod& lod = mesh.lods[r.lodIndex];
			emitted.push_back(Emit{ di, r.lodIndex, dv, (lod.meshletCount + NVC_TASK_WGSIZE - 1) / NVC_TASK_WGSIZE }); // dv = drawVisibility[di] before the :154 write
		}
		if (late)
			dvb[di] = r.visible ? 1 : 0; // :154-155
	}
}

struct ClusterOut
{
	std::vector<uint32_t> indices;
};

// clustercull.comp.glsl:56-149 / meshlet.task.glsl:53-149 — one lane.  Returns true when the lane appends
// its cluster index.  Updates the visibility bit when LATE (atomically: words are shared between threads).
inline bool clusterLane(const NvcCullData& cd, bool late, const NvcMeshTaskCommand& command, const NvcMeshDraw& meshDraw,
    const NvcMeshlet* meshlets, uint32_t* mvb, const HiZView* hiz, uint32_t mgi)
{
	bool valid = mgi < command.taskCount;
	if (!valid)
		return false; // Appendix C.5: out-of-range lanes have no observable effect

	uint32_t mi = mgi + command.taskOffset;
	uint32_t mvi = mgi + command.meshletVisibilityOffset;

	// early pass: a meshlet whose bit is clear cannot become visible (:91-92) and nothing else of the lane is observable,
	// so the arithmetic below is skipped for it (the CPU baseline should not do work the result cannot depend on)
	if (!late && cd.clusterOcclusionEnabled == 1 && cd.postPass == 0 && (__atomic_load_n(&mvb[mvi >> 5], __ATOMIC_RELAXED) & (1u << (mvi & 31))) == 0)
		return false;

	const NvcMeshlet& ml = meshlets[mi];

	vec3 lc = { halfToFloat(ml.center[0]), halfToFloat(ml.center[1]), halfToFloat(ml.center[2]) };
	vec3 rc = rotateQuat(lc, meshDraw.orientation);
	vec3 center = { rc.x * meshDraw.scale + meshDraw.position[0], rc.y * meshDraw.scale + meshDraw.position[1], rc.z * meshDraw.scale + meshDraw.position[2] };
	center = transformPoint(cd.view, center);

	float radius = halfToFloat(ml.radius) * meshDraw.scale;

	vec3 la = { float(int(ml.cone_axis[0])) / 127.0f, float(int(ml.cone_axis[1])) / 127.0f, float(int(ml.cone_axis[2])) / 127.0f };
	vec3 cone_axis = transformVector(cd.view, rotateQuat(la, meshDraw.orientation));
	float cone_cutoff = float(int(ml.cone_cutoff)) / 127.0f;

	bool visible = valid;
	bool skip = false;

	if (cd.clusterOcclusionEnabled == 1 && cd.postPass == 0) // :86-99
	{
		uint32_t bit = __atomic_load_n(&mvb[mvi >> 5], __ATOMIC_RELAXED) & (1u << (mvi & 31));
		if (!late && bit == 0)
			visible = false;
		if (late && command.lateDrawVisibility == 1 && bit != 0)
			skip = true;
	}

	// :102  coneCull(center, radius, cone_axis, cone_cutoff, vec3(0))  math.h:41-44
	bool backface = dot3(center, cone_axis) >= cone_cutoff * length3(center) + radius;
	visible = visible && (cd.clusterBackfaceEnabled == 0 || !backface);
	visible = visible && frustumVisible(cd, center, radius); // :104-108

	if (late && cd.clusterOcclusionEnabled == 1 && visible) // :110-124
		visible = visible && occlusionVisible(cd, *hiz, center, radius);

	if (late && cd.clusterOcclusionEnabled == 1) // :126-131 (valid is true here)
	{
		if (visible)
			__atomic_fetch_or(&mvb[mvi >> 5], 1u << (mvi & 31), __ATOMIC_RELAXED);
		else
			__atomic_fetch_and(&mvb[mvi >> 5], ~(1u << (mvi & 31)), __ATOMIC_RELAXED);
	}

	return visible && !skip;
}

void clusterRange(const NvcCullData& cd, bool late, const NvcMeshTaskCommand* cmds, const NvcMeshDraw* draws, const NvcMeshlet* meshlets,
    uint32_t* mvb, const HiZView* hiz, uint32_t begin, uint32_t end, std::vector<uint32_t>& out)
{
	for (uint32_t commandId = begin; commandId < end; ++commandId)
	{
		const NvcMeshTaskCommand& command = cmds[commandId];
		if (command.taskCount == 0)
			continue; // padding command (tasksubmit.comp.glsl:42-46): no lane is valid
		const NvcMeshDraw& meshDraw = draws[command.drawId];
		const uint32_t lanes = std::min(command.taskCount, NVC_TASK_WGSIZE); // lanes >= taskCount are invalid (:84)
		for (uint32_t mgi = 0; mgi < lanes; ++mgi)
			if (clusterLane(cd, late, command, meshDraw, meshlets, mvb, hiz, mgi))
				out.push_back(commandId | (mgi << 24)); // :138
	}
}

// Minimal persistent worker pool: the multi-threaded oracle is also the reported CPU baseline, so it should not pay
// for creating (cores) threads in every pass.
// ORC_PIN=1 (bench.py's CPU arms): every worker stays on ONE of the CPUs the process may use (worker i -> i-th allowed
// CPU), so that the timing does not depend on where the scheduler happens to migrate 128 threads.
inline void pinWorker(int index)
{
	static const bool enabled = [] { const char* e = getenv("ORC_PIN"); return e && atoi(e) != 0; }();
	if (!enabled)
		return;
	cpu_set_t allowed;
	CPU_ZERO(&allowed);
	if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0)
		return;
	int n = CPU_COUNT(&allowed);
	if (n <= 0)
		return;
	int want = index % n, seen = 0;
	for (int c = 0; c < CPU_SETSIZE; ++c)
		if (CPU_ISSET(c, &allowed))
		{
			if (seen == want)
			{
				cpu_set_t one;
				CPU_ZERO(&one);
				CPU_SET(c, &one);
				pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
				return;
			}
			++seen;
		}
}

class Pool
{
public:
	static Pool& get()
	{
		static Pool p;
		return p;
	}

	// runs job(0..count-1) on up to `count` workers (the caller takes index 0) and waits for all of them
	void run(int count, const std::function<void(int)>& job)
	{
		if (count <= 1)
		{
			job(0);
			return;
		}
		std::unique_lock<std::mutex> lock(mutex_);
		while (int(workers_.size()) < count - 1)
			workers_.emplace_back([this, idx = int(workers_.size())]() { loop(idx); });
		job_ = &job;
		active_ = count - 1;
		pending_ = count - 1;
		++generation_;
		lock.unlock();
		wake_.notify_all();
		job(0);
		lock.lock();
		done_.wait(lock, [this]() { return pending_ == 0; });
		job_ = nullptr;
	}

private:
	Pool() {}
	~Pool()
	{
		{
			std::lock_guard<std::mutex> lock(mutex_);
			stop_ = true;
			++generation_;
		}
		wake_.notify_all();
		for (auto& w : workers_)
			w.join();
	}

	void loop(int idx)
	{
		pinWorker(idx + 1);
		uint64_t seen = 0;
		for (;;)
		{
			std::unique_lock<std::mutex> lock(mutex_);
			wake_.wait(lock, [&]() { return generation_ != seen; });
			seen = generation_;
			if (stop_)
				return;
			if (idx >= active_)
				continue;
			const std::function<void(int)>* job = job_;
			lock.unlock();
			(*job)(idx + 1);
			lock.lock();
			if (--pending_ == 0)
				done_.notify_all();
		}
	}

	std::mutex mutex_;
	std::condition_variable wake_, done_;
	std::vector<std::thread> workers_;
	const std::function<void(int)>* job_ = nullptr;
	int active_ = 0, pending_ = 0;
	uint64_t generation_ = 0;
	bool stop_ = false;
};

// Over-decomposed variant: `chunks` contiguous ranges handed out dynamically (an atomic counter), so that a slow or late
// worker does not hold the pass back; fn(chunk, begin, end).  Results stay deterministic as long as the caller keeps
// per-CHUNK outputs and concatenates them in chunk order.
template <typename F>
void parallelChunks(uint32_t n, int threads, uint32_t chunks, F&& fn)
{
	chunks = std::max(1u, std::min(chunks, std::max(1u, n)));
	int nt = int(std::min<uint32_t>(uint32_t(std::max(1, threads)), chunks));
	auto range = [&](uint32_t c) { fn(c, uint32_t(uint64_t(n) * c / chunks), uint32_t(uint64_t(n) * (c + 1) / chunks)); };
	if (nt == 1)
	{
		for (uint32_t c = 0; c < chunks; ++c)
			range(c);
		return;
	}
	std::atomic<uint32_t> next{ 0 };
	Pool::get().run(nt, [&](int) {
		for (uint32_t c = next.fetch_add(1); c < chunks; c = next.fetch_add(1))
			range(c);
	});
}

inline uint32_t chunkCount(uint32_t n, int threads)
{
	// ~8 chunks per worker, at least 1024 items each
	return std::max(1u, std::min(uint32_t(std::max(1, threads)) * 8u, (n + 1023u) / 1024u));
}

template <typename F>
void parallelRanges(uint32_t n, int threads, F&& fn)
{
	int nt = std::max(1, threads);
	nt = int(std::min<uint32_t>(uint32_t(nt), std::max(1u, n)));
	if (nt == 1)
	{
		fn(0, 0u, n);
		return;
	}
	Pool::get().run(nt, [&](int t) {
		uint32_t b = uint32_t(uint64_t(n) * t / nt), e = uint32_t(uint64_t(n) * (t + 1) / nt);
		fn(t, b, e);
	});
}

} // namespace

extern "C"
{

// ---- drawcull.comp.glsl:54-156 (+ tasksubmit.comp.glsl:27-47 when task != 0) --------------------------------
// threads <= 1: serial.  Output order = ascending di (per-thread lists concatenated in range order), which is one
// legal outcome of the GLSL's atomics and makes the multi-threaded result identical to the serial one.
// lod_out (may be NULL): per draw selected lodIndex, 0xff when no command was emitted for the draw.
int orc_drawcull(const NvcCullData* cull, int late, int task, const NvcMeshDraw* draws, const NvcMesh* meshes,
    uint32_t* draw_visibility, void* commands, uint32_t* command_count4, const NvcHiZ* hiz, const NvcLimits* limits,
    uint8_t* lod_out, int threads)
{
	const NvcCullData& cd = *cull;
	uint32_t wglimit = limits ? limits->task_wglimit : NVC_TASK_WGLIMIT;
	HiZView hv = {};
	if (hiz)
		hv = HiZView{ hiz->texels, hiz->width, hiz->height, hiz->levels, hiz->level_offset };
	if (late && cd.occlusionEnabled == 1 && !hiz)
		return NVC_ERROR_INVALID_ARGUMENT;

	uint32_t n = cd.drawCount;
	int nt = std::max(1, threads);
	std::vector<std::vector<Emit>> parts(size_t(nt == 1 ? 1u : chunkCount(n, nt)));
	parallelChunks(n, nt, uint32_t(parts.size()), [&](uint32_t c, uint32_t b, uint32_t e) { drawcullRange(cd, late != 0, draws, meshes, draw_visibility, hiz ? &hv : nullptr, b, e, parts[c], lod_out); });

	// commandCount starts at 0 (vkCmdFillBuffer(dccb, 0, 4, 0) niagara.cpp:1541).  Slot of every emitted draw = exclusive
	// prefix sum in ascending di order (= the serial atomicAdd sequence); the per-thread lists are written in parallel.
	const size_t np = parts.size();
	std::vector<uint32_t> part_base(np + 1, 0);
	for (size_t t = 0; t < np; ++t)
	{
		uint64_t units = 0;
		if (task)
			for (const Emit& e : parts[t])
				units += e.groups; // :122
		else
			units = parts[t].size();
		part_base[t + 1] = uint32_t(part_base[t] + units);
	}
	const uint32_t count = part_base[np];

	if (task)
	{
		NvcMeshTaskCommand* out = static_cast<NvcMeshTaskCommand*>(commands);
		parallelChunks(uint32_t(np), nt, uint32_t(np), [&](uint32_t, uint32_t pb, uint32_t pe) {
			for (uint32_t t = pb; t < pe; ++t)
			{
				uint32_t dci = part_base[t]; // :123 atomicAdd
				for (const Emit& e : parts[t])
				{
					const NvcMeshDraw& draw = draws[e.di];
					const NvcMeshLod& lod = meshes[draw.meshIndex].lods[e.lod];
					uint32_t taskGroups = (lod.meshletCount + NVC_TASK_WGSIZE - 1) / NVC_TASK_WGSIZE;
					if (uint64_t(dci) + taskGroups <= wglimit) // :129 drop on overflow, counter still advances
						for (uint32_t i = 0; i < taskGroups; ++i)
						{
							NvcMeshTaskCommand& c = out[dci + i];
							c.drawId = e.di;
							c.taskOffset = lod.meshletOffset + i * NVC_TASK_WGSIZE;
							c.taskCount = std::min(NVC_TASK_WGSIZE, lod.meshletCount - i * NVC_TASK_WGSIZE);
							c.lateDrawVisibility = e.dv;
							c.meshletVisibilityOffset = draw.meshletVisibilityOffset + i * NVC_TASK_WGSIZE;
						}
					dci += taskGroups;
				}
			}
		});

		// tasksubmit.comp.glsl:27-47
		uint32_t clamped = std::min(count, wglimit);
		command_count4[0] = count;
		command_count4[1] = std::min((clamped + 63) / 64, NVC_MAX_DISPATCH_GROUPS);
		command_count4[2] = 64;
		command_count4[3] = 1;
		uint32_t boundary = (clamped + 63) & ~63u;
		for (uint32_t i = clamped; i < boundary; ++i)
			memset(&out[i], 0, sizeof(NvcMeshTaskCommand));
	}
	else
	{
		NvcMeshDrawCommand* out = static_cast<NvcMeshDrawCommand*>(commands);
		parallelChunks(uint32_t(np), nt, uint32_t(np), [&](uint32_t, uint32_t pb, uint32_t pe) {
			for (uint32_t t = pb; t < pe; ++t)
			{
				uint32_t dci = part_base[t];
				for (const Emit& e : parts[t])
				{
					const NvcMesh& mesh = meshes[draws[e.di].meshIndex];
					const NvcMeshLod& lod = mesh.lods[e.lod];
					NvcMeshDrawCommand& c = out[dci++]; // :143-150
					c.drawId = e.di;
					c.indexCount = lod.indexCount;
					c.instanceCount = 1;
					c.firstIndex = lod.indexOffset;
					c.vertexOffset = mesh.vertexOffset;
					c.firstInstance = 0;
				}
			}
		});
		command_count4[0] = count;
		command_count4[1] = command_count4[2] = command_count4[3] = 0;
	}
	return NVC_OK;
}

// ---- clustercull.comp.glsl:56-149 + clustersubmit.comp.glsl:25-45 --------------------------------------------
// Processes commandId < command_count4[1] * 64 (the indirect dispatch (X,64,1) of niagara.cpp:1599).
// Output order = ascending (commandId, mgi).
int orc_clustercull(const NvcCullData* cull, int late, const NvcMeshTaskCommand* task_commands, const uint32_t* command_count4,
    const NvcMeshDraw* draws, const NvcMeshlet* meshlets, uint32_t* meshlet_visibility, uint32_t* cluster_indices,
    uint32_t* cluster_count4, const NvcHiZ* hiz, const NvcLimits* limits, int threads)
{
	const NvcCullData& cd = *cull;
	uint32_t climit = limits ? limits->cluster_limit : NVC_CLUSTER_LIMIT;
	HiZView hv = {};
	if (hiz)
		hv = HiZView{ hiz->texels, hiz->width, hiz->height, hiz->levels, hiz->level_offset };
	if (late && cd.clusterOcclusionEnabled == 1 && !hiz)
		return NVC_ERROR_INVALID_ARGUMENT;

	uint32_t ncmd = command_count4[1] * 64;
	int nt = std::max(1, threads);
	std::vector<std::vector<uint32_t>> parts(size_t(nt == 1 ? 1u : chunkCount(ncmd, nt)));
	parallelChunks(ncmd, nt, uint32_t(parts.size()), [&](uint32_t c, uint32_t b, uint32_t e) { clusterRange(cd, late != 0, task_commands, draws, meshlets, meshlet_visibility, hiz ? &hv : nullptr, b, e, parts[c]); });

	// clusterCount starts at 0 (vkCmdFillBuffer(ccb, 0, 4, 0) niagara.cpp:1586); index = position in ascending
	// (commandId, mgi) order (:135 atomicAdd), entries past CLUSTER_LIMIT are dropped (:137)
	const size_t np = parts.size();
	std::vector<uint64_t> part_base(np + 1, 0);
	for (size_t t = 0; t < np; ++t)
		part_base[t + 1] = part_base[t] + parts[t].size();
	const uint32_t count = uint32_t(part_base[np]);
	parallelChunks(uint32_t(np), nt, uint32_t(np), [&](uint32_t, uint32_t pb, uint32_t pe) {
		for (uint32_t t = pb; t < pe; ++t)
		{
			uint64_t index = part_base[t];
			for (uint32_t ci : parts[t])
			{
				if (index < climit)
					cluster_indices[index] = ci;
				++index;
			}
		}
	});

	// clustersubmit.comp.glsl:25-45
	uint32_t clamped = std::min(count, climit);
	cluster_count4[0] = count;
	cluster_count4[1] = NVC_CLUSTER_TILE;
	cluster_count4[2] = std::min((clamped + 255) / 256, NVC_MAX_DISPATCH_GROUPS);
	cluster_count4[3] = 256 / NVC_CLUSTER_TILE;
	uint32_t boundary = (clamped + 255) & ~255u;
	for (uint32_t i = clamped; i < boundary; ++i)
		cluster_indices[i] = ~0u;
	return NVC_OK;
}

// ---- meshlet.task.glsl:53-149: per-command payload compaction (ascending lane order) --------------------------
int orc_taskcull(const NvcCullData* cull, int late, const NvcMeshTaskCommand* task_commands, const uint32_t* command_count4,
    const NvcMeshDraw* draws, const NvcMeshlet* meshlets, uint32_t* meshlet_visibility, NvcMeshTaskPayload* payloads,
    uint32_t* emit_counts, const NvcHiZ* hiz, int threads)
{
	const NvcCullData& cd = *cull;
	HiZView hv = {};
	if (hiz)
		hv = HiZView{ hiz->texels, hiz->width, hiz->height, hiz->levels, hiz->level_offset };
	if (late && cd.clusterOcclusionEnabled == 1 && !hiz)
		return NVC_ERROR_INVALID_ARGUMENT;

	uint32_t ncmd = command_count4[1] * 64;
	parallelRanges(ncmd, threads, [&](int, uint32_t b, uint32_t e) {
		for (uint32_t commandId = b; commandId < e; ++commandId)
		{
			const NvcMeshTaskCommand& command = task_commands[commandId];
			uint32_t sharedCount = 0; // :71
			if (command.taskCount != 0)
			{
				const NvcMeshDraw& meshDraw = draws[command.drawId];
				for (uint32_t mgi = 0; mgi < NVC_TASK_WGSIZE; ++mgi)
					if (clusterLane(cd, late != 0, command, meshDraw, meshlets, meshlet_visibility, hiz ? &hv : nullptr, mgi))
						payloads[commandId].clusterIndices[sharedCount++] = commandId | (mgi << 24); // :137-139
			}
			emit_counts[commandId] = sharedCount; // EmitMeshTasksEXT(sharedCount, 1, 1)  :143
		}
	});
	return NVC_OK;
}

// ---- meshlet.mesh.glsl:89-105: what the mesh workgroups of vkCmdDrawMeshTasksIndirectEXT(ccb, 4) decode ----------
int orc_decode_clusters(const uint32_t* cluster_indices, const uint32_t* cluster_count4, const NvcMeshTaskCommand* task_commands,
    const NvcMeshlet* meshlets, NvcClusterRecord* records, uint32_t* stats4)
{
	uint32_t gx = cluster_count4[1], gy = cluster_count4[2], gz = cluster_count4[3];
	stats4[0] = stats4[1] = stats4[2] = stats4[3] = 0;
	for (uint32_t y = 0; y < gy; ++y)
		for (uint32_t z = 0; z < gz; ++z)
			for (uint32_t x = 0; x < gx; ++x)
			{
				uint32_t slot = x + y * 256 + z * NVC_CLUSTER_TILE; // :94
				uint32_t ci = cluster_indices[slot];
				NvcClusterRecord r = { ~0u, ~0u, ~0u, ~0u };
				if (ci == ~0u) // :96
					stats4[1]++;
				else
				{
					const NvcMeshTaskCommand& command = task_commands[ci & 0xffffff]; // :102
					uint32_t mi = command.taskOffset + (ci >> 24);                   // :103
					if ((ci >> 24) >= command.taskCount)
						stats4[2]++;
					r.drawId = command.drawId;
					r.meshletIndex = mi;
					r.vertexCount = meshlets[mi].vertexCount;
					r.triangleCount = meshlets[mi].triangleCount;
					stats4[0]++;
					stats4[3] += r.triangleCount;
				}
				if (records)
					records[slot] = r;
			}
	return NVC_OK;
}

// ---- depthreduce.comp.glsl:14-22 + niagara.cpp:1703-1733: level i = MIN-sample of level i-1 (depth for i = 0)
int orc_depth_pyramid(const float* depth, uint32_t depth_width, uint32_t depth_height, const NvcHiZ* hiz, int threads)
{
	const float* src = depth;
	uint32_t sw = depth_width, sh = depth_height;
	for (uint32_t l = 0; l < hiz->levels; ++l)
	{
		uint32_t lw = std::max(1u, hiz->width >> l), lh = std::max(1u, hiz->height >> l);
		float* dst = hiz->texels + hiz->level_offset[l];
		parallelRanges(lh, threads, [&](int, uint32_t y0, uint32_t y1) {
			for (uint32_t y = y0; y < y1; ++y)
				for (uint32_t x = 0; x < lw; ++x)
				{
					// texture(..., (vec2(pos) + vec2(0.5)) / imageSize)   depthreduce.comp.glsl:19
					float u = (float(x) + 0.5f) / float(lw);
					float v = (float(y) + 0.5f) / float(lh);
					dst[size_t(y) * lw + x] = sampleMin(src, sw, sh, u, v);
				}
		});
		src = dst;
		sw = lw;
		sh = lh;
	}
	return NVC_OK;
}

// ---- scalar helpers exported for unit tests (hand-computed cases, cross-check against the numpy restatement) ----
float orc_half_to_float(uint16_t h) { return halfToFloat(h); }

void orc_rotate_quat(const float v[3], const float q[4], float out[3])
{
	vec3 r = rotateQuat(vec3{ v[0], v[1], v[2] }, q);
	out[0] = r.x, out[1] = r.y, out[2] = r.z;
}

int orc_project_sphere(const float c[3], float r, float znear, float P00, float P11, float aabb[4])
{
	vec4 a = {};
	bool ok = projectSphere(vec3{ c[0], c[1], c[2] }, r, znear, P00, P11, a);
	aabb[0] = a.x, aabb[1] = a.y, aabb[2] = a.z, aabb[3] = a.w;
	return ok;
}

float orc_occlusion_mip(const float aabb[4], float pw, float ph) { return getOcclusionMip(vec4{ aabb[0], aabb[1], aabb[2], aabb[3] }, pw, ph); }

float orc_sample_min(const float* img, uint32_t w, uint32_t h, float u, float v) { return sampleMin(img, w, h, u, v); }

int orc_ceil_log2(float x) { return ceilLog2Exact(x); }

int orc_cone_cull(const float center[3], float radius, const float axis[3], float cutoff)
{
	vec3 c = { center[0], center[1], center[2] }, a = { axis[0], axis[1], axis[2] };
	return dot3(c, a) >= cutoff * length3(c) + radius;
}

void orc_transform_point(const float m[16], const float p[3], float out[3])
{
	vec3 r = transformPoint(m, vec3{ p[0], p[1], p[2] });
	out[0] = r.x, out[1] = r.y, out[2] = r.z;
}

int orc_hardware_threads(void) { return int(std::thread::hardware_concurrency()); }

// ---- host-side helpers of the CHECKER (so that bench.py's CPU arms never load the product library) --------------------
// niagara.cpp:1002-1020: meshletVisibilityOffset = running sum over draws of the max-over-LODs meshlet count
uint32_t orc_visibility_offsets(NvcMeshDraw* draws, uint32_t draw_count, const NvcMesh* meshes)
{
	uint32_t total = 0;
	for (uint32_t i = 0; i < draw_count; ++i)
	{
		const NvcMesh& mesh = meshes[draws[i].meshIndex];
		draws[i].meshletVisibilityOffset = total;
		uint32_t most = 0;
		for (uint32_t l = 0; l < mesh.lodCount && l < NVC_MAX_LODS; ++l)
			most = std::max(most, mesh.lods[l].meshletCount);
		total += most;
	}
	return total;
}

// niagara.cpp:1339-1342 + resources.cpp:280-292: pyramid = previous power of two of the depth target, full mip chain,
// packed level after level
int orc_hiz_layout(uint32_t depth_width, uint32_t depth_height, NvcHiZ* out)
{
	if (!out || depth_width == 0 || depth_height == 0)
		return NVC_ERROR_INVALID_ARGUMENT;
	auto pow2 = [](uint32_t v) { uint32_t r = 1; while (r * 2 < v) r *= 2; return r; }; // niagara.cpp:439-447 (strictly below v)
	memset(out, 0, sizeof(*out));
	out->width = pow2(depth_width);
	out->height = pow2(depth_height);
	uint32_t w = out->width, h = out->height, levels = 0, total = 0;
	while (levels < NVC_MAX_HIZ_LEVELS)
	{
		out->level_offset[levels++] = total;
		total += w * h;
		if (w == 1 && h == 1)
			break;
		w = std::max(1u, w / 2);
		h = std::max(1u, h / 2);
	}
	out->levels = levels;
	out->total_texels = total;
	return NVC_OK;
}

// niagara.cpp:424-432,1487-1516 in double precision, rounded once: view = scale(1,1,-1) * inverse(T * R(q)),
// infinite-far reverse-Z projection, normalised frustum side planes, 1-pixel LOD target, pyramid size
void orc_cull_data(const NvcCamera* camera, uint32_t screen_width, uint32_t screen_height, uint32_t draw_count, const NvcCullOptions* options, NvcCullData* out)
{
	const double x = camera->orientation[0], y = camera->orientation[1], z = camera->orientation[2], w = camera->orientation[3];
	// rotation matrix of the unit quaternion, R[row][col]
	const double R[3][3] = { { 1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y) }, { 2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x) }, { 2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y) } };
	NvcCullData cd;
	memset(&cd, 0, sizeof(cd));
	for (int r = 0; r < 3; ++r)
	{
		const double flip = r == 2 ? -1.0 : 1.0;
		double t = 0;
		for (int c = 0; c < 3; ++c)
		{
			cd.view[c * 4 + r] = float(flip * R[c][r]) + 0.0f; // (R^T)[r][c], column-major storage; + 0: no negative zeros
			t += R[c][r] * camera->position[c];
		}
		cd.view[12 + r] = float(-flip * t) + 0.0f;
	}
	cd.view[15] = 1.f;
	// the projection terms in binary32 like the reference's glm code (niagara.cpp:424-432)
	const float aspect = float(screen_width) / float(screen_height);
	const float f = 1.0f / tanf(camera->fovY / 2.0f);
	cd.P00 = f / aspect;
	cd.P11 = f;
	cd.znear = camera->znear;
	cd.zfar = options->draw_distance;
	// planes row3 + row0 = (P00, 0, 1, 0), row3 + row1 = (0, P11, 1, 0), normalised by the length of their xyz
	const float nx = sqrtf(cd.P00 * cd.P00 + 1.0f), ny = sqrtf(cd.P11 * cd.P11 + 1.0f);
	cd.frustum[0] = cd.P00 / nx;
	cd.frustum[1] = 1.0f / nx;
	cd.frustum[2] = cd.P11 / ny;
	cd.frustum[3] = 1.0f / ny;
	cd.drawCount = draw_count;
	cd.cullingEnabled = options->culling;
	cd.lodEnabled = options->lod;
	cd.occlusionEnabled = options->occlusion;
	cd.lodTarget = (2 / cd.P11) * (1.f / float(screen_height)) * float(1 << options->debug_lod_step);
	auto pow2 = [](uint32_t v) { uint32_t r = 1; while (r * 2 < v) r *= 2; return r; }; // niagara.cpp:439-447 (strictly below v)
	cd.pyramidWidth = float(pow2(screen_width));
	cd.pyramidHeight = float(pow2(screen_height));
	cd.clusterOcclusionEnabled = options->occlusion && options->cluster_occlusion && options->mesh_shading;
	*out = cd;
}

// niagara.cpp:1547-1550,1595-1596: the per-pass copy of the frame's CullData
void orc_pass_data(const NvcCullData* frame, int for_drawcull, uint32_t post_pass, NvcCullData* out)
{
	*out = *frame;
	if (for_drawcull)
		out->clusterBackfaceEnabled = post_pass == 0;
	out->postPass = post_pass;
}

} // extern "C"
