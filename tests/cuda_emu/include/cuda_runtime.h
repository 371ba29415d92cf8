// TEST INFRASTRUCTURE — a CPU stand-in for the CUDA runtime + SIMT execution model, so that the PRODUCT's kernel source
// (niagara_b200/csrc/nvc_kernels.cu, nvc_api.cu) can be compiled by g++ and executed on the host when no GPU is at hand
// (tests/test_kernels_emulated.py).  It checks kernel LOGIC (indexing, scans, compaction, epilogues, arithmetic order), not
// hardware behaviour; the real parity tests are the -m gpu ones.
//
// Execution model (tests/cuda_emu/emu.cpp): blocks run one after another on one OS thread; the threads of a block are fibers
// (ucontext) resumed round-robin.  Warp collectives (__shfl_sync, __ballot_sync, ...) and __syncthreads are rendezvous
// points: a fiber yields until every participating lane / thread has arrived.  Device memory is host memory, atomics are
// plain read-modify-writes (single OS thread), the __f*_rn intrinsics are single IEEE operations (build with
// -ffp-contract=off), __ldg & co. are plain loads.  Kernel launches are rewritten by gen_emu.py into emu::launch(...).
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

#define NVC_EMU 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static // one block at a time: a function-level static is that block's shared memory

struct uint3
{
	unsigned int x, y, z;
};
struct dim3
{
	unsigned int x, y, z;
	dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2
{
	float x, y;
};
struct alignas(16) float4
{
	float x, y, z, w;
};
struct uint2
{
	unsigned int x, y;
};
struct alignas(16) uint4
{
	unsigned int x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }
inline float2 make_float2(float x, float y) { return float2{ x, y }; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{ x, y, z, w }; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{ x, y }; }

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

// ---- runtime API subset --------------------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum
{
	cudaSuccess = 0,
	cudaErrorMemoryAllocation = 2,
	cudaDevAttrMultiProcessorCount = 16
};
inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 4; return cudaSuccess; } // "4 SMs": small persistent grids
template <typename T>
inline cudaError_t cudaMalloc(T** p, size_t n)
{
	*p = static_cast<T*>(calloc(1, n ? n : 1));
	return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
enum { cudaMemcpyDeviceToHost = 2, cudaMemcpyHostToDevice = 1 };
inline cudaError_t cudaMemcpy(void* d, const void* s_, size_t n, int) { memcpy(d, s_, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
template <typename K>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* blocks, K, int, size_t) { *blocks = 2; return cudaSuccess; }

// ---- SIMT emulation ----------------------------------------------------------------------------------------------
namespace emu
{
void launch(dim3 grid, dim3 block, size_t dynamic_smem, const std::function<void()>& body);
void* dynamic_smem();
const uint32_t* warp_gather(uint32_t mask, uint32_t value); // rendezvous of the lanes in `mask`; returns the 32 lane values
void block_barrier();
uint32_t lane();
} // namespace emu

inline void __syncthreads() { emu::block_barrier(); }
inline void __syncwarp(uint32_t mask = 0xffffffffu) { emu::warp_gather(mask, 0); }
inline void __threadfence() {}
inline void __threadfence_block() {}

template <typename T>
inline T __shfl_sync(uint32_t mask, T v, int src)
{
	static_assert(sizeof(T) == 4, "32-bit shuffles only");
	uint32_t bits;
	memcpy(&bits, &v, 4);
	uint32_t out = emu::warp_gather(mask, bits)[uint32_t(src) & 31u];
	T r;
	memcpy(&r, &out, 4);
	return r;
}
template <typename T>
inline T __shfl_up_sync(uint32_t mask, T v, unsigned delta)
{
	uint32_t l = emu::lane();
	uint32_t bits;
	memcpy(&bits, &v, 4);
	const uint32_t* all = emu::warp_gather(mask, bits);
	uint32_t out = l >= delta ? all[l - delta] : bits;
	T r;
	memcpy(&r, &out, 4);
	return r;
}
template <typename T>
inline T __shfl_down_sync(uint32_t mask, T v, unsigned delta)
{
	uint32_t l = emu::lane();
	uint32_t bits;
	memcpy(&bits, &v, 4);
	const uint32_t* all = emu::warp_gather(mask, bits);
	uint32_t out = l + delta < 32 ? all[l + delta] : bits;
	T r;
	memcpy(&r, &out, 4);
	return r;
}
template <typename T>
inline T __shfl_xor_sync(uint32_t mask, T v, int lanemask)
{
	uint32_t l = emu::lane();
	uint32_t bits;
	memcpy(&bits, &v, 4);
	uint32_t out = emu::warp_gather(mask, bits)[(l ^ uint32_t(lanemask)) & 31u];
	T r;
	memcpy(&r, &out, 4);
	return r;
}
inline uint32_t __ballot_sync(uint32_t mask, int pred)
{
	const uint32_t* all = emu::warp_gather(mask, pred ? 1u : 0u);
	uint32_t r = 0;
	for (int i = 0; i < 32; ++i)
		if ((mask >> i) & 1u)
			r |= (all[i] & 1u) << i;
	return r;
}
inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) == mask; }
inline uint32_t __reduce_or_sync(uint32_t mask, uint32_t v)
{
	const uint32_t* all = emu::warp_gather(mask, v);
	uint32_t r = 0;
	for (int i = 0; i < 32; ++i)
		if ((mask >> i) & 1u)
			r |= all[i];
	return r;
}
inline uint32_t __match_any_sync(uint32_t mask, uint32_t v)
{
	uint32_t l = emu::lane();
	const uint32_t* all = emu::warp_gather(mask, v);
	uint32_t r = 0;
	for (int i = 0; i < 32; ++i)
		if (((mask >> i) & 1u) && all[i] == all[l])
			r |= 1u << i;
	return r;
}
inline uint32_t __activemask() { return 0xffffffffu; }

// ---- memory ----------------------------------------------------------------------------------------------------------
template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline T __ldcg(const T* p) { return *p; }
template <typename T>
inline T __ldcs(const T* p) { return *p; }
template <typename T>
inline T __ldca(const T* p) { return *p; }
template <typename T>
inline void __stcs(T* p, T v) { *p = v; }
template <typename T>
inline void __stcg(T* p, T v) { *p = v; }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
inline uint32_t atomicOr(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
inline uint32_t atomicAnd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o & v; return o; }
inline int atomicMax(int* p, int v) { int o = *p; *p = o > v ? o : v; return o; }
inline uint32_t atomicMax(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o > v ? o : v; return o; }
inline uint32_t atomicMin(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o < v ? o : v; return o; }
inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }

// ---- arithmetic intrinsics: single IEEE operations (the TU is built with -ffp-contract=off) -------------------------------
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz(uint32_t(v)) : 32; }
inline uint32_t __brev(uint32_t v)
{
	uint32_t r = 0;
	for (int i = 0; i < 32; ++i)
		r |= ((v >> i) & 1u) << (31 - i);
	return r;
}
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return uint32_t((uint64_t(a) * b) >> 32); }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) { return uint32_t(((uint64_t(hi) << 32) | lo) >> (shift & 31u)); }
inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t shift) { return uint32_t((((uint64_t(hi) << 32) | lo) << (shift & 31u)) >> 32); }
inline void __nanosleep(unsigned) {}
using ::fmaxf;
using ::fminf;
template <typename T>
inline T min(T a, T b) { return a < b ? a : b; }
template <typename T>
inline T max(T a, T b) { return a > b ? a : b; }
inline uint32_t min(uint32_t a, int b) { return a < uint32_t(b) ? a : uint32_t(b); }
inline uint32_t min(int a, uint32_t b) { return uint32_t(a) < b ? uint32_t(a) : b; }
inline uint32_t max(uint32_t a, int b) { return a > uint32_t(b) ? a : uint32_t(b); }
