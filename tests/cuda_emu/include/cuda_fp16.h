// TEST INFRASTRUCTURE — the two fp16 helpers the kernels use (see cuda_runtime.h in this directory)
#pragma once
#include <stdint.h>
#include <string.h>
struct __half
{
	unsigned short bits;
};
inline __half __ushort_as_half(unsigned short b) { return __half{ b }; }
inline float __half2float(__half h) // exact IEEE binary16 -> binary32 (subnormals preserved)
{
	uint32_t s = uint32_t(h.bits & 0x8000u) << 16, e = (h.bits >> 10) & 0x1fu, m = h.bits & 0x3ffu, out;
	if (e == 0x1fu)
		out = s | 0x7f800000u | (m << 13);
	else if (e != 0)
		out = s | ((e + 112u) << 23) | (m << 13);
	else if (m == 0)
		out = s;
	else
	{
		int shift = __builtin_clz(m) - 21;
		m = (m << shift) & 0x3ffu;
		out = s | (uint32_t(113 - shift) << 23) | (m << 13);
	}
	float f;
	memcpy(&f, &out, 4);
	return f;
}
