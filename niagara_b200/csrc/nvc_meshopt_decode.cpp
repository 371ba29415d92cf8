// nvc_meshopt_decode.cpp — decoders for the two remaining meshopt streams of a compressed scene cache (row N2):
// the vertex codec (scenecache.cpp:64-72,244-248: vertices and the RT position stream) and the index codec
// (:74-82,250-254).  Host only.  Both are lossless, so the check is simple: decoding the compressed cache the
// reference wrote must give the bytes of the uncompressed cache it wrote (tests/test_scene_cache.py).
//
// Formats (meshoptimizer v1.0, extern/meshoptimizer/src/vertexcodec.cpp, indexcodec.cpp — restated, not copied):
//
// VERTEX stream   [0xa0 | version] blocks... tail
//   tail  = the last max(tail, 24 (v1) / 32 (v0)) bytes; its final `stride` (+ stride/4 for v1) bytes hold the
//           baseline vertex and, for v1, one channel byte per 4-byte group
//   block = min(256, (8192 / stride) & ~15) vertices; [v1: stride/4 control bytes]; then per 4-byte group its four
//           byte planes, each: control 3 = `count` literal bytes, 2 = all zero, 0/1 = byte groups
//   byte groups: 2-bit selector per 16 values ((groups + 3) / 4 header bytes) into the width table
//           v0 {0,2,4,8}, v1 control 0 {0,1,2,4}, control 1 {1,2,4,8}; width 0 = zeros, 8 = 16 literals, else 16 packed
//           values (1-bit: LSB first, 2/4-bit: MSB first) where the all-ones value means "next escape byte"
//   deltas per 4-byte group by channel & 3: 0 = per byte zigzag delta, 1 = per 16-bit zigzag delta,
//           2 = 32-bit xor with the previous value after rotating left by (32 - (channel >> 4)) & 31
//
// INDEX stream    [0xe0 | version] one code byte per triangle, data bytes, 16-byte aux table at the very end
//   code < 0xf0 : high nibble = edge FIFO slot (a, b), low nibble = c: 0 next new vertex, 1..12 (v1) / 1..14 (v0)
//                 vertex FIFO slot, 13 / 14 (v1) = last -/+ 1, 15 = explicit zigzag varint delta from `last`
//   0xf0..0xfd  : restart, (b, c) slots from aux[code & 15]; a = next new vertex
//   0xfe / 0xff : restart with an explicit aux byte (0xff: a explicit too); aux == 0 resets the vertex counter
#include "../../include/niagara_cull.h"

#include <string.h>

namespace nvc
{

namespace
{

struct ByteReader
{
	const uint8_t* at;
	const uint8_t* end;
	bool take(size_t n, const uint8_t*& out)
	{
		if (size_t(end - at) < n)
			return false;
		out = at;
		at += n;
		return true;
	}
};

// 16 values of one byte plane
bool read_byte_group(ByteReader& r, int width, uint8_t* out)
{
	const uint8_t* packed;
	if (width == 0)
	{
		memset(out, 0, 16);
		return true;
	}
	if (width == 8)
	{
		if (!r.take(16, packed))
			return false;
		memcpy(out, packed, 16);
		return true;
	}
	if (!r.take(size_t(2 * width), packed))
		return false;
	// unpack without escapes first; a value equal to the all-ones code is then replaced by the next escape byte, in order
	const uint8_t escape = uint8_t((1u << width) - 1u);
	if (width == 4)
		for (int i = 0; i < 8; ++i)
		{
			out[2 * i] = packed[i] >> 4;
			out[2 * i + 1] = packed[i] & 15u;
		}
	else if (width == 2)
		for (int i = 0; i < 4; ++i)
		{
			uint8_t b = packed[i];
			out[4 * i] = b >> 6;
			out[4 * i + 1] = (b >> 4) & 3u;
			out[4 * i + 2] = (b >> 2) & 3u;
			out[4 * i + 3] = b & 3u;
		}
	else // width 1: least significant bit first
		for (int i = 0; i < 16; ++i)
			out[i] = (packed[i >> 3] >> (i & 7)) & 1u;
	for (int i = 0; i < 16; ++i)
		if (out[i] == escape)
		{
			const uint8_t* e;
			if (!r.take(1, e))
				return false;
			out[i] = *e;
		}
	return true;
}

bool read_plane(ByteReader& r, int control, int version, uint32_t count, uint8_t* out /* room for count rounded up to 16 */)
{
	if (control == 3)
	{
		const uint8_t* lit;
		if (!r.take(count, lit))
			return false;
		memcpy(out, lit, count);
		return true;
	}
	if (control == 2)
	{
		memset(out, 0, count);
		return true;
	}
	static const int kWidthsV0[4] = { 0, 2, 4, 8 };
	static const int kWidthsV1[5] = { 0, 1, 2, 4, 8 };
	const int* widths = version == 0 ? kWidthsV0 : kWidthsV1 + control;
	uint32_t groups = (count + 15u) / 16u;
	const uint8_t* selectors;
	if (!r.take((groups + 3u) / 4u, selectors))
		return false;
	for (uint32_t g = 0; g < groups; ++g)
	{
		int sel = (selectors[g / 4] >> ((g % 4) * 2)) & 3;
		if (!read_byte_group(r, widths[sel], out + g * 16))
			return false;
	}
	return true;
}

inline uint32_t rotl(uint32_t v, int r) { return (v << r) | (v >> ((32 - r) & 31)); }

} // namespace

// 0 ok, negative NvcStatus otherwise
int decode_vertex_stream(const uint8_t* stream, size_t stream_size, uint32_t vertex_count, uint32_t stride, uint8_t* out)
{
	if (stride == 0 || stride > 256 || stride % 4 != 0)
		return NVC_ERROR_INVALID_ARGUMENT;
	if (stream_size < 1 || (stream[0] & 0xf0) != 0xa0)
		return NVC_ERROR_CORRUPT;
	const int version = stream[0] & 0x0f;
	if (version > 1)
		return NVC_ERROR_UNSUPPORTED;
	const size_t tail = stride + (version == 0 ? 0 : stride / 4);
	const size_t tail_min = version == 0 ? 32 : 24;
	const size_t tail_padded = tail < tail_min ? tail_min : tail;
	if (stream_size - 1 < tail_padded)
		return NVC_ERROR_CORRUPT;

	uint8_t previous[256];
	memcpy(previous, stream + stream_size - tail, stride);
	const uint8_t* channels = version == 0 ? nullptr : stream + stream_size - tail + stride;

	ByteReader r = { stream + 1, stream + stream_size - tail_padded };
	uint32_t per_block = uint32_t((8192u / stride) & ~15u);
	if (per_block > 256)
		per_block = 256;

	uint8_t planes[4][256 + 16];
	for (uint32_t first = 0; first < vertex_count; first += per_block)
	{
		uint32_t count = vertex_count - first < per_block ? vertex_count - first : per_block;
		const uint8_t* control = nullptr;
		if (version != 0 && !r.take(stride / 4, control))
			return NVC_ERROR_CORRUPT;
		uint8_t* block = out + size_t(first) * stride;
		for (uint32_t k = 0; k < stride; k += 4)
		{
			uint32_t ctrl = control ? control[k / 4] : 0u;
			for (int j = 0; j < 4; ++j)
				if (!read_plane(r, int((ctrl >> (j * 2)) & 3u), version, count, planes[j]))
					return NVC_ERROR_CORRUPT;
			int channel = channels ? channels[k / 4] : 0;
			switch (channel & 3)
			{
			case 0:
				for (int j = 0; j < 4; ++j)
				{
					uint8_t p = previous[k + j];
					for (uint32_t i = 0; i < count; ++i)
					{
						uint8_t v = planes[j][i];
						p = uint8_t(((0u - (v & 1u)) ^ (v >> 1)) + p);
						block[size_t(i) * stride + k + j] = p;
					}
				}
				break;
			case 1:
				for (int half = 0; half < 2; ++half)
				{
					uint16_t p = uint16_t(previous[k + half * 2] | (previous[k + half * 2 + 1] << 8));
					for (uint32_t i = 0; i < count; ++i)
					{
						uint16_t v = uint16_t(planes[half * 2][i] | (planes[half * 2 + 1][i] << 8));
						p = uint16_t(((0u - (v & 1u)) ^ (v >> 1)) + p);
						block[size_t(i) * stride + k + half * 2] = uint8_t(p);
						block[size_t(i) * stride + k + half * 2 + 1] = uint8_t(p >> 8);
					}
				}
				break;
			case 2:
			{
				int rot = (32 - (channel >> 4)) & 31;
				uint32_t p = uint32_t(previous[k]) | (uint32_t(previous[k + 1]) << 8) | (uint32_t(previous[k + 2]) << 16) | (uint32_t(previous[k + 3]) << 24);
				for (uint32_t i = 0; i < count; ++i)
				{
					uint32_t v = uint32_t(planes[0][i]) | (uint32_t(planes[1][i]) << 8) | (uint32_t(planes[2][i]) << 16) | (uint32_t(planes[3][i]) << 24);
					p = rotl(v, rot) ^ p;
					uint8_t* dst = block + size_t(i) * stride + k;
					dst[0] = uint8_t(p);
					dst[1] = uint8_t(p >> 8);
					dst[2] = uint8_t(p >> 16);
					dst[3] = uint8_t(p >> 24);
				}
				break;
			}
			default:
				return NVC_ERROR_CORRUPT;
			}
		}
		memcpy(previous, block + size_t(count - 1) * stride, stride);
	}
	return r.at == r.end ? NVC_OK : NVC_ERROR_CORRUPT;
}

namespace
{

struct Ring16
{
	uint32_t slot[16];
	uint32_t head = 0;
	Ring16() { memset(slot, 0xff, sizeof(slot)); }
	// the slot is written even when the head does not advance — the encoder does the same, and both must agree
	void put(uint32_t v, bool advance = true)
	{
		slot[head] = v;
		head = (head + (advance ? 1u : 0u)) & 15u;
	}
	uint32_t back(uint32_t n) const { return slot[(head - n) & 15u]; } // back(1) = most recent
};

bool read_varint_delta(ByteReader& r, uint32_t& last)
{
	uint32_t value = 0;
	for (int i = 0; i < 5; ++i)
	{
		const uint8_t* b;
		if (!r.take(1, b))
			return false;
		value |= uint32_t(*b & 127u) << (7 * i);
		if (*b < 128)
			break;
	}
	last += (value >> 1) ^ (0u - (value & 1u));
	return true;
}

} // namespace

int decode_index_stream(const uint8_t* stream, size_t stream_size, uint32_t index_count, uint32_t* out)
{
	if (index_count % 3 != 0)
		return NVC_ERROR_INVALID_ARGUMENT;
	const uint32_t triangles = index_count / 3;
	if (stream_size < size_t(1) + triangles + 16)
		return NVC_ERROR_CORRUPT;
	if ((stream[0] & 0xf0) != 0xe0)
		return NVC_ERROR_CORRUPT;
	const int version = stream[0] & 0x0f;
	if (version > 1)
		return NVC_ERROR_UNSUPPORTED;
	const uint32_t direct_slots = version >= 1 ? 13u : 15u;

	const uint8_t* codes = stream + 1;
	const uint8_t* aux = stream + stream_size - 16;
	ByteReader r = { codes + triangles, aux };

	Ring16 verts, edge_a, edge_b;
	auto push_edge = [&](uint32_t a, uint32_t b) {
		edge_a.put(a);
		edge_b.put(b);
	};
	uint32_t next = 0, last = 0;

	for (uint32_t t = 0; t < triangles; ++t)
	{
		const uint32_t code = codes[t];
		uint32_t a, b, c;
		if (code < 0xf0)
		{
			uint32_t e = (code >> 4) + 1u;
			a = edge_a.back(e);
			b = edge_b.back(e);
			uint32_t sel = code & 15u;
			if (sel < direct_slots)
			{
				c = sel == 0 ? next : verts.back(sel + 1u);
				next += sel == 0 ? 1u : 0u;
				verts.put(c, sel == 0);
			}
			else
			{
				if (sel == 15u)
				{
					if (!read_varint_delta(r, last))
						return NVC_ERROR_CORRUPT;
				}
				else
					last += sel == 13u ? 0xffffffffu : 1u; // 13 -> -1, 14 -> +1
				c = last;
				verts.put(c);
			}
			push_edge(c, b);
			push_edge(a, c);
		}
		else
		{
			uint32_t sel_a = 0, sel_b, sel_c;
			if (code < 0xfe)
			{
				uint32_t x = aux[code & 15u];
				sel_b = x >> 4;
				sel_c = x & 15u;
			}
			else
			{
				const uint8_t* x;
				if (!r.take(1, x))
					return NVC_ERROR_CORRUPT;
				sel_a = code == 0xfe ? 0u : 15u;
				sel_b = *x >> 4;
				sel_c = *x & 15u;
				if (*x == 0)
					next = 0;
			}
			// FIFO lookups use the ring as it is BEFORE this triangle's vertices are pushed
			a = sel_a == 0 ? next++ : 0u;
			b = sel_b == 0 ? next++ : verts.back(sel_b);
			c = sel_c == 0 ? next++ : verts.back(sel_c);
			if (sel_a == 15u)
			{
				if (!read_varint_delta(r, last))
					return NVC_ERROR_CORRUPT;
				a = last;
			}
			if (sel_b == 15u)
			{
				if (!read_varint_delta(r, last))
					return NVC_ERROR_CORRUPT;
				b = last;
			}
			if (sel_c == 15u)
			{
				if (!read_varint_delta(r, last))
					return NVC_ERROR_CORRUPT;
				c = last;
			}
			verts.put(a);
			verts.put(b, sel_b == 0 || sel_b == 15u);
			verts.put(c, sel_c == 0 || sel_c == 15u);
			push_edge(b, a);
			push_edge(c, b);
			push_edge(a, c);
		}
		out[size_t(t) * 3 + 0] = a;
		out[size_t(t) * 3 + 1] = b;
		out[size_t(t) * 3 + 2] = c;
	}
	return r.at == r.end ? NVC_OK : NVC_ERROR_CORRUPT;
}

} // namespace nvc
