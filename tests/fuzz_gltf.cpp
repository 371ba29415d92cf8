// TEST INFRASTRUCTURE: sanitizer fuzz of the glTF importer (niagara_b200/csrc/nvc_gltf.cpp), built by tests/test_gltf_import.py with
// g++ -fsanitize=address,undefined from the product source.  Random byte flips, truncations and JSON-token edits of a valid file;
// every primitive of an accepted scene is extracted.  Any invalid access aborts; only documented status codes may come back.
#include "../include/niagara_cull.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

static uint64_t rng_state = 0x243f6a8885a308d3ull;
static uint32_t rnd()
{
	rng_state ^= rng_state << 13;
	rng_state ^= rng_state >> 7;
	rng_state ^= rng_state << 17;
	return uint32_t(rng_state >> 32);
}

int main(int argc, char** argv)
{
	if (argc < 3)
		return 2;
	FILE* f = fopen(argv[1], "rb");
	if (!f)
		return 2;
	std::vector<unsigned char> file;
	unsigned char buf[65536];
	size_t n;
	while ((n = fread(buf, 1, sizeof(buf), f)) > 0)
		file.insert(file.end(), buf, buf + n);
	fclose(f);
	const int iterations = atoi(argv[2]);
	long accepted = 0, rejected = 0, extracted = 0;
	for (int it = 0; it < iterations; ++it)
	{
		std::vector<unsigned char> m = file;
		switch (rnd() % 4)
		{
		case 0: // byte flips
			for (uint32_t k = 0, e = 1 + rnd() % 8; k < e; ++k)
				m[rnd() % m.size()] = (unsigned char)rnd();
			break;
		case 1: // truncation
			m.resize(rnd() % m.size());
			break;
		case 2: // digit edits inside the JSON chunk (counts, offsets, indices)
			for (uint32_t k = 0, e = 1 + rnd() % 6; k < e; ++k)
			{
				size_t at = 20 + rnd() % (m.size() > 4000 ? 3800 : m.size() - 21);
				if (m[at] >= '0' && m[at] <= '9')
					m[at] = (unsigned char)('0' + rnd() % 10);
			}
			break;
		default: // splice a random window of the file over another place
		{
			size_t len = 1 + rnd() % 64, a = rnd() % (m.size() - len), b = rnd() % (m.size() - len);
			memmove(&m[a], &m[b], len);
			break;
		}
		}
		if (m.empty())
			continue;
		NvcGltfScene* scene = nullptr;
		int st = nvc_gltf_import(m.data(), m.size(), "/nonexistent", 0, 1, &scene);
		if (st != NVC_OK)
		{
			if (st != NVC_ERROR_CORRUPT && st != NVC_ERROR_UNSUPPORTED && st != NVC_ERROR_INVALID_ARGUMENT)
			{
				fprintf(stderr, "unexpected status %d\n", st);
				return 3;
			}
			++rejected;
			continue;
		}
		++accepted;
		NvcGltfInfo info;
		nvc_gltf_info(scene, &info);
		std::vector<NvcMeshDraw> draws(info.draw_count + 1);
		std::vector<NvcAnimation> anims(info.animation_count + 1);
		std::vector<NvcKeyframe> keys(info.keyframe_count + 1);
		std::vector<float> scale(info.primitive_count + 1);
		nvc_gltf_scene_arrays(scene, draws.data(), anims.data(), keys.data(), scale.data());
		for (uint32_t p = 0; p < info.primitive_count; ++p)
		{
			uint32_t vc = 0, ic = 0;
			if (nvc_gltf_primitive_size(scene, p, &vc, &ic) != NVC_OK || vc > (1u << 22) || ic > (1u << 24))
				continue; // an edited count asking for an absurd output: the caller would refuse to allocate it
			std::vector<NvcVertex> v(vc + 1);
			std::vector<uint32_t> ix(ic + 1);
			st = nvc_gltf_primitive_data(scene, p, v.data(), ix.data());
			if (st != NVC_OK && st != NVC_ERROR_CORRUPT)
			{
				fprintf(stderr, "unexpected primitive status %d\n", st);
				return 3;
			}
			extracted += st == NVC_OK;
		}
		nvc_gltf_free(scene);
	}
	printf("mutations %d accepted %ld rejected %ld primitives extracted %ld\n", iterations, accepted, rejected, extracted);
	return 0;
}
