// TEST INFRASTRUCTURE: sanitizer fuzz of the host-side decoders (scene cache parse / read, meshlet / vertex / index codecs).
// Built by tests/test_scene_cache.py with g++ -fsanitize=address,undefined directly from the product sources (no CUDA needed:
// the decoders are plain C++), run over a reference-written cache with random byte flips, truncations and count edits.
// Any out-of-bounds access, overflow or leak aborts the process; the decoders may only answer OK / CORRUPT / UNSUPPORTED / INVALID.
#include "../include/niagara_cull.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint32_t rnd()
{
	rng_state ^= rng_state << 13;
	rng_state ^= rng_state >> 7;
	rng_state ^= rng_state << 17;
	return uint32_t(rng_state >> 32);
}

static int decode_all(const std::vector<unsigned char>& file, long counts[4])
{
	NvcSceneCacheInfo info;
	int st = nvc_scene_cache_parse(file.data(), file.size(), &info);
	if (st != NVC_OK)
	{
		counts[1]++;
		return st;
	}
	for (int s = 0; s < NVC_CACHE_SECTION_COUNT; ++s)
	{
		if (info.sections[s].decoded_bytes > (64u << 20))
			continue; // a fuzzed count field asking for an absurd output: the caller would refuse to allocate it
		std::vector<unsigned char> out(size_t(info.sections[s].decoded_bytes) + 1);
		st = nvc_scene_cache_read(file.data(), file.size(), &info, s, out.data(), out.size() - 1);
		if (st != NVC_OK && st != NVC_ERROR_CORRUPT && st != NVC_ERROR_UNSUPPORTED && st != NVC_ERROR_INVALID_ARGUMENT)
		{
			fprintf(stderr, "unexpected status %d for section %d\n", st, s);
			exit(3);
		}
		counts[st == NVC_OK ? 2 : 3]++;
	}
	return NVC_OK;
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
	int iterations = atoi(argv[2]);
	long counts[4] = { 0, 0, 0, 0 };
	if (decode_all(file, counts) != NVC_OK)
		return 4; // the pristine file must parse
	for (int it = 0; it < iterations; ++it)
	{
		std::vector<unsigned char> bad = file;
		switch (rnd() % 4)
		{
		case 0: // byte flips anywhere (header included)
			for (uint32_t k = 0, m = 1 + rnd() % 8; k < m; ++k)
				bad[rnd() % bad.size()] = (unsigned char)rnd();
			break;
		case 1: // byte flips in the payload only
			for (uint32_t k = 0, m = 1 + rnd() % 8; k < m; ++k)
				bad[160 + rnd() % (bad.size() - 160)] = (unsigned char)rnd();
			break;
		case 2: // truncation
			bad.resize(rnd() % bad.size());
			break;
		default: // a count / size field of the header
			bad[28 + (rnd() % 20) * 4 + rnd() % 2] = (unsigned char)rnd();
			break;
		}
		decode_all(bad, counts);
		counts[0]++;
	}
	printf("mutations %ld: rejected by parse %ld, sections decoded ok %ld, sections rejected %ld\n", counts[0], counts[1], counts[2], counts[3]);
	return 0;
}
