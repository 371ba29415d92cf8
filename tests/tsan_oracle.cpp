// TEST INFRASTRUCTURE: the multi-threaded oracle (oracle/oracle.cpp) under ThreadSanitizer — two full frames over an instanced
// scene with 8 worker threads.  Built and run by tests/test_oracle_units.py::test_oracle_is_race_free.
// Input: an .nvcg geometry dump (tests/golden) + the library's own PCG32 scene; output: exit code 0 and no TSan report.
#include "../include/niagara_cull.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

extern "C"
{
int orc_drawcull(const NvcCullData*, int, int, const NvcMeshDraw*, const NvcMesh*, uint32_t*, void*, uint32_t*, const NvcHiZ*, const NvcLimits*, uint8_t*, int);
int orc_clustercull(const NvcCullData*, int, const NvcMeshTaskCommand*, const uint32_t*, const NvcMeshDraw*, const NvcMeshlet*, uint32_t*, uint32_t*, uint32_t*, const NvcHiZ*,
    const NvcLimits*, int);
int orc_depth_pyramid(const float*, uint32_t, uint32_t, const NvcHiZ*, int);
}

int main(int argc, char** argv)
{
	if (argc < 3)
		return 2;
	FILE* f = fopen(argv[1], "rb");
	if (!f)
		return 2;
	uint32_t header[8];
	if (fread(header, sizeof(header), 1, f) != 1)
		return 2;
	std::vector<NvcMesh> meshes(header[2]);
	std::vector<NvcMeshlet> meshlets(header[3]);
	if (fread(meshes.data(), sizeof(NvcMesh), meshes.size(), f) != meshes.size() || fread(meshlets.data(), sizeof(NvcMeshlet), meshlets.size(), f) != meshlets.size())
		return 2;
	fclose(f);

	const uint32_t n = uint32_t(atoi(argv[2])), width = 640, height = 480;
	std::vector<NvcMeshDraw> draws(n);
	// a deterministic cloud of draws in front of the camera (the camera looks down -Z)
	uint32_t state = 12345u;
	auto rnd = [&]() { state = state * 1664525u + 1013904223u; return float(state >> 8) / float(1 << 24); };
	uint32_t bits = 0;
	for (uint32_t i = 0; i < n; ++i)
	{
		NvcMeshDraw& d = draws[i];
		memset(&d, 0, sizeof(d));
		d.position[0] = rnd() * 80.f - 40.f;
		d.position[1] = rnd() * 50.f - 25.f;
		d.position[2] = -(5.f + rnd() * 150.f);
		d.scale = 1.f + rnd() * 3.f;
		d.orientation[3] = 1.f;
		d.meshIndex = i % uint32_t(meshes.size());
		d.meshletVisibilityOffset = bits;
		uint32_t most = 0;
		for (uint32_t l = 0; l < meshes[d.meshIndex].lodCount; ++l)
			most = meshes[d.meshIndex].lods[l].meshletCount > most ? meshes[d.meshIndex].lods[l].meshletCount : most;
		bits += most;
	}
	NvcCamera cam = { { 0, 0, 0 }, { 0, 0, 0, 1 }, 1.2217305f, 0.1f };
	NvcCullOptions opt = { 200.f, 1, 1, 1, 1, 1, 0 };
	NvcCullData cd;
	nvc_host_cull_data(&cam, width, height, n, &opt, &cd, nullptr);
	NvcHiZ hiz;
	nvc_hiz_layout(width, height, &hiz);
	std::vector<float> pyramid(hiz.total_texels), depth(size_t(width) * height);
	for (size_t i = 0; i < depth.size(); ++i)
		depth[i] = (i * 2654435761u % 97u) < 30u ? 0.002f : 0.f;
	hiz.texels = pyramid.data();
	std::vector<uint32_t> dvb(n, 0), mvb((bits + 31) / 32 + 1, 0), cib(size_t(n) * 8 * 64 + 256), dccb(4), ccb(4);
	std::vector<NvcMeshTaskCommand> dcb(size_t(n) * 8 + 64);
	const int threads = 8;
	unsigned long long emitted = 0;
	for (int frame = 0; frame < 2; ++frame)
		for (int late = 0; late < 2; ++late)
		{
			NvcCullData pass;
			nvc_host_pass_data(&cd, 1, 0, &pass);
			if (late)
				orc_depth_pyramid(depth.data(), width, height, &hiz, threads);
			if (orc_drawcull(&pass, late, 1, draws.data(), meshes.data(), dvb.data(), dcb.data(), dccb.data(), &hiz, nullptr, nullptr, threads) != 0)
				return 3;
			nvc_host_pass_data(&cd, 0, 0, &pass);
			pass.clusterBackfaceEnabled = 1;
			if (orc_clustercull(&pass, late, dcb.data(), dccb.data(), draws.data(), meshlets.data(), mvb.data(), cib.data(), ccb.data(), &hiz, nullptr, threads) != 0)
				return 3;
			emitted += ccb[0];
		}
	printf("emitted %llu clusters\n", emitted);
	return emitted > 1000 ? 0 : 4;
}
