// examples/frame_loop.cpp — the reference's frame loop for the visibility path, host side in C++ (as in niagara),
// with the Vulkan compute dispatches replaced by the C ABI of include/niagara_cull.h.  No Python, no torch.
//
// Mirrors src/niagara.cpp: random scene (969-998) -> visibility offsets (1002-1020) -> buffers (1027-1090) ->
// per frame CullData (1487-1516) and the pass order cull / clusters / pyramid / late cull / late clusters (1765-1788).
// Geometry is synthetic here (the real program would pass the arrays scene.cpp cooked).
//
// build: make -C examples        run: examples/frame_loop [draws] [frames]
#include "../include/niagara_cull.h"

#include <cuda_runtime.h>

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CHECK_CUDA(x) \
	do \
	{ \
		cudaError_t e_ = (x); \
		if (e_ != cudaSuccess) \
		{ \
			fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
			exit(1); \
		} \
	} while (0)

#define CHECK_NVC(ctx, x) \
	do \
	{ \
		int s_ = (x); \
		if (s_ != NVC_OK) \
		{ \
			fprintf(stderr, "%s:%d: %s (%s)\n", __FILE__, __LINE__, nvc_status_string(s_), nvc_last_error(ctx)); \
			exit(1); \
		} \
	} while (0)

template <typename T>
static T* upload(const std::vector<T>& v)
{
	T* d = nullptr;
	CHECK_CUDA(cudaMalloc(&d, v.size() * sizeof(T) + 16));
	CHECK_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
	return d;
}

static uint32_t lcg(uint32_t& s)
{
	s = s * 1664525u + 1013904223u;
	return s >> 8;
}

int main(int argc, char** argv)
{
	uint32_t drawCount = argc > 1 ? uint32_t(atoi(argv[1])) : 1000000;
	int frames = argc > 2 ? atoi(argv[2]) : 10;
	const uint32_t meshCount = 1024, lodCount = 4, lod0Meshlets = 64;
	const uint32_t width = 1920, height = 1080;

	// ---- geometry tables: Mesh[] with a LOD chain, Meshlet[] with cull bounds ----
	std::vector<NvcMesh> meshes(meshCount);
	std::vector<NvcMeshlet> meshlets;
	uint32_t seed = 7;
	for (uint32_t m = 0; m < meshCount; ++m)
	{
		NvcMesh& mesh = meshes[m];
		memset(&mesh, 0, sizeof(mesh));
		mesh.radius = 0.3f + 0.7f * float(lcg(seed) & 0xffff) / 65535.f;
		mesh.lodCount = lodCount;
		for (uint32_t l = 0; l < lodCount; ++l)
		{
			uint32_t n = lod0Meshlets >> l;
			mesh.lods[l].meshletOffset = uint32_t(meshlets.size());
			mesh.lods[l].meshletCount = n;
			mesh.lods[l].indexCount = n * 96 * 3;
			mesh.lods[l].error = (l == 0 ? 0.f : 5e-4f * float(1 << (2 * (l - 1)))) * mesh.radius;
			for (uint32_t i = 0; i < n; ++i)
			{
				NvcMeshlet ml;
				memset(&ml, 0, sizeof(ml));
				ml.center[0] = ml.center[1] = ml.center[2] = 0x3000 + uint16_t(lcg(seed) & 0x3ff); // ~0.125..0.25 as binary16
				ml.radius = 0x2e66;                                                                 // ~0.1
				ml.cone_axis[0] = int8_t(lcg(seed) % 255 - 127);
				ml.cone_axis[1] = int8_t(lcg(seed) % 255 - 127);
				ml.cone_axis[2] = int8_t(lcg(seed) % 255 - 127);
				ml.cone_cutoff = int8_t(16 + lcg(seed) % 112);
				meshlets.push_back(ml);
			}
		}
	}

	// ---- scene: niagara.cpp:969-998 + 1002-1020 ----
	std::vector<NvcMeshDraw> draws(drawCount);
	nvc_host_random_draws(draws.data(), drawCount, meshCount, 300.f);
	uint32_t postMask = 0;
	uint32_t visibilityBits = nvc_host_visibility_offsets(draws.data(), drawCount, meshes.data(), &postMask);

	NvcContext* ctx = nullptr;
	int s = nvc_create(0, nullptr, &ctx);
	if (s != NVC_OK)
	{
		fprintf(stderr, "nvc_create: %s\n", nvc_status_string(s));
		return 1;
	}
	cudaStream_t stream;
	CHECK_CUDA(cudaStreamCreate(&stream));

	// ---- buffers: niagara.cpp:1027-1090 (same names) ----
	NvcMeshDraw* db = upload(draws);
	NvcMesh* mb = upload(meshes);
	NvcMeshlet* mlb = upload(meshlets);
	uint32_t *dvb, *dccb, *mvb, *cib, *ccb;
	void* dcb;
	CHECK_CUDA(cudaMalloc(&dvb, drawCount * 4));
	CHECK_CUDA(cudaMemset(dvb, 0, drawCount * 4)); // niagara.cpp:1450-1458 first-frame clear
	CHECK_CUDA(cudaMalloc(&dcb, size_t(NVC_TASK_WGLIMIT) * sizeof(NvcMeshTaskCommand)));
	CHECK_CUDA(cudaMalloc(&dccb, 16));
	size_t mvbBytes = (visibilityBits + 31) / 32 * 4;
	CHECK_CUDA(cudaMalloc(&mvb, mvbBytes));
	CHECK_CUDA(cudaMemset(mvb, 0, mvbBytes)); // niagara.cpp:1460-1468
	CHECK_CUDA(cudaMalloc(&cib, size_t(NVC_CLUSTER_LIMIT) * 4));
	CHECK_CUDA(cudaMalloc(&ccb, 16));

	// depth target stand-in + pyramid (niagara.cpp:1339-1350)
	std::vector<float> depth(size_t(width) * height, 0.f);
	for (uint32_t y = height / 3; y < height / 2; ++y)
		for (uint32_t x = width / 4; x < width / 2; ++x)
			depth[size_t(y) * width + x] = 0.1f / 40.f; // an occluder at view depth 40
	float* depthTarget = upload(depth);
	NvcHiZ hiz;
	CHECK_NVC(ctx, nvc_hiz_layout(width, height, &hiz));
	CHECK_CUDA(cudaMalloc(&hiz.texels, size_t(hiz.total_texels) * 4));

	NvcCamera camera = { { 0, 0, 0 }, { 0, 0, 0, 1 }, 70.f * 3.14159265f / 180.f, 0.1f }; // niagara.cpp:833-837
	NvcCullOptions options = { 200.f, 1, 1, 1, 1, 1, 0 };                                   // niagara.cpp:31-44, 1000

	cudaEvent_t e0, e1;
	CHECK_CUDA(cudaEventCreate(&e0));
	CHECK_CUDA(cudaEventCreate(&e1));

	for (int frame = 0; frame < frames; ++frame)
	{
		camera.position[2] = -0.5f * float(frame); // fly forward (view z = -world z)
		NvcCullData cullData, passData;
		nvc_host_cull_data(&camera, width, height, drawCount, &options, &cullData, nullptr); // niagara.cpp:1487-1516

		CHECK_CUDA(cudaEventRecord(e0, stream));
		// early cull + early clusters (niagara.cpp:1766-1769)
		nvc_host_pass_data(&cullData, 1, 0, &passData);
		CHECK_NVC(ctx, nvc_drawcull(ctx, stream, &passData, 0, 1, db, mb, dvb, dcb, dccb, &hiz));
		nvc_host_pass_data(&cullData, 0, 0, &passData);
		CHECK_NVC(ctx, nvc_clustercull(ctx, stream, &passData, 0, (const NvcMeshTaskCommand*)dcb, dccb, db, mlb, mvb, cib, ccb, &hiz));
		// ... early render would rasterise here; pyramid from its depth (niagara.cpp:1772)
		CHECK_NVC(ctx, nvc_depth_pyramid(ctx, stream, depthTarget, width, height, &hiz));
		// late cull + late clusters (niagara.cpp:1775-1778)
		nvc_host_pass_data(&cullData, 1, 0, &passData);
		CHECK_NVC(ctx, nvc_drawcull(ctx, stream, &passData, 1, 1, db, mb, dvb, dcb, dccb, &hiz));
		nvc_host_pass_data(&cullData, 0, 0, &passData);
		CHECK_NVC(ctx, nvc_clustercull(ctx, stream, &passData, 1, (const NvcMeshTaskCommand*)dcb, dccb, db, mlb, mvb, cib, ccb, &hiz));
		CHECK_CUDA(cudaEventRecord(e1, stream));

		uint32_t hd[4], hc[4];
		CHECK_CUDA(cudaMemcpyAsync(hd, dccb, 16, cudaMemcpyDeviceToHost, stream));
		CHECK_CUDA(cudaMemcpyAsync(hc, ccb, 16, cudaMemcpyDeviceToHost, stream));
		CHECK_CUDA(cudaStreamSynchronize(stream));
		float ms = 0;
		CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
		printf("frame %2d: %.3f ms  late task commands %u (dispatch %u x %u x %u)  late clusters %u (dispatch %u x %u x %u)\n", frame, ms, hd[0], hd[1], hd[2], hd[3], hc[0], hc[1], hc[2], hc[3]);
	}

	nvc_destroy(ctx);
	return 0;
}
