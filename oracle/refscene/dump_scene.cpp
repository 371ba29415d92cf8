// TEST INFRASTRUCTURE (oracle side): drives the *reference's own* geometry pipeline
// (/root/reference/src/scene.cpp: loadMesh / loadScene -> appendMesh -> meshoptimizer) and dumps
// the arrays the visibility hot path consumes (Mesh[], Meshlet[], MeshDraw[]) as raw binary.
// The reference sources are compiled where they lie (see Makefile); nothing is copied.
//
// File format "NVCG" v1 (little endian):
//   u32 magic 'NVCG', u32 version, u32 meshCount, u32 meshletCount, u32 drawCount, u32 reserved[3]
//   Mesh[meshCount] (208 B), Meshlet[meshletCount] (24 B), MeshDraw[drawCount] (48 B)
#include "scene_arrays.h"

#include <stdio.h>

static_assert(sizeof(Mesh) == 208, "Mesh layout");
static_assert(sizeof(Meshlet) == 24, "Meshlet layout");
static_assert(sizeof(MeshDraw) == 48, "MeshDraw layout");
static_assert(sizeof(MeshLod) == 20, "MeshLod layout");

int main(int argc, char** argv)
{
	if (argc < 3)
	{
		fprintf(stderr, "usage: %s out.nvcg input.{obj,gltf,glb}...\n", argv[0]);
		return 2;
	}

	SceneArrays in;
	for (int i = 2; i < argc; ++i)
		if (!in.load(argv[i]))
		{
			fprintf(stderr, "failed to load %s\n", argv[i]);
			return 1;
		}
	Geometry& geometry = in.geo;
	std::vector<MeshDraw>& draws = in.draws;

	FILE* f = fopen(argv[1], "wb");
	if (!f)
		return 1;
	uint32_t header[8] = { 0x4743564eu, 1u, uint32_t(geometry.meshes.size()), uint32_t(geometry.meshlets.size()), uint32_t(draws.size()), 0, 0, 0 };
	fwrite(header, sizeof(header), 1, f);
	fwrite(geometry.meshes.data(), sizeof(Mesh), geometry.meshes.size(), f);
	fwrite(geometry.meshlets.data(), sizeof(Meshlet), geometry.meshlets.size(), f);
	fwrite(draws.data(), sizeof(MeshDraw), draws.size(), f);
	fclose(f);

	printf("%s: meshes %zu meshlets %zu draws %zu\n", argv[1], geometry.meshes.size(), geometry.meshlets.size(), draws.size());
	for (size_t m = 0; m < geometry.meshes.size() && m < 4; ++m)
	{
		const Mesh& mesh = geometry.meshes[m];
		printf(" mesh %zu: radius %g lods %u:", m, mesh.radius, mesh.lodCount);
		for (uint32_t l = 0; l < mesh.lodCount; ++l)
			printf(" %u(%g)", mesh.lods[l].meshletCount, mesh.lods[l].error);
		printf("\n");
	}
	return 0;
}
