// examples/scene_cache_info.cpp — host-only C++ use of the scene-cache and animation entry points (rows N2 / N3 of
// DESIGN.md §7), the way niagara's main() uses loadSceneCache (src/niagara.cpp:857-876) and its animation block
// (:1362-1390).  Needs no GPU: maps a .cache file, prints the header and the section table, decodes the per-meshlet
// stream, and evaluates the keyframe tracks at a few times.
//
// build: make -C examples scene_cache_info        run: examples/scene_cache_info scene.cache [time ...]
#include "../include/niagara_cull.h"

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <vector>

static const char* kSectionNames[NVC_CACHE_SECTION_COUNT] = { "vertices", "indices", "meshlets", "meshletdata", "meshletvtx0", "meshes", "materials", "draws", "lights",
	"animations", "keyframes", "omm_data", "omm_indices", "omm_descs", "texture_paths" };

int main(int argc, char** argv)
{
	if (argc < 2)
	{
		fprintf(stderr, "usage: %s scene.cache [animation time ...]\n", argv[0]);
		return 2;
	}
	int fd = open(argv[1], O_RDONLY);
	struct stat st;
	if (fd < 0 || fstat(fd, &st) != 0)
	{
		perror(argv[1]);
		return 1;
	}
	size_t size = size_t(st.st_size);
	void* file = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
	if (file == MAP_FAILED)
	{
		perror("mmap");
		return 1;
	}

	NvcSceneCacheInfo info;
	int status = nvc_scene_cache_parse(file, size, &info);
	if (status != NVC_OK)
	{
		fprintf(stderr, "%s: %s\n", argv[1], nvc_status_string(status));
		return 1;
	}
	const NvcSceneCacheHeader& h = info.header;
	printf("version %u  hashMeta %016llx  meshlet limits %u/%u  clrt %u  compressed %u\n", h.version, (unsigned long long)h.hashMeta, h.meshletMaxVertices, h.meshletMaxTriangles,
	    h.clrtMode, h.compressed);
	printf("camera position (%g %g %g) fovY %g znear %g\n", h.camera.position[0], h.camera.position[1], h.camera.position[2], h.camera.fovY, h.camera.znear);
	for (int s = 0; s < NVC_CACHE_SECTION_COUNT; ++s)
	{
		const NvcSceneCacheSection& sec = info.sections[s];
		printf("%-13s count %9u  offset %10llu  stored %10llu  decoded %10llu%s\n", kSectionNames[s], sec.count, (unsigned long long)sec.offset, (unsigned long long)sec.stored_bytes,
		    (unsigned long long)sec.decoded_bytes, sec.compressed ? "  (meshopt stream)" : "");
	}

	// the arrays the visibility path uploads are raw in the file: this is the whole "load"
	const NvcMeshlet* meshlets = reinterpret_cast<const NvcMeshlet*>(static_cast<const char*>(file) + info.sections[NVC_CACHE_MESHLETS].offset);
	const NvcMesh* meshes = reinterpret_cast<const NvcMesh*>(static_cast<const char*>(file) + info.sections[NVC_CACHE_MESHES].offset);
	unsigned long long triangles = 0;
	for (uint32_t i = 0; i < h.meshletCount; ++i)
		triangles += meshlets[i].triangleCount;
	printf("meshes %u (first: radius %g, %u lods)  meshlets %u  triangles in meshlets %llu\n", h.meshCount, h.meshCount ? meshes[0].radius : 0.f, h.meshCount ? meshes[0].lodCount : 0u,
	    h.meshletCount, triangles);

	std::vector<uint32_t> meshletdata(h.meshletdataCount);
	status = nvc_scene_cache_read(file, size, &info, NVC_CACHE_MESHLETDATA, meshletdata.data(), meshletdata.size() * sizeof(uint32_t));
	unsigned long long digest = 1469598103934665603ull;
	for (uint32_t w : meshletdata)
		digest = (digest ^ w) * 1099511628211ull;
	printf("meshletdata: %s, %zu words, fnv %016llx\n", nvc_status_string(status), meshletdata.size(), digest);

	if (argc > 2 && h.animationCount)
	{
		std::vector<NvcMeshDraw> draws(h.drawCount);
		nvc_scene_cache_read(file, size, &info, NVC_CACHE_DRAWS, draws.data(), draws.size() * sizeof(NvcMeshDraw));
		const NvcAnimation* animations = reinterpret_cast<const NvcAnimation*>(static_cast<const char*>(file) + info.sections[NVC_CACHE_ANIMATIONS].offset);
		const NvcKeyframe* keyframes = reinterpret_cast<const NvcKeyframe*>(static_cast<const char*>(file) + info.sections[NVC_CACHE_KEYFRAMES].offset);
		std::vector<uint32_t> indices(h.animationCount);
		std::vector<NvcMeshDraw> values(h.animationCount);
		for (int a = 2; a < argc; ++a)
		{
			double t = atof(argv[a]);
			int n = nvc_host_animate(animations, h.animationCount, keyframes, h.keyframeCount, t, draws.data(), h.drawCount, indices.data(), values.data(), h.animationCount);
			printf("t = %g: %d draws move", t, n);
			for (int i = 0; i < n && i < 3; ++i)
				printf("  [%u] -> (%.9g %.9g %.9g) scale %.9g", indices[i], values[i].position[0], values[i].position[1], values[i].position[2], values[i].scale);
			printf("\n");
		}
	}
	munmap(file, size);
	close(fd);
	return 0;
}
