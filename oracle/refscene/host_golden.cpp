// TEST INFRASTRUCTURE: produces golden vectors for the host-side helpers (niagara_b200/csrc/nvc_host.cpp) with the
// reference's own math library (glm, compiled from /root/reference/extern/glm with the reference's defines).
// niagara.cpp's main() cannot be compiled here (Vulkan/GLFW), so what its few lines involved compute is written out
// again, in our own terms, around the same glm calls: PCG32 + random scene (niagara.cpp:449-481, 969-998) and the CullData fill (niagara.cpp:424-437, 1487-1516).
//
// Output "NVCH" v1: u32 magic, u32 version, u32 drawCountA, u32 meshCountA, u32 drawCountB, u32 meshCountB, u32 cameraCount, u32 pad
//   MeshDraw[drawCountA], MeshDraw[drawCountB], then per camera: {float pos[3], quat xyzw[4], fovY, znear, u32 w, u32 h, u32 drawCount, u32 lodStep} + CullData(144 B)
#include "host_golden.h"

#include <stdio.h>

int main(int argc, char** argv)
{
	if (argc < 2)
		return 2;
	std::vector<MeshDraw> a = randomScene(4096, 1), b = randomScene(2048, 7);
	CameraCase cams[] = {
		{ { 0, 0, 0 }, { 0, 0, 0, 1 }, glm::radians(70.f), 0.1f, 1024, 768, 4096, 0 },
		{ { 10.5f, -3.25f, 42.f }, { 0.1825742f, 0.3651484f, 0.5477226f, 0.7302967f }, glm::radians(70.f), 0.1f, 1920, 1080, 1000000, 0 },
		{ { -120.f, 60.f, -250.f }, { -0.3f, 0.1f, 0.2f, 0.9273618f }, glm::radians(50.f), 0.5f, 4096, 4096, 1000000, 2 },
		{ { 1.f, 2.f, 3.f }, { 0.f, 0.7071068f, 0.f, 0.7071068f }, glm::radians(90.f), 1.f, 2560, 1440, 12345, 1 },
	};
	uint32_t ncam = sizeof(cams) / sizeof(cams[0]);
	FILE* f = fopen(argv[1], "wb");
	if (!f)
		return 1;
	uint32_t header[8] = { 0x4843564eu, 1u, uint32_t(a.size()), 1u, uint32_t(b.size()), 7u, ncam, 0 };
	fwrite(header, sizeof(header), 1, f);
	fwrite(a.data(), sizeof(MeshDraw), a.size(), f);
	fwrite(b.data(), sizeof(MeshDraw), b.size(), f);
	for (uint32_t i = 0; i < ncam; ++i)
	{
		CullData cd = fillCullData(cams[i]);
		fwrite(&cams[i], sizeof(CameraCase), 1, f);
		fwrite(&cd, sizeof(cd), 1, f);
	}
	fclose(f);
	printf("%s: %zu + %zu draws, %u cameras\n", argv[1], a.size(), b.size(), ncam);
	return 0;
}
