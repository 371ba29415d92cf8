// TEST INFRASTRUCTURE (oracle side): drives the *reference's own* importer and scene-cache writer
// (/root/reference/src/scene.cpp loadScene / loadMesh, src/scenecache.cpp saveSceneCache, vendored meshoptimizer codecs)
// to produce the fixtures our scene-cache reader (csrc/nvc_scene_cache.cpp) and animation evaluation
// (nvc_host_animate) are checked against.  The reference sources are compiled where they lie; nothing is copied.
//
//   write_cache <prefix> input.{obj,gltf,glb}...
//     <prefix>.raw.cache   saveSceneCache(..., compressed = false)
//     <prefix>.z.cache     saveSceneCache(..., compressed = true)
//     <prefix>.z.meshletdata   geometry.meshletdata (uint32[]) as the reference's own loadSceneCache decodes <prefix>.z.cache
//                          (the meshlet codec rotates triangles, so this — not the raw cache — is what a reader of the
//                          compressed cache must reproduce; bytes past triangleCount * 3 in a meshlet's last word are
//                          whatever the reference's SIMD decoder left there)
//     <prefix>.z.indices       geometry.indices (uint32[]) as the reference's loadSceneCache decodes <prefix>.z.cache
//     <prefix>.nvca        animation golden: the frame loop's update (niagara.cpp:1362-1390, glm::mix / glm::slerp) applied
//                          at a list of animationTime values; format below
//
// NVCA v1 (little endian): u32 magic 'NVCA', u32 version, u32 animationCount, u32 keyframeCount, u32 drawCount, u32 timeCount,
//   u32 reserved[2]; Animation[animationCount] (24 B); Keyframe[keyframeCount] (32 B); MeshDraw[drawCount] (initial);
//   then timeCount records { f64 animationTime; MeshDraw[drawCount] (draws after the update at that time) }
#include "scene_arrays.h"

#include <math.h>
#include <stdio.h>

static_assert(sizeof(Animation) == 24 && sizeof(Keyframe) == 32 && sizeof(MeshDraw) == 48, "layouts");

int main(int argc, char** argv)
{
	if (argc < 3)
	{
		fprintf(stderr, "usage: %s prefix input.{obj,gltf,glb}...\n", argv[0]);
		return 2;
	}

	SceneArrays in;
	in.cam.orientation = quat(1, 0, 0, 0);
	in.cam.fovY = glm::radians(70.f);
	in.cam.znear = 0.1f;
	in.sun = normalize(vec3(1.0f, 1.0f, 1.0f));
	for (int i = 2; i < argc; ++i)
	{
		size_t firstMesh = in.geo.meshes.size();
		bool scene = false;
		if (!in.load(argv[i], &scene))
		{
			fprintf(stderr, "failed to load %s\n", argv[i]);
			return 1;
		}
		if (!scene) // a bare mesh: one identity draw per mesh
			for (size_t m = firstMesh; m < in.geo.meshes.size(); ++m)
			{
				MeshDraw draw = {};
				draw.scale = 1.f;
				draw.orientation = quat(1, 0, 0, 0);
				draw.meshIndex = uint32_t(m);
				in.draws.push_back(draw);
			}
	}
	Geometry& geometry = in.geo;
	std::vector<Material>& materials = in.mats;
	std::vector<MeshDraw>& draws = in.draws;
	std::vector<Light>& lights = in.lights;
	std::vector<std::string>& texturePaths = in.textures;
	std::vector<Animation>& animations = in.tracks;
	std::vector<Keyframe>& keyframes = in.keys;
	Camera& camera = in.cam;
	vec3& sun = in.sun;

	// meshletVisibilityOffset as niagara.cpp:1003-1020 assigns it
	uint32_t meshletVisibilityCount = 0;
	for (MeshDraw& draw : draws)
	{
		const Mesh& mesh = geometry.meshes[draw.meshIndex];
		draw.meshletVisibilityOffset = meshletVisibilityCount;
		uint32_t meshletCount = 0;
		for (uint32_t l = 0; l < mesh.lodCount; ++l)
			meshletCount = std::max(meshletCount, mesh.lods[l].meshletCount);
		meshletVisibilityCount += meshletCount;
	}

	std::string prefix = argv[1];
	const uint64_t hashMeta = 0x0123456789abcdefull;
	if (!saveSceneCache((prefix + ".raw.cache").c_str(), geometry, materials, draws, lights, texturePaths, animations, keyframes, camera, sun, hashMeta, false, false, false) ||
	    !saveSceneCache((prefix + ".z.cache").c_str(), geometry, materials, draws, lights, texturePaths, animations, keyframes, camera, sun, hashMeta, false, true, true))
	{
		fprintf(stderr, "saveSceneCache failed\n");
		return 1;
	}

	// the compressed cache read back by the reference's own loader
	{
		Geometry g2;
		std::vector<Material> m2;
		std::vector<MeshDraw> d2;
		std::vector<Light> l2;
		std::vector<std::string> t2;
		std::vector<Animation> a2;
		std::vector<Keyframe> k2;
		Camera c2 = {};
		vec3 s2(0.f);
		if (!loadSceneCache((prefix + ".z.cache").c_str(), g2, m2, d2, l2, t2, a2, k2, c2, s2, hashMeta, false, 0))
		{
			fprintf(stderr, "loadSceneCache failed\n");
			return 1;
		}
		if (g2.meshlets.size() != geometry.meshlets.size() || memcmp(g2.meshlets.data(), geometry.meshlets.data(), g2.meshlets.size() * sizeof(Meshlet)) != 0 ||
		    d2.size() != draws.size() || memcmp(d2.data(), draws.data(), d2.size() * sizeof(MeshDraw)) != 0 || g2.vertices.size() != geometry.vertices.size() ||
		    memcmp(g2.vertices.data(), geometry.vertices.data(), g2.vertices.size() * sizeof(Vertex)) != 0)
		{
			fprintf(stderr, "round trip mismatch\n");
			return 1;
		}
		FILE* mf = fopen((prefix + ".z.meshletdata").c_str(), "wb");
		if (!mf)
			return 1;
		fwrite(g2.meshletdata.data(), sizeof(uint32_t), g2.meshletdata.size(), mf);
		fclose(mf);
		// the index codec rotates triangles as well: what the reference's loader returns is the golden, not the raw cache
		FILE* xf = fopen((prefix + ".z.indices").c_str(), "wb");
		if (!xf)
			return 1;
		fwrite(g2.indices.data(), sizeof(uint32_t), g2.indices.size(), xf);
		fclose(xf);
	}

	// animation golden
	std::vector<double> times;
	for (int k = 0; k < 40; ++k)
		times.push_back(0.13 + 0.173 * k); // starts before the first keyframe (index < 0), wraps around several periods
	times.push_back(0.5);
	times.push_back(1.0);   // exactly on keyframes
	times.push_back(3.4999);
	times.push_back(1234.5678);

	FILE* f = fopen((prefix + ".nvca").c_str(), "wb");
	if (!f)
		return 1;
	uint32_t header[8] = { 0x4143564eu, 1u, uint32_t(animations.size()), uint32_t(keyframes.size()), uint32_t(draws.size()), uint32_t(times.size()), 0, 0 };
	fwrite(header, sizeof(header), 1, f);
	fwrite(animations.data(), sizeof(Animation), animations.size(), f);
	fwrite(keyframes.data(), sizeof(Keyframe), keyframes.size(), f);
	fwrite(draws.data(), sizeof(MeshDraw), draws.size(), f);

	for (double animationTime : times)
	{
		// What the frame loop does for draw animations (niagara.cpp:1366-1390), written out with the SAME glm calls
		// (glm::mix for translation / scale, glm::slerp for rotation) so that glm's arithmetic — not ours — defines the golden:
		// track position in keyframe units (double), wrapped with fmod; integer part picks the key pair, fraction blends.
		for (const Animation& track : animations)
		{
			if (track.drawIndex < 0)
				continue;
			double cursor = (animationTime - track.startTime) / track.period;
			if (cursor < 0)
				continue; // not started yet
			cursor = fmod(cursor, double(track.keyframeCount));
			int from = int(cursor) % track.keyframeCount;
			int to = (from + 1) % track.keyframeCount;
			float blend = float(cursor - floor(cursor));
			const Keyframe& k0 = keyframes[track.keyframeOffset + from];
			const Keyframe& k1 = keyframes[track.keyframeOffset + to];
			MeshDraw& target = draws[track.drawIndex];
			target.position = glm::mix(k0.translation, k1.translation, blend);
			target.scale = glm::mix(k0.scale, k1.scale, blend);
			target.orientation = glm::slerp(k0.rotation, k1.rotation, blend);
		}
		fwrite(&animationTime, sizeof(double), 1, f);
		fwrite(draws.data(), sizeof(MeshDraw), draws.size(), f);
	}
	fclose(f);

	printf("%s: meshes %zu meshlets %zu meshletdata %zu vertices %zu indices %zu draws %zu animations %zu keyframes %zu\n", prefix.c_str(),
	    geometry.meshes.size(), geometry.meshlets.size(), geometry.meshletdata.size(), geometry.vertices.size(), geometry.indices.size(), draws.size(), animations.size(), keyframes.size());
	return 0;
}
