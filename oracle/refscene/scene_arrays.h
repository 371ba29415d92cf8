// TEST INFRASTRUCTURE: everything the reference's importer fills, in one place, for the two tools of this directory.
#pragma once
#include "common.h"
#include "scene.h"

#include <string.h>

// scene.cpp references this from textures.cpp only inside buildSceneOmm (ray-tracing feature, out of scope); each tool is one TU
unsigned char* decodeImageRGBA(const char*, int, unsigned int&, unsigned int&, unsigned int&) { return nullptr; }

struct SceneArrays
{
	Geometry geo;
	std::vector<Material> mats;
	std::vector<MeshDraw> draws;
	std::vector<Light> lights;
	std::vector<std::string> textures;
	std::vector<Animation> tracks;
	std::vector<Keyframe> keys;
	Camera cam = {};
	vec3 sun = vec3(0.f);

	SceneArrays() { mats.push_back(Material()); } // slot 0 is the reference's dummy material (niagara.cpp seeds it before loadScene)

	// .gltf / .glb go through loadScene, anything else through loadMesh; returns false when the reference's loader does
	bool load(const char* path, bool* was_scene = nullptr)
	{
		const char* ext = strrchr(path, '.');
		bool scene = ext && (strcmp(ext, ".gltf") == 0 || strcmp(ext, ".glb") == 0);
		if (was_scene)
			*was_scene = scene;
		return scene ? loadScene(geo, mats, draws, lights, textures, tracks, keys, cam, sun, path) : loadMesh(geo, path);
	}
};
