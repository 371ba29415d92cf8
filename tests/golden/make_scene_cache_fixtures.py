#!/usr/bin/env python3
"""Regenerates the scene-cache / animation fixtures with the REFERENCE's own importer, cache writer and cache loader
(oracle/_ref/write_cache, built by `make -C oracle ref` from /root/reference; see oracle/refscene/write_cache.cpp):

  animated.gltf            our own asset (make_animated_gltf.py)
  animated.raw.cache       saveSceneCache(compressed = false)
  animated.z.cache         saveSceneCache(compressed = true)
  animated.z.meshletdata   meshletdata[] as the reference's loadSceneCache decodes animated.z.cache
  animated.nvca            animation golden (niagara.cpp:1362-1390 with glm::mix / glm::slerp at 44 times)
  kitten.z.cache           data/kitten.obj, compressed (792 meshlets: every code path of the meshlet codec)
  kitten_cook.npz          fp16 positions + raw meshletdata of the cooked kitten (inputs of nvc_cook_meshlet_bounds)
  scene_cache_expected.json  sha256 of the reference-decoded kitten meshletdata (padding bytes masked) + counts

Only runs where /root/reference exists.  Run: python tests/golden/make_scene_cache_fixtures.py"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from niagara_b200 import scene_cache
    from test_scene_cache import meaningful_bytes

    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    subprocess.run([sys.executable, os.path.join(HERE, "make_animated_gltf.py")], check=True)
    tool = os.path.join(ROOT, "oracle", "_ref", "write_cache")
    tmp = tempfile.mkdtemp()
    subprocess.run([tool, os.path.join(tmp, "animated"), os.path.join(HERE, "animated.gltf")], check=True)
    subprocess.run([tool, os.path.join(tmp, "kitten"), "/root/reference/data/kitten.obj"], check=True)
    for name in ("animated.raw.cache", "animated.z.cache", "animated.z.meshletdata", "animated.nvca", "kitten.z.cache"):
        shutil.copy(os.path.join(tmp, name), os.path.join(HERE, name))
    c = scene_cache.SceneCache(os.path.join(HERE, "kitten.z.cache"))
    ref = np.fromfile(os.path.join(tmp, "kitten.z.meshletdata"), dtype="<u4")
    mask = meaningful_bytes(c.section("meshlets"), len(ref))
    def digest(path):
        return hashlib.sha256(np.fromfile(path, dtype="<u4").tobytes()).hexdigest()

    expected = {
        "animated": {"indices_sha256": digest(os.path.join(tmp, "animated.z.indices"))},
        "kitten": {
            "indices_sha256": digest(os.path.join(tmp, "kitten.z.indices")),
            "vertices_sha256": hashlib.sha256(scene_cache.SceneCache(os.path.join(tmp, "kitten.raw.cache")).section("vertices").tobytes()).hexdigest(),
            "meshletdata_words": int(len(ref)),
            "meshletdata_sha256_masked": hashlib.sha256((ref.view(np.uint8) * mask).tobytes()).hexdigest(),
            "meshlets": int(c.header.meshletCount),
            "vertices": int(c.header.vertexCount),
            "indices": int(c.header.indexCount),
        }
    }
    # inputs of the meshlet-bounds kernel for the cooked kitten (N4): fp16 positions + meshletdata in COOKED triangle order
    r = scene_cache.SceneCache(os.path.join(tmp, "kitten.raw.cache"))
    verts = r.section("vertices").view(np.uint16).reshape(-1, 8)
    np.savez_compressed(os.path.join(HERE, "kitten_cook.npz"), positions=np.ascontiguousarray(verts[:, :3]), meshletdata=r.section("meshletdata"))
    with open(os.path.join(HERE, "scene_cache_expected.json"), "w") as f:
        json.dump(expected, f, indent=1)
    shutil.rmtree(tmp)
    print(expected)


if __name__ == "__main__":
    main()
