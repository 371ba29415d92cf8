"""CPU-only: the glTF importer (nvc_gltf_*, row N3) against what the REFERENCE's loadScene produced from the same files
(tests/golden/animated.raw.cache and hierarchy_expected.npz, written by the reference's importer + cache writer through
oracle/_ref/write_cache; generators: make_scene_cache_fixtures.py, make_gltf_fixtures.py) — MeshDraw[], Animation[], Keyframe[]
and the camera bit for bit; the vertex data of every primitive as a set (the reference's cooker de-duplicates and reorders vertices)."""
import ctypes
import json
import os
import struct

import numpy as np
import pytest

from niagara_b200 import gltf, layout, scene_cache
from niagara_b200.lib import NvcError, load_library


def _same_draws(mine, ref):
    mine = mine.copy()
    mine["meshletVisibilityOffset"] = ref["meshletVisibilityOffset"]  # assigned later by the application (niagara.cpp:1002-1020)
    return mine.tobytes() == ref.tobytes()


def test_animated_gltf_matches_the_reference_importer(golden_dir):
    g = gltf.GltfScene(os.path.join(golden_dir, "animated.gltf"))
    c = scene_cache.SceneCache(os.path.join(golden_dir, "animated.raw.cache"))
    assert (g.info.draw_count, g.info.animation_count, g.info.keyframe_count) == (14, 3, 18)
    assert _same_draws(g.draws, c.section("draws"))
    assert g.animations.tobytes() == c.section("animations").tobytes()
    assert g.keyframes.tobytes() == c.section("keyframes").tobytes()
    assert set(g.draws["postPass"]) == {0, 1}
    cam = c.header.camera
    assert g.info.has_camera == 1 and list(g.info.camera.position) == list(cam.position) and list(g.info.camera.orientation) == list(cam.orientation) and g.info.camera.fovY == cam.fovY
    c.close()
    g.close()


def test_hierarchy_glb_matches_the_reference_importer(golden_dir):
    """three-level hierarchy with non-uniform parents, a mirroring matrix node, camera / lights below parents, keys baked through the
    parents, skipped primitives (points, non-indexed), skipped animations (STEP, node without draw or light), MASK / transmission"""
    g = gltf.GltfScene(os.path.join(golden_dir, "hierarchy.glb"))
    ref = np.load(os.path.join(golden_dir, "hierarchy_expected.npz"))
    assert g.info.primitive_count == 3 and g.info.draw_count == 6 and g.info.point_light_count == 1 and g.info.has_sun == 1
    assert _same_draws(g.draws, ref["draws"])
    assert g.animations.tobytes() == ref["animations"].tobytes()
    assert g.keyframes.tobytes() == ref["keyframes"].tobytes()
    assert sorted(set(g.draws["postPass"])) == [0, 1, 2]
    assert (g.animations["lightIndex"] >= 0).sum() == 1 and (g.animations["drawIndex"] >= 0).sum() == 2
    cam = np.array(list(g.info.camera.position) + list(g.info.camera.orientation) + [g.info.camera.fovY], np.float32)
    assert cam.tobytes() == ref["camera"].tobytes()
    # geometry: loadVertices output == the reference's cooked vertices as a set, per mesh
    verts = ref["vertices"].view(gltf.VERTEX_DTYPE) if ref["vertices"].dtype != gltf.VERTEX_DTYPE else ref["vertices"]
    for m in range(g.info.primitive_count):
        v, ix = g.primitive(m)
        assert ix.max() < len(v) and len(ix) % 3 == 0
        used = np.unique(v[np.unique(ix)].view(np.uint8).reshape(-1, 16), axis=0)
        lo, n = int(ref["mesh_vertex_offset"][m]), int(ref["mesh_vertex_count"][m])
        want = np.unique(np.ascontiguousarray(verts[lo : lo + n]).view(np.uint8).reshape(-1, 16), axis=0)
        assert np.array_equal(used, want), m
    g.close()


def test_malformed_files_are_rejected_not_read(golden_dir, tmp_path):
    lib = load_library()
    good = open(os.path.join(golden_dir, "hierarchy.glb"), "rb").read()

    def status(data, base=b"."):
        scene = ctypes.c_void_p()
        s = lib.nvc_gltf_import(data, len(data), base, 0, 1, ctypes.byref(scene))
        if s == 0:
            lib.nvc_gltf_free(scene)
        return s

    assert status(good) == 0
    assert status(good[:40]) == layout_status("corrupt")
    assert status(b"{\"asset\":") == layout_status("corrupt")
    # JSON chunk edits: accessor running past its view, view past its buffer, node cycle, sparse accessor
    jlen = struct.unpack_from("<I", good, 12)[0]
    doc = json.loads(good[20 : 20 + jlen])
    binchunk = good[20 + jlen :]

    def rebuild(d):
        js = json.dumps(d, separators=(",", ":")).encode()
        js += b" " * (-len(js) % 4)
        return struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(js) + len(binchunk)) + struct.pack("<II", len(js), 0x4E4F534A) + js + binchunk

    d = json.loads(json.dumps(doc)); d["accessors"][0]["count"] = 10**6
    scene = ctypes.c_void_p()
    data = rebuild(d)
    assert lib.nvc_gltf_import(data, len(data), b".", 0, 1, ctypes.byref(scene)) == 0  # positions are only read on request
    vc, ic = ctypes.c_uint32(), ctypes.c_uint32()
    assert lib.nvc_gltf_primitive_size(scene, 0, ctypes.byref(vc), ctypes.byref(ic)) == 0 and vc.value == 10**6
    v = np.zeros(vc.value, dtype=gltf.VERTEX_DTYPE)
    assert lib.nvc_gltf_primitive_data(scene, 0, v.ctypes.data_as(ctypes.c_void_p), None) == layout_status("corrupt")
    lib.nvc_gltf_free(scene)
    d = json.loads(json.dumps(doc)); d["bufferViews"][0]["byteLength"] = 10**9
    assert status(rebuild(d)) == layout_status("corrupt")
    d = json.loads(json.dumps(doc)); d["nodes"][5]["children"] = [0]
    assert status(rebuild(d)) == layout_status("corrupt")
    d = json.loads(json.dumps(doc)); d["accessors"][1]["sparse"] = {"count": 1}
    assert status(rebuild(d)) == layout_status("unsupported")
    d = json.loads(json.dumps(doc)); d["nodes"][2]["mesh"] = 7
    assert status(rebuild(d)) == layout_status("corrupt")
    # byte-level fuzz of the whole file: every outcome is a status code
    rng = np.random.default_rng(5)
    for _ in range(300):
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        assert status(bytes(b)) in (0, layout_status("corrupt"), layout_status("unsupported"), layout_status("invalid"))


def layout_status(name):
    return {"corrupt": -8, "unsupported": -7, "invalid": -1}[name] if False else _STATUS[name]


_STATUS = {}


def setup_module(module):
    lib = load_library()
    for code in range(-12, 0):
        text = lib.nvc_status_string(code).decode()
        if text == "corrupt scene cache":
            _STATUS["corrupt"] = code
        elif text == "unsupported input":
            _STATUS["unsupported"] = code
        elif text == "invalid argument":
            _STATUS["invalid"] = code


def test_importer_under_address_and_ub_sanitizers(golden_dir, tmp_path):
    """tests/fuzz_gltf.cpp: the importer compiled from the product source with ASan + UBSan over 4000 mutated files"""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fuzz_gltf")
    srcs = [os.path.join(root, "tests", "fuzz_gltf.cpp"), os.path.join(root, "niagara_b200", "csrc", "nvc_gltf.cpp")]
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-o", exe] + srcs, capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("this toolchain has no sanitizer runtime: " + build.stderr.splitlines()[0])
    assert build.returncode == 0, build.stderr
    for name in ("hierarchy.glb", "animated.gltf"):
        run = subprocess.run([exe, os.path.join(golden_dir, name), "2000"], capture_output=True, text=True)
        assert run.returncode == 0, (run.stdout[-500:], run.stderr[-3000:])
        assert "mutations 2000" in run.stdout
