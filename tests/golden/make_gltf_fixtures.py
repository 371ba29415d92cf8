#!/usr/bin/env python3
"""Authors tests/golden/hierarchy.glb — OUR OWN binary glTF (no third-party asset) that exercises what the importer must get
right beyond animated.gltf: a node HIERARCHY three levels deep (rotated / non-uniformly scaled parents), a node given as a
MATRIX (with a mirroring, i.e. negative determinant), the camera and a directional light below a parent, an animated point
light and an animated child (keys baked through its parents), u16 indices, a primitive with TANGENTs, a points primitive and a
non-indexed primitive (both skipped by the reference), materials: opaque / MASK / transmission.
Then runs the REFERENCE's importer + cache writer on it (oracle/_ref/write_cache, built from /root/reference) and stores what
the frame loop would get in hierarchy_expected.npz: MeshDraw[], Animation[], Keyframe[], camera, per-mesh vertex ranges and the
reference's cooked Vertex[] (for a set comparison with the importer's loadVertices output).
Only runs where /root/reference exists.  Run: python tests/golden/make_gltf_fixtures.py"""
import json
import math
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def quat(axis, angle):
    axis = np.array(axis, dtype=np.float64)
    axis /= np.linalg.norm(axis)
    s = math.sin(angle / 2)
    return [float(axis[0] * s), float(axis[1] * s), float(axis[2] * s), float(math.cos(angle / 2))]


def main():
    from make_animated_gltf import icosphere

    pos, tri = icosphere(2)  # 162 vertices, 320 triangles
    nrm = pos.copy()
    # tangents: any unit vector orthogonal to the normal, w = +-1
    t = np.cross(nrm, np.array([0.3, 1.0, 0.2], np.float32))
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    tan = np.concatenate([t, np.where(np.arange(len(t)) % 3 == 0, -1.0, 1.0)[:, None]], 1).astype(np.float32)
    uv = np.stack([pos[:, 0] * 0.5 + 0.5, pos[:, 1] * 0.5 + 0.5], 1).astype(np.float32)
    blob, views, accessors = bytearray(), [], []

    def add(data, comp, typ, count, normalized=False):
        while len(blob) % 4:
            blob.append(0)
        views.append({"buffer": 0, "byteOffset": len(blob), "byteLength": len(data)})
        blob.extend(data)
        acc = {"bufferView": len(views) - 1, "componentType": comp, "count": count, "type": typ}
        if normalized:
            acc["normalized"] = True
        accessors.append(acc)
        return len(accessors) - 1

    a_pos = add(pos.tobytes(), 5126, "VEC3", len(pos))
    accessors[a_pos]["min"], accessors[a_pos]["max"] = pos.min(0).tolist(), pos.max(0).tolist()
    a_nrm = add(nrm.tobytes(), 5126, "VEC3", len(nrm))
    a_tan = add(tan.tobytes(), 5126, "VEC4", len(tan))
    a_uv16 = add((np.clip(uv, 0, 1) * 65535 + 0.5).astype(np.uint16).tobytes(), 5123, "VEC2", len(uv), normalized=True)
    third = len(tri) // 3
    a_i0 = add(tri[:third].astype(np.uint16).tobytes(), 5123, "SCALAR", third * 3)
    a_i1 = add(tri[third : 2 * third].astype(np.uint32).tobytes(), 5125, "SCALAR", third * 3)
    a_i2 = add(tri[2 * third :].astype(np.uint16).tobytes(), 5123, "SCALAR", (len(tri) - 2 * third) * 3)

    rng = np.random.default_rng(23)
    mirror = np.array([[0.0, 0.0, 1.5, 0.0], [0.0, 1.5, 0.0, 0.0], [1.5, 0.0, 0.0, 0.0], [2.0, -1.0, 4.0, 1.0]], np.float32)  # column-major rows = columns; det < 0
    nodes = [
        {"name": "root", "children": [1, 2, 7], "translation": [1.0, 2.0, -3.0], "rotation": quat((0.2, 1.0, 0.1), 0.7), "scale": [1.25, 1.25, 1.25]},
        {"name": "arm", "children": [3, 4], "translation": [3.0, 0.5, 0.0], "rotation": quat((1.0, 0.0, 0.3), 1.9), "scale": [0.5, 0.75, 1.5]},
        {"name": "ballA", "mesh": 0, "translation": [-2.0, 0.0, 1.0], "rotation": quat((0.0, 0.0, 1.0), 2.5)},
        {"name": "ballB", "mesh": 1, "translation": [0.0, 4.0, 0.0], "scale": [2.0, 2.0, 2.0]},
        {"name": "hand", "children": [5], "rotation": quat((0.5, 0.5, 0.7), 0.4)},
        {"name": "ballC", "mesh": 0, "translation": [1.0, 1.0, 1.0], "rotation": quat((0.1, 0.9, 0.2), 3.0), "scale": [0.6, 0.6, 0.6]},
        {"name": "mirrored", "mesh": 1, "matrix": [float(v) for v in mirror.reshape(-1)]},
        {"name": "rig", "children": [8, 9, 10], "translation": [0.0, 6.0, 10.0], "rotation": quat((1.0, 0.0, 0.0), -0.3)},
        {"name": "camera", "camera": 0, "translation": [0.0, 0.5, 2.0], "rotation": quat((0.0, 1.0, 0.0), 0.15)},
        {"name": "sun", "rotation": quat((1.0, 0.2, 0.0), -1.1), "extensions": {"KHR_lights_punctual": {"light": 0}}},
        {"name": "lamp", "translation": [1.0, 1.0, 1.0], "extensions": {"KHR_lights_punctual": {"light": 1}}},
        {"name": "empty-animated", "translation": [9.0, 9.0, 9.0]},
    ]
    times = np.array([0.0, 0.25, 0.5, 0.75, 1.0], dtype=np.float32)
    a_time = add(times.tobytes(), 5126, "SCALAR", len(times))
    accessors[a_time]["min"], accessors[a_time]["max"] = [0.0], [1.0]
    samplers, channels = [], []

    def track(node, paths):
        for path in paths:
            if path == "translation":
                data, typ = rng.uniform(-4, 4, (len(times), 3)).astype(np.float32), "VEC3"
            elif path == "rotation":
                data, typ = np.array([quat(rng.uniform(-1, 1, 3), float(rng.uniform(0, 6))) for _ in times], dtype=np.float32), "VEC4"
            else:
                data, typ = rng.uniform(0.5, 2.0, (len(times), 3)).astype(np.float32), "VEC3"
            out = add(np.ascontiguousarray(data).tobytes(), 5126, typ, len(times))
            samplers.append({"input": a_time, "output": out})  # interpolation defaults to LINEAR
            channels.append({"sampler": len(samplers) - 1, "target": {"node": node, "path": path}})

    track(5, ("translation", "rotation", "scale"))  # ballC: three levels deep
    track(3, ("rotation",))  # ballB
    track(10, ("translation",))  # the point light
    track(11, ("translation",))  # a node without draw or light: skipped by the reference
    step_out = add(rng.uniform(-1, 1, (len(times), 3)).astype(np.float32).tobytes(), 5126, "VEC3", len(times))
    samplers.append({"input": a_time, "output": step_out, "interpolation": "STEP"})
    channels.append({"sampler": len(samplers) - 1, "target": {"node": 2, "path": "translation"}})  # STEP: skipped

    gltf = {
        "asset": {"version": "2.0", "generator": "niagara_b200 tests/golden/make_gltf_fixtures.py"},
        "extensionsUsed": ["KHR_lights_punctual", "KHR_materials_transmission"],
        "extensions": {"KHR_lights_punctual": {"lights": [{"type": "directional", "intensity": 3.0}, {"type": "point", "intensity": 50.0, "range": 20.0, "color": [1.0, 0.5, 0.25]}]}},
        "scene": 0,
        "scenes": [{"nodes": [0]}],
        "nodes": nodes,
        "cameras": [{"type": "perspective", "perspective": {"yfov": 1.1, "znear": 0.1}}],
        "materials": [
            {"name": "opaque"},
            {"name": "cutout", "alphaMode": "MASK", "alphaCutoff": 0.4},
            {"name": "glass", "extensions": {"KHR_materials_transmission": {"transmissionFactor": 0.9}}},
        ],
        "meshes": [
            {"primitives": [
                {"attributes": {"POSITION": a_pos, "NORMAL": a_nrm, "TANGENT": a_tan, "TEXCOORD_0": a_uv16}, "indices": a_i0, "material": 1},
                {"attributes": {"POSITION": a_pos}, "mode": 0},  # points: not a mesh for the reference
                {"attributes": {"POSITION": a_pos, "NORMAL": a_nrm}, "indices": a_i1, "material": 2},
            ]},
            {"primitives": [
                {"attributes": {"POSITION": a_pos, "NORMAL": a_nrm}},  # non-indexed: skipped
                {"attributes": {"POSITION": a_pos, "NORMAL": a_nrm, "TEXCOORD_0": a_uv16}, "indices": a_i2},
            ]},
        ],
        "animations": [{"name": "motion", "samplers": samplers, "channels": channels}],
        "accessors": accessors,
        "bufferViews": views,
        "buffers": [{"byteLength": len(blob)}],
    }
    js = json.dumps(gltf, separators=(",", ":")).encode()
    js += b" " * (-len(js) % 4)
    while len(blob) % 4:
        blob.append(0)
    glb = struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(js) + 8 + len(blob)) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blob), 0x004E4942) + bytes(blob)
    out = os.path.join(HERE, "hierarchy.glb")
    open(out, "wb").write(glb)
    print(out, len(glb), "bytes")

    # ---- the reference's importer on it ----
    from niagara_b200 import scene_cache

    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    tmp = tempfile.mkdtemp()
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "write_cache"), os.path.join(tmp, "hierarchy"), out], check=True)
    c = scene_cache.SceneCache(os.path.join(tmp, "hierarchy.raw.cache"))
    cam = c.header.camera
    np.savez_compressed(
        os.path.join(HERE, "hierarchy_expected.npz"),
        draws=c.section("draws"), animations=c.section("animations"), keyframes=c.section("keyframes"),
        camera=np.array(list(cam.position) + list(cam.orientation) + [cam.fovY], np.float32),
        mesh_vertex_offset=c.section("meshes")["vertexOffset"], mesh_vertex_count=c.section("meshes")["vertexCount"],
        vertices=c.section("vertices"),
    )
    print({k: len(c.section(k)) for k in ("draws", "animations", "keyframes", "meshes")})


if __name__ == "__main__":
    main()
