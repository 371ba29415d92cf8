#!/usr/bin/env python3
"""Authors tests/golden/animated.gltf — OUR OWN small glTF scene (no third-party asset): one icosphere mesh (two
primitives: opaque + alpha-blended material, so the importer produces postPass 0 and 1 draws), seven nodes with
translation / rotation / uniform scale, three of them animated with LINEAR samplers over 6 keyframes, one camera.
It exists to drive the reference's glTF importer + scene-cache writer (oracle/refscene/write_cache.cpp) and through
them our cache reader and animation evaluation.  Run: python tests/golden/make_animated_gltf.py"""
import base64
import json
import math
import os
import struct

import numpy as np


def icosphere(subdiv):
    t = (1.0 + 5.0**0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v, dtype=np.float32), np.array(f, dtype=np.uint32)


def quat(axis, angle):
    axis = np.array(axis, dtype=np.float64)
    axis /= np.linalg.norm(axis)
    s = math.sin(angle / 2)
    return [float(axis[0] * s), float(axis[1] * s), float(axis[2] * s), float(math.cos(angle / 2))]


def main():
    pos, tri = icosphere(3)  # 642 vertices, 1280 triangles
    nrm = pos.copy()
    uv = np.stack([np.arctan2(pos[:, 0], pos[:, 2]) / (2 * math.pi) + 0.5, np.arcsin(np.clip(pos[:, 1], -1, 1)) / math.pi + 0.5], 1).astype(np.float32)
    half = len(tri) // 2
    blob, views, accessors = bytearray(), [], []

    def add(data, target, comp, typ, count, minmax=None):
        while len(blob) % 4:
            blob.append(0)
        views.append({"buffer": 0, "byteOffset": len(blob), "byteLength": len(data), **({"target": target} if target else {})})
        blob.extend(data)
        acc = {"bufferView": len(views) - 1, "componentType": comp, "count": count, "type": typ}
        if minmax:
            acc["min"], acc["max"] = minmax
        accessors.append(acc)
        return len(accessors) - 1

    a_pos = add(pos.tobytes(), 34962, 5126, "VEC3", len(pos), (pos.min(0).tolist(), pos.max(0).tolist()))
    a_nrm = add(nrm.tobytes(), 34962, 5126, "VEC3", len(nrm))
    a_uv = add(uv.tobytes(), 34962, 5126, "VEC2", len(uv))
    a_i0 = add(tri[:half].tobytes(), 34963, 5125, "SCALAR", half * 3)
    a_i1 = add(tri[half:].tobytes(), 34963, 5125, "SCALAR", (len(tri) - half) * 3)

    rng = np.random.default_rng(17)
    nodes = []
    for i in range(7):
        nodes.append({
            "name": "ball%d" % i,
            "mesh": 0,
            "translation": [float(x) for x in rng.uniform(-6, 6, 3).astype(np.float32)],
            "rotation": quat(rng.uniform(-1, 1, 3), float(rng.uniform(0, 3))),
            "scale": [float(np.float32(rng.uniform(0.5, 2.0)))] * 3,
        })
    nodes.append({"name": "camera", "camera": 0, "translation": [0.0, 1.0, 12.0]})

    times = np.array([0.5, 1.0, 1.5, 2.0, 2.5, 3.0], dtype=np.float32)
    a_time = add(times.tobytes(), None, 5126, "SCALAR", len(times), ([float(times[0])], [float(times[-1])]))
    samplers, channels = [], []
    for node in (1, 3, 4):
        tr = rng.uniform(-8, 8, (len(times), 3)).astype(np.float32)
        ro = np.array([quat(rng.uniform(-1, 1, 3), float(rng.uniform(0, 6))) for _ in times], dtype=np.float32)
        sc = np.repeat(rng.uniform(0.4, 2.5, (len(times), 1)).astype(np.float32), 3, axis=1)
        for path, data, typ in (("translation", tr, "VEC3"), ("rotation", ro, "VEC4"), ("scale", sc, "VEC3")):
            if node == 4 and path == "scale":
                continue  # one node without a scale channel
            out = add(np.ascontiguousarray(data).tobytes(), None, 5126, typ, len(times))
            samplers.append({"input": a_time, "output": out, "interpolation": "LINEAR"})
            channels.append({"sampler": len(samplers) - 1, "target": {"node": node, "path": path}})

    gltf = {
        "asset": {"version": "2.0", "generator": "niagara_b200 tests/golden/make_animated_gltf.py"},
        "scene": 0,
        "scenes": [{"nodes": list(range(len(nodes)))}],
        "nodes": nodes,
        "cameras": [{"type": "perspective", "perspective": {"yfov": 0.9, "znear": 0.25, "aspectRatio": 1.5}}],
        "materials": [
            {"name": "opaque", "pbrMetallicRoughness": {"baseColorFactor": [0.8, 0.7, 0.6, 1.0]}},
            {"name": "blend", "alphaMode": "BLEND", "pbrMetallicRoughness": {"baseColorFactor": [0.2, 0.4, 0.9, 0.5]}},
        ],
        "meshes": [{"primitives": [
            {"attributes": {"POSITION": a_pos, "NORMAL": a_nrm, "TEXCOORD_0": a_uv}, "indices": a_i0, "material": 0},
            {"attributes": {"POSITION": a_pos, "NORMAL": a_nrm, "TEXCOORD_0": a_uv}, "indices": a_i1, "material": 1},
        ]}],
        "animations": [{"name": "wobble", "samplers": samplers, "channels": channels}],
        "accessors": accessors,
        "bufferViews": views,
        "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(bytes(blob)).decode()}],
    }
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "animated.gltf")
    with open(out, "w") as f:
        json.dump(gltf, f, separators=(",", ":"))
    print(out, os.path.getsize(out), "bytes;", len(pos), "vertices", len(tri), "triangles")


if __name__ == "__main__":
    main()
