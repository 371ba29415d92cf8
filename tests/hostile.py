"""Shared by the GPU fuzz test and the reference-shader fuzz test: a scene full of non-finite / degenerate inputs."""
import os

import numpy as np

from niagara_b200 import layout, scenes


def hostile_scene(golden_dir, n, screen=(1024, 512)):
    """zero, negative and huge scales, NaN / inf positions, un-normalised and zero quaternions, fp16 inf / NaN /
    subnormal meshlet bounds, extreme cone bytes, inf and denormal depth texels."""
    meshes, meshlets, _ = layout.load_nvcg(os.path.join(golden_dir, "kitten_pirate.nvcg"))
    meshlets = meshlets.copy()
    rng = np.random.default_rng(99)
    k, zq = max(1, 3000 * n // 40000), max(1, 500 * n // 40000)  # 3000 / 500 at the GPU test's n = 40000
    s = scenes.reference_random_scene(meshes, meshlets, n, screen=screen)
    d = s.draws
    d["position"] *= 0.25
    special = np.array([0.0, -0.0, 1e-30, -1e-30, 1e30, -1e30, np.inf, -np.inf, np.nan, 1e-45, 3.4e38], dtype=np.float32)
    for field, cols in (("position", 3), ("orientation", 4)):
        idx = rng.integers(0, n, k)
        d[field][idx, rng.integers(0, cols, k)] = special[rng.integers(0, len(special), k)]
    idx = rng.integers(0, n, k)
    d["scale"][idx] = special[rng.integers(0, len(special), k)]
    d["orientation"][rng.integers(0, n, zq)] = 0.0
    # meshlet bounds: fp16 specials and extreme s8 cones
    m = s.meshlets
    h_special = np.array([0x0000, 0x8000, 0x0001, 0x03FF, 0x7BFF, 0xFBFF, 0x7C00, 0xFC00, 0x7E00, 0x3C00], dtype=np.uint16)
    mi = rng.integers(0, len(m), 300)
    m["center"][mi, rng.integers(0, 3, 300)] = h_special[rng.integers(0, len(h_special), 300)]
    mi = rng.integers(0, len(m), 150)
    m["radius"][mi] = h_special[rng.integers(0, len(h_special), 150)]
    mi = rng.integers(0, len(m), 300)
    m["cone_axis"][mi] = rng.choice(np.array([-128, -127, 0, 127], dtype=np.int8), (300, 3))
    m["cone_cutoff"][rng.integers(0, len(m), 300)] = rng.choice(np.array([-128, -127, 0, 127], dtype=np.int8), 300)
    s.meshes = s.meshes.copy()
    depth = s.depth
    depth[::7, ::5] = np.float32(np.inf)
    depth[3::11, 2::13] = np.float32(1e-38)
    return s
