"""Seeded synthetic inputs for the BASELINE.json configurations.

Pure numpy; produces the reference's buffer layouts (layout.*_DTYPE).  Nothing here is on the product path — these
are the scenes bench.py and the parity tests feed to both the CUDA path and the oracle."""
import math

import numpy as np

from . import host, layout


def _unit_vectors(rng, n):
    v = rng.standard_normal((n, 3), dtype=np.float32)
    v /= np.sqrt((v * v).sum(axis=1, keepdims=True))
    return v


def synthetic_meshlets(count, seed=1, chunk=1 << 20):
    """Meshlet cull bounds with kitten-like statistics: center fp16 U(-0.5,0.5)^3, radius fp16 in
    [0.02, 0.45] skewed to small values (mean ~0.1), s8 unit cone axis, s8 cutoff U(16,127) with ~6 % == 127.
    Generated in chunks (float32 temporaries) so that 10M records do not need GBs of scratch."""
    rng = np.random.default_rng(seed)
    m = np.zeros(count, dtype=layout.MESHLET_DTYPE)
    for lo in range(0, count, chunk):
        n = min(chunk, count - lo)
        part = m[lo : lo + n]
        part["center"] = (rng.random((n, 3), dtype=np.float32) - np.float32(0.5)).astype(np.float16).view(np.uint16)
        part["radius"] = (np.float32(0.02) + np.float32(0.43) * rng.random(n, dtype=np.float32) ** 4).astype(np.float16).view(np.uint16)
        part["cone_axis"] = np.rint(_unit_vectors(rng, n) * np.float32(127.0)).astype(np.int8)
        cutoff = rng.integers(16, 128, n, dtype=np.int16)
        cutoff[rng.random(n, dtype=np.float32) < 0.06] = 127
        part["cone_cutoff"] = cutoff.astype(np.int8)
    m["vertexCount"] = 64
    m["triangleCount"] = 96
    m["dataOffset"] = (np.arange(count, dtype=np.uint64) * 40 % (1 << 32)).astype(np.uint32)
    return m


def synthetic_meshes(num_meshes, lod_counts, lod0_meshlets, seed=2):
    """Mesh table + matching meshlet ranges.
    lod_counts: LODs per mesh (int); lod0_meshlets: meshlets in LOD 0; each further LOD has half (at least 1).
    Mesh: center U(-0.1,0.1)^3, radius U(0.3,1.0), error = {0, 5e-4, 2e-3, 8e-3, ...} * radius (monotone).
    Returns (meshes, total_meshlets)."""
    rng = np.random.default_rng(seed)
    meshes = np.zeros(num_meshes, dtype=layout.MESH_DTYPE)
    meshes["center"] = rng.uniform(-0.1, 0.1, (num_meshes, 3)).astype(np.float32)
    meshes["radius"] = rng.uniform(0.3, 1.0, num_meshes).astype(np.float32)
    meshes["lodCount"] = lod_counts
    meshes["vertexCount"] = 64 * lod0_meshlets
    meshes["vertexOffset"] = (np.arange(num_meshes, dtype=np.uint64) * 64 * lod0_meshlets % (1 << 31)).astype(np.uint32)
    errors = [0.0, 5e-4, 2e-3, 8e-3, 2e-2, 5e-2, 1e-1, 2e-1]
    per_lod = [max(1, lod0_meshlets >> l) for l in range(lod_counts)]
    per_mesh = sum(per_lod)
    base = np.arange(num_meshes, dtype=np.uint64) * per_mesh
    off = 0
    for l in range(lod_counts):
        meshes["lods"]["meshletOffset"][:, l] = (base + off).astype(np.uint32)
        meshes["lods"]["meshletCount"][:, l] = per_lod[l]
        meshes["lods"]["indexCount"][:, l] = per_lod[l] * 96 * 3
        meshes["lods"]["indexOffset"][:, l] = ((base + off) * 288 % (1 << 32)).astype(np.uint32)
        meshes["lods"]["error"][:, l] = (errors[l] * meshes["radius"]).astype(np.float32)
        off += per_lod[l]
    return meshes, int(num_meshes) * per_mesh


def frustum_draws(count, mesh_indices, fov_y=math.radians(70.0), aspect=1.0, zmin=4.0, zmax=190.0, seed=3, fill=0.9):
    """Draws placed INSIDE the view frustum of the reference's default camera (origin, identity orientation; view =
    scale(1,1,-1), so view-space z = -world z), scale/orientation as in the reference's random scene
    (niagara.cpp:984-993: scale in [2,4), rotation <= 90 degrees about a random axis)."""
    rng = np.random.default_rng(seed)
    d = np.zeros(count, dtype=layout.MESHDRAW_DTYPE)
    # uniform in volume: z^3 uniform
    z = (rng.uniform(zmin**3, zmax**3, count)) ** (1.0 / 3.0)
    ty = math.tan(fov_y / 2) * fill
    tx = ty * aspect
    d["position"][:, 0] = (rng.uniform(-1, 1, count) * tx * z).astype(np.float32)
    d["position"][:, 1] = (rng.uniform(-1, 1, count) * ty * z).astype(np.float32)
    d["position"][:, 2] = (-z).astype(np.float32)
    d["scale"] = ((rng.random(count) + 1.0) * 2.0).astype(np.float32)
    axis = _unit_vectors(rng, count).astype(np.float64)
    angle = np.radians(rng.random(count) * 90.0)
    d["orientation"][:, :3] = (axis * np.sin(angle * 0.5)[:, None]).astype(np.float32)
    d["orientation"][:, 3] = np.cos(angle * 0.5).astype(np.float32)
    d["meshIndex"] = mesh_indices
    return d


def synthetic_depth(width, height, znear=0.1, occluders=200, seed=4, zrange=(5.0, 150.0), max_extent=0.18):
    """Prior-frame depth target stand-in: reverse-Z (clear = 0 = infinitely far, niagara.cpp:1763), `occluders` random
    rectangles / discs at view depth z in U(zrange) written as depth = znear / z, nearest wins (max)."""
    rng = np.random.default_rng(seed)
    depth = np.zeros((height, width), dtype=np.float32)
    for i in range(occluders):
        z = rng.uniform(*zrange)
        dval = np.float32(znear / z)
        cx, cy = rng.uniform(0, width), rng.uniform(0, height)
        ex, ey = rng.uniform(0.02, max_extent) * width * 0.5, rng.uniform(0.02, max_extent) * height * 0.5
        x0, x1 = int(max(0, cx - ex)), int(min(width, cx + ex))
        y0, y1 = int(max(0, cy - ey)), int(min(height, cy + ey))
        if x1 <= x0 or y1 <= y0:
            continue
        region = depth[y0:y1, x0:x1]
        if i % 2 == 0:
            np.maximum(region, dval, out=region)
        else:
            yy, xx = np.ogrid[y0:y1, x0:x1]
            mask = ((xx - cx) / max(ex, 1.0)) ** 2 + ((yy - cy) / max(ey, 1.0)) ** 2 <= 1.0
            region[mask] = np.maximum(region[mask], dval)
    return depth


class Scene:
    """Everything one configuration needs, in host memory."""

    def __init__(self, name, meshes, meshlets, draws, depth, camera, screen, visibility_bits, note="", helpers=None):
        self.helpers = helpers  # None = niagara_b200.host; bench.py's CPU arms pass the checker's own helpers
        self.name = name
        self.meshes = meshes
        self.meshlets = meshlets
        self.draws = draws
        self.depth = depth
        self.camera = camera
        self.screen = screen  # (width, height) of the depth target / swapchain
        self.visibility_bits = visibility_bits
        self.note = note

    def cull_data(self, **toggles):
        return (self.helpers or host).cull_data(self.camera, self.screen[0], self.screen[1], len(self.draws), **toggles)

    def __getstate__(self):
        st = dict(self.__dict__)
        st["helpers"] = None
        c = self.camera
        st["camera"] = (tuple(c.position), tuple(c.orientation), float(c.fovY), float(c.znear))
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self.camera = host.make_camera(*st["camera"])


def reference_random_scene(meshes, meshlets, draw_count, screen=(1024, 768), depth_seed=4, occluders=60, name="reference-random"):
    """C1/C2 recipe: the reference's own PCG32 scene (niagara.cpp:969-998) over a given geometry table, default camera."""
    draws = host.random_draws(draw_count, len(meshes))
    bits, _ = host.visibility_offsets(draws, meshes)
    cam = host.make_camera()
    depth = synthetic_depth(screen[0], screen[1], cam.znear, occluders=occluders, seed=depth_seed)
    return Scene(name, meshes, meshlets, draws, depth, cam, screen, bits)


def config2_scene(draw_count=1_000_000, num_meshes=1024, screen=(4096, 4096), seed=11, lod0_meshlets=64):
    """BASELINE configs[1]: 1M synthetic MeshDraws (reference PCG32 recipe), 4 LODs, 4Kx4K synthetic prior-frame depth.
    num_meshes is SURVEY §8(d)'s knob (1024: the mesh table is L2 resident; 1 000 000: every draw gathers its own Mesh)."""
    meshes, nmeshlets = synthetic_meshes(num_meshes, 4, lod0_meshlets, seed=seed)
    meshlets = synthetic_meshlets(nmeshlets, seed=seed + 1)
    s = reference_random_scene(meshes, meshlets, draw_count, screen=screen, depth_seed=seed + 2, occluders=200, name="C2")
    s.note = "%d draws (reference PCG32 scene), %d meshes x 4 LODs, %dx%d depth" % (draw_count, num_meshes, screen[0], screen[1])
    return s


def config4_scene(draw_count=1_000_000, meshlets_per_draw=10, screen=(4096, 4096), seed=21, occluders=120, helpers=None):
    """BASELINE configs[3]: 10M synthetic meshlets / 1M draws: one UNIQUE mesh per draw (so Meshlet[] = 240 MB >> L2,
    SURVEY F9), every draw inside the frustum so that the cluster pass really tests ~all meshlet instances."""
    h = helpers or host
    meshes, nmeshlets = synthetic_meshes(draw_count, 1, meshlets_per_draw, seed=seed)
    meshlets = synthetic_meshlets(nmeshlets, seed=seed + 1)
    cam = h.make_camera()
    aspect = screen[0] / screen[1]
    draws = frustum_draws(draw_count, np.arange(draw_count, dtype=np.uint32), fov_y=cam.fovY, aspect=aspect, seed=seed + 2)
    bits, _ = h.visibility_offsets(draws, meshes)
    depth = synthetic_depth(screen[0], screen[1], cam.znear, occluders=occluders, seed=seed + 3, zrange=(60.0, 190.0), max_extent=0.12)
    s = Scene("C4", meshes, meshlets, draws, depth, cam, screen, bits, helpers=helpers)
    s.note = "%d draws x %d unique meshlets each (%d meshlet instances), all inside the frustum, %dx%d depth" % (
        draw_count,
        meshlets_per_draw,
        nmeshlets,
        screen[0],
        screen[1],
    )
    return s


def config3_scene(golden_dir, draw_count=10000, screen=(4096, 4096), seed=31):
    """BASELINE configs[2] stand-in (no bistro.gltf in the image): the reference-cooked kitten.obj (309 LOD-0 meshlets, 792 over its LOD chain,
    tests/golden/kitten.nvcg + its vertex / meshlet data) instanced `draw_count` times inside the frustum = ~3.1M LOD-0 meshlet
    instances, with the geometry the device rasteriser needs (scene.vertices: uint8 view of 16-byte Vertex records,
    scene.meshletdata: uint32).  The prior-frame depth is produced on the device, so scene.depth is all zeros."""
    import os

    meshes, meshlets, _ = layout.load_nvcg(os.path.join(golden_dir, "kitten.nvcg"))
    z = np.load(os.path.join(golden_dir, "kitten_cook.npz"))
    positions, meshletdata = z["positions"], z["meshletdata"]
    vertices = np.zeros((len(positions), 8), dtype=np.uint16)
    vertices[:, :3] = positions
    rng = np.random.default_rng(seed)
    cam = host.make_camera()
    ty = math.tan(cam.fovY / 2) * 0.95
    tx = ty * screen[0] / screen[1]
    depth_z = rng.uniform(8.0**3, 190.0**3, draw_count) ** (1.0 / 3.0)
    d = np.zeros(draw_count, dtype=layout.MESHDRAW_DTYPE)
    d["position"][:, 0] = (rng.uniform(-1, 1, draw_count) * tx * depth_z).astype(np.float32)
    d["position"][:, 1] = (rng.uniform(-1, 1, draw_count) * ty * depth_z).astype(np.float32)
    d["position"][:, 2] = (-depth_z).astype(np.float32)  # the camera looks down -Z (niagara.cpp:1492)
    d["scale"] = rng.uniform(1.5, 4.5, draw_count).astype(np.float32)
    q = rng.standard_normal((draw_count, 4))
    d["orientation"] = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    bits, _ = host.visibility_offsets(d, meshes)
    s = Scene("C3", meshes, meshlets, d, np.zeros((screen[1], screen[0]), np.float32), cam, screen, bits)
    s.vertices = np.ascontiguousarray(vertices).view(np.uint8).reshape(-1)
    s.meshletdata = np.ascontiguousarray(meshletdata, dtype=np.uint32)
    s.note = "%d draws of the reference-cooked kitten (%d LOD-0 meshlets each), %dx%d, depth produced on the device" % (draw_count, int(meshes["lods"]["meshletCount"][0, 0]), screen[0], screen[1])
    return s


def instanced_scene(nvcg_path, draw_count, screen=(1024, 768), name="instanced"):
    """C1 / C3 stand-in: geometry cooked by the reference's own scene.cpp (tests/golden/*.nvcg) instanced with the
    reference's random-scene recipe."""
    meshes, meshlets, _ = layout.load_nvcg(nvcg_path)
    return reference_random_scene(meshes, meshlets, draw_count, screen=screen, name=name)
