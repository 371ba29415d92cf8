"""Reader of the reference's scene cache (.cache v7, src/scenecache.cpp) through the C ABI (nvc_scene_cache_*):
SURVEY §8(f) row N2 — the on-disk format that feeds the visibility path.  Host only."""
import ctypes
import mmap

import numpy as np

from . import host, layout
from .lib import NvcError, load_library

_DTYPES = {
    "meshlets": layout.MESHLET_DTYPE,
    "meshes": layout.MESH_DTYPE,
    "draws": layout.MESHDRAW_DTYPE,
    "animations": layout.ANIMATION_DTYPE,
    "keyframes": layout.KEYFRAME_DTYPE,
    "meshletdata": np.dtype("<u4"),
    "indices": np.dtype("<u4"),
    "omm_descs": np.dtype("<u4"),
    "meshletvtx0": np.dtype("<u2"),
}


class SceneCache:
    """A parsed cache file.  `section(name)` returns a numpy array (a copy; the meshopt streams of a compressed cache —
    vertices, indices, meshlet data, RT positions — are decoded)."""

    def __init__(self, path):
        self.path = path
        self._file = open(path, "rb")
        self._map = mmap.mmap(self._file.fileno(), 0, access=mmap.ACCESS_READ)
        self._buf = (ctypes.c_char * len(self._map)).from_buffer_copy(self._map) if len(self._map) < (1 << 20) else None
        self.size = len(self._map)
        self.info = layout.SceneCacheInfo()
        self._lib = load_library()
        status = self._lib.nvc_scene_cache_parse(self._address(), self.size, ctypes.byref(self.info))
        if status != 0:
            raise NvcError("%s: %s" % (path, self._lib.nvc_status_string(status).decode()))
        self.header = self.info.header

    def _address(self):
        if self._buf is not None:
            return ctypes.addressof(self._buf)
        # large files: hand the mapping itself to the library (read-only mmap -> numpy view -> address)
        self._view = np.frombuffer(self._map, dtype=np.uint8)
        return self._view.ctypes.data

    def section_info(self, name):
        return self.info.sections[layout.CACHE_SECTIONS.index(name)]

    def section(self, name):
        idx = layout.CACHE_SECTIONS.index(name)
        sec = self.info.sections[idx]
        raw = np.zeros(int(sec.decoded_bytes), dtype=np.uint8)
        status = self._lib.nvc_scene_cache_read(self._address(), self.size, ctypes.byref(self.info), idx, raw.ctypes.data, raw.nbytes)
        if status != 0:
            raise NvcError("%s[%s]: %s" % (self.path, name, self._lib.nvc_status_string(status).decode()))
        if name == "texture_paths":
            return [bytes(raw[i * 256 : (i + 1) * 256]).split(b"\0", 1)[0].decode() for i in range(sec.count)]
        return raw.view(_DTYPES[name]) if name in _DTYPES else raw

    def camera(self):
        c = self.header.camera
        return host.make_camera(tuple(c.position), tuple(c.orientation), float(c.fovY), float(c.znear))

    def close(self):
        self._view = None
        self._map.close()
        self._file.close()


def load_scene(path, screen=(1024, 768), depth=None, name=None):
    """Scene (niagara_b200.scenes.Scene) from a cache file: Mesh[] / Meshlet[] / MeshDraw[] exactly as stored, the
    cache's camera, meshletVisibilityOffset as stored (niagara.cpp:1003-1020 assigns it before saving)."""
    from . import scenes

    c = SceneCache(path)
    meshes, meshlets, draws = c.section("meshes"), c.section("meshlets"), c.section("draws")
    bits = 0
    if len(draws):
        counts = meshes["lods"]["meshletCount"].max(axis=1)[draws["meshIndex"]]
        bits = int((draws["meshletVisibilityOffset"].astype(np.int64) + counts).max())
    cam = c.camera()
    if depth is None:
        depth = scenes.synthetic_depth(screen[0], screen[1], cam.znear, occluders=40, seed=9)
    s = scenes.Scene(name or path, meshes, meshlets, draws, depth, cam, screen, bits)
    s.animations, s.keyframes = c.section("animations"), c.section("keyframes")
    c.close()
    return s
