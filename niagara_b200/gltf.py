"""glTF 2.0 import for the visibility path through the C ABI (nvc_gltf_*): SURVEY §8(f) row N3 — what the reference's loadScene
(src/scene.cpp:473-853) hands to the frame loop.  Host only."""
import ctypes
import os

import numpy as np

from . import layout
from .lib import NvcError, load_library

VERTEX_DTYPE = np.dtype([("vx", "<u2"), ("vy", "<u2"), ("vz", "<u2"), ("tp", "<u2"), ("np", "<u4"), ("tu", "<u2"), ("tv", "<u2")])
assert VERTEX_DTYPE.itemsize == 16


class GltfScene:
    def __init__(self, path, first_mesh_index=0, material_offset=1):
        """material_offset: index 0 is the reference's dummy material (scene.cpp:551), so imported materials start at 1"""
        self._lib = load_library()
        data = open(path, "rb").read()
        self._scene = ctypes.c_void_p()
        status = self._lib.nvc_gltf_import(data, len(data), os.path.dirname(os.path.abspath(path)).encode(), int(first_mesh_index), int(material_offset), ctypes.byref(self._scene))
        if status != 0:
            raise NvcError("%s: %s" % (path, self._lib.nvc_status_string(status).decode()))
        self.info = layout.GltfInfo()
        self._lib.nvc_gltf_info(self._scene, ctypes.byref(self.info))
        i = self.info
        self.draws = np.zeros(i.draw_count, dtype=layout.MESHDRAW_DTYPE)
        self.animations = np.zeros(i.animation_count, dtype=layout.ANIMATION_DTYPE)
        self.keyframes = np.zeros(i.keyframe_count, dtype=layout.KEYFRAME_DTYPE)
        self.mesh_scale = np.zeros(i.primitive_count, dtype=np.float32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if len(a) else None
        self._lib.nvc_gltf_scene_arrays(self._scene, p(self.draws), p(self.animations), p(self.keyframes), p(self.mesh_scale))

    def primitive(self, index):
        """(Vertex[] as VERTEX_DTYPE, uint32 indices) of one triangle primitive = one reference mesh before cooking"""
        vc, ic = ctypes.c_uint32(), ctypes.c_uint32()
        status = self._lib.nvc_gltf_primitive_size(self._scene, int(index), ctypes.byref(vc), ctypes.byref(ic))
        if status != 0:
            raise NvcError(self._lib.nvc_status_string(status).decode())
        v = np.zeros(vc.value, dtype=VERTEX_DTYPE)
        ix = np.zeros(ic.value, dtype=np.uint32)
        status = self._lib.nvc_gltf_primitive_data(self._scene, int(index), v.ctypes.data_as(ctypes.c_void_p), ix.ctypes.data_as(ctypes.c_void_p))
        if status != 0:
            raise NvcError(self._lib.nvc_status_string(status).decode())
        return v, ix

    def close(self):
        if self._scene:
            self._lib.nvc_gltf_free(self._scene)
            self._scene = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
