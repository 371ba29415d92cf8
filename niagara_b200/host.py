"""Host-side helpers mirroring the reference's CPU code around the path (implemented in C++: csrc/nvc_host.cpp)."""
import ctypes
import math

import numpy as np

from . import layout
from .lib import load_library


def previous_pow2(v):
    return int(load_library().nvc_previous_pow2(int(v)))


def image_mip_levels(w, h):
    return int(load_library().nvc_image_mip_levels(int(w), int(h)))


def hiz_layout(depth_width, depth_height):
    hiz = layout.HiZ()
    status = load_library().nvc_hiz_layout(int(depth_width), int(depth_height), ctypes.byref(hiz))
    if status != 0:
        raise ValueError("invalid depth size %dx%d" % (depth_width, depth_height))
    return hiz


def random_draws(draw_count, mesh_count, scene_radius=300.0):
    """The reference's built-in random scene (niagara.cpp:969-998, PCG32 state 0x42)."""
    draws = np.zeros(int(draw_count), dtype=layout.MESHDRAW_DTYPE)
    load_library().nvc_host_random_draws(draws.ctypes.data_as(ctypes.c_void_p), int(draw_count), int(mesh_count), float(scene_radius))
    return draws


def visibility_offsets(draws, meshes):
    """niagara.cpp:1002-1020 — fills draws['meshletVisibilityOffset'] in place; returns (bit count, postPass mask)."""
    mask = ctypes.c_uint32(0)
    meshes = np.ascontiguousarray(meshes)
    count = load_library().nvc_host_visibility_offsets(draws.ctypes.data_as(ctypes.c_void_p), len(draws), meshes.ctypes.data_as(ctypes.c_void_p), ctypes.byref(mask))
    return int(count), int(mask.value)


def make_camera(position=(0.0, 0.0, 0.0), orientation=(0.0, 0.0, 0.0, 1.0), fov_y=math.radians(70.0), znear=0.1):
    """Default camera of the reference: origin, identity, fovY 70 deg, znear 0.1 (niagara.cpp:833-837)."""
    cam = layout.Camera()
    cam.position[:] = position
    cam.orientation[:] = orientation
    cam.fovY = fov_y
    cam.znear = znear
    return cam


def cull_data(camera, screen_width, screen_height, draw_count, draw_distance=200.0, culling=True, lod=True, occlusion=True, cluster_occlusion=True, mesh_shading=True, debug_lod_step=0):
    """CullData exactly as niagara.cpp:1487-1516 builds it (runtime toggles niagara.cpp:31-44)."""
    opts = layout.CullOptions(float(draw_distance), int(culling), int(lod), int(occlusion), int(cluster_occlusion), int(mesh_shading), int(debug_lod_step))
    out = layout.CullData()
    load_library().nvc_host_cull_data(ctypes.byref(camera), int(screen_width), int(screen_height), int(draw_count), ctypes.byref(opts), ctypes.byref(out), None)
    return out


def projection(camera, screen_width, screen_height):
    """The infinite reverse-Z projection of niagara.cpp:424-432 for this camera / aspect, column-major float32[16]."""
    opts = layout.CullOptions(200.0, 1, 1, 1, 1, 1, 0)
    out = layout.CullData()
    proj = (ctypes.c_float * 16)()
    load_library().nvc_host_cull_data(ctypes.byref(camera), int(screen_width), int(screen_height), 0, ctypes.byref(opts), ctypes.byref(out), proj)
    return np.array(proj, dtype=np.float32)


def quat_from_axis_angle(axis, angle):
    ax = np.asarray(axis, dtype=np.float64)
    ax = ax / np.linalg.norm(ax)
    s = math.sin(angle * 0.5)
    return (float(ax[0] * s), float(ax[1] * s), float(ax[2] * s), float(math.cos(angle * 0.5)))


def animate(animations, keyframes, animation_time, draws):
    """niagara.cpp:1362-1390: evaluates the keyframe tracks at animation_time into draws (in place) and returns the
    packed update list (indices uint32[n], values MeshDraw[n]) for VisibilityPath.update_draws."""
    import ctypes

    import numpy as np

    from . import layout
    from .lib import NvcError, load_library

    n = len(animations)
    idx = np.zeros(max(n, 1), dtype=np.uint32)
    val = np.zeros(max(n, 1), dtype=layout.MESHDRAW_DTYPE)
    lib = load_library()
    got = lib.nvc_host_animate(animations.ctypes.data if n else None, n, keyframes.ctypes.data if len(keyframes) else None, len(keyframes), float(animation_time), draws.ctypes.data, len(draws), idx.ctypes.data, val.ctypes.data, len(idx))
    if got < 0:
        raise NvcError("nvc_host_animate: " + lib.nvc_status_string(got).decode())
    return idx[:got], val[:got]
