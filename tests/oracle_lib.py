"""ctypes wrapper of oracle/liboracle.so — the CPU restatement used ONLY as the checker by the tests, smoke() and
bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

from niagara_b200 import layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

vp = ctypes.c_void_p


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    lib = ctypes.CDLL(path)
    lib.orc_drawcull.restype = ctypes.c_int
    lib.orc_drawcull.argtypes = [ctypes.POINTER(layout.CullData), ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.POINTER(layout.HiZ), ctypes.POINTER(layout.Limits), vp, ctypes.c_int]
    lib.orc_clustercull.restype = ctypes.c_int
    lib.orc_clustercull.argtypes = [ctypes.POINTER(layout.CullData), ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, ctypes.POINTER(layout.HiZ), ctypes.POINTER(layout.Limits), ctypes.c_int]
    lib.orc_taskcull.restype = ctypes.c_int
    lib.orc_taskcull.argtypes = [ctypes.POINTER(layout.CullData), ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, ctypes.POINTER(layout.HiZ), ctypes.c_int]
    lib.orc_depth_pyramid.restype = ctypes.c_int
    lib.orc_depth_pyramid.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(layout.HiZ), ctypes.c_int]
    lib.orc_decode_clusters.restype = ctypes.c_int
    lib.orc_decode_clusters.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.orc_half_to_float.restype = ctypes.c_float
    lib.orc_half_to_float.argtypes = [ctypes.c_uint16]
    lib.orc_rotate_quat.argtypes = [vp, vp, vp]
    lib.orc_project_sphere.restype = ctypes.c_int
    lib.orc_project_sphere.argtypes = [vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, vp]
    lib.orc_occlusion_mip.restype = ctypes.c_float
    lib.orc_occlusion_mip.argtypes = [vp, ctypes.c_float, ctypes.c_float]
    lib.orc_sample_min.restype = ctypes.c_float
    lib.orc_sample_min.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, ctypes.c_float]
    lib.orc_ceil_log2.restype = ctypes.c_int
    lib.orc_ceil_log2.argtypes = [ctypes.c_float]
    lib.orc_cone_cull.restype = ctypes.c_int
    lib.orc_cone_cull.argtypes = [vp, ctypes.c_float, vp, ctypes.c_float]
    lib.orc_transform_point.argtypes = [vp, vp, vp]
    lib.orc_hardware_threads.restype = ctypes.c_int
    lib.orc_visibility_offsets.restype = ctypes.c_uint32
    lib.orc_visibility_offsets.argtypes = [vp, ctypes.c_uint32, vp]
    lib.orc_hiz_layout.restype = ctypes.c_int
    lib.orc_hiz_layout.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(layout.HiZ)]
    lib.orc_cull_data.restype = None
    lib.orc_cull_data.argtypes = [ctypes.POINTER(layout.Camera), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(layout.CullOptions), ctypes.POINTER(layout.CullData)]
    lib.orc_pass_data.restype = None
    lib.orc_pass_data.argtypes = [ctypes.POINTER(layout.CullData), ctypes.c_int, ctypes.c_uint32, ctypes.POINTER(layout.CullData)]
    _LIB = lib
    return lib


class CheckerHost:
    """Host-side helpers served by the ORACLE library (orc_*), with the surface of niagara_b200.host that the scene
    generators and OraclePath need.  bench.py's CPU arms use it so that they never load the product library."""

    @staticmethod
    def visibility_offsets(draws, meshes):
        meshes = np.ascontiguousarray(meshes)
        return int(load().orc_visibility_offsets(draws.ctypes.data_as(vp), len(draws), meshes.ctypes.data_as(vp))), 1

    @staticmethod
    def make_camera(position=(0.0, 0.0, 0.0), orientation=(0.0, 0.0, 0.0, 1.0), fov_y=1.2217304763960306, znear=0.1):
        cam = layout.Camera()
        cam.position[:] = position
        cam.orientation[:] = orientation
        cam.fovY = fov_y
        cam.znear = znear
        return cam

    @staticmethod
    def cull_data(camera, screen_width, screen_height, draw_count, draw_distance=200.0, culling=True, lod=True, occlusion=True, cluster_occlusion=True, mesh_shading=True, debug_lod_step=0):
        opts = layout.CullOptions(float(draw_distance), int(culling), int(lod), int(occlusion), int(cluster_occlusion), int(mesh_shading), int(debug_lod_step))
        out = layout.CullData()
        load().orc_cull_data(ctypes.byref(camera), int(screen_width), int(screen_height), int(draw_count), ctypes.byref(opts), ctypes.byref(out))
        return out

    @staticmethod
    def hiz_layout(depth_width, depth_height):
        hiz = layout.HiZ()
        assert load().orc_hiz_layout(int(depth_width), int(depth_height), ctypes.byref(hiz)) == 0
        return hiz

    @staticmethod
    def pass_data(cull_data, for_drawcull, post_pass):
        out = layout.CullData()
        load().orc_pass_data(ctypes.byref(cull_data), int(for_drawcull), int(post_pass), ctypes.byref(out))
        return out


def _p(a):
    return a.ctypes.data_as(vp) if a is not None else None


def _round_up(v, m):
    return (v + m - 1) // m * m


class OraclePath:
    """Same surface as niagara_b200.path.VisibilityPath, on host arrays, through the oracle."""

    def __init__(self, meshes, meshlets, draws, depth_width, depth_height, task_wglimit=layout.TASK_WGLIMIT, cluster_limit=layout.CLUSTER_LIMIT, mesh_shading=True, threads=1, cmd_capacity=None, cluster_capacity=None, helpers=None):
        """helpers: None = the product's host helpers (niagara_b200.host, what the tests compare against), or CheckerHost."""
        if helpers is None:
            from niagara_b200 import host
        else:
            host = helpers
        self.helpers = helpers

        self.lib = load()
        self.threads = threads
        self.mesh_shading = mesh_shading
        self.meshes = np.ascontiguousarray(meshes)
        self.meshlets = np.ascontiguousarray(meshlets)
        self.draws = np.ascontiguousarray(draws)
        self.limits = layout.Limits(task_wglimit, cluster_limit)
        n = len(draws)
        self.dvb = np.zeros(max(1, n), dtype=np.uint32)
        ncmd = cmd_capacity if cmd_capacity is not None else min(_round_up(task_wglimit, 64), max(64, _round_up(n * 8, 64)))
        self.cmd_capacity = ncmd
        self.dcb = np.zeros(max(ncmd * 20, n * 24), dtype=np.uint8)
        self.dccb = np.zeros(4, dtype=np.uint32)
        ncl = cluster_capacity if cluster_capacity is not None else min(_round_up(cluster_limit, 256), _round_up(ncmd * 64, 256))
        self.cib = np.zeros(ncl, dtype=np.uint32)
        self.ccb = np.zeros(4, dtype=np.uint32)
        self.mvb = None
        self.depth_width, self.depth_height = depth_width, depth_height
        self.hiz = host.hiz_layout(depth_width, depth_height)
        self.pyramid_texels = np.zeros(self.hiz.total_texels, dtype=np.float32)
        self.hiz.texels = self.pyramid_texels.ctypes.data
        self.lod_out = np.zeros(max(1, n), dtype=np.uint8)

    def set_visibility_bits(self, count):
        self.mvb = np.zeros(max(1, (int(count) + 31) // 32), dtype=np.uint32)

    def _pass_data(self, cull_data, for_drawcull, post_pass):
        if self.helpers is not None:
            return self.helpers.pass_data(cull_data, for_drawcull, post_pass)
        from niagara_b200.lib import load_library

        out = layout.CullData()
        load_library().nvc_host_pass_data(ctypes.byref(cull_data), for_drawcull, post_pass, ctypes.byref(out))
        return out

    def cull(self, cull_data, late, post_pass=0, task=None):
        task = self.mesh_shading if task is None else task
        pd = self._pass_data(cull_data, 1, post_pass)
        s = self.lib.orc_drawcull(ctypes.byref(pd), int(late), int(task), _p(self.draws), _p(self.meshes), _p(self.dvb), _p(self.dcb), _p(self.dccb), ctypes.byref(self.hiz), ctypes.byref(self.limits), _p(self.lod_out), self.threads)
        assert s == 0, s

    def render_clusters(self, cull_data, late, post_pass=0, cluster_backface=None):
        pd = self._pass_data(cull_data, 0, post_pass)
        if cluster_backface is not None:
            pd.clusterBackfaceEnabled = int(cluster_backface)
        s = self.lib.orc_clustercull(ctypes.byref(pd), int(late), _p(self.dcb), _p(self.dccb), _p(self.draws), _p(self.meshlets), _p(self.mvb), _p(self.cib), _p(self.ccb), ctypes.byref(self.hiz), ctypes.byref(self.limits), self.threads)
        assert s == 0, s

    def task_shading(self, cull_data, late, payloads, emit_counts, post_pass=0, cluster_backface=None):
        pd = self._pass_data(cull_data, 0, post_pass)
        if cluster_backface is not None:
            pd.clusterBackfaceEnabled = int(cluster_backface)
        s = self.lib.orc_taskcull(ctypes.byref(pd), int(late), _p(self.dcb), _p(self.dccb), _p(self.draws), _p(self.meshlets), _p(self.mvb), _p(payloads), _p(emit_counts), ctypes.byref(self.hiz), self.threads)
        assert s == 0, s

    def pyramid(self, depth):
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        s = self.lib.orc_depth_pyramid(_p(depth), self.depth_width, self.depth_height, ctypes.byref(self.hiz), self.threads)
        assert s == 0, s

    def frame(self, cull_data, depth, post_passes=False, cluster_backface=None):
        self.cull(cull_data, late=False)
        if self.mesh_shading:
            self.render_clusters(cull_data, late=False, cluster_backface=cluster_backface)
        self.pyramid(depth)
        self.cull(cull_data, late=True)
        if self.mesh_shading:
            self.render_clusters(cull_data, late=True, cluster_backface=cluster_backface)
        if post_passes:
            self.cull(cull_data, late=True, post_pass=1)
            if self.mesh_shading:
                self.render_clusters(cull_data, late=True, post_pass=1, cluster_backface=cluster_backface)

    def decode_clusters(self, want_records=True):
        slots = int(self.ccb[2]) * 256
        records = np.zeros((max(slots, 1), 4), dtype=np.uint32) if want_records else None
        stats = np.zeros(4, dtype=np.uint32)
        s = self.lib.orc_decode_clusters(_p(self.cib), _p(self.ccb), _p(self.dcb), _p(self.meshlets), _p(records), _p(stats))
        assert s == 0
        return (records[:slots] if want_records else None), stats

    # readback in the same shape as VisibilityPath
    def read_counts(self):
        return self.dccb.copy(), self.ccb.copy()

    def read_task_commands(self, count):
        return self.dcb[: int(count) * 20].view(layout.MESHTASKCOMMAND_DTYPE).copy()

    def read_draw_commands(self, count):
        return self.dcb[: int(count) * 24].view(layout.MESHDRAWCOMMAND_DTYPE).copy()

    def read_cluster_indices(self, count):
        return self.cib[: int(count)].copy()

    def level(self, l):
        w, h = self.hiz.level_size(l)
        off = self.hiz.level_offset[l]
        return self.pyramid_texels[off : off + w * h].reshape(h, w)


def cluster_pairs(cluster_indices, task_commands):
    """Order-independent view of a cluster pass result: sorted (drawId, meshlet index) pairs — the reference's
    consumer decodes exactly this (meshlet.mesh.glsl:94-103: command = taskCommands[ci & 0xffffff],
    mi = command.taskOffset + (ci >> 24))."""
    ci = np.asarray(cluster_indices, dtype=np.uint32)
    cmd = task_commands[ci & 0xFFFFFF]
    mi = cmd["taskOffset"].astype(np.uint64) + (ci >> 24)
    pairs = (cmd["drawId"].astype(np.uint64) << 32) | mi
    return np.sort(pairs)


def sorted_commands(cmds):
    """Order-independent view of a command list."""
    return np.sort(cmds, order=list(cmds.dtype.names))
