"""Widening N4 — meshlet bounds + normal cones (csrc/nvc_cook.cuh / nvc_cook_meshlet_bounds).
The checker is the REFERENCE itself: the Meshlet[] its cooker wrote (scene.cpp appendMeshlet ->
meshopt_computeMeshletBounds) into caches produced by oracle/refscene/write_cache.  CPU tests run the product's header
compiled for the host (tests/cook_host.cpp); the -m gpu test runs the kernel through the C ABI."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from niagara_b200 import layout, scene_cache

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
CULL_FIELDS = 12  # center[3] + radius (fp16) + cone_axis[3] + cone_cutoff (int8): the first 12 bytes of a Meshlet


def _host_lib():
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libcookhost.so")
    srcs = [os.path.join(HERE, "cook_host.cpp"), os.path.join(ROOT, "niagara_b200", "csrc", "nvc_cook.cuh"), os.path.join(ROOT, "include", "niagara_cull.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-x", "c++", srcs[0], "-o", so], check=True)
    return ctypes.CDLL(so)


def _inputs(cache_path):
    c = scene_cache.SceneCache(cache_path)
    assert not c.header.compressed
    vertices = np.ascontiguousarray(c.section("vertices"))  # raw bytes, 16 per vertex
    return vertices, c.section("meshletdata"), c.section("meshlets")


def _scrub(meshlets):
    m = meshlets.copy()
    m.view(np.uint8).reshape(len(m), 24)[:, :CULL_FIELDS] = 0xCD
    return m


def _check(got, want, what):
    g = got.view(np.uint8).reshape(len(got), 24)
    w = want.view(np.uint8).reshape(len(want), 24)
    bad = np.nonzero((g != w).any(1))[0]
    assert len(bad) == 0, (what, len(bad), bad[:5], got[bad[:3]], want[bad[:3]])


def test_host_instantiation_reproduces_reference_cooker_animated():
    lib = _host_lib()
    vertices, data, want = _inputs(os.path.join(GOLDEN, "animated.raw.cache"))
    got = _scrub(want)
    lib.cookhost_meshlet_bounds(vertices.ctypes.data, len(vertices) // 16, data.ctypes.data, len(data), got.ctypes.data, len(got))
    _check(got, want, "animated")
    assert len(np.unique(want["cone_cutoff"])) > 3 and (want["radius"] != 0).all()


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference checkout to cook kitten.obj / pirate.obj now")
def test_host_instantiation_reproduces_reference_cooker_real_meshes(tmp_path):
    """792 + ~150 meshlets cooked by the reference right now (all LODs, short and long reference lists, wide cones)."""
    lib = _host_lib()
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    tool = os.path.join(ROOT, "oracle", "_ref", "write_cache")
    total = 0
    for name, src in (("kitten", "/root/reference/data/kitten.obj"), ("pirate", "/root/reference/extern/meshoptimizer/demo/pirate.obj"), ("pirateglb", "/root/reference/extern/meshoptimizer/demo/pirate.glb")):
        subprocess.run([tool, str(tmp_path / name), src], check=True, stdout=subprocess.DEVNULL)
        vertices, data, want = _inputs(str(tmp_path / (name + ".raw.cache")))
        got = _scrub(want)
        lib.cookhost_meshlet_bounds(vertices.ctypes.data, len(vertices) // 16, data.ctypes.data, len(data), got.ctypes.data, len(got))
        _check(got, want, name)
        total += len(want)
    assert total > 900


def test_committed_kitten_inputs_match():
    """kitten_cook.npz (positions + raw meshletdata of the cooked kitten, for the GPU box where the reference is absent)
    reproduces kitten.nvcg's Meshlet[] through the host instantiation."""
    lib = _host_lib()
    z = np.load(os.path.join(GOLDEN, "kitten_cook.npz"))
    _, want, _ = layout.load_nvcg(os.path.join(GOLDEN, "kitten.nvcg"))
    positions, data = z["positions"], z["meshletdata"]  # keep the arrays alive while the C code reads them
    vertices = np.zeros((len(positions), 8), dtype=np.uint16)
    vertices[:, :3] = positions
    got = _scrub(want)
    lib.cookhost_meshlet_bounds(vertices.ctypes.data, len(vertices), data.ctypes.data, len(data), got.ctypes.data, len(got))
    _check(got, want, "kitten")


@pytest.mark.gpu
def test_gpu_kernel_reproduces_reference_cooker():
    import torch

    from niagara_b200.lib import check, load_library

    assert torch.cuda.is_available()
    lib = load_library()
    ctx = ctypes.c_void_p()
    check(lib.nvc_create(0, None, ctypes.byref(ctx)), None, "nvc_create")
    try:
        cases = []
        vertices, data, want = _inputs(os.path.join(GOLDEN, "animated.raw.cache"))
        cases.append(("animated", vertices, data, want))
        z = np.load(os.path.join(GOLDEN, "kitten_cook.npz"))
        _, kw, _ = layout.load_nvcg(os.path.join(GOLDEN, "kitten.nvcg"))
        kv = np.zeros((len(z["positions"]), 8), dtype=np.uint16)
        kv[:, :3] = z["positions"]
        cases.append(("kitten", kv.view(np.uint8).reshape(-1), z["meshletdata"], kw))
        for name, vertices, data, want in cases:
            dv = torch.from_numpy(np.ascontiguousarray(vertices).view(np.uint8).reshape(-1).copy()).cuda()
            dd = torch.from_numpy(np.ascontiguousarray(data).view(np.int32).copy()).cuda()
            dm = torch.from_numpy(_scrub(want).view(np.uint8).reshape(-1).copy()).cuda()
            rej = torch.full((1,), 77, dtype=torch.int32, device="cuda")
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            check(lib.nvc_cook_meshlet_bounds(ctx, stream, dv.data_ptr(), dv.numel() // 16, dd.data_ptr(), dd.numel(), dm.data_ptr(), len(want), rej.data_ptr()), ctx, "nvc_cook_meshlet_bounds")
            torch.cuda.synchronize()
            assert int(rej.item()) == 0
            _check(dm.cpu().numpy().view(layout.MESHLET_DTYPE), want, name)
            # a meshlet pointing past the arrays is rejected and left alone
            broken = _scrub(want)
            broken["dataOffset"][0] = len(data)
            broken["baseVertex"][1] = dv.numel() // 16
            dm = torch.from_numpy(broken.view(np.uint8).reshape(-1).copy()).cuda()
            check(lib.nvc_cook_meshlet_bounds(ctx, stream, dv.data_ptr(), dv.numel() // 16, dd.data_ptr(), dd.numel(), dm.data_ptr(), len(want), rej.data_ptr()), ctx, "nvc_cook_meshlet_bounds")
            torch.cuda.synchronize()
            assert int(rej.item()) == 2
            out = dm.cpu().numpy().view(layout.MESHLET_DTYPE)
            assert np.array_equal(out[:2], broken[:2])
            _check(out[2:], want[2:], name + " rest")
    finally:
        lib.nvc_destroy(ctx)
