"""CPU-only: pins the oracle's scalar pieces with hand-computed / analytic cases, exhaustive conversions and the
independent numpy restatement (tests/numpy_ref.py).  The reference holds no golden vectors for this path
(SURVEY F3); the strongest anchor is tests/test_refshader.py (the reference's own shaders run on the host), these are
the independent ones; see DESIGN.md §2."""
import ctypes
import math
import os
from fractions import Fraction

import numpy as np
import pytest

import numpy_ref as nr
import oracle_lib
from niagara_b200 import host, layout, scenes

F = np.float32


def arr(*v):
    return np.array(v, dtype=np.float32)


def test_half_to_float_exhaustive():
    lib = oracle_lib.load()
    bits = np.arange(65536, dtype=np.uint16)
    want = bits.view(np.float16).astype(np.float32)
    got = np.array([lib.orc_half_to_float(int(b)) for b in bits], dtype=np.float32)
    nan = np.isnan(want)
    assert np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32)) and np.isnan(got[nan]).all()


def _round_f32(fr):
    """exact round-to-nearest-even of a Fraction to binary32"""
    if fr == 0:
        return F(0.0)
    f = F(float(fr))
    cands = [f, np.nextafter(f, F(np.inf)), np.nextafter(f, F(-np.inf))]
    best = min(cands, key=lambda c: (abs(Fraction(float(c)) - fr), int(np.float32(c).view(np.uint32)) & 1))
    return F(best)


def test_div127_exact():
    """csrc/nvc_math.cuh s8_div127: q0 = a*r; rem = fma(-q0,127,a); q = fma(rem,r,q0) == correctly rounded a/127 for all
    s8 inputs (emulated with exact rational arithmetic)."""
    r = F(0.00787401574803149606)
    for i in range(-128, 128):
        a = F(i)
        q0 = F(a * r)
        rem = _round_f32(Fraction(float(-q0)) * 127 + Fraction(float(a)))
        q = _round_f32(Fraction(float(rem)) * Fraction(float(r)) + Fraction(float(q0)))
        assert q == F(a / F(127.0)), i
        assert _round_f32(Fraction(i, 127)) == F(a / F(127.0))


def test_ceil_log2_exact():
    lib = oracle_lib.load()
    for e in range(-140, 120):
        x = math.ldexp(1.0, e)
        assert lib.orc_ceil_log2(x) == e
        if -126 <= e < 120:
            assert lib.orc_ceil_log2(float(np.nextafter(F(x), F(np.inf)))) == e + 1
            assert lib.orc_ceil_log2(float(np.nextafter(F(x), F(0)))) == e
    rng = np.random.default_rng(0)
    xs = np.exp(rng.uniform(-20, 20, 2000)).astype(np.float32)
    want = nr.ceil_log2_exact(xs)
    got = np.array([lib.orc_ceil_log2(float(x)) for x in xs])
    assert np.array_equal(got, want)


def test_rotate_quat_identity_and_axis():
    lib = oracle_lib.load()
    out = arr(0, 0, 0)
    lib.orc_rotate_quat(arr(1, 2, 3).ctypes.data, arr(0, 0, 0, 1).ctypes.data, out.ctypes.data)
    assert np.array_equal(out, arr(1, 2, 3))
    # 90 degrees about +z maps x -> y (math.h:46-49 with q = (0,0,sin45,cos45))
    s = math.sqrt(0.5)
    lib.orc_rotate_quat(arr(1, 0, 0).ctypes.data, arr(0, 0, s, s).ctypes.data, out.ctypes.data)
    assert np.allclose(out, [0, 1, 0], atol=1e-6)


def test_project_sphere_on_axis_analytic():
    """c = (0,0,10), r = 1: vx = sqrt(99); minx = -10 / (10 vx) = -1/sqrt(99); symmetric in x and y."""
    lib = oracle_lib.load()
    aabb = arr(0, 0, 0, 0)
    assert lib.orc_project_sphere(arr(0, 0, 10).ctypes.data, 1.0, 0.1, 1.5, 2.0, aabb.ctypes.data) == 1
    t = 1.0 / math.sqrt(99.0)
    want = [-t * 1.5 * 0.5 + 0.5, -(t * 2.0) * 0.5 + 0.5, t * 1.5 * 0.5 + 0.5, (t * 2.0) * 0.5 + 0.5]
    assert np.allclose(aabb, want, atol=2e-7)
    # crossing the near plane: c.z < r + znear -> false (math.h:4-5)
    assert lib.orc_project_sphere(arr(0, 0, 1.05).ctypes.data, 1.0, 0.1, 1.5, 2.0, aabb.ctypes.data) == 0
    assert lib.orc_project_sphere(arr(0, 0, 1.1).ctypes.data, 1.0, 0.1, 1.5, 2.0, aabb.ctypes.data) == int(F(1.1) >= F(1.0) + F(0.1))


def test_occlusion_mip_cases():
    lib = oracle_lib.load()

    def mip(x0, y0, x1, y1, pw=512.0, ph=512.0):
        return lib.orc_occlusion_mip(arr(x0, y0, x1, y1).ctypes.data, pw, ph)

    # 8 texels wide exactly: ceil(log2(8)) = 3; at mip 2 (fmip = 128) the footprint is 2 texels starting on a texel
    # boundary (fract 0 + 2 <= 2) so the finer mip is taken: 2
    assert mip(0.25, 0.25, 0.25 + 8 / 512, 0.25 + 8 / 512) == 2.0
    # same size, offset by half a mip-2 texel: fract 0.5 + 2 > 2 -> stays at 3
    assert mip(0.25 + 2 / 512, 0.25, 0.25 + 10 / 512, 0.25 + 8 / 512) == 3.0
    # just above 8 texels: ceil(log2) = 4; finer mip 3: 8.0x/8 = 1.0x texels + fract 0 <= 2 -> 3
    assert mip(0.25, 0.25, 0.25 + 8.5 / 512, 0.25 + 1 / 512) == 3.0
    # sub-texel: level <= 0 -> 0
    assert mip(0.5, 0.5, 0.5 + 0.3 / 512, 0.5 + 0.2 / 512) == 0.0
    # degenerate / inverted box: log2 of non-positive -> max(level, 0) = 0
    assert mip(0.5, 0.5, 0.5, 0.5) == 0.0 and mip(0.6, 0.6, 0.5, 0.5) == 0.0
    # whole screen: 512 texels -> 9, finer mip 8 has 2 texels: fract(0) + 2 <= 2 -> 8
    assert mip(0.0, 0.0, 1.0, 1.0) == 8.0


def test_sampler_min_reduction_rules():
    lib = oracle_lib.load()
    img = np.arange(16, dtype=np.float32).reshape(4, 4) + 1
    img[1, 1] = 0.5

    def s(u, v):
        return lib.orc_sample_min(img.ctypes.data, 4, 4, u, v)

    # texel centre (x = 1.0 exactly after -0.5): weights (1,0) -> ONLY texel (1,1), the zero-weight neighbours excluded
    assert s(1.5 / 4, 1.5 / 4) == 0.5
    assert s(2.5 / 4, 2.5 / 4) == img[2, 2]
    # between four texels: min of the 2x2
    assert s(2.0 / 4, 2.0 / 4) == min(img[1, 1], img[1, 2], img[2, 1], img[2, 2])
    assert s(3.0 / 4, 3.0 / 4) == min(img[2, 2], img[2, 3], img[3, 2], img[3, 3])
    # clamp to edge
    assert s(0.0, 0.0) == img[0, 0] and s(1.0, 1.0) == img[3, 3] and s(-3.0, 0.5 / 4) == img[0, 0] and s(7.0, 3.5 / 4) == img[3, 3]


@pytest.mark.parametrize("size", [(64, 64), (100, 60), (30, 17), (129, 257), (2, 2), (1, 1), (256, 8)])
def test_pyramid_matches_numpy_and_blockmin(size):
    w, h = size
    rng = np.random.default_rng(w * 1000 + h)
    depth = rng.random((h, w), dtype=np.float32)
    o = oracle_lib.OraclePath(np.zeros(1, layout.MESH_DTYPE), np.zeros(1, layout.MESHLET_DTYPE), np.zeros(1, layout.MESHDRAW_DTYPE), w, h)
    o.pyramid(depth)
    levels = nr.build_pyramid(depth, o.hiz.width, o.hiz.height, o.hiz.levels)
    for l in range(o.hiz.levels):
        assert np.array_equal(o.level(l), levels[l]), l
    # above level 0 every texel is the exact min of its 2x2 (or 2x1 / 1x2) parent block: conservative pyramid
    for l in range(1, o.hiz.levels):
        p = o.level(l - 1)
        ph, pw = p.shape
        c = o.level(l)
        for y in range(c.shape[0]):
            for x in range(c.shape[1]):
                blk = p[min(2 * y, ph - 1) : min(2 * y + 1, ph - 1) + 1, min(2 * x, pw - 1) : min(2 * x + 1, pw - 1) + 1]
                assert c[y, x] == blk.min()
    # power-of-two exact halving at level 0
    if w == 2 * o.hiz.width and h == 2 * o.hiz.height:
        want = depth.reshape(h // 2, 2, w // 2, 2).min(axis=(1, 3))
        assert np.array_equal(o.level(0), want)
    # the top of the pyramid is the global minimum only when every source texel is covered
    assert o.level(o.hiz.levels - 1).shape == (1, 1)


def _scene(golden_dir, n=3000, screen=(640, 480)):
    s = scenes.instanced_scene(os.path.join(golden_dir, "kitten_pirate.nvcg"), n, screen=screen)
    return s


@pytest.mark.parametrize("toggles", [dict(), dict(lod=False), dict(culling=False), dict(occlusion=False), dict(cluster_occlusion=False)])
def test_oracle_agrees_with_numpy_restatement(golden_dir, toggles):
    """Two frames of the full path: every per-draw and per-meshlet decision of the C++ oracle must equal the
    independent numpy restatement bit for bit (visible sets, lod, dvb, mvb, counters)."""
    s = _scene(golden_dir)
    cd = s.cull_data(**toggles)
    o = oracle_lib.OraclePath(s.meshes, s.meshlets, s.draws, *s.screen)
    o.set_visibility_bits(s.visibility_bits)
    o.dvb[: len(s.draws) // 2] = 1  # mixed history so that early/late both have work in frame 0
    rng = np.random.default_rng(5)
    o.mvb[:] = rng.integers(0, 1 << 32, len(o.mvb), dtype=np.uint64).astype(np.uint32)
    levels = nr.build_pyramid(s.depth, o.hiz.width, o.hiz.height, o.hiz.levels)
    from niagara_b200.lib import load_library

    def pass_data(for_draw, backface=None):
        pd = layout.CullData()
        load_library().nvc_host_pass_data(ctypes.byref(cd), for_draw, 0, ctypes.byref(pd))
        if backface is not None:
            pd.clusterBackfaceEnabled = backface
        return pd.to_numpy()

    for frame in range(2):
        for late in (False, True):
            if late:
                o.pyramid(s.depth)
                for l in range(o.hiz.levels):
                    assert np.array_equal(o.level(l), levels[l])
            dvb_before = o.dvb.copy()
            reached, visible, emit, lod = nr.drawcull_decisions(pass_data(1), late, s.draws, s.meshes, dvb_before, levels)
            o.cull(cd, late)
            assert np.array_equal(o.lod_out[: len(s.draws)] != 0xFF, emit)
            assert np.array_equal(o.lod_out[: len(s.draws)][emit], lod[emit].astype(np.uint8))
            if late:
                want_dvb = np.where(reached, visible.astype(np.uint32), dvb_before[: len(s.draws)])
                assert np.array_equal(o.dvb[: len(s.draws)], want_dvb)
            else:
                assert np.array_equal(o.dvb, dvb_before)
            dccb, _ = o.read_counts()
            cmds = o.read_task_commands(dccb[1] * 64)
            groups = (s.meshes["lods"]["meshletCount"][s.draws["meshIndex"], lod] + 63) // 64
            assert dccb[0] == groups[emit].sum()
            assert np.array_equal(np.unique(cmds["drawId"][: dccb[0]]), np.nonzero(emit & (groups > 0))[0])

            mvb_before = o.mvb.copy()
            cid, mgi, mi, mvi, vis, out = nr.cluster_decisions(pass_data(0, 1), late, cmds, s.draws, s.meshlets, mvb_before, levels)
            o.render_clusters(cd, late, cluster_backface=True)
            _, ccb = o.read_counts()
            assert ccb[0] == out.sum()
            got = o.read_cluster_indices(ccb[0])
            want = (cid[out] | (mgi[out] << 24)).astype(np.uint32)
            assert np.array_equal(got, want)  # the oracle's order is ascending (command, lane)
            if late and cd.clusterOcclusionEnabled == 1:
                want_mvb = mvb_before.copy()
                np.bitwise_or.at(want_mvb, mvi[vis] >> 5, (np.uint32(1) << (mvi[vis] & 31).astype(np.uint32)))
                np.bitwise_and.at(want_mvb, mvi[~vis] >> 5, ~(np.uint32(1) << (mvi[~vis] & 31).astype(np.uint32)))
                assert np.array_equal(o.mvb, want_mvb)
            else:
                assert np.array_equal(o.mvb, mvb_before)


def test_multithreaded_oracle_equals_serial(golden_dir):
    s = _scene(golden_dir, n=5000)
    cd = s.cull_data()
    res = []
    for threads in (1, 5):
        o = oracle_lib.OraclePath(s.meshes, s.meshlets, s.draws, *s.screen, threads=threads)
        o.set_visibility_bits(s.visibility_bits)
        for f in range(2):
            o.frame(cd, s.depth, cluster_backface=True)
        dccb, ccb = o.read_counts()
        res.append((dccb.copy(), ccb.copy(), o.read_task_commands(dccb[1] * 64), o.read_cluster_indices((ccb[0] + 255) // 256 * 256), o.dvb.copy(), o.mvb.copy(), o.pyramid_texels.copy()))
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)


def test_decode_clusters_matches_consumer_rule(golden_dir):
    """oracle decode (meshlet.mesh.glsl:89-105 walk over the (16, Y, 16) dispatch) == direct numpy decode"""
    s = _scene(golden_dir, n=4000)
    cd = s.cull_data()
    o = oracle_lib.OraclePath(s.meshes, s.meshlets, s.draws, *s.screen)
    o.set_visibility_bits(s.visibility_bits)
    o.frame(cd, s.depth, cluster_backface=True)  # frame 0: the late pass emits everything visible
    _, ccb = o.read_counts()
    n = int(ccb[0])
    assert n > 100
    rec, stats = o.decode_clusters()
    assert stats[0] == n and stats[2] == 0 and stats[1] == int(ccb[2]) * 256 - n
    cmds = o.read_task_commands(int(o.dccb[1]) * 64)
    ci = o.read_cluster_indices(n)
    want_draw = cmds["drawId"][ci & 0xFFFFFF]
    want_mi = cmds["taskOffset"][ci & 0xFFFFFF] + (ci >> 24)
    assert np.array_equal(rec[:n, 0], want_draw) and np.array_equal(rec[:n, 1], want_mi)
    assert np.array_equal(rec[:n, 3], s.meshlets["triangleCount"][want_mi]) and stats[3] == s.meshlets["triangleCount"][want_mi].sum()
    assert (rec[n:] == 0xFFFFFFFF).all()


def test_two_phase_invariants(golden_dir):
    """Size-independent properties of the two-phase scheme, on a moving camera:
    * no meshlet instance is emitted by both the early and the late pass of a frame (no double draw);
    * every meshlet instance the late pass finds visible was emitted by one of them (nothing missed);
    * after the late pass dvb / mvb hold exactly the late verdicts; the early pass never changes them."""
    s = _scene(golden_dir, n=6000, screen=(800, 600))
    o = oracle_lib.OraclePath(s.meshes, s.meshlets, s.draws, *s.screen)
    o.set_visibility_bits(s.visibility_bits)
    cams = [host.make_camera((0, 0, 0)), host.make_camera((12, -4, 9), host.quat_from_axis_angle((0, 1, 0), 0.35)), host.make_camera((25, 3, -14), host.quat_from_axis_angle((0.1, 1, 0), 0.9))]
    for f in range(4):
        s.camera = cams[f % len(cams)]
        cd = s.cull_data()
        dvb0, mvb0 = o.dvb.copy(), o.mvb.copy()
        o.cull(cd, late=False)
        o.render_clusters(cd, late=False, cluster_backface=True)
        assert np.array_equal(o.dvb, dvb0) and np.array_equal(o.mvb, mvb0)
        early = oracle_lib.cluster_pairs(o.read_cluster_indices(int(o.ccb[0])), o.read_task_commands(int(o.dccb[1]) * 64))
        o.pyramid(s.depth)
        o.cull(cd, late=True)
        o.render_clusters(cd, late=True, cluster_backface=True)
        cmds = o.read_task_commands(int(o.dccb[1]) * 64)
        late = oracle_lib.cluster_pairs(o.read_cluster_indices(int(o.ccb[0])), cmds)
        assert len(np.intersect1d(early, late)) == 0
        # late-visible set = set bits of the meshlets the late pass processed
        live = cmds[cmds["taskCount"] > 0]
        vis = []
        for c in live:
            mvi = c["meshletVisibilityOffset"] + np.arange(c["taskCount"], dtype=np.uint64)
            bits = (o.mvb[(mvi >> 5).astype(np.int64)] >> (mvi & 31).astype(np.uint32)) & 1
            mi = c["taskOffset"] + np.nonzero(bits)[0]
            vis.append((np.uint64(c["drawId"]) << np.uint64(32)) | mi.astype(np.uint64))
        vis = np.sort(np.concatenate(vis)) if vis else np.zeros(0, np.uint64)
        both = np.union1d(early, late)
        assert np.isin(vis, both).all()
        assert len(late) > 0 or f > 0


def test_oracle_is_race_free(tmp_path, golden_dir):
    """The multi-threaded oracle (dynamic chunk scheduling, atomics on the shared visibility words) under ThreadSanitizer:
    two full frames with 8 threads, no report (tests/tsan_oracle.cpp)."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "tsan_oracle")
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-ffp-contract=off", "-o", exe, os.path.join(root, "tests", "tsan_oracle.cpp"),
                            os.path.join(root, "oracle", "oracle.cpp"), os.path.join(root, "niagara_b200", "csrc", "nvc_host.cpp"), "-lpthread"], capture_output=True, text=True)
    if build.returncode != 0 and ("tsan" in build.stderr or "sanitize" in build.stderr):
        pytest.skip("this toolchain has no ThreadSanitizer runtime")
    assert build.returncode == 0, build.stderr
    run = subprocess.run([exe, os.path.join(golden_dir, "kitten_pirate.nvcg"), "3000"], capture_output=True, text=True, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66"))
    assert run.returncode == 0 and "ThreadSanitizer" not in run.stderr, run.stderr[-3000:]
    assert "emitted" in run.stdout
