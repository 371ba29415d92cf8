"""CPU-only: the oracle's scalar building blocks against the independent numpy restatement on random inputs, including
zeros, denormals, huge values, infinities and NaNs (the GLSL's behaviour on those is what IEEE arithmetic gives)."""
import numpy as np

import numpy_ref as nr
import oracle_lib

F = np.float32
SPECIAL = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-38, 1.0, -1.0, 0.5, 2.0, 3.4e38, -3.4e38, np.inf, -np.inf, np.nan], dtype=np.float32)


def _mixed(rng, n, scale):
    v = (rng.standard_normal(n) * scale).astype(np.float32)
    idx = rng.integers(0, n, n // 25)
    v[idx] = SPECIAL[rng.integers(0, len(SPECIAL), len(idx))]
    return v


def _same(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    both_nan = np.isnan(a) & np.isnan(b)
    return np.array_equal(a.view(np.uint32)[~both_nan], b.view(np.uint32)[~both_nan]) and np.array_equal(np.isnan(a), np.isnan(b))


def test_project_sphere_random():
    lib = oracle_lib.load()
    rng = np.random.default_rng(1)
    n = 20000
    cx, cy, cz = _mixed(rng, n, 30), _mixed(rng, n, 30), np.abs(_mixed(rng, n, 60))
    r = np.abs(_mixed(rng, n, 3))
    ok_np, aabb_np = nr.project_sphere((cx, cy, cz), r, F(0.1), F(1.07), F(1.43))
    out = np.zeros(4, np.float32)
    for i in range(n):
        ok = lib.orc_project_sphere(np.array([cx[i], cy[i], cz[i]], np.float32).ctypes.data, float(r[i]), 0.1, float(F(1.07)), float(F(1.43)), out.ctypes.data)
        assert bool(ok) == bool(ok_np[i]), i
        if ok:
            assert _same(out, [aabb_np[k][i] for k in range(4)]), (i, out, [aabb_np[k][i] for k in range(4)])


def test_occlusion_mip_random():
    lib = oracle_lib.load()
    rng = np.random.default_rng(2)
    n = 20000
    x0, y0 = _mixed(rng, n, 0.6), _mixed(rng, n, 0.6)
    w, h = np.abs(_mixed(rng, n, 0.05)), np.abs(_mixed(rng, n, 0.05))
    # a tenth of the boxes exactly a power-of-two number of texels wide: the ceil(log2) boundary
    k = rng.integers(0, n, n // 10)
    w[k] = (2.0 ** rng.integers(-3, 11, len(k)) / 2048.0).astype(np.float32)
    aabb = (x0, y0, (x0 + w).astype(np.float32), (y0 + h).astype(np.float32))
    want = nr.occlusion_mip(aabb, F(2048.0), F(1024.0))
    for i in range(n):
        got = lib.orc_occlusion_mip(np.array([aabb[0][i], aabb[1][i], aabb[2][i], aabb[3][i]], np.float32).ctypes.data, 2048.0, 1024.0)
        assert got == want[i] or (got >= 1e8 and want[i] >= 1e8), (i, got, want[i], [a[i] for a in aabb])


def test_sample_min_random():
    lib = oracle_lib.load()
    rng = np.random.default_rng(3)
    for (w, h) in [(1, 1), (2, 1), (5, 3), (64, 64), (37, 128)]:
        img = rng.random((h, w), dtype=np.float32)
        n = 4000
        u, v = _mixed(rng, n, 0.7) + F(0.5), _mixed(rng, n, 0.7) + F(0.5)
        # some coordinates exactly on texel centres / edges (zero-weight neighbours must be ignored)
        k = rng.integers(0, n, n // 5)
        u[k] = ((rng.integers(0, 2 * w + 1, len(k)) * 0.5) / w).astype(np.float32)
        v[k] = ((rng.integers(0, 2 * h + 1, len(k)) * 0.5) / h).astype(np.float32)
        finite = np.isfinite(u) & np.isfinite(v)
        want = nr.sample_min_level(img, u[finite], v[finite])
        got = np.array([lib.orc_sample_min(img.ctypes.data, w, h, float(a), float(b)) for a, b in zip(u[finite], v[finite])], np.float32)
        assert np.array_equal(got, want), (w, h)
        # non-finite coordinates: defined (clamped), never out of bounds
        for a, b in zip(u[~finite], v[~finite]):
            r = lib.orc_sample_min(img.ctypes.data, w, h, float(a), float(b))
            assert img.min() <= r <= img.max()
