#!/usr/bin/env python
"""CPU study for the round-2 "filtered predicates" idea: how far is the strict-float32 screen-space AABB of
projectSphere from the exact (float64) value, and how often does a meshlet sit within eps of a discrete decision
boundary of the Hi-Z lookup (mip level selection, texel footprint)?  Run: python tools/filtered_predicate_study.py"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy_ref as nr  # noqa: E402
import oracle_lib  # noqa: E402
from niagara_b200 import layout, scenes  # noqa: E402
from niagara_b200.lib import load_library  # noqa: E402


def main(draws=30000):
    s = scenes.config4_scene(draws, screen=(4096, 4096))
    cd = s.cull_data()
    o = oracle_lib.OraclePath(s.meshes, s.meshlets, s.draws, *s.screen, threads=8, cmd_capacity=2 * draws)
    o.set_visibility_bits(s.visibility_bits)
    o.pyramid(s.depth)
    o.cull(cd, True)
    cmds = o.read_task_commands(int(o.dccb[1]) * 64)
    pd = layout.CullData()
    load_library().nvc_host_pass_data(ctypes.byref(cd), 0, 0, ctypes.byref(pd))
    c = pd.to_numpy()

    def centers(F):
        tc = np.minimum(cmds["taskCount"], 64).astype(np.int64)
        cid = np.repeat(np.arange(len(cmds)), tc)
        mgi = np.arange(int(tc.sum())) - np.repeat(np.cumsum(tc) - tc, tc)
        cc = cmds[cid]
        mi = cc["taskOffset"].astype(np.int64) + mgi
        d, ml = s.draws[cc["drawId"]], s.meshlets[mi]
        q = tuple(d["orientation"][:, i].astype(F) for i in range(4))
        lc = tuple(nr.half_to_float(ml["center"][:, i]).astype(F) for i in range(3))
        rc = nr.rotate_quat(lc, q) if F is np.float32 else tuple(np.asarray(v) for v in nr.rotate_quat(lc, q))
        cen = tuple(rc[i] * d["scale"].astype(F) + d["position"][:, i].astype(F) for i in range(3))
        m = [F(x) for x in c["view"]]
        cen = tuple(((m[i] * cen[0] + m[4 + i] * cen[1]) + m[8 + i] * cen[2]) + m[12 + i] for i in range(3))
        return cen, nr.half_to_float(ml["radius"]).astype(F) * d["scale"].astype(F)

    def aabb(cn, r, F):
        P00, P11 = F(c["P00"]), F(c["P11"])
        crx, cry, crz = cn[0] * r, cn[1] * r, cn[2] * r
        czr2 = cn[2] * cn[2] - r * r
        vx = np.sqrt(cn[0] * cn[0] + czr2)
        vy = np.sqrt(cn[1] * cn[1] + czr2)
        minx, maxx = (vx * cn[0] - crz) / (vx * cn[2] + crx), (vx * cn[0] + crz) / (vx * cn[2] - crx)
        miny, maxy = (vy * cn[1] - crz) / (vy * cn[2] + cry), (vy * cn[1] + crz) / (vy * cn[2] - cry)
        return (minx * P00 * F(0.5) + F(0.5), maxy * P11 * F(-0.5) + F(0.5), maxx * P00 * F(0.5) + F(0.5), miny * P11 * F(-0.5) + F(0.5))

    with np.errstate(all="ignore"):
        c32, r32 = centers(np.float32)
        c64, r64 = centers(np.float64)
        ok = c32[2] >= r32 + np.float32(c["znear"])
        a32, a64 = aabb(c32, r32, np.float32), aabb(c64, r64, np.float64)
        err = np.max([np.abs(a32[i].astype(np.float64) - a64[i])[ok] for i in range(4)], axis=0)
        print("items %d; |aabb_f32 - aabb_f64| (uv units): median %.2e  p99 %.2e  max %.2e" % (ok.sum(), np.median(err), np.percentile(err, 99), err.max()))
        pw = float(c["pyramidWidth"])
        m = np.maximum((a64[2] - a64[0])[ok], (a64[3] - a64[1])[ok]) * pw
        L = np.ceil(np.log2(np.maximum(m, 1e-30)))
        near_pow2 = np.minimum(np.abs(m / 2.0**L - 1.0), np.abs(m / 2.0 ** (L - 1) - 1.0))
        w = pw / 2.0 ** np.clip(L, 0, 11)
        x = ((a64[0] + a64[2]) * 0.5)[ok] * w - 0.5
        y = ((a64[1] + a64[3]) * 0.5)[ok] * w - 0.5
        for eps in (1e-6, 4e-6, 1.6e-5):
            amb_level = (near_pow2 < eps * pw / np.maximum(m, 1)).mean()
            amb_texel = ((np.abs(x - np.round(x)) < eps * w) | (np.abs(y - np.round(y)) < eps * w)).mean()
            print("eps %.1e (uv): ambiguous mip level %.3f%%, ambiguous texel footprint %.3f%%, P(warp of 32 needs the exact path) %.1f%%" % (eps, 100 * amb_level, 100 * amb_texel, 100 * (1 - (1 - amb_level - amb_texel) ** 32)))


if __name__ == "__main__":
    main()
