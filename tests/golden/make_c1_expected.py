#!/usr/bin/env python
"""Regenerates tests/golden/c1_kitten_expected.npz: the oracle's output for BASELINE configs[0]
("kitten.obj, 4096 instanced MeshDraws, frustum-only cull on host CPU") plus the two-phase occlusion variant.

Inputs: tests/golden/kitten.nvcg (cooked by the reference's own scene.cpp, see oracle/refscene/) and the reference's
PCG32 random scene (niagara.cpp:969-998).  The reference holds no expected outputs for this path, so these
vectors pin OUR oracle against accidental drift; they are re-derived by tests/numpy_ref.py and reproduced byte for byte by
the reference's own shaders compiled for the host (oracle/refshader; tests/test_golden_outputs.py).

    python tests/golden/make_c1_expected.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def compute(path_class=None):
    """path_class: oracle_lib.OraclePath (default) or refshader_lib.RefShaderPath — with one host thread the reference's
    shaders run their workgroups in order, so even the order inside dcb equals the oracle's ascending order (cib is compared sorted:
    the (X, 64, 1) grid is walked x-fastest, i.e. not in ascending commandId)."""
    import oracle_lib
    from niagara_b200 import scenes

    path_class = path_class or oracle_lib.OraclePath

    s = scenes.instanced_scene(os.path.join(ROOT, "tests", "golden", "kitten.nvcg"), 4096)
    out = {}

    # (a) configs[0]: frustum + LOD only, draw-command path, every draw "visible last frame" (steady state)
    cd = s.cull_data(occlusion=False, cluster_occlusion=False, mesh_shading=False)
    o = path_class(s.meshes, s.meshlets, s.draws, *s.screen, mesh_shading=False)
    o.dvb[:] = 1
    o.cull(cd, late=False, task=False)
    n = int(o.dccb[0])
    out["a_count"] = np.uint32(n)
    out["a_commands"] = o.read_draw_commands(n).view(np.uint32).reshape(n, 6)
    out["a_lod"] = o.lod_out[: len(s.draws)].copy()

    # (b) two frames of the full two-phase path with cone culling
    cd = s.cull_data()
    o = path_class(s.meshes, s.meshlets, s.draws, *s.screen)
    o.set_visibility_bits(s.visibility_bits)
    for f in range(2):
        o.frame(cd, s.depth, cluster_backface=True)
        out["b%d_dccb" % f] = o.dccb.copy()
        out["b%d_ccb" % f] = o.ccb.copy()
        out["b%d_dvb" % f] = np.packbits(o.dvb[: len(s.draws)].astype(np.uint8))
        out["b%d_mvb" % f] = o.mvb.copy()
        out["b%d_commands" % f] = o.read_task_commands(int(o.dccb[1]) * 64).view(np.uint32).reshape(-1, 5)
        out["b%d_clusters_crc" % f] = np.uint64(int(np.bitwise_xor.reduce(np.sort(o.read_cluster_indices(int(o.ccb[0]))).astype(np.uint64) * np.arange(1, int(o.ccb[0]) + 1, dtype=np.uint64)))) if int(o.ccb[0]) else np.uint64(0)  # order-free
    out["b_pyramid_top"] = o.pyramid_texels[-341:].copy()  # the 5 coarsest mips
    out["b_pyramid_crc"] = np.uint64(int(np.bitwise_xor.reduce(o.pyramid_texels.view(np.uint32).astype(np.uint64) * np.arange(1, len(o.pyramid_texels) + 1, dtype=np.uint64))))
    return out


if __name__ == "__main__":
    data = compute()
    path = os.path.join(ROOT, "tests", "golden", "c1_kitten_expected.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), "bytes;", "frustum-only visible draws:", int(data["a_count"]))
