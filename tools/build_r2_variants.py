#!/usr/bin/env python
"""Builds the round-2 tuning variants A/B'd by tools/r2_gpu1.sh (niagara_b200/variant_<name>.so, git-ignored, shipped by gpurun)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from niagara_b200 import _build  # noqa: E402

VARIANTS = {
    "fb3": ["NVC_FILTER_MIN_BLOCKS=3"],  # filtered cluster kernel: resident CTAs per SM (register cap 80 / 64 / 48 / 40)
    "fb5": ["NVC_FILTER_MIN_BLOCKS=5"],
    "fb6": ["NVC_FILTER_MIN_BLOCKS=6"],
    "dpt1": ["NVC_DRAW_PER_THREAD=1"],  # drawcull: draws per thread
    "dpt4": ["NVC_DRAW_PER_THREAD=4"],
    "pdl": ["NVC_PDL=1"],  # programmatic dependent launch of every frame kernel
    "smem_items": ["NVC_SMEM_ITEMS=1"],  # exact early cluster kernel: per-batch item table (round-1 experiment)
    # second GPU call (tools/r2_gpu2.sh): the software pipeline, the L2 eviction hints and the batch prefetch, one at a time
    "nopipe": ["NVC_FILTER_PIPELINE=0"],
    "nohints": ["NVC_STREAM_HINTS=0"],
    "bpf": ["NVC_FILTER_BATCH_PREFETCH=1"],
    "bpf_fb3": ["NVC_FILTER_BATCH_PREFETCH=1", "NVC_FILTER_MIN_BLOCKS=3"],
    "dpt2": ["NVC_DRAW_PER_THREAD=2"],
}

SETS = {"r2a": ["fb3", "fb5", "fb6", "dpt1", "dpt4", "pdl", "smem_items"], "r2b": ["nopipe", "nohints", "bpf", "bpf_fb3", "fb3", "fb5", "pdl"]}

if __name__ == "__main__":
    names = SETS[sys.argv[1]] if len(sys.argv) > 1 else list(VARIANTS)
    for name in names:
        defines = VARIANTS[name]
        print(_build.build(force=True, defines=defines, out=os.path.join(_build.HERE, "variant_%s.so" % name)))
