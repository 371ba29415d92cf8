#!/usr/bin/env python
"""Builds tuning variants A/B'd on the GPU box (niagara_b200/variant_<name>.so, git-ignored, shipped by gpurun).
usage: tools/build_r2_variants.py name [name ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from niagara_b200 import _build  # noqa: E402

VARIANTS = {
    "fb3": ["NVC_FILTER_MIN_BLOCKS=3"],  # filtered cluster kernel: resident CTAs per SM (register cap 80 / 64 / 48 / 40)
    "fb5": ["NVC_FILTER_MIN_BLOCKS=5"],
    "fb6": ["NVC_FILTER_MIN_BLOCKS=6"],
    "pdl": ["NVC_PDL=1"],  # programmatic dependent launch of every frame kernel
    "smem_items": ["NVC_SMEM_ITEMS=1"],  # exact early cluster kernel: per-batch item table (round-1 experiment)
    "f55": ["NVC_FILTER_MIN_BLOCKS=5", "NVC_FILTER_MIN_BLOCKS_EARLY=5", "NVC_FILTER_ITEMS=512"],
    "fe5": ["NVC_FILTER_MIN_BLOCKS_EARLY=5", "NVC_FILTER_ITEMS=512"],  # early filtered kernel: 5 / 6 CTAs per SM (smaller item table)
    "fe6": ["NVC_FILTER_MIN_BLOCKS_EARLY=6", "NVC_FILTER_ITEMS=512"],
    "pipe0": ["NVC_FILTER_PIPELINE=0"],  # meshlet prefetch distance of the filtered cluster kernels, in chunks
    "pipe1": ["NVC_FILTER_PIPELINE=1"],
    "pipe2": ["NVC_FILTER_PIPELINE=2"],
    "nohints": ["NVC_STREAM_HINTS=0"],
    "bpf": ["NVC_FILTER_BATCH_PREFETCH=1"],
}

if __name__ == "__main__":
    for name in sys.argv[1:]:
        print(_build.build(force=True, defines=VARIANTS[name], out=os.path.join(_build.HERE, "variant_%s.so" % name)))
