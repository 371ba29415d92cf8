#!/usr/bin/env python
"""Builds the tuning variants that tools/variants.sh A/Bs on the GPU box (niagara_b200/variant_<name>.so, git-ignored but
shipped by gpurun).  Every variant listed here is bit-exact under the CPU emulation (tests/test_kernels_emulated.py).
Usage (here, before gpurun):  python tools/build_variants.py
       (on the box)           tools/variants.sh base smem_items; NVC_LIB_PATH=$PWD/niagara_b200/variant_smem_items.so python -m pytest tests -m gpu -q -x"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from niagara_b200 import _build  # noqa: E402

VARIANTS = {
    "base": [],
    "smem_items": ["NVC_SMEM_ITEMS=1"],  # early cluster pass: per-batch item table in shared memory
    "pdl": ["NVC_PDL=1"],  # programmatic dependent launch of the five frame kernels (compiles; semantics need the GPU parity run)
    "pdl_smem_items": ["NVC_PDL=1", "NVC_SMEM_ITEMS=1"],
}

if __name__ == "__main__":
    for name, defines in VARIANTS.items():
        out = os.path.join(_build.HERE, "variant_%s.so" % name)
        print(_build.build(force=True, defines=defines, out=out))
