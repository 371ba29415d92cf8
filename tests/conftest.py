import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_artifacts():
    """The oracle (CPU checker) is rebuilt from source when stale; the CUDA library is rebuilt only where nvcc exists
    (the GPU box uses the prebuilt .so that travelled with the snapshot)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    from niagara_b200 import _build

    if _build.needs_build() and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        _build.build()
    yield


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) on a machine without a CUDA device."""
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
