"""niagara_b200 — B200-native (sm_100a CUDA) implementation of zeux/niagara's GPU-driven visibility path.

Only what the hot path needs: csrc/ (kernels + C ABI, built into libniagara_cull.so), ctypes bindings (lib),
the reference's struct layouts (layout), host-side helpers (host), the call-site mirror (path.VisibilityPath)
and synthetic scene generators for the BASELINE.json configurations (scenes)."""
from . import layout  # noqa: F401
from .lib import NvcError, load_library  # noqa: F401

__all__ = ["layout", "load_library", "NvcError"]
