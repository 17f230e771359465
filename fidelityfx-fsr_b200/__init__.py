"""fidelityfx-fsr_b200 — B200-native FSR 1.0 hot path (EASU + RCAS) behind the reference's entry points.

The directory name carries a hyphen (it mirrors the reference repository's name), so import it through
`import fsr1_b200` (the alias module at the repo root) or importlib; see fsr1_b200.py.
Nothing here computes pixels on the CPU: the product is lib/libfsr1_b200.so (csrc/*.cu, C ABI in
include/fsr1_b200.h); this package is the reference-shaped host layer over it.
"""
from . import _lib  # noqa: F401
from .frames import structured, to_half, uniform  # noqa: F401
from .sharded import SlabPlan, exchange_halo, exchange_halo_many  # noqa: F401


def __getattr__(name):  # torch-dependent parts load lazily so geometry/frames work without torch
    if name in ("api", "filter"):
        import importlib
        return importlib.import_module("." + name, __name__)
    if name in ("FSR_Filter", "State"):
        from . import filter as _f
        return getattr(_f, name)
    if name == "ShardedUpscaler":
        from .sharded import ShardedUpscaler
        return ShardedUpscaler
    raise AttributeError(name)
