"""hypergen_b200 — B200-native (sm_100a) implementation of HyperGen's sketch -> dist hot path.

The product is the CUDA shared library behind ``include/hypergen_b200.h``; this package is
the thin host side: the ctypes binding (``ffi``), host glue mirroring the reference's
drivers and file formats (``sketch``, ``dist``, ``fileio``), synthetic workloads (``synth``)
and the one-process-per-GPU sharding (``multigpu``).  The directory is named
``hyper-gen_b200`` (not importable as such); import it as ``hypergen_b200`` through the
shim module at the repository root.
"""
from . import _build, ffi  # noqa: F401
from .ffi import Context, Group, HyperGenError, Peer, make_params  # noqa: F401

__all__ = ["Context", "Group", "Peer", "HyperGenError", "make_params", "ffi", "build"]


def build(force: bool = False) -> str:
    return _build.build(force=force)
