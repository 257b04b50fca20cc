"""nbody6ppgpu_b200 -- B200-native regular-force library for NBODY6++GPU (hot path only).

The product is ``libgpunb_b200.so`` (CUDA, sm_100a) behind the reference's Fortran-callable
C-ABI (``include/gpunb_b200.h``).  This package is the thin host-side mirror of that interface
(``gpunb.ForceLib``: open / send / regf / profile / close / gpupot with the reference's argument
meaning) plus the synthetic-snapshot generators used by the tests and the benchmark.

There is no CPU fallback: loading fails loudly when the CUDA library has not been built.
"""
from .gpunb import ForceLib, load, lib_path, LibraryMissing  # noqa: F401
from . import snapshots  # noqa: F401

__all__ = ["ForceLib", "load", "lib_path", "LibraryMissing", "snapshots"]
