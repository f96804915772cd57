"""surtr_b200 -- B200-native fracture engine for Surtr's convex-piece cutting path.

Product code only: the CUDA library (csrc/ -> libsurtr_b200.so, C ABI in include/surtr_b200.h), its ctypes
binding (engine.py) and the host-side mirror of the reference interfaces (host/).  Nothing here imports oracle/.
"""
from .engine import FractureContext, Fragments, SurtrError, FRAGMENT_DTYPE, LIB_PATH, load_library  # noqa: F401

__all__ = ["FractureContext", "Fragments", "SurtrError", "FRAGMENT_DTYPE", "LIB_PATH", "load_library"]
