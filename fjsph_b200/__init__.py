"""fjsph_b200 — B200-native WCSPH time-step engine behind FJSPH's step boundary.

Only the hot path lives here: csrc/ (hand-written sm_100a kernels + the C ABI of include/fjsph_b200.h),
engine.py (host mirror of the reference's function interface) and cases.py (input decks of the configs).
"""
from . import cases  # noqa: F401

__all__ = ["cases", "engine"]
