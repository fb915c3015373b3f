"""phaserotate.lv2_b200 — B200 (sm_100a) backend for the phase-rotation hot path
of x42/phaserotate.lv2.

The product is `libphaserot_cuda.so` (csrc/, C ABI in include/phaserot_cuda.h)
plus the host programs above it (host/: the `phase-rotate` CLI and the LV2
plugin).  This Python package only builds the library (`build`) and binds it
with ctypes (`capi`) for tests and bench.py.
"""
from . import build, capi  # noqa: F401
from .capi import Phaserot, PhaserotError  # noqa: F401
