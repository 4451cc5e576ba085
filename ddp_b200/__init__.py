"""ddp_b200 — B200-native (sm_100a) reverse-diffusion decode head for DDP.

The hot path lives in libddp_b200.so (ddp_b200/csrc, C ABI in include/ddp_b200.h); this package is
the thin Python host: ctypes binding, the engine that holds PyTorch tensors, and the MMSegmentation
plug-in surface (DDP / SelfAlignedDDP / DeformableHeadWithTime) of the reference.
"""
__version__ = "0.1.0"

from .engine import DecodeEngine  # noqa: F401
from .neck import NeckEngine  # noqa: F401
