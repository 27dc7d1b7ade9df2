"""bmfr_b200 — B200-native BMFR denoiser hot path (sm_100a CUDA behind a C ABI).

The product is ``libbmfr_b200.so`` (include/bmfr_b200.h).  This package is the thin host-side mirror
of the reference's frame loop on top of it; importing it never falls back to a CPU or PyTorch path.
"""
from ._lib import BmfrError, load  # noqa: F401
from .denoiser import Denoiser, block_offset  # noqa: F401

__all__ = ["Denoiser", "BmfrError", "block_offset", "load"]
