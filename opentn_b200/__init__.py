"""opentn_b200: B200-native (sm_100a, FP64) implementation of OpenTN's Stiefel trust-region hot path.

Mirrors the reference's module layout for that path only:
    opentn_b200.transformations / optimization / stiefel / trust_region_rcopt
plus `StiefelFit` (opentn_b200.fit), the recognisable cost object that replaces the reference's opaque closure.
All arithmetic of the hot path runs in libotn.so (hand-written CUDA, include/otn.h); there is no CPU fallback.
"""
from .fit import StiefelFit  # noqa: F401
from ._lib import OtnError  # noqa: F401

__all__ = ["StiefelFit", "OtnError"]
