"""B200-native drop-in for the sampling path of DCM-UPB/MCIntegratorPlusPlus.

The product is libmcig.so (hand-written sm_100a CUDA behind the C-ABI of include/mcig.h, with the C++ facade in
include/mci/). This package holds its sources (csrc/), the build script and a thin ctypes mirror of the reference's
`mci::MCI` interface used by tests/ and bench.py. There is no CPU fallback: without the built library or without a
B200 every compute call raises.
"""
from . import _capi  # noqa: F401
from .mci import *  # noqa: F401,F403
