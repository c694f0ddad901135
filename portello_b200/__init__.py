"""portello_b200 — B200-native (sm_100a) implementation of portello's read-mapping transfer (liftover) hot path.

The product is the C-ABI shared library `portello_b200/csrc/libportello_b200.so` (hand-written CUDA kernels + C++
host), declared in `include/portello_b200.h`.  This package is the thin ctypes harness around it.
"""
from .abi import (Batch, ContigSegments, Context, LiftLib, PtlError, Result, cigar_from_string, cigar_to_string,  # noqa: F401
                  pack_seq4, STAGE_ALL, STAGE_LEFT_SHIFT, STAGE_LIFTOVER, STAGE_SIMPLIFY)

__version__ = "0.1.0"
