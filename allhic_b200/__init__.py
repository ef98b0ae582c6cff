"""allhic_b200 -- B200-native engine for the `allhic optimize` hot path.

Python mirror of the reference's interface for this path (CLM / Tour scoring /
GARun / flips) on top of the C ABI in include/ahx.h.  The compute lives in
allhic_b200/lib/libahx.so (hand-written sm_100a CUDA); nothing here falls back
to the CPU.
"""
from ._ffi import AhxError, GAParams, GAStats, LIB_PATH  # noqa: F401
from .clm import CLM  # noqa: F401
from .engine import Engine, PinnedArray  # noqa: F401
