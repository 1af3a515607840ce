"""gfpl — Python (ctypes) binding of the B200-native GF-PL-SLAM tracking hot path.

The compute lives in gf-pl-slam_b200/lib/libgfpl.so (hand-written sm_100a CUDA behind the C ABI of
include/gfpl_c.h).  There is no CPU fallback: creating a Context without a CUDA device raises.
"""
from ._structs import *  # noqa: F401,F403
from . import _structs as structs  # noqa: F401
