"""frankenz_b200: B200-native (sm_100a CUDA) implementation of frankenz's brute-force
photometric likelihood path, behind the reference's Python API.

    from frankenz_b200.fitting import BruteForce, NearestNeighbors
    from frankenz_b200 import pdf

The arithmetic runs in `lib/libfzb200.so` (hand-written CUDA, C ABI in
`include/frankenz_b200.h`).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import pdf  # noqa: F401
from . import fitting  # noqa: F401
from .bruteforce import BruteForce  # noqa: F401
from .knn import NearestNeighbors  # noqa: F401
