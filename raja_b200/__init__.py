"""raja_b200 -- host-side mirror of the RAJA front-end for the b200 back-end.

The product is the CUDA library (raja_b200/csrc -> lib/libraja_b200.so, C ABI in
include/raja_b200.h) and the C++ policy adapter (include/RAJA_b200/).  This package
is the thin Python door onto the C ABI used by the tests and bench.py; PyTorch is
used only for device memory, streams and torch.distributed.
"""
from ._lib import B200Error, LIB_PATH  # noqa: F401
from .api import *  # noqa: F401,F403
from . import api as _api

__all__ = list(_api.__all__) + ["B200Error", "LIB_PATH"]
