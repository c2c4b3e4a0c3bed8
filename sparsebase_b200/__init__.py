"""sparsebase_b200 -- B200-native (sm_100a) implementation of SparseBase's preprocessing hot
path: order-two format conversion, Degree/RCM reordering, Permute1D/2D and the degree
features, behind a C ABI (include/sb200.h, libsb200.so).

Python is plumbing only (device memory via torch, streams, torch.distributed); the product is
the CUDA library.  There is no CPU fallback: importing works without a GPU (so that the build
can be checked), every compute call requires one.
"""
from . import lib  # noqa: F401
from .lib import Sb200Error, load  # noqa: F401
