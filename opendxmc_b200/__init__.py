"""opendxmc_b200 — B200-native replacement for the DXMClib `dxmc::Transport` photon-history path of OpenDXMC.

The product is `lib/libdxmc_b200.so` (hand-written CUDA for sm_100a behind the C ABI in include/dxb.h).
This package is the thin Python view of that ABI used by tests/ and bench.py; the C++ drop-in shim is
include/dxmc/.  There is no CPU fallback: importing `api` works without a GPU (materials, tubes, beams are
host code), creating a World requires a CUDA device.
"""
from . import _capi  # noqa: F401
from .api import *  # noqa: F401,F403
from . import workloads  # noqa: F401
