"""ugs_b200 -- host side of the B200-native UGS k-graphlet sampler (ctypes over libugs_b200.so).

The drop-in modules `ugs_sampler`, `epsilon_uniform_sampler` and `apx_ugs_sampler` next to this package
mirror the reference's pybind11 modules; this package is what they are built on.
"""
from ._lib import build, lib, LibraryMissing  # noqa: F401
from .engine import DeviceBatch, sample_batch_device  # noqa: F401
