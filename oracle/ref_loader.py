"""Test infrastructure only (oracle/): load the reference samplers built by oracle/build_ref.py.

`load(name)` imports oracle/_ref/<variant>/<name>.so WITHOUT registering it in sys.modules under
its public name, so it never shadows the B200 drop-in package of the same name.
`dump_preproc(mod, handle)` reads the reference's internal tables through oracle/ref_dump_shim.cpp.
"""
from __future__ import annotations

import ctypes
import importlib.machinery
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def have_driver() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def ref_path(name: str, variant: str | None = None) -> str:
    if variant is None:
        variant = "pinned" if have_driver() else "nopin"
    return os.path.join(HERE, "_ref", variant, name + ".so")


def available(name: str = "ugs_sampler", variant: str | None = None) -> bool:
    return os.path.exists(ref_path(name, variant))


def load(name: str = "ugs_sampler", variant: str | None = None):
    """Import a built reference module (pybind11) by file path."""
    import torch  # noqa: F401  (libtorch must be loaded first)
    path = ref_path(name, variant)
    key = path
    if key in _cache:
        return _cache[key]
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `python oracle/build_ref.py` where /root/reference exists")
    loader = importlib.machinery.ExtensionFileLoader(name, path)
    spec = importlib.util.spec_from_file_location(name, path, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    mod.__ref_path__ = path
    _cache[key] = mod
    return mod


def dump_preproc(mod, handle: int) -> dict:
    """All fields of the reference's `Preproc` for `handle` (sampler.hpp:81-93) as numpy arrays."""
    lib = ctypes.CDLL(mod.__ref_path__)
    sizes = (ctypes.c_int64 * 3)()
    lib.refdump_sizes.argtypes = [ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
    if lib.refdump_sizes(handle, sizes) != 0:
        raise KeyError(handle)
    n, nnz, an = int(sizes[0]), int(sizes[1]), int(sizes[2])
    out = dict(
        indptr=np.zeros(n + 1, np.int64), indices=np.zeros(nnz, np.int32), edge_col=np.zeros(nnz, np.int32),
        order=np.zeros(n, np.int32), index_of=np.zeros(n, np.int32), suffix_deg=np.zeros(n, np.int32),
        bucket_b=np.zeros(n, np.float64), Z=np.zeros(1, np.float64),
        alias_prob=np.zeros(an, np.float64), alias_alias=np.zeros(an, np.int32))
    order = ["indptr", "indices", "edge_col", "order", "index_of", "suffix_deg", "bucket_b", "Z",
             "alias_prob", "alias_alias"]
    lib.refdump_copy.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * len(order)
    rc = lib.refdump_copy(handle, *[out[k].ctypes.data_as(ctypes.c_void_p) for k in order])
    if rc != 0:
        raise KeyError(handle)
    out["Z"] = float(out["Z"][0])
    return out
