"""GPU: the pybind11/libtorch binding of the C ABI shown in INTEGRATION.md (ss-gnn_b200/binding/) gives exactly
the outputs of the ctypes drop-in -- i.e. the C ABI is bindable the way the reference binds its own sampler."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from helpers import random_batch

pytestmark = pytest.mark.gpu


def _binding():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("ugs_binding_build", os.path.join(root, "ss-gnn_b200", "binding", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.load()


@pytest.mark.parametrize("mode", ["sample", "graph", "global"])
def test_cpp_binding_equals_ctypes_dropin(cuda_lib, mode):
    import ugs_sampler
    b = _binding()
    rng = np.random.default_rng(3)
    ei, ptr = random_batch(rng, 40, 1, 45, deg=2.5, big=[(5, 120)])
    tei, tptr = torch.from_numpy(ei), torch.from_numpy(ptr)
    a = b.sample_batch(tei, tptr, 9, 4, mode, 7)
    c = ugs_sampler.sample_batch(tei, tptr, 9, 4, mode=mode, seed=7)
    assert len(a) == 5
    for x, y in zip(a, c):
        assert x.dtype == torch.int64 and not x.is_cuda and x.is_pinned()
        assert torch.equal(x, y)
    # unsorted columns go through the slow path inside the binding as well
    perm = rng.permutation(ei.shape[1])
    a = b.sample_batch(torch.from_numpy(np.ascontiguousarray(ei[:, perm])), tptr, 5, 3, mode, 1)
    c = ugs_sampler.sample_batch(torch.from_numpy(np.ascontiguousarray(ei[:, perm])), tptr, 5, 3, mode=mode, seed=1)
    for x, y in zip(a, c):
        assert torch.equal(x, y)
    with pytest.raises(RuntimeError, match="mode must be one of"):
        b.sample_batch(tei, tptr, 2, 3, "bogus", 1)
