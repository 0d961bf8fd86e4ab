"""GPU: randomized cross-checks of every device path against its oracle -- many small random batches (weird columns,
empty graphs, hubs, both preprocessing paths), random k / m / mode / seeds. Bit-exact everywhere."""
import numpy as np
import pytest
import torch

from helpers import random_batch

pytestmark = pytest.mark.gpu
MODES = ["sample", "graph", "global"]


def _batch(ei, ptr, k):
    from ugs_b200.engine import DeviceBatch
    return DeviceBatch(torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda(), k, int(ptr[0]), int(ptr[-1] - ptr[0]))


def _eq(outs, ref, tag):
    for a, r, nm in zip(outs, ref, ["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"]):
        assert np.array_equal(a.cpu().numpy(), r), (tag, nm)


def _cases(seed, n_cases, n_hi=45):
    rng = np.random.default_rng(seed)
    for c in range(n_cases):
        G = int(rng.integers(1, 30))
        big = [(int(rng.integers(0, G)), int(rng.integers(65, 160)))] if rng.random() < 0.4 else []
        ei, ptr = random_batch(rng, G, 1, n_hi, deg=float(rng.uniform(1.0, 3.5)), weird=bool(rng.random() < 0.7), big=big)
        if rng.random() < 0.25:  # columns not grouped by graph: the slow segmentation path
            ei = np.ascontiguousarray(ei[:, rng.permutation(ei.shape[1])])
        if rng.random() < 0.3:   # a batch that does not start at node 0
            off = int(rng.integers(1, 50))
            ei, ptr = ei + off, ptr + off
        yield c, rng, ei, ptr


def test_fuzz_ugs(cuda_lib):
    from oracle import ugs_oracle as uo
    for c, rng, ei, ptr in _cases(101, 40):
        k, m, mode, seed = int(rng.integers(1, 13)), int(rng.integers(1, 70)), MODES[int(rng.integers(0, 3))], int(rng.integers(0, 2 ** 40))
        b = _batch(ei, ptr, k)
        outs, _ = b.sample(m, k, mode, seed=seed)
        _eq(outs, uo.sample_batch(ei, ptr, m, k, mode, seed), ("ugs", c, k, m, mode))
        # extraction of the same rows through the caller-provided-rows path (reads int64 nodes, not the packed rows)
        eidx, eptr, esrc = b.extract(outs[0], m, k, mode)
        assert torch.equal(eidx, outs[1]) and torch.equal(eptr, outs[2]) and torch.equal(esrc, outs[4]), ("extract", c)


def test_fuzz_epsilon(cuda_lib):
    from oracle import eps_oracle as eo
    for c, rng, ei, ptr in _cases(202, 25):
        k, m, mode, seed = int(rng.integers(1, 9)), int(rng.integers(1, 40)), MODES[int(rng.integers(0, 3))], int(rng.integers(0, 2 ** 40))
        eps = float(rng.choice([0.05, 0.3, 1.0]))
        outs, _ = _batch(ei, ptr, k).sample_epsilon(m, k, mode, seed=seed, epsilon=eps)
        _eq(outs, eo.sample_batch(ei, ptr, m, k, mode, seed, eps), ("eps", c, k, m, mode, eps))


def test_fuzz_rwr(cuda_lib):
    from oracle import walk_oracle as wo
    for c, rng, ei, ptr in _cases(303, 25):
        k, m, mode, seed = int(rng.integers(1, 11)), int(rng.integers(1, 40)), MODES[int(rng.integers(0, 3))], int(rng.integers(0, 2 ** 40))
        p = float(rng.choice([0.0, 0.2, 0.9]))
        outs, _ = _batch(ei, ptr, k).sample_rwr(m, k, mode, seed=seed, p_restart=p)
        _eq(outs, wo.rwr_sample_batch(ei, ptr, m, k, mode, seed, p), ("rwr", c, k, m, mode, p))


def test_fuzz_uniform_enumeration(cuda_lib):
    from oracle import walk_oracle as wo
    rng0 = np.random.default_rng(404)
    for c in range(25):
        G = int(rng0.integers(1, 30))
        big = [(int(rng0.integers(0, G)), int(rng0.integers(65, 150)))] if rng0.random() < 0.4 else []
        ei, ptr = random_batch(rng0, G, 1, 40, deg=float(rng0.uniform(0.8, 1.8)), weird=bool(rng0.random() < 0.7), big=big)
        k, m, mode, seed = int(rng0.integers(1, 6 if big else 8)), int(rng0.integers(1, 40)), MODES[int(rng0.integers(0, 3))], int(rng0.integers(0, 2 ** 40))
        outs, _ = _batch(ei, ptr, k).sample_uniform_enum(m, k, mode, seed=seed)
        _eq(outs, wo.uniform_sample_batch(ei, ptr, m, k, mode, seed), ("enum", c, k, m, mode))


def test_fuzz_wl_and_encoder_inputs(cuda_lib):
    from oracle import wl_oracle as wlo
    from ugs_b200 import encode, wl
    for c, rng, ei, ptr in _cases(505, 12, n_hi=30):
        k, m, seed = int(rng.integers(1, 9)), int(rng.integers(1, 20)), int(rng.integers(0, 2 ** 40))
        if ptr[0] != 0:
            ei, ptr = ei - ptr[0], ptr - ptr[0]
        b = _batch(ei, ptr, k)
        (nodes, eidx, eptr, sptr, esrc), _ = b.sample(m, k, "sample", seed=seed)
        it = int(rng.integers(1, 4))
        assert wl.wl_hashes(nodes, eidx, eptr, it) == wlo.batch_hashes(nodes.cpu().numpy(), eidx.cpu().numpy(),
                                                                      eptr.cpu().numpy(), it), ("wl", c)
        N, E = int(ptr[-1]), ei.shape[1]
        x = torch.randn(max(N, 1), 3).cuda()
        ea = torch.randn(max(E, 1), 2).cuda()
        f = encode.sample_encoder_inputs(b, x, ea, m, k, seed=seed)
        a = encode.encoder_inputs(x, ea, nodes, eidx, eptr, esrc)
        for u, v in zip(a, f[:4]):
            assert torch.equal(u, v), ("encode", c)
