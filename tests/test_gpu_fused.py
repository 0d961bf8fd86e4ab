"""The fused small-batch path (csrc/ugs_fused.cu: one library call, one host round trip, worst-case-sized emission
buffers, general-path item counts read on the device) against the staged path and the oracle: bit-identical."""
import numpy as np
import pytest
import torch

from helpers import bidir, random_batch

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _staged(ei, ptr, m, k, mode, seed, **kw):
    from ugs_b200.engine import DeviceBatch
    b = DeviceBatch(_dev(ei), _dev(ptr), k, int(ptr[0]), int(ptr[-1] - ptr[0]))
    return b.sample(m, k, mode, seed=seed, **kw)


def _fused(ei, ptr, m, k, mode, seed, **kw):
    from ugs_b200.engine import sample_batch_fused
    return sample_batch_fused(_dev(ei), _dev(ptr), m, k, mode, seed, int(ptr[0]), int(ptr[-1] - ptr[0]), **kw)


NAMES = ["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"]


@pytest.mark.parametrize("mode", ["sample", "graph", "global"])
@pytest.mark.parametrize("seed,k,m", [(1, 3, 7), (2, 6, 64), (3, 8, 33), (4, 1, 5), (5, 20, 9), (6, 5, 100)])
def test_fused_equals_staged_and_oracle(cuda_lib, mode, seed, k, m):
    from oracle import ugs_oracle as uo
    rng = np.random.default_rng(seed)
    ei, ptr = random_batch(rng, 40, 1, 45, deg=2.5, big=[(5, 120), (9, 70), (11, 700)])
    f = _fused(ei, ptr, m, k, mode, seed)
    assert f is not None
    (fo, fs), (so, ss) = f, _staged(ei, ptr, m, k, mode, seed)
    ref = uo.sample_batch(ei, ptr, m, k, mode, seed)
    for name, a, b, r in zip(NAMES, fo, so, ref):
        assert a.dtype == torch.int64 and a.is_contiguous(), name
        assert torch.equal(a, b), f"{name}: fused != staged"
        assert np.array_equal(a.cpu().numpy(), r), f"{name}: fused != oracle"
    assert fs == ss


def test_fused_large_graph_in_a_small_batch(cuda_lib):
    """A general-path graph above UGS_LARGE_N = 4096 nodes: the cooperative preprocessing kernel reads n_large on the device."""
    from ugs_b200 import synth
    ba, pba = synth.barabasi_albert(6000, 3, seed=2)
    rng = np.random.default_rng(3)
    wei, wptr = random_batch(rng, 6, 5, 30, deg=2.0)
    ei, ptr = synth.concat_batches([(wei, wptr), (ba, pba), (wei, wptr)])
    (fo, fs), (so, ss) = _fused(ei, ptr, 20, 6, "sample", 9), _staged(ei, ptr, 20, 6, "sample", 9)
    assert all(torch.equal(a, b) for a, b in zip(fo, so)) and fs == ss and fs["n_general_graphs"] >= 1


def test_fused_sharding_arguments_and_per_graph_seed(cuda_lib):
    rng = np.random.default_rng(8)
    ei, ptr = random_batch(rng, 25, 3, 40, deg=2.5, weird=False)
    for kw in (dict(graph_index_base=1000), dict(sample_index_base=17), dict(per_graph_seed=True)):
        (fo, _), (so, _) = _fused(ei, ptr, 12, 4, "graph", 5, **kw), _staged(ei, ptr, 12, 4, "graph", 5, **kw)
        assert all(torch.equal(a, b) for a, b in zip(fo, so)), kw


def test_fused_declines_batches_that_need_the_staged_path(cuda_lib):
    rng = np.random.default_rng(7)
    ei, ptr = random_batch(rng, 20, 3, 30, deg=2.0)
    perm = rng.permutation(ei.shape[1])
    assert _fused(np.ascontiguousarray(ei[:, perm]), ptr, 9, 4, "global", 1) is None  # columns not graph-contiguous
    bad = ptr.copy()
    bad[3], bad[4] = bad[4], bad[3]
    assert _fused(ei, bad, 9, 4, "global", 1) is None  # ptr decreasing: the staged path raises for it
    # degenerate but valid: no columns at all, graphs smaller than k
    e0 = np.zeros((2, 0), np.int64)
    f, s = _fused(e0, ptr, 3, 2, "sample", 1), _staged(e0, ptr, 3, 2, "sample", 1)
    assert all(torch.equal(a, b) for a, b in zip(f[0], s[0])) and f[0][1].shape == (2, 0)


def test_dropin_takes_the_fused_path_for_minibatches(cuda_lib, monkeypatch):
    """ugs_sampler.sample_batch on a DataLoader-sized batch (CPU tensors in, pinned CPU tensors out) equals the staged
    path (UGS_B200_FUSED=0) and the oracle; the kernels launched are the same ones."""
    import ugs_sampler
    from oracle import ugs_oracle as uo
    from ugs_b200 import _lib, synth
    ei, ptr = synth.make_batch("proteins", 32, seed=3)
    tei, tptr = torch.from_numpy(ei), torch.from_numpy(ptr)
    n0 = int(_lib.lib().ugs_launch_count())
    a = ugs_sampler.sample_batch(tei, tptr, 50, 6, mode="sample", seed=3)
    n1 = int(_lib.lib().ugs_launch_count())
    monkeypatch.setenv("UGS_B200_FUSED", "0")
    b = ugs_sampler.sample_batch(tei, tptr, 50, 6, mode="sample", seed=3)
    n2 = int(_lib.lib().ugs_launch_count())
    ref = uo.sample_batch(ei, ptr, 50, 6, "sample", 3)
    for x, y, r in zip(a, b, ref):
        assert not x.is_cuda and x.is_pinned() and torch.equal(x, y) and np.array_equal(x.numpy(), r)
    assert n1 - n0 >= 5 and n2 - n1 >= 5


def test_fused_guess_without_general_path_graphs_is_checked(cuda_lib):
    """After a batch without general-path graphs the next fused call is queued without the general-path kernels
    (assume_no_general); when that batch does hold a graph above 64 vertices the device count exposes the wrong guess and
    the call is repeated in full -- results stay bit-identical to the staged path in every order of batches."""
    from ugs_b200 import synth
    from ugs_b200.engine import _fused_tls
    small = synth.make_batch("mutag", 40, seed=3)
    mixed = synth.make_batch("proteins", 40, seed=3)
    assert np.diff(mixed[1]).max() > 64 and np.diff(small[1]).max() <= 64
    dev = torch.cuda.current_device()
    for order in ([small, small, mixed, mixed, small], [mixed, small, mixed]):
        for ei, ptr in order:
            fo, fs = _fused(ei, ptr, 12, 5, "sample", 9)
            so, ss = _staged(ei, ptr, 12, 5, "sample", 9)
            for name, a, b in zip(NAMES, fo, so):
                assert torch.equal(a, b), name
            assert fs == ss
            assert _fused_tls.info[("no_general", dev)] == (ss["n_general_graphs"] == 0)


def test_fused_compact_and_narrowed_results_agree(cuda_lib, monkeypatch):
    """Bound-sized edge arrays up to FUSED_COMPACT_MAX_BYTES are filled compactly by the emission kernels (rows behind
    one another at the device-side total; the results are views of that allocation); above it the bound-sized arrays are
    narrowed by ugs_fused_finish. Same tensors either way, for odd and even totals, with general-path graphs in the batch."""
    from ugs_b200 import engine, synth
    rng = np.random.default_rng(5)
    for ei, ptr in (random_batch(rng, 30, 1, 45, deg=2.5, big=[(5, 120)]), synth.make_batch("mutag", 25, seed=2),
                    synth.make_batch("proteins", 20, seed=4)):
        for m, k, mode in ((7, 4, "sample"), (16, 6, "global"), (3, 3, "graph")):
            monkeypatch.setattr(engine, "FUSED_COMPACT_MAX_BYTES", 1 << 40)
            co, cs = _fused(ei, ptr, m, k, mode, 11)
            assert co[1].untyped_storage().nbytes() >= 24 * 2 * ei.shape[1] * m  # a view of the bound-sized array
            monkeypatch.setattr(engine, "FUSED_COMPACT_MAX_BYTES", 0)
            no, ns = _fused(ei, ptr, m, k, mode, 11)
            assert no[1].untyped_storage().nbytes() == 24 * cs["n_edge_entries"]
            so, ss = _staged(ei, ptr, m, k, mode, 11)
            for name, a, b, c in zip(NAMES, co, no, so):
                assert torch.equal(a, b) and torch.equal(a, c), (name, m, k, mode)
                assert a.is_contiguous()
            assert cs == ns == ss
