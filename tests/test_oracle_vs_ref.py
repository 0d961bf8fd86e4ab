"""CPU: the C oracle against the REAL reference built in oracle/_ref (skipped where the reference tree was
never built, e.g. on the GPU box the prebuilt files travel with the snapshot)."""
import numpy as np
import pytest

from helpers import random_batch


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_loader
    if not ref_loader.available("ugs_sampler", "nopin"):
        pytest.skip("oracle/_ref not built")
    return ref_loader.load("ugs_sampler", "nopin")


def test_random_graph_tables(ref):
    import torch
    from oracle import ref_loader, ugs_oracle as uo
    rng = np.random.default_rng(0)
    for t in range(150):
        n = int(rng.integers(1, 40))
        E = int(rng.integers(0, 4 * n + 1))
        ei = rng.integers(-1 if t % 7 == 0 else 0, n + (2 if t % 5 == 0 else 0), size=(2, E)).astype(np.int64)
        k = int(rng.integers(1, 7))
        h = ref.create_preproc(torch.from_numpy(ei), n, k)
        dr = ref_loader.dump_preproc(ref, h)
        info = ref.get_preproc_info(h)
        ref.destroy_preproc(h)
        P = uo.Preproc(ei, n, k)
        do = P.dump()
        for key in ("indptr", "indices", "edge_col", "order", "index_of", "suffix_deg", "bucket_b"):
            assert np.array_equal(dr[key], do[key]), (t, key)
        assert dr["Z"] == do["Z"]
        assert dict(info) == P.info()


@pytest.mark.parametrize("mode", ["sample", "graph", "global"])
def test_batch_extraction_on_reference_samples(ref, mode):
    import torch
    from oracle import ugs_oracle as uo
    rng = np.random.default_rng(3)
    ei, ptr = random_batch(rng, 25, 1, 50, deg=2.5)
    outs = ref.sample_batch(torch.from_numpy(ei), torch.from_numpy(ptr), 7, 4, mode, 42)
    mine = uo.sample_batch(ei, ptr, 7, 4, mode, 42, given_nodes=outs[0].numpy())
    for a, b in zip(outs, mine):
        assert np.array_equal(a.numpy(), b)
