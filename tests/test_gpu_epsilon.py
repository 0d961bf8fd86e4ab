"""GPU: the epsilon-uniform variant (drop-in `epsilon_uniform_sampler`) against the oracle (bit-exact: shared
Philox rule), the reference's golden extraction vectors, and the enumerated law."""
import os
from collections import Counter

import numpy as np
import pytest
import torch
from scipy.stats import chi2

from helpers import bidir, random_batch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "eps_golden.npz")


def _batch(ei, ptr, k):
    from ugs_b200.engine import DeviceBatch
    return DeviceBatch(torch.from_numpy(np.ascontiguousarray(ei)).cuda(), torch.from_numpy(np.ascontiguousarray(ptr)).cuda(),
                       k, int(ptr[0]), int(ptr[-1] - ptr[0]))


@pytest.mark.parametrize("mode", ["sample", "global"])
@pytest.mark.parametrize("seed,k,m,eps", [(30, 3, 7, 0.1), (31, 6, 40, 0.5), (32, 8, 33, 0.01), (33, 4, 5, 1.0)])
def test_eps_bit_exact_vs_oracle(cuda_lib, mode, seed, k, m, eps):
    from oracle import eps_oracle as eo
    rng = np.random.default_rng(seed)
    ei, ptr = random_batch(rng, 40, 1, 45, deg=2.5, big=[(5, 120), (9, 70)])
    outs, stats = _batch(ei, ptr, k).sample_epsilon(m, k, mode, seed=seed, epsilon=eps)
    ref = eo.sample_batch(ei, ptr, m, k, mode, seed, eps)
    for name, a, r in zip(["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"], outs, ref):
        a = a.cpu().numpy()
        assert a.shape == r.shape, name
        if not np.array_equal(a, r):
            bad = np.nonzero((a != r).reshape(a.shape[0], -1).any(axis=1))[0] if a.ndim == 2 and name == "nodes" else []
            raise AssertionError(f"{name} differs (mode={mode}); first bad rows {list(bad[:5])}")


@pytest.mark.parametrize("mode", ["sample", "global"])
def test_eps_cuda_extraction_matches_reference_golden(cuda_lib, mode):
    gold = np.load(GOLD)
    for name in [str(x) for x in gold["names"]]:
        ei, ptr = gold[f"{name}/edge_index"], gold[f"{name}/ptr"]
        k, m = (int(x) for x in gold[f"{name}/km"])
        nodes = torch.from_numpy(gold[f"{name}/{mode}/nodes"]).cuda()
        outs, _ = _batch(ei, ptr, k).sample_epsilon(m, k, mode, given_nodes=nodes)
        assert np.array_equal(outs[1].cpu().numpy(), gold[f"{name}/{mode}/edge_index_out"]), name
        assert np.array_equal(outs[2].cpu().numpy(), gold[f"{name}/{mode}/edge_ptr"]), name
        assert np.array_equal(outs[3].cpu().numpy(), gold[f"{name}/{mode}/sample_ptr"]), name
        assert np.array_equal(outs[4].cpu().numpy(), gold[f"{name}/{mode}/edge_src"]), name


def test_eps_law_on_gpu(cuda_lib):
    import epsilon_uniform_sampler as eus
    from oracle import exact_law as xl, ugs_oracle as uo
    for und, n, k, eps in [([(0, 1), (1, 2), (0, 2), (2, 3)], 4, 3, 0.1), ([(i, (i + 1) % 6) for i in range(6)], 6, 3, 0.5)]:
        ei = bidir(und)
        rows = xl.rows_from_dump(uo.Preproc(ei, n, k).dump())
        law = {kk: v for kk, v in xl.p_epsilon(rows, k, eps).items() if v > 1e-12}
        M = 200000
        nodes = eus.sample_batch(torch.from_numpy(ei).cuda(), torch.tensor([0, n]).cuda(), M, k, seed=77, epsilon=eps)[0]
        cnt = Counter((frozenset(r) if r[0] >= 0 else ()) for r in nodes.cpu().numpy().tolist())
        stat, df, extra = xl.chi_square(cnt, law, M)
        assert extra == 0 and stat < chi2.ppf(0.999, df)


def test_eps_dropin_api(cuda_lib):
    """reference tests/test_epsilon_sampler.py:28-42: eps in {0.01, 0.1, 0.5}, CPU and CUDA inputs."""
    import epsilon_uniform_sampler as eus
    from ugs_b200 import synth
    ei, ptr = synth.make_batch("mutag", 20, seed=3)
    tei, tptr = torch.from_numpy(ei), torch.from_numpy(ptr)
    for eps in (0.01, 0.1, 0.5):
        out = eus.sample_batch(tei, tptr, 8, 4, mode="sample", seed=42, epsilon=eps)
        assert len(out) == 5 and all(not t.is_cuda and t.dtype == torch.int64 for t in out)
        nodes, eidx, eptr, sptr, esrc = out
        assert nodes.shape == (160, 4) and sptr.tolist() == list(range(0, 161, 8)) and int(eptr[-1]) == eidx.shape[1]
        assert bool((nodes[:, 1:] > nodes[:, :-1]).all())  # sorted rows
        sid = np.searchsorted(eptr.numpy()[1:], np.arange(eidx.shape[1]), side="right")
        assert np.array_equal(nodes.numpy()[sid, eidx[0].numpy()], ei[0][esrc.numpy()])
        assert np.array_equal(nodes.numpy()[sid, eidx[1].numpy()], ei[1][esrc.numpy()])
    out = eus.sample_batch(tei.cuda(), tptr.cuda(), 8, 4, epsilon=0.1)
    assert all(t.is_cuda for t in out)
    with pytest.raises(RuntimeError, match="epsilon must be in"):
        eus.sample_batch(tei, tptr, 8, 4, epsilon=1.5)
