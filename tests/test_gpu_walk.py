"""GPU: rwr_sampler and uniform_sampler drop-ins (csrc/ugs_rwr.cu, csrc/ugs_enum.cu) -- bit-exact against the oracle
(oracle/walk_oracle.c, itself pinned against the real reference), extraction of the reference's own samples against
the golden vectors, laws by chi-square, reference API behaviour."""
import os
from collections import Counter

import numpy as np
import pytest
import torch
from scipy.stats import chi2

from helpers import bidir, random_batch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "walk_golden.npz")


def _batch(ei, ptr, k):
    from ugs_b200.engine import DeviceBatch
    return DeviceBatch(torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda(), k, int(ptr[0]), int(ptr[-1] - ptr[0]))


@pytest.mark.parametrize("mode", ["sample", "global"])
@pytest.mark.parametrize("seed,k,m,p", [(1, 3, 7, 0.2), (2, 5, 33, 0.0), (3, 4, 64, 0.7), (4, 9, 5, 0.2)])
def test_rwr_bit_exact_vs_oracle(cuda_lib, mode, seed, k, m, p):
    from oracle import walk_oracle as wo
    rng = np.random.default_rng(seed)
    ei, ptr = random_batch(rng, 30, 1, 45, deg=2.5, big=[(4, 140), (9, 70)])  # bitmask-path and general-path graphs
    outs, stats = _batch(ei, ptr, k).sample_rwr(m, k, mode, seed=seed, p_restart=p)
    ref = wo.rwr_sample_batch(ei, ptr, m, k, mode, seed, p)
    for a, r, nm in zip(outs, ref, ["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"]):
        assert np.array_equal(a.cpu().numpy(), r), nm
    sampled = np.repeat(np.diff(ptr) >= k, m)  # rows of skipped graphs (n < k) are not counted as failures
    assert stats["n_incomplete"] == int((ref[0][sampled, 0] < 0).sum())
    assert (outs[4] == -1).all()


@pytest.mark.parametrize("mode", ["sample", "global"])
@pytest.mark.parametrize("seed,k,m", [(1, 3, 7), (2, 5, 33), (3, 4, 64), (4, 1, 5), (5, 2, 9), (6, 7, 3)])
def test_uniform_enum_bit_exact_vs_oracle(cuda_lib, mode, seed, k, m):
    from oracle import walk_oracle as wo
    rng = np.random.default_rng(seed)
    ei, ptr = random_batch(rng, 40, 1, 30, deg=1.4, big=[(4, 64), (9, 33)])
    b = _batch(ei, ptr, k)
    outs, stats = b.sample_uniform_enum(m, k, mode, seed=seed)
    ref = wo.uniform_sample_batch(ei, ptr, m, k, mode, seed)
    for a, r, nm in zip(outs, ref, ["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"]):
        assert np.array_equal(a.cpu().numpy(), r), nm
    # the enumerated lists themselves, graph by graph, in the oracle's order
    lst, es = b.enumerate_connected(k)
    off = b.enum_graph_offsets(es).cpu().numpy()
    lst = lst.cpu().numpy().view(np.uint64)[:, 0]
    for g in range(len(ptr) - 1):
        lo, n = int(ptr[g]), int(ptr[g + 1] - ptr[g])
        if n < k or n <= 0:
            assert off[g + 1] == off[g]
            continue
        msk = (ei[0] >= lo) & (ei[0] < lo + n) & (ei[1] >= lo) & (ei[1] < lo + n)
        want = wo.enumerate_connected(ei[:, msk] - lo, n, k)
        assert np.array_equal(lst[off[g]:off[g + 1]], want), g
    assert stats["n_enumerated"] == int(off[-1])


@pytest.mark.parametrize("tag", ["rwr", "uniform"])
@pytest.mark.parametrize("mode", ["sample", "global"])
def test_extraction_of_reference_samples_matches_golden(cuda_lib, tag, mode):
    """The real reference's node rows through the CUDA extraction kernels == the reference's edges / pointers."""
    gold = np.load(GOLD)
    for name in [str(x) for x in gold["names"]]:
        ei, ptr = gold[f"{name}/edge_index"], gold[f"{name}/ptr"]
        k, m = (int(x) for x in gold[f"{name}/km"])
        nodes = torch.from_numpy(gold[f"{name}/{tag}/{mode}/nodes"]).cuda()
        b = _batch(ei, ptr, k)
        if tag == "rwr":
            eidx, eptr, esrc = b.extract(nodes, m, k, "sample" if mode == "sample" else "global")
            esrc = torch.full_like(esrc, -1)
        else:
            (_, eidx, eptr, _, esrc), _ = b.sample_epsilon(m, k, mode, 1, 0.5, given_nodes=nodes)
        assert np.array_equal(eidx.cpu().numpy(), gold[f"{name}/{tag}/{mode}/edge_index_out"]), name
        assert np.array_equal(eptr.cpu().numpy(), gold[f"{name}/{tag}/{mode}/edge_ptr"]), name
        assert np.array_equal(esrc.cpu().numpy(), gold[f"{name}/{tag}/{mode}/edge_src"]), name


@pytest.mark.parametrize("mode", ["sample", "global"])
@pytest.mark.parametrize("seed,k,m", [(11, 3, 9), (12, 4, 5), (13, 1, 3), (14, 2, 40)])
def test_uniform_enum_big_graphs_bit_exact_vs_oracle(cuda_lib, mode, seed, k, m):
    """Batches mixing bitmask-path graphs, graphs with <= 64 nodes but too many columns for that path, and graphs of
    65..200 nodes (multi-word bitsets, records of k 16-bit ids): samples bit-exact vs the oracle, lists in its order."""
    from oracle import walk_oracle as wo
    rng = np.random.default_rng(seed)
    ei, ptr = random_batch(rng, 14, 1, 30, deg=1.5, weird=False, big=[(2, 130), (5, 66), (9, 200), (11, 40)])
    # a dense small graph: 40 nodes, > 384 columns -> general path, but still 64-bit masks
    lo, n = int(ptr[11]), int(ptr[12] - ptr[11])
    if n >= 8:
        extra = np.array([(lo + a, lo + b) for a in range(min(n, 30)) for b in range(a + 1, min(n, 30))], np.int64).T
        first = int(np.searchsorted(ei[0], ptr[12]))
        ei = np.ascontiguousarray(np.concatenate([ei[:, :first], extra, extra[::-1], ei[:, first:]], axis=1))
    b = _batch(ei, ptr, k)
    assert b.n_general >= 3
    outs, stats = b.sample_uniform_enum(m, k, mode, seed=seed)
    ref = wo.uniform_sample_batch(ei, ptr, m, k, mode, seed)
    for a, r, nm in zip(outs, ref, ["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"]):
        assert np.array_equal(a.cpu().numpy(), r), nm
    lst, es = b.enumerate_connected(k)
    off = b.enum_graph_offsets(es).cpu().numpy()
    rec = lst.cpu().numpy()
    assert rec.shape[1] == 2
    for g in (2, 9):
        lo, n = int(ptr[g]), int(ptr[g + 1] - ptr[g])
        msk = (ei[0] >= lo) & (ei[0] < lo + n) & (ei[1] >= lo) & (ei[1] < lo + n)
        want = wo.enumerate_connected_any(ei[:, msk] - lo, n, k)
        ids = rec[off[g]:off[g + 1]].copy().view(np.uint16).reshape(-1, 8)[:, :k]
        assert [tuple(int(x) for x in r) for r in ids] == want, g


def test_laws_on_the_device(cuda_lib):
    """200k device samples: uniform_sampler is uniform over the connected k-subsets (and complete: every one of them
    appears); rwr_sampler follows the Markov-chain law of the ordered rows."""
    from oracle import exact_law as xl, ugs_oracle as uo
    import rwr_sampler
    import uniform_sampler
    und, n, k = [(i, (i + 1) % 6) for i in range(6)] + [(0, 3), (3, 6), (6, 7)], 8, 4
    ei = bidir(und)
    rows = xl.rows_from_dump(uo.Preproc(ei, n, k).dump())
    M = 200000
    tei, tptr = torch.from_numpy(ei).cuda(), torch.tensor([0, n]).cuda()
    nodes = uniform_sampler.sample_batch(tei, tptr, M, k, "sample", 3)[0].cpu().numpy()
    sup = xl.connected_subsets(rows, k)
    law = {s: 1.0 / len(sup) for s in sup}
    obs = Counter(frozenset(int(x) for x in r) for r in nodes)
    stat, df, extra = xl.chi_square(obs, law, M)
    assert extra == 0 and set(obs) == set(sup) and stat < chi2.ppf(0.999, df), (stat, df)
    cv = np.std(list(obs.values())) / np.mean(list(obs.values()))
    assert cv < 0.15  # the reference's own uniformity threshold (tests/test_uniformity.py:155-160)
    nodes = rwr_sampler.sample_batch(tei, tptr, M, k, "sample", 3, 0.2)[0].cpu().numpy()
    law = {kk: v for kk, v in xl.p_rwr(rows, k, 0.2).items() if v > 1e-9}
    obs = Counter(tuple(int(x) for x in r) if r[0] >= 0 else () for r in nodes)
    stat, df, extra = xl.chi_square(obs, law, M)
    assert extra == 0 and stat < chi2.ppf(0.999, df), (stat, df)


def test_dropin_modules_behave_like_the_reference(cuda_lib):
    import rwr_sampler
    import uniform_sampler
    ei, ptr = random_batch(np.random.default_rng(3), 12, 4, 25, deg=2.0, weird=False)
    tei, tptr = torch.from_numpy(ei), torch.from_numpy(ptr)
    for mod, extra in ((rwr_sampler, dict(p_restart=0.3)), (uniform_sampler, {})):
        outs = mod.sample_batch(edge_index=tei, ptr=tptr, m_per_graph=5, k=3, mode="sample", seed=1, **extra)
        assert len(outs) == 5 and all(o.dtype == torch.int64 and not o.is_cuda and o.is_pinned() for o in outs)
        nodes, eidx, eptr, sptr, esrc = outs
        assert nodes.shape == (12 * 5, 3) and sptr.tolist() == [5 * g for g in range(13)]
        assert eptr.shape == (61,) and eptr[-1] == eidx.shape[1] == esrc.shape[0]
        assert (eidx >= 0).all() and (eidx < 3).all()
        g = torch.repeat_interleave(torch.arange(12), 5)
        ok = nodes[:, 0] >= 0
        assert ((nodes[ok] >= tptr[g[ok]].unsqueeze(1)) & (nodes[ok] < tptr[g[ok] + 1].unsqueeze(1))).all()
        again = mod.sample_batch(tei, tptr, 5, 3, "sample", 1, **extra)
        assert all(torch.equal(a, b) for a, b in zip(outs, again))  # reproducible
        cu = mod.sample_batch(tei.cuda(), tptr.cuda(), 5, 3, "sample", 1, **extra)
        assert all(c.is_cuda and torch.equal(c.cpu(), a) for c, a in zip(cu, outs))  # returns on the input device
        glob = mod.sample_batch(tei, tptr, 5, 3, "global", 1, **extra)
        assert torch.equal(glob[0], nodes) and glob[1].max() >= 3
        with pytest.raises(RuntimeError, match="int64"):
            mod.sample_batch(tei.int(), tptr, 5, 3)
    assert (rwr_sampler.sample_batch(tei, tptr, 5, 3)[4] == -1).all()
    with pytest.raises(RuntimeError, match="p_restart"):
        rwr_sampler.sample_batch(tei, tptr, 5, 3, p_restart=1.5)
    with pytest.raises(RuntimeError, match="k must be >= 1"):
        rwr_sampler.sample_batch(tei, tptr, 5, 0)
    # uniform enumeration: graphs above 64 nodes need k <= 8, above 2048 nodes are refused
    big, bptr = random_batch(np.random.default_rng(1), 2, 100, 120, deg=2.0, weird=False)
    with pytest.raises(RuntimeError, match="k <= 8"):
        uniform_sampler.sample_batch(torch.from_numpy(big), torch.from_numpy(bptr), 2, 9)
    huge = torch.stack([torch.arange(0, 2999), torch.arange(1, 3000)])
    with pytest.raises(RuntimeError, match="<= 2048 nodes"):
        uniform_sampler.sample_batch(huge, torch.tensor([0, 3000]), 2, 3)


def test_qm9_shaped_completeness(cuda_lib):
    """BASELINE configs[2]: QM9-shaped molecules, exhaustive k=4/5: no -1 rows, and with enough draws every connected
    k-subgraph of a molecule is sampled (reference tests/test_ugs_qm9_completeness.py:44-49 asserts the former)."""
    import uniform_sampler
    from oracle import walk_oracle as wo
    from ugs_b200 import synth
    ei, ptr = synth.make_batch("qm9", 300, seed=9)
    for k in (4, 5):
        nodes = uniform_sampler.sample_batch(torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda(), 10, k)[0]
        sizes = np.diff(ptr)
        bad = (nodes[:, 0] < 0).cpu().numpy().reshape(-1, 10).any(axis=1)
        assert not bad[sizes >= k].any()
    g = int(np.argmax(np.diff(ptr)))
    lo, n = int(ptr[g]), int(ptr[g + 1] - ptr[g])
    msk = (ei[0] >= lo) & (ei[0] < lo + n)
    sup = wo.enumerate_connected(ei[:, msk] - lo, n, 4)
    nodes = uniform_sampler.sample_batch(torch.from_numpy(ei[:, msk] - lo).cuda(), torch.tensor([0, n]).cuda(),
                                         200 * len(sup), 4)[0].cpu().numpy()
    seen = {sum(1 << int(x) for x in r) for r in nodes}
    assert seen == {int(x) for x in sup}
