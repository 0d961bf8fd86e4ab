"""GPU: WL graphlet classification (csrc/ugs_wl.cu) -- the hex strings equal the reference's compute_wl_hash
(golden) and the oracle's on real sampler output; ids and vocabulary building behave like the reference's loops."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import random_batch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "wl_golden.json")


def _as_samples(rows, k):
    """Pack golden cases as sampler-style output: nodes[B,k] (-1 padded), edge_index[2,E], edge_ptr[B+1]."""
    nodes = np.full((len(rows), k), -1, np.int64)
    eu, ev, eptr = [], [], [0]
    for s, r in enumerate(rows):
        nodes[s, :r["n"]] = np.arange(r["n"]) + 100 * s
        for u, v in r["edges"]:
            eu.append(u)
            ev.append(v)
        eptr.append(len(eu))
    ei = np.array([eu, ev], np.int64).reshape(2, -1)
    return torch.from_numpy(nodes).cuda(), torch.from_numpy(ei).cuda(), torch.tensor(eptr).cuda()


@pytest.mark.parametrize("it", [1, 2, 3])
def test_wl_hashes_equal_the_reference_golden(cuda_lib, it):
    from ugs_b200 import wl
    rows = json.load(open(GOLD))
    nodes, ei, eptr = _as_samples(rows, 12)
    got = wl.wl_hashes(nodes, ei, eptr, it)
    for r, h in zip(rows, got):
        assert h == r["hash"][str(it)], (r["name"], it)


def test_wl_hashes_with_node_features_equal_the_reference_golden(cuda_lib):
    """use_node_features=True: md5 of the feature rows on the device (1-, 28-, 56- and 80-byte rows: one md5 block,
    padding spilling into a second block, two data blocks)."""
    from ugs_b200 import wl
    rows = [r for r in json.load(open(GOLD)) if "feat" in r]
    nodes, ei, eptr = _as_samples(rows, 12)
    for kk in rows[0]["feat"]:
        dt = getattr(torch, rows[0]["feat"][kk]["dtype"])
        ncol = len(rows[0]["feat"][kk]["values"][0])
        x = torch.zeros((100 * len(rows) + 12, ncol), dtype=dt)
        for s, r in enumerate(rows):
            x[100 * s:100 * s + r["n"]] = torch.tensor(r["feat"][kk]["values"], dtype=dt)
        got = wl.wl_hashes(nodes, ei, eptr, 3, x=x.cuda())
        for r, h in zip(rows, got):
            assert h == r["feat"][kk]["hash"], (r["name"], kk)


@pytest.mark.parametrize("sampler", ["ugs", "epsilon", "uniform"])
def test_wl_on_sampler_output_equals_oracle(cuda_lib, sampler):
    """4x-multiplicity UGS entries and one-entry-per-column epsilon/uniform entries give the same nx.Graph."""
    from oracle import wl_oracle as wo
    from ugs_b200 import wl
    from ugs_b200.engine import DeviceBatch
    rng = np.random.default_rng(2)
    ei, ptr = random_batch(rng, 25, 1, 30, deg=2.3)
    k, m = 5, 9
    b = DeviceBatch(torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda(), k, 0, int(ptr[-1]))
    if sampler == "ugs":
        outs, _ = b.sample(m, k, "sample", seed=3)
    elif sampler == "epsilon":
        outs, _ = b.sample_epsilon(m, k, "sample", seed=3, epsilon=0.5)
    else:
        outs, _ = b.sample_uniform_enum(m, k, "sample", seed=3)
    nodes, eidx, eptr = outs[0], outs[1], outs[2]
    got = wl.wl_hashes(nodes, eidx, eptr, 3)
    want = wo.batch_hashes(nodes.cpu().numpy(), eidx.cpu().numpy(), eptr.cpu().numpy(), 3)
    assert got == want
    x = torch.randint(0, 4, (int(ptr[-1]), 3), generator=torch.Generator().manual_seed(1)).float()
    got = wl.wl_hashes(nodes, eidx, eptr, 2, x=x.cuda())
    want = wo.batch_hashes(nodes.cpu().numpy(), eidx.cpu().numpy(), eptr.cpu().numpy(), 2, x=x.numpy())
    assert got == want
    assert any(h is None for h in got) or (nodes >= 0).any(dim=1).all()


def test_wl_ids_and_vocabulary_like_the_reference_loops(cuda_lib):
    from oracle import wl_oracle as wo
    from ugs_b200 import synth, wl
    from ugs_b200.engine import DeviceBatch
    ei, ptr = synth.make_batch("qm9", 200, seed=4)
    k, m = 4, 10
    b = DeviceBatch(torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda(), k, 0, int(ptr[-1]))
    (nodes, eidx, eptr, _, _), _ = b.sample(m, k, "sample", seed=1)
    hashes = wo.batch_hashes(nodes.cpu().numpy(), eidx.cpu().numpy(), eptr.cpu().numpy(), 3)
    # build_wl_vocabulary_from_loader (wl_vocab.py:150-176): first-seen order
    want_vocab = {}
    for h in hashes:
        if h is not None and h not in want_vocab:
            want_vocab[h] = len(want_vocab)
    got_vocab = wl.extend_wl_vocabulary({}, nodes, eidx, eptr, 3)
    assert got_vocab == want_vocab and list(got_vocab) == list(want_vocab)
    assert 2 <= len(got_vocab) <= 6  # connected 4-node graphlets: path, star, cycle, paw, diamond, clique
    # _compute_wl_ids with a vocabulary that misses one class: unknown -> len(vocab)
    part = {h: i for h, i in want_vocab.items() if i != 1}
    ids = wl.compute_wl_ids(nodes, eidx, eptr, part, 3).cpu().tolist()
    want = [len(part) if h is None else part.get(h, len(part)) for h in hashes]
    assert ids == want
    # a second batch extends the ids without renumbering
    (n2, e2, p2, _, _), _ = b.sample(m, k, "sample", seed=2)
    v2 = wl.extend_wl_vocabulary(dict(got_vocab), n2, e2, p2, 3)
    assert all(v2[h] == i for h, i in got_vocab.items())
    with pytest.raises(ValueError):
        wl.wl_digests(nodes, eidx, eptr, 0)
    with pytest.raises(RuntimeError, match="k in"):
        wl.wl_digests(torch.zeros((1, 17), dtype=torch.long).cuda(), eidx, torch.zeros(2, dtype=torch.long).cuda(), 3)
