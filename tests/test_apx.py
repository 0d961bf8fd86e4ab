"""APX-UGS (reference apx_ugs_sampler): oracle sanity on CPU; on the GPU the CUDA path is compared bit-exactly with
the oracle (shared random-access Philox rule) and, in law, with frequencies recorded from the REAL reference
(tests/golden/apx_ref_counts.json: 240 requested samples, 29 CPU-minutes of the reference)."""
import json
import os
from collections import Counter

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "apx_ref_counts.json")
EDGES = [[0, 1], [1, 0], [1, 2], [2, 1], [2, 0], [0, 2], [2, 3], [3, 2], [3, 4], [4, 3]]  # test_apx_ugs.py:28-41


def _graph():
    ei = np.array(EDGES, np.int64).T.copy()
    return ei, np.array([0, ei.shape[1]], np.int64)


def test_apx_oracle_tables_on_the_reference_test_graph():
    from oracle import apx_oracle as ao
    ei, ptr = _graph()
    G = ao.ApxGraph(ei, ptr, 3, 42, 0.5)
    d = G.dump()
    assert G.n == 5 and d["rowptr"].tolist() == [0, 2, 4, 7, 9, 10]
    assert d["adj"].tolist() == [1, 2, 0, 2, 0, 1, 3, 2, 4, 3]  # sorted, deduplicated (:36-40)
    assert sorted(d["order"].tolist()) == [0, 1, 2, 3, 4] and d["position"][d["order"]].tolist() == [0, 1, 2, 3, 4]
    assert d["order"][0] == 2  # the only degree-3 vertex comes first
    # small-degree fix-up makes estimates exact: only vertex 2 can root a 3-graphlet in its suffix graph
    assert d["est"].tolist() == [0.0, 0.0, 243.0, 0.0, 0.0] and G.Z == 243.0
    # ptr is an EDGE range (:15-33): restricting it to the triangle's columns gives a 3-vertex graph
    G3 = ao.ApxGraph(ei, np.array([0, 6]), 3, 42, 0.5)
    assert G3.n == 3
    # n < k -> empty result (:473-478)
    s, p = ao.sample_batch(ei, np.array([0, 2]), 4, 3, "sample", 1, 0.5)
    assert s.shape == (3, 0) and p.tolist() == [0]


def test_apx_oracle_bounded_trials_are_deterministic():
    from oracle import apx_oracle as ao
    ei, ptr = _graph()
    G = ao.ApxGraph(ei, ptr, 3, 7, 0.5)
    rows, trials = G.sample(3, max_trials=3000)
    rows2, trials2 = ao.ApxGraph(ei, ptr, 3, 7, 0.5).sample(3, max_trials=3000)
    assert np.array_equal(rows, rows2) and np.array_equal(trials, trials2)
    for r in rows.tolist():
        assert 2 in r and len(set(r)) == 3


@pytest.mark.gpu
def test_apx_cuda_tables_bit_exact(cuda_lib):
    import torch
    import apx_ugs_sampler as aus
    from oracle import apx_oracle as ao
    from ugs_b200 import synth
    cases = [(_graph(), 3, 0.5), (_graph(), 4, 0.2)]
    ei, ptr = synth.make_batch("mutag", 1, seed=5)
    cases.append(((ei, np.array([0, ei.shape[1]], np.int64)), 3, 0.5))
    for (ei, ptr), k, eps in cases:
        d = aus._dump_tables(torch.from_numpy(ei), torch.from_numpy(ptr), k, seed=11, epsilon=eps)
        o = ao.ApxGraph(ei, ptr, k, 11, eps).dump()
        assert d["n"] == len(o["order"])
        assert np.array_equal(d["rowptr"], o["rowptr"]) and np.array_equal(d["adj"], o["adj"])
        assert np.array_equal(d["order"], o["order"]) and np.array_equal(d["position"], o["position"])
        assert np.array_equal(d["est"], o["est"])


@pytest.mark.gpu
def test_apx_cuda_samples_bit_exact_with_bounded_trials(cuda_lib):
    import torch
    import apx_ugs_sampler as aus
    from oracle import apx_oracle as ao
    ei, ptr = _graph()
    m, cap = 12, 30000
    s, p = aus.sample_batch(torch.from_numpy(ei), torch.from_numpy(ptr), m, 3, "sample", 5, 0.5, _max_trials=cap)
    rows, trials = ao.ApxGraph(ei, ptr, 3, 5, 0.5).sample(m, max_trials=cap)
    assert aus.sample_batch.last_trials == trials.tolist()
    assert s.shape == (3, rows.shape[0]) and np.array_equal(s.numpy().T, rows)
    assert p.tolist() == list(range(rows.shape[0] + 1)) and not s.is_cuda


@pytest.mark.gpu
def test_apx_cuda_law_matches_the_real_reference(cuda_lib):
    """Two-sample chi-square (alpha = 0.001) between the reference's recorded frequencies and 600 requested CUDA
    samples with the reference's 10^6-trial cap; the fraction of dropped samples must agree too."""
    import torch
    import apx_ugs_sampler as aus
    from scipy.stats import chi2
    gold = json.load(open(GOLD))
    ei, ptr = _graph()
    m = 600
    s, p = aus.sample_batch(torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda(), m, gold["k"], "sample", 2024,
                            gold["epsilon"])
    assert s.is_cuda and s.shape[0] == 3 and p.numel() == s.shape[1] + 1
    mine = Counter(",".join(map(str, sorted(c))) for c in s.t().cpu().tolist())
    ref = gold["counts"]
    assert set(mine) <= set(ref)
    n1, n2 = sum(ref.values()), sum(mine.values())
    stat = 0.0
    for key in ref:
        a, b = ref[key], mine.get(key, 0)
        stat += (a * np.sqrt(n2 / n1) - b * np.sqrt(n1 / n2)) ** 2 / (a + b)
    assert stat < chi2.ppf(0.999, len(ref) - 1), (stat, dict(mine), ref)
    f_ref, f_mine = gold["returned"] / gold["requested"], n2 / m
    se = np.sqrt(f_ref * (1 - f_ref) * (1 / gold["requested"] + 1 / m))
    assert abs(f_ref - f_mine) < 4 * se, (f_ref, f_mine)
