"""CPU: the oracle's sampling law against the exact reference law P_ref (oracle/exact_law.py), and -- where
oracle/_ref is built -- against the reference's own empirical frequencies. Chi-square at alpha = 0.001
(fixed seeds, so the test is deterministic; the threshold only documents the significance level)."""
from collections import Counter

import numpy as np
import pytest
from scipy.stats import chi2

from helpers import bidir

CASES = [
    ("triangle_tail", [(0, 1), (1, 2), (0, 2), (2, 3)], 4, 3),
    ("ring6", [(i, (i + 1) % 6) for i in range(6)], 6, 3),
    ("toluene_like", [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 0), (0, 6), (6, 7), (7, 8), (3, 9)], 10, 4),
    ("k5_plus_tail", [(a, b) for a in range(5) for b in range(a + 1, 5)] + [(4, 5), (5, 6)], 7, 4),
]


def _key(row):
    r = [int(x) for x in row]
    return frozenset(r) if -1 not in r else tuple(x for x in r if x >= 0)


@pytest.mark.parametrize("name,und,n,k", CASES)
def test_oracle_law_is_p_ref(name, und, n, k):
    from oracle import exact_law as xl, ugs_oracle as uo
    P = uo.Preproc(bidir(und), n, k)
    law = xl.p_ref(P.dump(), k)
    assert abs(sum(law.values()) - 1.0) < 1e-12
    M = 40000
    cnt = Counter(_key(r) for r in P.sample_nodes(M, seed=123))
    stat, df, extra = xl.chi_square(cnt, law, M)
    assert extra == 0
    assert stat < chi2.ppf(0.999, df), (name, stat, df)


def test_survey_probed_laws():
    """triangle+tail k=3 -> 0.8/0.1/0.1 ; 6-ring k=3 -> 1/7,2/7,1/7,1/7,1/7,1/7."""
    from oracle import exact_law as xl, ugs_oracle as uo
    law = xl.p_ref(uo.Preproc(bidir(CASES[0][1]), 4, 3).dump(), 3)
    assert law[frozenset({0, 1, 2})] == pytest.approx(0.8)
    assert law[frozenset({0, 2, 3})] == pytest.approx(0.1)
    assert law[frozenset({1, 2, 3})] == pytest.approx(0.1)
    law = xl.p_ref(uo.Preproc(bidir(CASES[1][1]), 6, 3).dump(), 3)
    assert sorted(law.values()) == pytest.approx([1 / 7] * 5 + [2 / 7])


def test_relaxation_levels_partial_rows():
    from oracle import exact_law as xl, ugs_oracle as uo
    P = uo.Preproc(bidir([(0, 1)]), 3, 3)  # Z = 0 -> level 1 (suffix_deg > 0), partial rows
    assert P.level == 1 and P.Z == 0.0
    law = xl.p_ref(P.dump(), 3)
    rows = P.sample_nodes(2000, seed=5)
    assert (rows[:, 2] == -1).all()
    cnt = Counter(_key(r) for r in rows)
    assert set(cnt) == set(law)
    P2 = uo.Preproc(np.zeros((2, 0), np.int64), 3, 2)  # no edges -> level 2, uniform over all vertices
    assert P2.level == 2
    c = Counter(int(r[0]) for r in P2.sample_nodes(3000, seed=1))
    assert set(c) == {0, 1, 2} and min(c.values()) > 850


def test_reference_frequencies_follow_the_same_law():
    """The REAL reference (oracle/_ref) sampled with distinct seeds follows P_ref too (its i-th sample depends
    on the sample index only, so one call with m = M gives M distinct streams)."""
    from oracle import exact_law as xl, ref_loader, ugs_oracle as uo
    if not ref_loader.available("ugs_sampler", "nopin"):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    import torch
    mod = ref_loader.load("ugs_sampler", "nopin")
    for name, und, n, k in CASES[:3]:
        ei = bidir(und)
        law = xl.p_ref(uo.Preproc(ei, n, k).dump(), k)
        h = mod.create_preproc(torch.from_numpy(ei), n, k)
        M = 30000
        rows = mod.sample(h, M, k, "local", 0, 99)[0].numpy()
        mod.destroy_preproc(h)
        stat, df, extra = xl.chi_square(Counter(_key(r) for r in rows), law, M)
        assert extra == 0 and stat < chi2.ppf(0.999, df), (name, stat, df)


@pytest.mark.parametrize("name,und,n,k", CASES[:3] + [("toluene_k5", CASES[2][1], 10, 5)])
def test_uniform_mode_oracle_is_uniform(name, und, n, k):
    """oracle/uni_oracle.c (SURVEY A.8 accept/reject mode): chi-square against the uniform law over connected
    induced k-subgraphs -- the property the reference's own UGS does not have."""
    from oracle import exact_law as xl, ugs_oracle as uo
    ei = bidir(und)
    rows = xl.rows_from_dump(uo.Preproc(ei, n, k).dump())
    subs = xl.connected_subsets(rows, k)
    law = {s: 1.0 / len(subs) for s in subs}
    M = 30000
    nodes, trials = uo.UniformTables(ei, n, k).sample_nodes(M, seed=3)
    assert (nodes >= 0).all() and trials >= M
    stat, df, extra = xl.chi_square(Counter(_key(r) for r in nodes), law, M)
    assert extra == 0 and stat < chi2.ppf(0.999, df), (name, stat, df)
    # the reference law on the same graph is NOT uniform
    stat_ref = xl.chi_square(Counter(_key(r) for r in uo.Preproc(ei, n, k).sample_nodes(M, seed=3)), law, M)[0]
    assert stat_ref > chi2.ppf(0.999, df)
