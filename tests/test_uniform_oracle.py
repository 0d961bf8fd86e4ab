"""CPU: the exact-uniform accept/reject mode's oracle (oracle/uni_oracle.c). The adjacency-list statement (any graph
size; the CUDA general path is compared with it) reproduces the 64-bit-mask statement bit for bit on small graphs and
passes a chi-square against the UNIFORM law on a PROTEINS-shaped 100-node graph."""
from collections import Counter

import numpy as np
import pytest
from scipy.stats import chi2

from helpers import random_batch, per_graph_columns


@pytest.mark.parametrize("seed,k", [(1, 3), (2, 4), (3, 5), (4, 6)])
def test_uniform_general_oracle_equals_mask_oracle(seed, k):
    from oracle import ugs_oracle as uo
    rng = np.random.default_rng(seed)
    ei, ptr = random_batch(rng, 12, 4, 40, deg=2.5, weird=True)
    for g in range(len(ptr) - 1):
        n = int(ptr[g + 1] - ptr[g])
        if n < k:
            continue
        lc, _ = per_graph_columns(ei, ptr, g)
        a, b = uo.UniformTables(lc, n, k, general=False), uo.UniformTables(lc, n, k, general=True)
        da, db = a.dump(), b.dump()
        for key in ("order", "dsuf", "b", "cum"):
            assert np.array_equal(da[key], db[key]), (g, key)
        assert da["Z"] == db["Z"] and da["C"] == db["C"]
        ra, ta = a.sample_nodes(40, seed, g)
        rb, tb = b.sample_nodes(40, seed, g)
        assert np.array_equal(ra, rb) and ta == tb, g


def uniform_test_graph(n=100, seed=5):
    """PROTEINS-shaped graph above the bitmask path's 64 vertices: random tree + extra edges, average degree ~3.7."""
    from ugs_b200 import synth
    rng = np.random.default_rng(seed)
    ei, ptr = synth.tree_plus_edges(np.array([n]), 3.7, rng)
    return ei, ptr


def test_uniform_general_oracle_passes_chi_square_against_uniform():
    from oracle import exact_law as xl, ugs_oracle as uo
    ei, ptr = uniform_test_graph()
    n, k, M = int(ptr[1]), 4, 120000
    rows = xl.rows_from_dump(uo.Preproc(ei, n, k).dump())
    # connected 4-subsets by growing from every vertex (C(100,4) combinations would take minutes in Python)
    adj = [set(r) - {v} for v, r in enumerate(rows)]
    subs = set()
    frontier = {frozenset([v]) for v in range(n)}
    for _ in range(k - 1):
        frontier = {s | {w} for s in frontier for u in s for w in adj[u] if w not in s}
    subs = frontier
    T = uo.UniformTables(ei, n, k)
    assert T.general
    nodes, trials = T.sample_nodes(M, 77, 0)
    assert (nodes >= 0).all()
    cnt = Counter(frozenset(r) for r in nodes.tolist())
    law = {s: 1.0 / len(subs) for s in subs}
    stat, df, extra = xl.chi_square(cnt, law, M)
    assert extra == 0 and stat < chi2.ppf(0.999, df), (stat, df, len(subs), trials / M)
