"""CPU: the epsilon-uniform oracle (oracle/eps_oracle.c) -- deterministic extraction against golden vectors from the
real reference, law against the exact enumerator (oracle/exact_law.py::p_epsilon) and the reference's frequencies."""
import os
from collections import Counter

import numpy as np
import pytest
from scipy.stats import chi2

from helpers import bidir

GOLD = os.path.join(os.path.dirname(__file__), "golden", "eps_golden.npz")
LAW_CASES = [
    ("triangle_tail", [(0, 1), (1, 2), (0, 2), (2, 3)], 4, 3, 0.1),
    ("ring6", [(i, (i + 1) % 6) for i in range(6)], 6, 3, 0.5),
    ("toluene_like", [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 0), (0, 6), (6, 7), (7, 8), (3, 9)], 10, 4, 0.01),
]


def _key(r):
    return frozenset(int(x) for x in r) if r[0] >= 0 else ()


@pytest.mark.parametrize("mode", ["sample", "global"])
def test_eps_extraction_matches_reference_golden(mode):
    from oracle import eps_oracle as eo
    gold = np.load(GOLD)
    for name in [str(x) for x in gold["names"]]:
        ei, ptr = gold[f"{name}/edge_index"], gold[f"{name}/ptr"]
        k, m = (int(x) for x in gold[f"{name}/km"])
        nodes = gold[f"{name}/{mode}/nodes"]
        out = eo.sample_batch(ei, ptr, m, k, mode, 42, 0.1, given_nodes=nodes)
        for a, nm in zip(out, ("nodes", "edge_index_out", "edge_ptr", "sample_ptr", "edge_src")):
            assert np.array_equal(a, gold[f"{name}/{mode}/{nm}"]), (name, nm)
        # sorted rows, one entry per column inside the sample (no 4x multiplicity)
        ok = nodes[nodes[:, 0] >= 0]
        assert (np.diff(ok, axis=1) > 0).all()


@pytest.mark.parametrize("name,und,n,k,eps", LAW_CASES)
def test_eps_oracle_law(name, und, n, k, eps):
    from oracle import eps_oracle as eo, exact_law as xl, ugs_oracle as uo
    ei = bidir(und)
    rows = xl.rows_from_dump(uo.Preproc(ei, n, k).dump())
    law = {kk: v for kk, v in xl.p_epsilon(rows, k, eps).items() if v > 1e-12}
    M = 40000
    nodes = eo.sample_batch(ei, np.array([0, n]), M, k, "sample", 5, eps)[0]
    stat, df, extra = xl.chi_square(Counter(_key(r) for r in nodes), law, M)
    assert extra == 0 and stat < chi2.ppf(0.999, df), (name, stat, df)


def test_reference_eps_frequencies_follow_the_enumerated_law():
    from oracle import exact_law as xl, ref_loader, ugs_oracle as uo
    if not ref_loader.available("epsilon_uniform_sampler", "nopin"):
        pytest.skip("oracle/_ref not built")
    import torch
    torch.set_num_threads(1)
    mod = ref_loader.load("epsilon_uniform_sampler", "nopin")
    for name, und, n, k, eps in LAW_CASES[:2]:
        ei = bidir(und)
        rows = xl.rows_from_dump(uo.Preproc(ei, n, k).dump())
        law = {kk: v for kk, v in xl.p_epsilon(rows, k, eps).items() if v > 1e-12}
        M = 30000
        rn = mod.sample_batch(torch.from_numpy(ei), torch.tensor([0, n]), M, k, "sample", 7, eps)[0].numpy()
        stat, df, extra = xl.chi_square(Counter(_key(r) for r in rn), law, M)
        assert extra == 0 and stat < chi2.ppf(0.999, df), (name, stat, df)


def test_eps_oracle_edge_cases():
    from oracle import eps_oracle as eo
    # n < k and disconnected graphs: all -1 rows, sample_ptr still advances (A.9)
    ei = bidir([(2, 3), (5, 6)])
    ptr = np.array([0, 2, 4, 8])
    nodes, eidx, eptr, sptr, esrc = eo.sample_batch(ei, ptr, 3, 3, "sample", 1, 0.5)
    assert (nodes == -1).all() and eidx.shape == (2, 0) and sptr.tolist() == [0, 3, 6, 9] and eptr.max() == 0
    with pytest.raises(RuntimeError, match="epsilon must be in"):
        eo.sample_batch(ei, ptr, 3, 3, "sample", 1, 0.0)
