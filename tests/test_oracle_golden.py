"""CPU: the C oracle (oracle/ugs_oracle.c) against golden vectors generated from the REAL reference
(tests/golden/make_golden.py -> ugs_golden.npz). Bit-exact for every deterministic stage."""
import os

import numpy as np
import pytest

from helpers import per_graph_columns

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ugs_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _names(gold):
    return [str(x) for x in gold["names"]]


def test_golden_has_all_cases(gold):
    assert len(_names(gold)) == 13


def test_oracle_tables_match_reference_dump(gold):
    from oracle import ugs_oracle as uo
    checked = 0
    for name in _names(gold):
        ei, ptr = gold[f"{name}/edge_index"], gold[f"{name}/ptr"]
        k = int(gold[f"{name}/km"][0])
        for g in range(len(ptr) - 1):
            n = int(ptr[g + 1] - ptr[g])
            if n <= 0:
                continue
            lcols, _ = per_graph_columns(ei, ptr, g)
            P = uo.Preproc(lcols, n, k)
            d = P.dump()
            for key in ("indptr", "indices", "edge_col", "order", "index_of", "suffix_deg", "bucket_b"):
                assert np.array_equal(d[key], gold[f"{name}/g{g}/{key}"]), (name, g, key)
            Z = float(gold[f"{name}/g{g}/Z"][0])
            assert d["Z"] == Z, (name, g)
            if Z > 0:  # the reference builds its alias table only when Z > 0 (preproc.cpp:241-243)
                assert np.array_equal(d["alias_prob"], gold[f"{name}/g{g}/alias_prob"]), (name, g)
                assert np.array_equal(d["alias_alias"], gold[f"{name}/g{g}/alias_alias"]), (name, g)
            info = P.info()
            gi = gold[f"{name}/g{g}/info"]
            assert [info["num_nodes"], info["num_edges_stored"], info["bucket_count_nonzero"]] == gi.tolist()
            # the peeling order is a stable sort by (multidegree, id)
            deg = np.diff(d["indptr"])
            assert np.array_equal(d["order"], np.lexsort((np.arange(n), deg)))
            checked += 1
    assert checked >= 50


@pytest.mark.parametrize("mode", ["sample", "graph", "global"])
def test_oracle_extraction_matches_reference_for_its_own_samples(gold, mode):
    """Feed the reference's node rows to the oracle's extractor: edge_index / edge_ptr / edge_src /
    sample_ptr must come out exactly as the reference emitted them (A.7, all three modes)."""
    from oracle import ugs_oracle as uo
    for name in _names(gold):
        ei, ptr = gold[f"{name}/edge_index"], gold[f"{name}/ptr"]
        k, m = (int(x) for x in gold[f"{name}/km"])
        nodes = gold[f"{name}/{mode}/nodes"]
        out = uo.sample_batch(ei, ptr, m, k, mode, 42, given_nodes=nodes)
        assert np.array_equal(out[0], nodes), name
        assert np.array_equal(out[1], gold[f"{name}/{mode}/edge_index_out"]), name
        assert np.array_equal(out[2], gold[f"{name}/{mode}/edge_ptr"]), name
        assert np.array_equal(out[3], gold[f"{name}/{mode}/sample_ptr"]), name
        assert np.array_equal(out[4], gold[f"{name}/{mode}/edge_src"]), name


def test_survey_golden_triangle_tail(gold):
    """order=[3,0,1,2] suffix_deg=[2,4,2,0] bucket_b=[4,16,0,0] Z=20."""
    assert gold["triangle_tail/g0/order"].tolist() == [3, 0, 1, 2]
    assert gold["triangle_tail/g0/suffix_deg"].tolist() == [2, 4, 2, 0]
    assert gold["triangle_tail/g0/bucket_b"].tolist() == [4.0, 16.0, 0.0, 0.0]
    assert float(gold["triangle_tail/g0/Z"][0]) == 20.0
    assert gold["triangle_tail/g0/alias_prob"].tolist() == [0.8, 1.0, 0.0, 0.0]
    assert int(gold["triangle_tail/g0/info"][1]) == 16  # 4 undirected edges -> 16 CSR entries (fact 3)


def test_reference_edge_case_outputs(gold):
    """SURVEY.md section 4 probed golden outputs (seed 42, mode "sample")."""
    # triangle + isolated node, k=3: 12 entries per sample
    ep = gold["triangle_isolated/sample/edge_ptr"]
    assert (np.diff(ep) == 12).all()
    # 5-chain one-direction, k=4: 6 entries per sample, samples [0,1,2,3] or [4,3,2,1]
    assert (np.diff(gold["chain5_onedir/sample/edge_ptr"]) == 6).all()
    rows = {tuple(r) for r in gold["chain5_onedir/sample/nodes"].tolist()}
    assert rows <= {(0, 1, 2, 3), (4, 3, 2, 1)}
    # single edge n=3 k=3: partial rows, no edges
    nodes = gold["single_edge_n3_k3/sample/nodes"]
    assert (nodes[:, 2] == -1).all() and gold["single_edge_n3_k3/sample/edge_ptr"].max() == 0
    # n<k graph followed by a 4-cycle: first m rows all -1, sample_ptr = arange*m
    nodes = gold["small_then_cycle/sample/nodes"]
    assert (nodes[:4] == -1).all() and gold["small_then_cycle/sample/sample_ptr"].tolist() == [0, 4, 8]
