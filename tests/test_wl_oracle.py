"""CPU: the WL-hash oracle (oracle/wl_oracle.py, a NetworkX-free restatement) against hashes produced by the
reference's own compute_wl_hash (tests/golden/wl_golden.json) and against networkx itself where it is installed."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "wl_golden.json")


def test_wl_oracle_matches_reference_golden():
    from oracle import wl_oracle as wo
    rows = json.load(open(GOLD))
    assert len(rows) >= 70
    for r in rows:
        for it in (1, 2, 3):
            assert wo.compute_wl_hash(r["edges"], r["n"], it) == r["hash"][str(it)], (r["name"], it)


def test_wl_oracle_with_node_features_matches_reference_golden():
    """use_node_features=True: attribute = md5(feature row bytes).hexdigest()[:8] (wl_vocab.py:45-49)."""
    from oracle import wl_oracle as wo
    rows = [r for r in json.load(open(GOLD)) if "feat" in r]
    assert len(rows) >= 20
    for r in rows:
        for kk, f in r["feat"].items():
            x = np.array(f["values"], dtype=f["dtype"]).reshape(r["n"], -1)
            attrs = [wo.node_attr(x[i].tobytes()) for i in range(r["n"])]
            assert wo.compute_wl_hash(r["edges"], r["n"], 3, attrs) == f["hash"], (r["name"], kk)


def test_wl_oracle_matches_networkx():
    nx = pytest.importorskip("networkx")
    from oracle import wl_oracle as wo
    rng = np.random.default_rng(5)
    for _ in range(40):
        n = int(rng.integers(1, 10))
        edges = [(int(rng.integers(0, n)), int(rng.integers(0, n))) for _ in range(int(rng.integers(0, 2 * n + 1)))]
        G = nx.Graph()
        G.add_nodes_from(range(n))
        G.add_edges_from(edges)
        for i in range(n):
            G.nodes[i]["attr"] = str(G.degree(i))
        for it in (1, 3, 4):
            assert wo.compute_wl_hash(edges, n, it) == nx.weisfeiler_lehman_graph_hash(G, node_attr="attr", iterations=it)
    with pytest.raises(ValueError):
        wo.compute_wl_hash([], 2, 0)
