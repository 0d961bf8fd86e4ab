"""Test infrastructure only (oracle/): restatement of the reference's WL graphlet hash without NetworkX.

compute_wl_hash (gps/utils/wl_vocab.py:21-67) builds an nx.Graph from the sample's local edge list, sets
attr = str(degree) and calls nx.weisfeiler_lehman_graph_hash(G, node_attr='attr', iterations). This file restates the
published NetworkX algorithm (networkx/algorithms/graph_hashing.py, node_attr branch; networkx is a pip dependency of
the reference, not vendored) with hashlib.blake2b.
Parity status: PINNED -- tests/golden/wl_golden.json holds hashes produced by the reference's own compute_wl_hash
(imported from /root/reference with the installed networkx 3.6.1; generator tests/golden/make_golden_wl.py), and
tests/test_wl_oracle.py also compares against networkx directly where it is importable.
"""
from __future__ import annotations

from collections import Counter
from hashlib import blake2b


def _h(label: str) -> str:
    return blake2b(label.encode("ascii"), digest_size=16).hexdigest()


def node_attr(row_bytes: bytes) -> str:
    """wl_vocab.py:45-49: hashlib.md5(node_features[i].cpu().numpy().tobytes()).hexdigest()[:8]."""
    from hashlib import md5
    return md5(row_bytes).hexdigest()[:8]


def compute_wl_hash(edges, num_nodes: int, num_iterations: int = 3, attrs=None) -> str:
    """edges: iterable of (u, v) local ids. nx.Graph semantics: undirected, parallel edges merge, a self loop makes
    the node its own neighbour once and adds 2 to its degree. attrs: per-node attribute strings (use_node_features)
    or None for the degree attribute."""
    if num_iterations <= 0:
        raise ValueError("The WL algorithm requires that `iterations` be positive")
    nbrs = [set() for _ in range(num_nodes)]
    for u, v in edges:
        u, v = int(u), int(v)
        nbrs[u].add(v)
        nbrs[v].add(u)
    labels = list(attrs) if attrs is not None else [str(len(nb - {i}) + (2 if i in nb else 0)) for i, nb in enumerate(nbrs)]
    counts = []
    for _ in range(num_iterations):
        labels = [_h(labels[i] + "".join(sorted(labels[j] for j in nbrs[i]))) for i in range(num_nodes)]
        counts.extend(sorted(Counter(labels).items(), key=lambda x: x[0]))
    return _h(str(tuple(counts)))


def batch_hashes(nodes, edge_index, edge_ptr, num_iterations: int = 3, x=None):
    """Per sample: extract_subgraph_from_batch (wl_vocab.py:70-108) + compute_wl_hash; None when no valid node.
    x: numpy feature matrix of the batch (use_node_features=True) or None."""
    out = []
    for s in range(nodes.shape[0]):
        valid = nodes[s][nodes[s] >= 0]
        nn = len(valid)
        if nn == 0:
            out.append(None)
            continue
        a, b = int(edge_ptr[s]), int(edge_ptr[s + 1])
        attrs = [node_attr(x[int(v)].tobytes()) for v in valid] if x is not None else None
        out.append(compute_wl_hash(zip(edge_index[0, a:b], edge_index[1, a:b]), nn, num_iterations, attrs))
    return out
