"""Seeded synthetic graph batches in the shapes BASELINE.json names.

Graphs are a random spanning tree plus extra uniform edges up to a target average degree, emitted as
bidirectional, coalesced (sorted) columns and concatenated graph-contiguously, i.e. the layout
`torch_geometric.data.Batch.from_data_list` produces. Pure numpy; no reference code involved.
"""
from __future__ import annotations

import numpy as np

WORKLOADS = {
    # name: (default G, k, m)
    "mutag": (188, 6, 64),
    "proteins": (1113, 6, 100),
    "qm9": (10000, 4, 10),
    "zinc": (131072, 8, 64),
    "powerlaw": (1, 6, 100000),
}


def _sizes(kind: str, G: int, rng: np.random.Generator) -> np.ndarray:
    if kind == "mutag":
        return rng.integers(10, 29, size=G)
    if kind == "proteins":
        # log-normal clipped to [4, 620], mean ~39 (PROTEINS: mean 39.1, max 620)
        n = np.exp(rng.normal(3.25, 0.85, size=G))
        return np.clip(np.rint(n), 4, 620).astype(np.int64)
    if kind == "qm9":
        return rng.integers(3, 30, size=G)
    if kind == "zinc":
        return rng.integers(9, 39, size=G)
    raise ValueError(kind)


_AVG_DEG = {"mutag": 2.2, "proteins": 3.7, "qm9": 2.1, "zinc": 2.15}


def tree_plus_edges(sizes: np.ndarray, avg_deg: float, rng: np.random.Generator, max_extra=None):
    """Vectorised: one random recursive tree per graph + extra random intra-graph edges (simple graph).
    Returns (edge_index[2,E] int64, ptr[G+1] int64) with bidirectional sorted columns."""
    sizes = np.asarray(sizes, dtype=np.int64)
    G = sizes.shape[0]
    ptr = np.zeros(G + 1, np.int64)
    np.cumsum(sizes, out=ptr[1:])
    NT = int(ptr[-1])
    gid = np.repeat(np.arange(G, dtype=np.int64), sizes)
    local = np.arange(NT, dtype=np.int64) - ptr[gid]
    # tree: node i>=1 attaches to a uniform earlier node of its graph
    nz = local > 0
    child = np.nonzero(nz)[0].astype(np.int64)
    parent = ptr[gid[child]] + np.floor(rng.random(child.shape[0]) * local[child]).astype(np.int64)
    # extra edges
    target_und = np.rint(sizes * avg_deg / 2.0).astype(np.int64)
    extra = np.maximum(target_und - (sizes - 1), 0)
    extra = np.where(sizes >= 3, extra, 0)
    if max_extra is not None:
        extra = np.minimum(extra, max_extra)
    eg = np.repeat(np.arange(G, dtype=np.int64), extra)
    a = np.floor(rng.random(eg.shape[0]) * sizes[eg]).astype(np.int64)
    b = np.floor(rng.random(eg.shape[0]) * (sizes[eg] - 1)).astype(np.int64)
    b = np.where(b >= a, b + 1, b)  # b != a
    ua = np.concatenate([child, ptr[eg] + a])
    ub = np.concatenate([parent, ptr[eg] + b])
    lo = np.minimum(ua, ub)
    hi = np.maximum(ua, ub)
    key = np.unique(lo * np.int64(NT + 1) + hi)  # simple graph: drop duplicate undirected edges
    lo = key // np.int64(NT + 1)
    hi = key % np.int64(NT + 1)
    src = np.concatenate([lo, hi])
    dst = np.concatenate([hi, lo])
    o = np.lexsort((dst, src))
    return np.stack([src[o], dst[o]]).astype(np.int64), ptr


def make_batch(kind: str, G: int, seed: int = 1234):
    """(edge_index, ptr) for one of the named molecule/protein-shaped workloads."""
    rng = np.random.default_rng(seed)
    sizes = _sizes(kind, G, rng)
    return tree_plus_edges(sizes, _AVG_DEG[kind], rng, max_extra=2 if kind == "qm9" else None)


def barabasi_albert(n: int, m0: int = 3, seed: int = 1234):
    """One Barabasi-Albert graph (n nodes, m0 edges per new node); bidirectional sorted columns."""
    rng = np.random.default_rng(seed)
    targets = list(range(m0))
    repeated: list[int] = []
    src, dst = [], []
    for v in range(m0, n):
        for t in targets:
            src.append(v)
            dst.append(t)
        repeated.extend(targets)
        repeated.extend([v] * m0)
        chosen: set[int] = set()
        while len(chosen) < m0:
            chosen.add(repeated[int(rng.integers(0, len(repeated)))])
        targets = list(chosen)
    s = np.asarray(src, np.int64)
    d = np.asarray(dst, np.int64)
    a = np.concatenate([s, d])
    b = np.concatenate([d, s])
    o = np.lexsort((b, a))
    return np.stack([a[o], b[o]]), np.asarray([0, n], np.int64)


def concat_batches(batches):
    """Concatenate (edge_index, ptr) pairs into one graph-contiguous batch."""
    eis, ptrs, off = [], [np.zeros(1, np.int64)], 0
    for ei, p in batches:
        eis.append(ei + off)
        ptrs.append(p[1:] + off)
        off += int(p[-1])
    return np.concatenate(eis, axis=1), np.concatenate(ptrs)
