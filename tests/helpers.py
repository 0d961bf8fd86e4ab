"""Shared test inputs: random graph batches with every edge case the reference handles."""
import numpy as np


def bidir(und):
    c = []
    for a, b in und:
        c += [(a, b), (b, a)]
    return np.array(c, np.int64).reshape(-1, 2).T.copy()


def random_batch(rng, G, n_lo=1, n_hi=40, deg=3.0, weird=True, big=()):
    """Graph-contiguous batch with arbitrary columns: multi-edges, self loops, one-direction edges,
    cross-graph and out-of-range columns (dropped by the reference), empty graphs."""
    sizes = [int(rng.integers(n_lo, n_hi + 1)) for _ in range(G)]
    for i, nb in big:
        sizes[i] = nb
    if weird and G > 2:
        sizes[int(rng.integers(0, G))] = 0
    ptr = np.zeros(G + 1, np.int64)
    ptr[1:] = np.cumsum(sizes)
    cols = []
    for g, n in enumerate(sizes):
        lo = int(ptr[g])
        if n == 0:
            continue
        E = int(rng.integers(0, max(1, int(n * deg)) + 1))
        for _ in range(E):
            a = int(rng.integers(0, n))
            b = int(rng.integers(0, n))
            r = rng.random()
            if weird and r < 0.03:
                b = a  # self loop
            cols.append((lo + a, lo + b))
            if r > 0.3:
                cols.append((lo + b, lo + a))  # mostly bidirectional, sometimes one-direction
            if weird and r > 0.95:
                cols.append((lo + a, lo + b))  # duplicate column
        if weird and rng.random() < 0.2 and ptr[-1] > n:
            cols.append((lo, int((lo + n) % max(int(ptr[-1]), 1))))  # cross-graph / maybe in-graph column
        if weird and rng.random() < 0.1:
            cols.append((lo, int(ptr[-1]) + 3))  # out of range
    ei = np.array(cols, np.int64).reshape(-1, 2).T.copy() if cols else np.zeros((2, 0), np.int64)
    return ei, ptr


def per_graph_columns(ei, ptr, g):
    """A.1: the reference's slice_and_renumber (ugs_sampler_batch_extension.cpp:41-75) in numpy."""
    lo, hi = int(ptr[g]), int(ptr[g + 1])
    m = (ei[0] >= lo) & (ei[0] < hi) & (ei[1] >= lo) & (ei[1] < hi)
    idx = np.nonzero(m)[0]
    return ei[:, idx] - lo, idx
