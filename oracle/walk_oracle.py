"""Test infrastructure only (oracle/): ctypes wrappers around oracle/walk_oracle.c -- the rwr_sampler and
uniform_sampler restatements. Edge extraction reuses the already pinned extractors: the UGS one (rwr emits UGS-style
entries with edge_src = -1, rwr_sampler.cpp:217-256) and the column-scan one of eps_oracle (uniform_sampler.cpp:216-245)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import eps_oracle as _eo
from . import ugs_oracle as _uo

_ready = False


def _lib():
    global _ready
    L = _uo.lib()
    if not _ready:
        vp, i64, ci, u64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint64
        L.rwr_oracle_sample_nodes.argtypes = [vp, vp, ci, ci, ci, u64, u64, ctypes.c_double, vp]
        L.enum_oracle_list.restype = i64
        L.enum_oracle_list.argtypes = [vp, vp, ci, ci, vp, i64]
        L.enum_oracle_sample_nodes.argtypes = [vp, i64, ci, ci, u64, u64, vp]
        _ready = True
    return L


def _graph_csr(ei, ptr, g, k):
    lo, hi = int(ptr[g]), int(ptr[g + 1])
    n = hi - lo
    msk = (ei[0] >= lo) & (ei[0] < hi) & (ei[1] >= lo) & (ei[1] < hi)
    d = _uo.Preproc(ei[:, msk] - lo, n, max(k, 1)).dump()
    return n, lo, np.ascontiguousarray(d["indptr"]), np.ascontiguousarray(d["indices"])


def rwr_sample_batch(edge_index, ptr, m_per_graph: int, k: int, mode: str = "sample", seed: int = 42,
                     p_restart: float = 0.2, graph_index_base: int = 0):
    """Oracle counterpart of rwr_sampler.sample_batch (rwr_sampler.cpp:72-295)."""
    if k < 1:
        raise RuntimeError("k must be >= 1")
    if not (0.0 <= p_restart <= 1.0):
        raise RuntimeError("p_restart in [0,1]")
    ei = np.asarray(edge_index, dtype=np.int64)
    ptr = np.asarray(ptr, dtype=np.int64)
    G, m = len(ptr) - 1, int(m_per_graph)
    nodes = np.full((G * m, k), -1, np.int64)
    L, p = _lib(), _uo._p
    for g in range(G):
        n = int(ptr[g + 1] - ptr[g])
        if n <= 0 or n < k or m == 0:
            continue
        n, lo, indptr, indices = _graph_csr(ei, ptr, g, k)
        loc = np.full((m, k), -1, np.int32)
        L.rwr_oracle_sample_nodes(p(indptr), p(indices), n, m, k, seed & (2 ** 64 - 1), graph_index_base + g,
                                  float(p_restart), p(loc))
        nodes[g * m:(g + 1) * m] = np.where(loc >= 0, loc.astype(np.int64) + lo, -1)
    out = _uo.sample_batch(ei, ptr, m, k, "sample" if mode == "sample" else "global", seed, given_nodes=nodes)
    out[4][:] = -1  # src_edge "not tracked" (:243)
    return out


def enumerate_connected_any(edge_index, num_nodes: int, k: int):
    """The same enumeration for a graph of any size, in pure Python with ints as bitsets: list of vertex tuples in
    the order they joined the set (roots ascending; always the smallest candidate first -- the order of
    walk_oracle.c::esu_rec and of the CUDA kernels). Small cases only."""
    ei = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    n, lo, indptr, indices = _graph_csr(ei, np.array([0, num_nodes], np.int64), 0, k)
    adj = [0] * n
    for v in range(n):
        for u in indices[indptr[v]:indptr[v + 1]]:
            if int(u) != v:
                adj[v] |= 1 << int(u)
    out = []

    def rec(S, ext, NS, gt):
        if len(S) == k:
            out.append(tuple(S))
            return
        while ext:
            low = ext & -ext
            ext ^= low
            w = low.bit_length() - 1
            rec(S + [w], ext | (adj[w] & gt & ~NS), NS | adj[w], gt)

    full = (1 << n) - 1
    for v in range(n):
        gt = full & ~((1 << (v + 1)) - 1)
        rec([v], adj[v] & gt, adj[v] | (1 << v), gt)
    return out


def enumerate_connected(edge_index, num_nodes: int, k: int):
    """All connected induced k-subsets of one graph with <= 64 nodes (local ids) as a uint64 mask array, in the
    oracle's list order."""
    ei = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    n, lo, indptr, indices = _graph_csr(ei, np.array([0, num_nodes], np.int64), 0, k)
    L, p = _lib(), _uo._p
    c = int(L.enum_oracle_list(p(indptr), p(indices), n, k, None, 0))
    if c < 0:
        raise RuntimeError("enumerate_connected: more than 64 nodes, use enumerate_connected_any")
    lst = np.zeros(max(c, 1), np.uint64)
    L.enum_oracle_list(p(indptr), p(indices), n, k, p(lst), c)
    return lst[:c]


def uniform_sample_batch(edge_index, ptr, m_per_graph: int, k: int, mode: str = "sample", seed: int = 42,
                         graph_index_base: int = 0):
    """Oracle counterpart of uniform_sampler.sample_batch (uniform_sampler.cpp:86-282)."""
    ei = np.asarray(edge_index, dtype=np.int64)
    ptr = np.asarray(ptr, dtype=np.int64)
    G, m = len(ptr) - 1, int(m_per_graph)
    nodes = np.full((G * m, k), -1, np.int64)
    L, p = _lib(), _uo._p
    for g in range(G):
        lo, n = int(ptr[g]), int(ptr[g + 1] - ptr[g])
        if n <= 0 or n < k or m == 0:
            continue
        msk = (ei[0] >= lo) & (ei[0] < lo + n) & (ei[1] >= lo) & (ei[1] < lo + n)
        loc = np.full((m, k), -1, np.int32)
        if n <= 64:
            lst = enumerate_connected(ei[:, msk] - lo, n, k)
            L.enum_oracle_sample_nodes(p(lst) if len(lst) else None, len(lst), m, k, seed & (2 ** 64 - 1),
                                       graph_index_base + g, p(loc))
        else:  # same pick rule on the Python list (uniform_sampler.cpp:196-210; rows ascending)
            tl = enumerate_connected_any(ei[:, msk] - lo, n, k)
            w = np.zeros(4, np.uint32)
            for i in range(m if tl else 0):
                L.oracle_philox(i, 0, (graph_index_base + g) & 0xFFFFFFFF, 0xE0000000, seed & 0xFFFFFFFF,
                                (seed >> 32) & 0xFFFFFFFF, p(w))
                idx = (((int(w[0]) << 32) | int(w[1])) * len(tl)) >> 64
                loc[i] = sorted(tl[idx])
        nodes[g * m:(g + 1) * m] = np.where(loc >= 0, loc.astype(np.int64) + lo, -1)
    return _eo.sample_batch(ei, ptr, m, k, mode, seed, 0.5, given_nodes=nodes)
