"""Test infrastructure only (oracle/): ctypes wrapper around oracle/eps_oracle.c (epsilon_uniform_sampler)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import ugs_oracle as _uo

MODES = {"sample": 0, "graph": 1, "global": 2}
_ready = False


def _lib():
    global _ready
    L = _uo.lib()
    if not _ready:
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        L.eps_oracle_sample_batch.restype = vp
        L.eps_oracle_sample_batch.argtypes = [vp, vp, i64, vp, i64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_uint64, ctypes.c_uint64, ctypes.c_double, vp]
        L.eps_oracle_batch_total_edges.restype = i64
        L.eps_oracle_batch_total_edges.argtypes = [vp]
        L.eps_oracle_batch_copy.argtypes = [vp] * 6
        L.eps_oracle_batch_free.argtypes = [vp]
        _ready = True
    return L


def sample_batch(edge_index, ptr, m_per_graph: int, k: int, mode: str = "sample", seed: int = 42,
                 epsilon: float = 0.1, graph_index_base: int = 0, given_nodes=None):
    """Oracle counterpart of epsilon_uniform_sampler.sample_batch (epsilon_uniform_sampler.cpp:122-362)."""
    if not (epsilon > 0.0 and epsilon <= 1.0):
        raise RuntimeError("epsilon must be in (0, 1]")
    _, r0, r1 = _uo._rows(edge_index)
    ptr = np.ascontiguousarray(np.asarray(ptr, dtype=np.int64))
    G = ptr.shape[0] - 1
    gn = None if given_nodes is None else np.ascontiguousarray(np.asarray(given_nodes, dtype=np.int64))
    L = _lib()
    p = _uo._p
    h = L.eps_oracle_sample_batch(p(r0), p(r1), r0.shape[0], p(ptr), G, int(m_per_graph), int(k), MODES.get(mode, 2),
                                  seed & (2 ** 64 - 1), graph_index_base, float(epsilon), p(gn) if gn is not None else None)
    try:
        E = int(L.eps_oracle_batch_total_edges(h))
        nodes = np.zeros((G * m_per_graph, k), np.int64)
        ei = np.zeros((2, E), np.int64)
        eptr = np.zeros(G * m_per_graph + 1, np.int64)
        sptr = np.zeros(G + 1, np.int64)
        esrc = np.zeros(E, np.int64)
        L.eps_oracle_batch_copy(h, p(nodes), p(ei), p(eptr), p(sptr), p(esrc))
    finally:
        L.eps_oracle_batch_free(h)
    return nodes, ei, eptr, sptr, esrc
