"""Test infrastructure only (oracle/): ctypes wrapper around oracle/apx_oracle.c (reference apx_ugs_sampler)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import ugs_oracle as _uo

_ready = False


def _lib():
    global _ready
    L = _uo.lib()
    if not _ready:
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        L.apx_oracle_create.restype = vp
        L.apx_oracle_create.argtypes = [vp, vp, i64, i64, ctypes.c_int, ctypes.c_uint64, ctypes.c_double]
        L.apx_oracle_free.argtypes = [vp]
        L.apx_oracle_n.argtypes = [vp]
        L.apx_oracle_nnz.restype = i64
        L.apx_oracle_nnz.argtypes = [vp]
        L.apx_oracle_Z.restype = ctypes.c_double
        L.apx_oracle_Z.argtypes = [vp]
        L.apx_oracle_dump.argtypes = [vp] * 6
        L.apx_oracle_sample.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp]
        _ready = True
    return L


class ApxGraph:
    """Graph + approximate degree-dominating order + estimates (apx_ugs_sampler.cpp:10-169)."""

    def __init__(self, edge_index, ptr, k: int, seed: int = 42, epsilon: float = 0.1):
        _, r0, r1 = _uo._rows(edge_index)
        ptr = np.asarray(ptr, dtype=np.int64)
        self.k = int(k)
        self._h = _lib().apx_oracle_create(_uo._p(r0), _uo._p(r1), int(ptr[0]), int(ptr[1]), self.k,
                                           seed & (2 ** 64 - 1), float(epsilon))
        self.n = int(_lib().apx_oracle_n(self._h))
        self.Z = float(_lib().apx_oracle_Z(self._h))

    def __del__(self):
        try:
            if self._h:
                _lib().apx_oracle_free(self._h)
                self._h = None
        except Exception:
            pass

    def dump(self):
        L = _lib()
        n, nnz = self.n, int(L.apx_oracle_nnz(self._h))
        d = dict(rowptr=np.zeros(n + 1, np.int64), adj=np.zeros(nnz, np.int32), order=np.zeros(n, np.int32),
                 position=np.zeros(n, np.int32), est=np.zeros(n, np.float64))
        L.apx_oracle_dump(self._h, *[_uo._p(d[k]) for k in ("rowptr", "adj", "order", "position", "est")])
        d["Z"] = self.Z
        return d

    def sample(self, m: int, max_trials: int = 0):
        out = np.zeros((m, self.k), np.int32)
        trials = np.zeros(m, np.int64)
        got = _lib().apx_oracle_sample(self._h, int(m), int(max_trials), _uo._p(out), _uo._p(trials))
        return out[:got].copy(), trials


def sample_batch(edge_index, ptr, m_per_graph: int, k: int, mode: str = "sample", seed: int = 42, epsilon: float = 0.1,
                 max_trials: int = 0):
    """Oracle counterpart of apx_ugs_sampler.sample_batch (:461-521): (samples[k,S], ptr[S+1])."""
    G = ApxGraph(edge_index, ptr, k, seed, epsilon)
    rows, _ = G.sample(m_per_graph, max_trials)
    S = rows.shape[0]
    if S == 0:
        return np.zeros((k, 0), np.int64), np.zeros(1, np.int64)
    return np.ascontiguousarray(rows.T.astype(np.int64)), np.arange(S + 1, dtype=np.int64)
