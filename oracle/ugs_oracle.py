"""Test infrastructure only (oracle/): ctypes wrapper around oracle/ugs_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this module. The product path (ss-gnn_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
MODES = {"sample": 0, "graph": 1, "global": 2}


def build(force: bool = False) -> str:
    out = os.path.join(HERE, "libugs_oracle.so")
    srcs = [os.path.join(HERE, f) for f in ("ugs_oracle.c", "eps_oracle.c", "uni_oracle.c", "apx_oracle.c", "walk_oracle.c")]
    srcs = [s for s in srcs if os.path.exists(s)]
    if force or not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-std=c11", "-fPIC", "-shared", "-o", out] + srcs + ["-lm"])
    return out


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp, i64, f64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double
        L.oracle_preproc_create.restype = vp
        L.oracle_preproc_create.argtypes = [vp, vp, i64, i64, ctypes.c_int]
        L.oracle_preproc_free.argtypes = [vp]
        for f in ("oracle_preproc_n", "oracle_preproc_nnz", "oracle_preproc_bucket_count_nonzero"):
            getattr(L, f).restype = i64
            getattr(L, f).argtypes = [vp]
        L.oracle_preproc_Z.restype = f64
        L.oracle_preproc_Z.argtypes = [vp]
        L.oracle_preproc_level.argtypes = [vp]
        L.oracle_preproc_copy.argtypes = [vp] * 11
        L.oracle_sample_nodes.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, vp]
        L.oracle_extract_one.restype = i64
        L.oracle_extract_one.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp]
        L.oracle_sample_batch.restype = vp
        L.oracle_sample_batch.argtypes = [vp, vp, i64, vp, i64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_uint64, ctypes.c_uint64, vp]
        L.oracle_batch_total_edges.restype = i64
        L.oracle_batch_total_edges.argtypes = [vp]
        L.oracle_batch_copy.argtypes = [vp] * 6
        L.oracle_batch_free.argtypes = [vp]
        L.oracle_philox.argtypes = [ctypes.c_uint32] * 6 + [vp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _rows(edge_index):
    ei = np.ascontiguousarray(np.asarray(edge_index, dtype=np.int64))
    if ei.ndim != 2 or ei.shape[0] != 2:
        raise ValueError("edge_index must be [2, E]")
    return ei, np.ascontiguousarray(ei[0]), np.ascontiguousarray(ei[1])


def philox(c0, c1, c2, c3, k0, k1):
    out = np.zeros(4, np.uint32)
    lib().oracle_philox(c0, c1, c2, c3, k0, k1, _p(out))
    return out


class Preproc:
    """Oracle counterpart of the reference's create_preproc handle (preproc.cpp:262-284)."""

    def __init__(self, edge_index, num_nodes: int, k: int):
        _, r0, r1 = _rows(edge_index)
        self.k = int(k)
        self._h = lib().oracle_preproc_create(_p(r0), _p(r1), r0.shape[0], int(num_nodes), int(k))
        self.n = int(lib().oracle_preproc_n(self._h))
        self.nnz = int(lib().oracle_preproc_nnz(self._h))

    def __del__(self):
        try:
            if self._h:
                lib().oracle_preproc_free(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def Z(self):
        return float(lib().oracle_preproc_Z(self._h))

    @property
    def level(self):
        return int(lib().oracle_preproc_level(self._h))

    def info(self):  # get_preproc_info, preproc.cpp:301-314
        return {"num_nodes": self.n, "num_edges_stored": self.nnz, "Z": self.Z,
                "bucket_count_nonzero": int(lib().oracle_preproc_bucket_count_nonzero(self._h))}

    def dump(self):
        n, nnz = self.n, self.nnz
        d = dict(indptr=np.zeros(n + 1, np.int64), indices=np.zeros(nnz, np.int32), edge_col=np.zeros(nnz, np.int32),
                 order=np.zeros(n, np.int32), index_of=np.zeros(n, np.int32), suffix_deg=np.zeros(n, np.int32),
                 bucket_b=np.zeros(n, np.float64), cum=np.zeros(n, np.float64),
                 alias_prob=np.zeros(n, np.float64), alias_alias=np.zeros(n, np.int32))
        lib().oracle_preproc_copy(self._h, *[_p(d[k]) for k in (
            "indptr", "indices", "edge_col", "order", "index_of", "suffix_deg", "bucket_b", "cum",
            "alias_prob", "alias_alias")])
        d["Z"] = self.Z
        d["level"] = self.level
        return d

    def sample_nodes(self, m: int, seed: int = 42, graph_index: int = 0):
        out = np.full((m, self.k), -1, np.int32)
        rc = lib().oracle_sample_nodes(self._h, m, self.k, seed & (2**64 - 1), graph_index, _p(out))
        if rc != 0:
            raise RuntimeError("No viable roots available")
        return out

    def extract(self, verts):
        verts = np.ascontiguousarray(np.asarray(verts, dtype=np.int32))
        cnt = int(lib().oracle_extract_one(self._h, self.k, _p(verts), None, None, None))
        ej, el, ec = (np.zeros(cnt, np.int32) for _ in range(3))
        if cnt:
            lib().oracle_extract_one(self._h, self.k, _p(verts), _p(ej), _p(el), _p(ec))
        return ej, el, ec


def sample_batch(edge_index, ptr, m_per_graph: int, k: int, mode: str = "sample", seed: int = 42,
                 graph_index_base: int = 0, given_nodes=None, sample_index_base: int = 0):
    """Oracle counterpart of ugs_sampler.sample_batch (ugs_sampler_batch_extension.cpp:77-300).

    Returns numpy (nodes[G*m,k], edge_index[2,E], edge_ptr[G*m+1], sample_ptr[G+1], edge_src[E])."""
    if mode not in MODES:
        raise RuntimeError("mode must be one of: 'sample', 'graph', 'global'")
    _, r0, r1 = _rows(edge_index)
    ptr = np.ascontiguousarray(np.asarray(ptr, dtype=np.int64))
    G = ptr.shape[0] - 1
    gn = None
    if given_nodes is not None:
        gn = np.ascontiguousarray(np.asarray(given_nodes, dtype=np.int64))
    L = lib()
    L.oracle_set_sample_index_base.argtypes = [ctypes.c_uint64]
    L.oracle_set_sample_index_base(int(sample_index_base))  # rows [base, base+m) of the single-call job
    h = L.oracle_sample_batch(_p(r0), _p(r1), r0.shape[0], _p(ptr), G, int(m_per_graph), int(k), MODES[mode],
                              seed & (2**64 - 1), graph_index_base, _p(gn) if gn is not None else None)
    try:
        E = int(L.oracle_batch_total_edges(h))
        nodes = np.zeros((G * m_per_graph, k), np.int64)
        ei = np.zeros((2, E), np.int64)
        eptr = np.zeros(G * m_per_graph + 1, np.int64)
        sptr = np.zeros(G + 1, np.int64)
        esrc = np.zeros(E, np.int64)
        L.oracle_batch_copy(h, _p(nodes), _p(ei), _p(eptr), _p(sptr), _p(esrc))
    finally:
        L.oracle_batch_free(h)
        L.oracle_set_sample_index_base(0)
    return nodes, ei, eptr, sptr, esrc


class UniformTables:
    """Tables of the exact-uniform accept/reject mode (oracle/uni_oracle.c) for one graph given its CSR: the 64-bit mask
    statement for n <= 64 (general=False), the adjacency-list statement for any n (general=True, the default above 64)."""

    def __init__(self, edge_index, num_nodes: int, k: int, general=None):
        self._P = Preproc(edge_index, num_nodes, k)
        d = self._P.dump()
        self._indptr, self._indices = d["indptr"], d["indices"]
        L = lib()
        self.n, self.k = int(num_nodes), int(k)
        self.general = (self.n > 64) if general is None else bool(general)
        vp = ctypes.c_void_p
        if self.general:
            L.uni_gen_create.restype = vp
            L.uni_gen_create.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int]
            L.uni_gen_free.argtypes = [vp]
            L.uni_gen_dump.argtypes = [vp] * 6
            L.uni_gen_sample_nodes.argtypes = [vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, vp, vp]
            self._h = L.uni_gen_create(_p(self._indptr), _p(self._indices), self.n, self.k)
        else:
            L.uni_oracle_create.restype = vp
            L.uni_oracle_create.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int]
            L.uni_oracle_free.argtypes = [vp]
            L.uni_oracle_dump.argtypes = [vp] * 7
            L.uni_oracle_sample_nodes.argtypes = [vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, vp, vp]
            self._h = L.uni_oracle_create(_p(self._indptr), _p(self._indices), self.n, self.k)
        if not self._h:
            raise RuntimeError("uniform mode needs k <= 8 (and n <= 64 for the mask statement)")

    def __del__(self):
        try:
            if self._h:
                (lib().uni_gen_free if self.general else lib().uni_oracle_free)(self._h)
                self._h = None
        except Exception:
            pass

    def dump(self):
        n = self.n
        d = dict(order=np.zeros(n, np.int32), dsuf=np.zeros(n, np.int32), suf=np.zeros(n, np.uint64),
                 b=np.zeros(n, np.float64), cum=np.zeros(n, np.float64))
        zc = np.zeros(2, np.float64)
        if self.general:
            lib().uni_gen_dump(self._h, _p(d["order"]), _p(d["dsuf"]), _p(d["b"]), _p(d["cum"]), _p(zc))
            del d["suf"]
        else:
            lib().uni_oracle_dump(self._h, _p(d["order"]), _p(d["dsuf"]), _p(d["suf"]), _p(d["b"]), _p(d["cum"]), _p(zc))
        d["Z"], d["C"] = float(zc[0]), float(zc[1])
        return d

    def sample_nodes(self, m: int, seed: int = 42, graph_index: int = 0):
        out = np.full((m, self.k), -1, np.int32)
        trials = np.zeros(1, np.int64)
        fn = lib().uni_gen_sample_nodes if self.general else lib().uni_oracle_sample_nodes
        fn(self._h, m, seed & (2 ** 64 - 1), graph_index, _p(out), _p(trials))
        return out, int(trials[0])


def sample_batch_uniform(edge_index, ptr, m_per_graph: int, k: int, mode: str = "sample", seed: int = 42,
                         graph_index_base: int = 0):
    """Batch driver of the uniform mode: nodes per graph from UniformTables, edges by the reference extraction."""
    ei = np.asarray(edge_index, dtype=np.int64)
    ptr = np.asarray(ptr, dtype=np.int64)
    G = len(ptr) - 1
    nodes = np.full((G * m_per_graph, k), -1, np.int64)
    for g in range(G):
        lo, hi = int(ptr[g]), int(ptr[g + 1])
        n = hi - lo
        if n <= 0 or n < k:
            continue
        msk = (ei[0] >= lo) & (ei[0] < hi) & (ei[1] >= lo) & (ei[1] < hi)
        T = UniformTables(ei[:, msk] - lo, n, k)
        rows, _ = T.sample_nodes(m_per_graph, seed, graph_index_base + g)
        nodes[g * m_per_graph:(g + 1) * m_per_graph] = np.where(rows >= 0, rows + lo, -1)
    return sample_batch(ei, ptr, m_per_graph, k, mode, seed, graph_index_base, given_nodes=nodes)
