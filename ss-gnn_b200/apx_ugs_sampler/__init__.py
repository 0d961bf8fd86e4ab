"""apx_ugs_sampler -- B200-native drop-in for SS-GNN's pybind11 module of the same name
(src/samplers/apx_ugs_sampler/src/apx_ugs_sampler.cpp:529-538; stubs __init__.pyi).

`sample_batch(edge_index, ptr, m_per_graph, k, mode="sample", seed=42, epsilon=0.1) -> (samples[k, S], ptr[S+1])`
with the reference's quirks kept: ONE graph per call, `ptr` is read as the edge-COLUMN range [ptr[0], ptr[1])
(:15-33), n = max vertex id + 1, `mode` is ignored, samples whose 10^6 trials all fail are dropped (S <= m).
Computed by sm_100a kernels (csrc/ugs_apx.cu): the independent trials of a sample run in parallel, one warp each,
and the first accepted trial is the sample. Limits: n <= 1024, 2 <= k <= 8. Randomness is Philox (law-level
parity with the reference's sequential mt19937_64; see oracle/apx_oracle.c).
"""
from __future__ import annotations

import ctypes

import torch

from ugs_b200 import _lib
from ugs_sampler import _device

MAX_TRIALS = 1_000_000  # apx_ugs_sampler.cpp:411


def _prepare(edge_index, ptr, m, k, seed, epsilon):
    L = _lib.lib()
    if edge_index.dtype != torch.int64 or ptr.dtype != torch.int64:
        raise RuntimeError("edge_index and ptr must be int64")
    dev = edge_index.device if edge_index.is_cuda else _device()
    e0, e1 = (int(x) for x in ptr[:2].tolist())
    ei = edge_index.to(dev).contiguous()
    E = int(ei.size(1))
    wsb = int(L.ugs_apx_workspace_bytes(max(e1 - e0, 0), m))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    info = (ctypes.c_int64 * 3)()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.ugs_apx_prepare(ctypes.c_void_p(ei.data_ptr()), E, e0, e1, m, k, seed & (2 ** 64 - 1), float(epsilon),
                                 ctypes.c_void_p(ws.data_ptr()), wsb, info, st))
    return L, dev, ws, (int(info[0]), int(info[1]), int(info[2])), e1 - e0, st


def sample_batch(edge_index: torch.Tensor, ptr: torch.Tensor, m_per_graph: int, k: int, mode: str = "sample",
                 seed: int = 42, epsilon: float = 0.1, _max_trials: int = MAX_TRIALS):
    m, k = int(m_per_graph), int(k)
    out_dev = edge_index.device
    with torch.cuda.device(edge_index.device if edge_index.is_cuda else _device()):
        L, dev, ws, (n, nnz, have_z), ncols, st = _prepare(edge_index, ptr, m, k, int(seed), float(epsilon))
        if n < k or not have_z or m <= 0:  # :473-478, :406-408, :502-506
            return torch.zeros((k, 0), dtype=torch.int64, device=out_dev), torch.zeros(1, dtype=torch.int64, device=out_dev)
        best = (ctypes.c_int32 * m)()
        per = max(256, min(65536, (1 << 20) // max(m, 1)))
        tbase = 0
        pending = m
        sd = seed & (2 ** 64 - 1)
        while tbase < _max_trials and pending > 0:
            _lib.check(L.ugs_apx_trials(n, ncols, m, k, sd, float(epsilon), tbase, per, int(_max_trials),
                                        ctypes.c_void_p(ws.data_ptr()), best, st))
            tbase += per
            pending = sum(1 for s in range(m) if best[s] >= min(tbase, _max_trials))
        slots, n_ok = [], 0
        for s in range(m):
            if best[s] < _max_trials:
                slots.append(n_ok)
                n_ok += 1
            else:
                slots.append(-1)
        sample_batch.last_trials = [int(b) if b < _max_trials else -1 for b in best]
        if n_ok == 0:
            return torch.zeros((k, 0), dtype=torch.int64, device=out_dev), torch.zeros(1, dtype=torch.int64, device=out_dev)
        slot_t = torch.tensor(slots, dtype=torch.int32, device=dev)
        samples = torch.empty((k, n_ok), dtype=torch.int64, device=dev)
        _lib.check(L.ugs_apx_write(n, ncols, m, k, sd, float(epsilon), ctypes.c_void_p(slot_t.data_ptr()), n_ok,
                                   ctypes.c_void_p(ws.data_ptr()), ctypes.c_void_p(samples.data_ptr()), st))
        torch.cuda.current_stream().synchronize()
        return samples.to(out_dev), torch.arange(0, n_ok + 1, dtype=torch.int64, device=out_dev)


sample_batch.last_trials = None


def _dump_tables(edge_index, ptr, k, seed=42, epsilon=0.1):
    """Test hook: the graph, approximate order and estimates as numpy arrays."""
    import numpy as np
    with torch.cuda.device(edge_index.device if edge_index.is_cuda else _device()):
        L, dev, ws, (n, nnz, have_z), ncols, st = _prepare(edge_index, ptr, 1, int(k), int(seed), float(epsilon))
        d = dict(rowptr=np.zeros(n + 1, np.int32), adj=np.zeros(max(nnz, 0), np.int32), order=np.zeros(n, np.int32),
                 position=np.zeros(n, np.int32), est=np.zeros(n, np.float64))
        if n >= k:
            _lib.check(L.ugs_apx_dump(n, nnz, ncols, 1, ctypes.c_void_p(ws.data_ptr()),
                                      *[ctypes.c_void_p(d[x].ctypes.data) for x in ("rowptr", "adj", "order", "position", "est")], st))
        d["n"], d["have_z"] = n, have_z
        return d
