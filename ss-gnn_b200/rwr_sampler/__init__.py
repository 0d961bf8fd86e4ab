"""rwr_sampler -- B200-native drop-in for SS-GNN's pybind11 module of the same name
(src/samplers/rwr_sampler/src/rwr_sampler.cpp:297-308; selected by `cfg.sampler == "rwr"`, gps/experiment.py:164-168).

`sample_batch(edge_index, ptr, m_per_graph, k, mode="sample", seed=42, p_restart=0.2)` returns the reference's 5-tuple
(nodes[G*m,k] in visiting order, edge_index[2,E], edge_ptr[G*m+1], sample_ptr[G+1], edge_src[E] = -1). Like the
reference it accepts tensors on any device and returns on the input device (CPU outputs are pinned, :102,282-288).
Computed by sm_100a kernels only (csrc/ugs_rwr.cu); there is no CPU fallback.

Differences: randomness is Philox keyed (seed, graph, sample) -- the reference seeds one SplitMix64 per graph with
seed + g + 104729*omp_thread (:129), so its values depend on OpenMP scheduling, and its OpenMP loop writes rows through
a shared `sample_idx` (:122-266), so its row order is only the graph order with one thread. Here row g*m+i is always
sample i of graph g.
"""
from __future__ import annotations

import torch

from ugs_b200 import engine as _engine
from ugs_sampler import _device, _to_host_batch


def sample_batch(edge_index: torch.Tensor, ptr: torch.Tensor, m_per_graph: int, k: int, mode: str = "sample",
                 seed: int = 42, p_restart: float = 0.2):
    if edge_index.dtype != torch.int64:
        raise RuntimeError("edge_index must be int64")
    if ptr.dtype != torch.int64:
        raise RuntimeError("ptr must be int64")
    if int(k) < 1:
        raise RuntimeError("k must be >= 1")
    if not (p_restart >= 0.0 and p_restart <= 1.0):
        raise RuntimeError("p_restart in [0,1]")
    dev = edge_index.device if edge_index.is_cuda else _device()
    pb = ptr[[0, -1]].tolist() if ptr.numel() > 1 else [0, 0]
    with torch.cuda.device(dev):
        ei = edge_index.to(dev, non_blocking=True)
        p = ptr.to(dev, non_blocking=True)
        b = _engine.DeviceBatch(ei, p, int(k), pb[0], pb[1] - pb[0])
        outs, stats = b.sample_rwr(int(m_per_graph), int(k), mode, int(seed), float(p_restart))
        sample_batch.last_stats = stats
        return tuple(outs if edge_index.is_cuda else _to_host_batch(outs, mode, int(k)))


sample_batch.last_stats = None
