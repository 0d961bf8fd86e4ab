"""epsilon_uniform_sampler -- B200-native drop-in for SS-GNN's pybind11 module of the same name
(src/samplers/epsilon_uniform_sampler/src/epsilon_uniform_sampler.cpp:367-377).

`sample_batch(edge_index, ptr, m_per_graph, k, mode="sample", seed=42, epsilon=0.1)` returns the same 5-tuple as
`ugs_sampler.sample_batch` with SORTED node rows, one edge entry per batch column inside the sample (no 4x
multiplicity), all -1 rows when every rejection attempt failed. Like the reference it accepts tensors on any device
and returns on the input device (CPU outputs are pinned, :182,353-359). Computed by sm_100a kernels only.

Differences: graphs are emitted in graph order (the reference's order depends on OpenMP
scheduling, :295-319); randomness is Philox keyed (seed, graph, sample, attempt) instead of per-thread mt19937_64.
"""
from __future__ import annotations

import torch

from ugs_b200 import engine as _engine
from ugs_sampler import _device, _to_host_batch


def sample_batch(edge_index: torch.Tensor, ptr: torch.Tensor, m_per_graph: int, k: int, mode: str = "sample",
                 seed: int = 42, epsilon: float = 0.1):
    if edge_index.dtype != torch.int64:
        raise RuntimeError("edge_index must be int64")
    if ptr.dtype != torch.int64:
        raise RuntimeError("ptr must be int64")
    if not (epsilon > 0.0 and epsilon <= 1.0):
        raise RuntimeError("epsilon must be in (0, 1]")
    dev = edge_index.device if edge_index.is_cuda else _device()
    pb = ptr[[0, -1]].tolist() if ptr.numel() > 1 else [0, 0]
    with torch.cuda.device(dev):
        ei = edge_index.to(dev, non_blocking=True)
        p = ptr.to(dev, non_blocking=True)
        b = _engine.DeviceBatch(ei, p, int(k), pb[0], pb[1] - pb[0])
        outs, stats = b.sample_epsilon(int(m_per_graph), int(k), mode, int(seed), float(epsilon))
        sample_batch.last_stats = stats
        return tuple(outs if edge_index.is_cuda else _to_host_batch(outs, mode, int(k)))


sample_batch.last_stats = None
