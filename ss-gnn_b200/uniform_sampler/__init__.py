"""uniform_sampler -- B200-native drop-in for SS-GNN's pybind11 module of the same name
(src/samplers/uniform_sampler/src/uniform_sampler.cpp:287-295; selected by `cfg.sampler == "uniform"`,
gps/experiment.py:173-175).

`sample_batch(edge_index, ptr, m_per_graph, k, mode="sample", seed=42)` returns the reference's 5-tuple: every row is
one of the graph's connected induced k-subgraphs chosen uniformly (ascending vertex order), all -1 rows for graphs
that have none; one edge entry per batch column inside the sample, in column order, `edge_src` = that column. Like the
reference it accepts tensors on any device and returns on the input device (CPU outputs are pinned, :149,273-279).

Computed by sm_100a kernels only (csrc/ugs_enum.cu): the connected k-subsets are enumerated directly (ESU recursion
on adjacency bitmasks), not filtered out of all C(n,k) combinations as in the reference (:47-81), so the cost follows
the number of connected subsets. Graphs with at most 64 nodes run on 64-bit masks (k <= 16), larger ones (up to 2048
nodes, k <= 8) on multi-word bitsets; beyond that the call raises RuntimeError -- the reference's exhaustive C(n,k)
scan is hopeless there as well.
Differences: Philox keyed (seed, graph, sample) instead of one mt19937_64 stream + std::uniform_int_distribution.
"""
from __future__ import annotations

import torch

from ugs_b200 import engine as _engine
from ugs_sampler import _device, _to_host_batch


def sample_batch(edge_index: torch.Tensor, ptr: torch.Tensor, m_per_graph: int, k: int, mode: str = "sample",
                 seed: int = 42):
    if edge_index.dtype != torch.int64:
        raise RuntimeError("edge_index must be int64")
    if ptr.dtype != torch.int64:
        raise RuntimeError("ptr must be int64")
    dev = edge_index.device if edge_index.is_cuda else _device()
    pb = ptr[[0, -1]].tolist() if ptr.numel() > 1 else [0, 0]
    with torch.cuda.device(dev):
        ei = edge_index.to(dev, non_blocking=True)
        p = ptr.to(dev, non_blocking=True)
        b = _engine.DeviceBatch(ei, p, int(k), pb[0], pb[1] - pb[0])
        outs, stats = b.sample_uniform_enum(int(m_per_graph), int(k), mode, int(seed))
        sample_batch.last_stats = stats
        return tuple(outs if edge_index.is_cuda else _to_host_batch(outs, mode, int(k)))


sample_batch.last_stats = None
