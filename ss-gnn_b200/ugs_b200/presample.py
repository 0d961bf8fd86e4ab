"""Device-resident presample store.

Replaces the reference's Python presample cache -- `Experiment._setup_presampling` (gps/experiment.py:379-440: one
sampler call per dataset graph with `ptr = [0, n]` and `seed = cfg.seed + index`, results kept in a dict of CPU
tensors) and `Experiment._load_from_presample_cache` (:895-995: a Python loop with `.item()` calls and `torch.cat` per
mini-batch). Here the whole split is sampled by ONE batch call whose rows are, graph by graph, exactly the rows the
per-graph loop over the drop-in `ugs_sampler.sample_batch` produces (Philox key = seed + index), the result stays in
HBM, and a mini-batch is stitched by three kernels (csrc/ugs_presample.cu).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .engine import DeviceBatch, _ptr, _stream


class PresampleStore:
    """`PresampleStore.build(edge_index, ptr, m, k, seed)` -- edge_index / ptr: the split's graphs concatenated in
    dataset order as one PyG batch (CUDA int64). `seed` = the seed of dataset graph 0 (the reference continues the
    count across splits: seed = cfg.seed + number of graphs presampled so far, experiment.py:411).
    `load(graph_idx, ptr, edge_index)` returns the 5-tuple of `_load_from_presample_cache` for a mini-batch made of
    dataset graphs `graph_idx` (in that order) whose PyG batch has `ptr` / `edge_index`."""

    def __init__(self, nodes, edge_index, edge_ptr, edge_src, s_ptr, s_cstart, m, k, stats):
        self.nodes, self.edge_index, self.edge_ptr, self.edge_src = nodes, edge_index, edge_ptr, edge_src
        self.s_ptr, self.s_cstart = s_ptr, s_cstart
        self.m, self.k = int(m), int(k)
        self.num_graphs = int(s_ptr.numel()) - 1
        self.stats = stats
        self.device = nodes.device

    @classmethod
    def build(cls, edge_index: torch.Tensor, ptr: torch.Tensor, m: int, k: int, seed: int = 42):
        if not (edge_index.is_cuda and ptr.is_cuda):
            raise RuntimeError("PresampleStore needs CUDA tensors (there is no CPU path)")
        pb = ptr[[0, -1]].tolist() if ptr.numel() > 1 else [0, 0]
        b = DeviceBatch(edge_index, ptr, int(k), pb[0], pb[1] - pb[0])
        (nodes, eidx, eptr, _sptr, esrc), stats = b.sample(int(m), int(k), "sample", int(seed), per_graph_seed=True)
        G = max(b.G, 0)
        cstart = b.view("cstart")[:G + 1].clone()
        return cls(nodes, eidx, eptr, esrc, ptr.contiguous().clone(), cstart, m, k, stats)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.nodes, self.edge_index, self.edge_ptr, self.edge_src, self.s_ptr, self.s_cstart))

    def load(self, graph_idx: torch.Tensor, ptr: torch.Tensor, edge_index: torch.Tensor, keep_padding: bool = False):
        """keep_padding=False reproduces the reference bit for bit, including `nodes_g + node_offset` being applied to
        the -1 padding of incomplete rows (experiment.py:957-958); keep_padding=True leaves those entries at -1."""
        L = _lib.lib()
        dev = self.device
        graph_idx = graph_idx.to(dev, torch.int64).contiguous().view(-1)
        ptr = ptr.to(dev, torch.int64).contiguous()
        row0 = edge_index.to(dev, torch.int64)[0].contiguous()
        Gb, m, k = int(graph_idx.numel()), self.m, self.k
        if int(ptr.numel()) != Gb + 1:
            raise RuntimeError("ptr must have len(graph_idx) + 1 entries")
        with torch.cuda.device(dev):
            sb = int(L.ugs_presample_scratch_bytes(Gb))
            scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
            total = ctypes.c_int64(0)
            _lib.check(L.ugs_presample_plan(_ptr(self.edge_ptr), self.num_graphs, m, _ptr(graph_idx), Gb, _ptr(ptr),
                                            _ptr(row0), int(row0.numel()), _ptr(scratch), sb, ctypes.byref(total),
                                            _stream()))
            T = int(total.value)
            nodes = torch.empty((Gb * m, k), dtype=torch.int64, device=dev)
            eidx = torch.empty((2, T), dtype=torch.int64, device=dev)
            eptr = torch.empty(Gb * m + 1, dtype=torch.int64, device=dev)
            sptr = torch.empty(Gb + 1, dtype=torch.int64, device=dev)
            esrc = torch.empty(T, dtype=torch.int64, device=dev)
            _lib.check(L.ugs_presample_load(_ptr(self.nodes), _ptr(self.edge_index), int(self.edge_index.shape[1]),
                                            _ptr(self.edge_ptr), _ptr(self.edge_src), _ptr(self.s_ptr),
                                            _ptr(self.s_cstart), self.num_graphs, m, k, _ptr(graph_idx), Gb, _ptr(ptr),
                                            _ptr(scratch), T, 1 if keep_padding else 0, _ptr(nodes), _ptr(eidx),
                                            _ptr(eptr), _ptr(sptr), _ptr(esrc), _stream()))
        return nodes, eidx, eptr, sptr, esrc
