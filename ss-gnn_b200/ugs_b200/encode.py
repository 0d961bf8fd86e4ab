"""Encoder inputs on the device: the gather/offset step that consumes the sampler's output.

Mirrors the tensor ops at the top of `encode_subgraphs` (gps/models/ss_gnn.py:427-471; the same code is in
ss_gnn_wl.py:249-306) -- everything up to the call of the GNN encoder:

    global_x, global_edge_attr, global_edge_index, global_batch_ptr = encoder_inputs(
        x, edge_attr, nodes_t, edge_index_t, edge_ptr_t, edge_src_global_t)

`sample_encoder_inputs` is the fused path: the sampler emits the batch-level edge index itself
(UGS_MODE_ENCODER), so `edge_index_t` is never materialised and re-read.
All work is done by libugs_b200.so kernels (csrc/ugs_gather.cu); there is no torch/CPU fallback.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib
from .engine import DeviceBatch, _ptr, _stream


def _rows(t: torch.Tensor):
    """(contiguous 2-D view, number of rows, bytes per row) of a feature tensor indexed along dim 0."""
    t = t.contiguous()
    return t, int(t.shape[0]), math.prod(int(d) for d in t.shape[1:]) * t.element_size()


def gather_rows(src: torch.Tensor, index: torch.Tensor, check: bool = False) -> torch.Tensor:
    """out[i] = src[index[i]] for index[i] >= 0, zeros for index[i] < 0 (the -1 padding of incomplete samples,
    ss_gnn.py:446-453). Bit-exact copy of whole rows, any dtype. check=True raises IndexError like torch indexing
    does when an index is >= src.shape[0] (costs one host sync)."""
    L = _lib.lib()
    if not (src.is_cuda and index.is_cuda):
        raise RuntimeError("gather_rows needs CUDA tensors (there is no CPU path)")
    if index.dtype != torch.int64:
        raise RuntimeError("index must be int64")
    src, n, row_bytes = _rows(src)
    index = index.contiguous().view(-1)
    out = torch.empty((index.numel(),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    bad = torch.zeros(1, dtype=torch.int64, device=src.device) if check else None
    with torch.cuda.device(src.device):
        _lib.check(L.ugs_gather_rows(_ptr(src), n, row_bytes, _ptr(index), index.numel(), _ptr(out), _ptr(bad),
                                     _stream()))
    if check and int(bad.item()) > 0:
        raise IndexError(f"index out of range: {int(bad.item())} rows >= {n}")
    return out


def encoder_inputs(x, edge_attr, nodes_t, edge_index_t, edge_ptr_t, edge_src_global_t):
    """Same arguments as `encode_subgraphs` (ss_gnn.py:427-435); returns the four tensors it hands to the encoder:
    (global_x [B*k, ...], global_edge_attr [Ebar, ...] or None, global_edge_index [2, Ebar], global_batch_ptr [B*k])."""
    L = _lib.lib()
    B, k = int(nodes_t.shape[0]), int(nodes_t.shape[1])
    dev = x.device
    global_x = gather_rows(x, nodes_t.reshape(-1))
    global_edge_attr = gather_rows(edge_attr, edge_src_global_t) if edge_attr is not None else None
    total = int(edge_index_t.shape[1])
    edge_index_t = edge_index_t.contiguous()
    edge_ptr_t = edge_ptr_t.contiguous()
    gei = torch.empty((2, total), dtype=torch.int64, device=dev)
    batch = torch.empty(B * k, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.ugs_encoder_offsets(_ptr(edge_index_t), _ptr(edge_ptr_t), B, k, total, _ptr(gei), _ptr(batch),
                                         _stream()))
    return global_x, global_edge_attr, gei, batch


def sample_encoder_inputs(batch: DeviceBatch, x, edge_attr, m: int, k: int | None = None, seed: int = 42,
                          graph_index_base: int = 0):
    """Fused path: sample m k-graphlets per graph and return what the encoder consumes, without materialising the
    sample-local edge_index. Returns (global_x, global_edge_attr, global_edge_index, global_batch_ptr,
    (nodes, edge_ptr, sample_ptr, edge_src), stats). Identical, value for value, to
    encoder_inputs(x, edge_attr, *sample(mode="sample"))."""
    L = _lib.lib()
    (nodes, gei, edge_ptr, sample_ptr, edge_src), stats = batch.sample(m, k, "encoder", seed, graph_index_base)
    B, kk = int(nodes.shape[0]), int(nodes.shape[1])
    global_x = gather_rows(x, nodes.reshape(-1))
    global_edge_attr = gather_rows(edge_attr, edge_src) if edge_attr is not None else None
    bvec = torch.empty(B * kk, dtype=torch.int64, device=nodes.device)
    with torch.cuda.device(nodes.device):
        _lib.check(L.ugs_encoder_offsets(ctypes.c_void_p(0), ctypes.c_void_p(0), B, kk, 0, ctypes.c_void_p(0),
                                         _ptr(bvec), _stream()))
    return global_x, global_edge_attr, gei, bvec, (nodes, edge_ptr, sample_ptr, edge_src), stats
