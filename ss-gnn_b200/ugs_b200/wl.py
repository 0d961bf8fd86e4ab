"""Weisfeiler-Lehman graphlet classification on the device.

Mirrors gps/utils/wl_vocab.py (compute_wl_hash :21-67, extract_subgraph_from_batch :70-108, hash_to_id :206-217,
build_wl_vocabulary_from_loader :111-183) and SubgraphSamplingGNNWithWL._compute_wl_ids (gps/models/ss_gnn_wl.py:210-247)
(both `use_node_features` settings): the reference converts every sampled subgraph to a NetworkX graph in Python and hashes
it; here one kernel (csrc/ugs_wl.cu) hashes all samples of a batch and produces the SAME hex strings, so vocabularies
built by either side are interchangeable.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import _ptr, _stream


def node_attributes(x: torch.Tensor) -> torch.Tensor:
    """int32[N]: the first four bytes of md5(x[v].tobytes()) for every node -- the reference's node attribute when
    `use_node_features=True` (wl_vocab.py:45-49: hexdigest()[:8])."""
    L = _lib.lib()
    if not x.is_cuda:
        raise RuntimeError("node_attributes needs a CUDA tensor (there is no CPU path)")
    x = x.contiguous()
    n = int(x.shape[0])
    row_bytes = (x.numel() // n if n else 0) * x.element_size()
    out = torch.empty(n, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(L.ugs_wl_node_attr(_ptr(x), n, row_bytes, _ptr(out), _stream()))
    return out


def wl_digests(nodes_t: torch.Tensor, edge_index_t: torch.Tensor, edge_ptr_t: torch.Tensor, num_iterations: int = 3,
               x: torch.Tensor | None = None):
    """(digests uint8 [B,16], num_nodes int32 [B]) on the device. digests[s].hex() == compute_wl_hash(sample s).
    x: node features of the batch (indexed by the global ids in nodes_t) for `use_node_features=True`, else None."""
    L = _lib.lib()
    if not (nodes_t.is_cuda and edge_index_t.is_cuda and edge_ptr_t.is_cuda):
        raise RuntimeError("wl_digests needs CUDA tensors (there is no CPU path)")
    if int(num_iterations) <= 0:
        raise ValueError("The WL algorithm requires that `iterations` be positive")
    nodes_t, edge_index_t, edge_ptr_t = nodes_t.contiguous(), edge_index_t.contiguous(), edge_ptr_t.contiguous()
    B, k = int(nodes_t.shape[0]), int(nodes_t.shape[1])
    dev = nodes_t.device
    dig = torch.empty((B, 2), dtype=torch.int64, device=dev)
    nn = torch.empty(B, dtype=torch.int32, device=dev)
    attr = node_attributes(x) if x is not None else None
    with torch.cuda.device(dev):
        _lib.check(L.ugs_wl_hash(_ptr(nodes_t), B, k, _ptr(edge_index_t), int(edge_index_t.shape[1]), _ptr(edge_ptr_t),
                                 int(num_iterations), _ptr(attr), int(attr.numel()) if attr is not None else 0,
                                 _ptr(dig), _ptr(nn), _stream()))
    return dig.view(torch.uint8).view(B, 16), nn


def wl_hashes(nodes_t, edge_index_t, edge_ptr_t, num_iterations: int = 3, x=None):
    """List of hash strings, one per sample (None where the sample has no valid node) == compute_wl_hash per sample."""
    dig, nn = wl_digests(nodes_t, edge_index_t, edge_ptr_t, num_iterations, x)
    raw = dig.cpu().numpy().tobytes()
    nn = nn.cpu().numpy()
    return [raw[16 * i:16 * i + 16].hex() if nn[i] > 0 else None for i in range(len(nn))]


def _vocab_tensors(wl_vocab: dict, device):
    keys = [h for h in wl_vocab if isinstance(h, str) and len(h) == 32]
    arr = np.zeros((len(keys), 2), dtype=np.uint64)
    ids = np.zeros(len(keys), dtype=np.int64)
    for i, h in enumerate(keys):
        try:
            arr[i, 0], arr[i, 1] = int(h[:16], 16), int(h[16:], 16)
        except ValueError:  # the reference's "deg_..._edges_..." fallback strings never match a digest
            arr[i] = 0
        ids[i] = int(wl_vocab[h])
    order = np.lexsort((arr[:, 1], arr[:, 0])) if len(keys) else np.zeros(0, np.int64)
    arr, ids = arr[order], ids[order]
    return (torch.from_numpy(arr.view(np.int64).copy()).to(device), torch.from_numpy(ids.copy()).to(device))


def compute_wl_ids(nodes_t, edge_index_t, edge_ptr_t, wl_vocab: dict, num_iterations: int = 3, x=None) -> torch.Tensor:
    """`_compute_wl_ids` (ss_gnn_wl.py:210-247): LongTensor[B] of vocabulary ids on nodes_t's device; unknown hashes
    and samples without a valid node get len(wl_vocab)."""
    L = _lib.lib()
    dig, nn = wl_digests(nodes_t, edge_index_t, edge_ptr_t, num_iterations, x)
    dev = nodes_t.device
    vk, vi = _vocab_tensors(wl_vocab, dev)
    B = int(nn.numel())
    ids = torch.empty(B, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.ugs_wl_lookup(_ptr(dig), _ptr(nn), B, _ptr(vk), _ptr(vi), int(vi.numel()), len(wl_vocab),
                                   _ptr(ids), _stream()))
    return ids


def extend_wl_vocabulary(wl_vocab: dict, nodes_t, edge_index_t, edge_ptr_t, num_iterations: int = 3, x=None) -> dict:
    """One batch of `build_wl_vocabulary_from_loader` (wl_vocab.py:150-176): new hashes get the next free ids in
    first-seen order (sample order), exactly like the reference's loop."""
    dig, nn = wl_digests(nodes_t, edge_index_t, edge_ptr_t, num_iterations, x)
    d = dig.cpu().numpy()
    valid = nn.cpu().numpy() > 0
    if not valid.any():
        return wl_vocab
    rows = np.ascontiguousarray(d[valid]).view([("a", "V16")]).ravel()
    _, first = np.unique(rows, return_index=True)
    raw = np.ascontiguousarray(d[valid])
    current_id = len(wl_vocab)
    for i in np.sort(first):
        h = raw[i].tobytes().hex()
        if h not in wl_vocab:
            wl_vocab[h] = current_id
            current_id += 1
    return wl_vocab
