"""Host-side partitioning of a batch by graph across ranks. Graphs are independent units:
no collective on the sampling path; the only exchange is an all-gather of a fixed int64 statistics vector."""
from __future__ import annotations

import numpy as np
import torch

STAT_KEYS = ("n_graphlets", "n_edge_entries", "n_incomplete", "n_relaxed_graphs", "n_skipped_graphs",
             "n_general_graphs")


def graph_cost(ptr: np.ndarray, col_start: np.ndarray, m: int, k: int) -> np.ndarray:
    """Cost proxy per graph: columns to preprocess + m*k*(mean degree) growth work (section 8e)."""
    n = np.maximum(np.diff(ptr), 1).astype(np.float64)
    c = np.diff(col_start).astype(np.float64)
    return c + m * k * (2.0 * c / n)


def partition(ptr, edge_index, world: int, m: int = 1, k: int = 1):
    """Contiguous graph ranges [g0, g1) per rank, balanced by `graph_cost`. Columns must be graph-contiguous
    (PyG batches are). Returns a list of (g0, g1, c0, c1)."""
    ptr = np.asarray(ptr, dtype=np.int64)
    row0 = np.asarray(edge_index[0], dtype=np.int64)
    G = len(ptr) - 1
    gid = np.searchsorted(ptr, row0, side="right") - 1
    if len(gid) and (np.diff(gid) < 0).any():
        raise ValueError("partition() needs graph-contiguous edge columns")
    col_start = np.searchsorted(gid, np.arange(G + 1), side="left")
    cost = np.cumsum(graph_cost(ptr, col_start, m, k)) if G else np.zeros(0)
    total = cost[-1] if G else 0.0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cost, total * r / world, side="left")) if G else 0)
    cuts.append(G)
    cuts = np.maximum.accumulate(np.asarray(cuts))
    return [(int(cuts[r]), int(cuts[r + 1]), int(col_start[cuts[r]]), int(col_start[cuts[r + 1]]))
            for r in range(world)]


def shard(ptr, edge_index, rank: int, world: int, m: int = 1, k: int = 1):
    """(edge_index_shard, ptr_shard, graph_index_base, first_column) of `rank`. Node ids keep their batch-global
    values (ptr_shard[0] is the node base), so emitted node ids / source columns only need `first_column` added
    to `edge_src` to be valid for the whole batch."""
    g0, g1, c0, c1 = partition(ptr, edge_index, world, m, k)[rank]
    ei = np.asarray(edge_index)[:, c0:c1]
    return np.ascontiguousarray(ei), np.ascontiguousarray(np.asarray(ptr)[g0:g1 + 1]), g0, c0


def shard_samples(m: int, rank: int, world: int):
    """(sample_index_base, m_local) of `rank` when ONE graph (or a small batch of huge graphs) is replicated on every
    rank and the m samples per graph are split instead of the graphs (section 8e, power-law configs). Passing them to
    DeviceBatch.sample(m_local, ..., sample_index_base=base) yields rows [base, base+m_local) of the single-rank call."""
    m, rank, world = int(m), int(rank), int(world)
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    q, r = divmod(max(m, 0), world)
    base = rank * q + min(rank, r)
    return base, q + (1 if rank < r else 0)


def gather_stats(stats: dict, group=None, device=None) -> dict:
    """All-gather the per-rank statistics vector (the design's only collective) and return summed totals plus
    the per-rank rows. Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    vec = torch.tensor([int(stats.get(kk, 0)) for kk in STAT_KEYS], dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()):
        return {"total": dict(zip(STAT_KEYS, vec.tolist())), "per_rank": [vec.tolist()]}
    world = dist.get_world_size(group)
    rows = [torch.empty_like(vec) for _ in range(world)]
    dist.all_gather(rows, vec, group=group)
    allr = torch.stack(rows)
    return {"total": dict(zip(STAT_KEYS, allr.sum(0).tolist())), "per_rank": allr.tolist()}


def global_edge_offsets(per_rank_rows) -> list:
    """Exclusive prefix of the ranks' emitted-entry counts: add offsets[r] to rank r's edge_ptr to obtain the
    edge_ptr of the concatenated batch."""
    off, out = 0, []
    for row in per_rank_rows:
        out.append(off)
        off += int(row[STAT_KEYS.index("n_edge_entries")])
    return out
