"""ugs_sampler -- B200-native drop-in for SS-GNN's pybind11 module of the same name.

Same functions, argument names, defaults, output layout and error behaviour as the reference module
(src/samplers/ugs_sampler/src/extension.cpp:4-14, stubs __init__.pyi:1-56), so that
`import ugs_sampler; ugs_sampler.sample_batch(edge_index, ptr, m, k, mode="sample", seed=...)`
(gps/experiment.py:169-172,882-883) works unchanged once this directory's parent is on sys.path.

Everything is computed by hand-written sm_100a kernels behind libugs_b200.so (include/ugs_b200.h).
There is no CPU fallback: without the library or a CUDA device the calls raise RuntimeError.

Differences from the reference, all deliberate:
  * sample VALUES come from Philox4x32-10 keyed (seed, graph, sample) instead of an xorshift stream keyed on the
    sample index only; the output LAW (A.5-A.6) is the reference's;
  * no LRU preproc cache (the reference's key ignores k); preprocessing is recomputed per call on the GPU;
  * CUDA input tensors are accepted (outputs then stay on the device). With CPU inputs the outputs are CPU
    pinned int64 tensors exactly like the reference's (sampler.cpp:190); set UGS_B200_OUTPUT=cuda to keep
    them on the GPU;
  * UGS_B200_VARIANT=uniform switches sample_batch to the exact-uniform accept/reject mode (not in the reference:
    its own law is not uniform).
"""
from __future__ import annotations

import os
import threading

import torch

from ugs_b200 import engine as _engine

__version__ = "2.2.1+b200"

_registry = {}
_registry_lock = threading.Lock()
_next_handle = 1

_EDGE_MODES = {"local": "sample", "flat": "graph", "global": "global"}


# ---- UGS_DEBUG=1 diagnostics on stderr, in the reference's formats ------------------------------------------------
def _debug_on() -> bool:
    return os.environ.get("UGS_DEBUG") == "1"  # preproc.cpp:246-247: exactly the string "1"


def _fmt_preproc(n, k, Z, viable) -> str:
    """preproc.cpp:253-254"""
    return "[UGS PREPROC] n=%d k=%d Z=%.2e viable=%d/%d" % (n, k, Z, viable, n)


def _fmt_sample(n, k, m, Z, level, incomplete) -> str:
    """sampler.cpp:184-186; weighted / uniform = rows whose root came from the weights (level 0) / from the uniform
    relaxation draw (levels 1, 2)"""
    return ("[UGS SAMPLE] n=%d k=%d m=%d Z=%.2e | relax=%d weighted=%d uniform=%d incomplete=%d"
            % (n, k, m, Z, level, m if level == 0 else 0, 0 if level == 0 else m, incomplete))


def _fmt_cache(n_graphs) -> str:
    """ugs_sampler_batch_extension.cpp:249-254. There is no preprocessing cache here: every graph
    is preprocessed on the device in every call, which the line says in the reference's own vocabulary."""
    return "[UGS CACHE] hits=0 misses=%d hit_rate=0.0%% cache_size=0 (cache: none, preprocessing is recomputed)" % n_graphs


def _debug_report(b, nodes, m, k, with_cache_line):
    """One PREPROC and one SAMPLE line per sampled graph like the reference's per-graph calls, then the cache line.
    Debug mode only: reads the per-graph records and the node rows back to the host."""
    import sys
    gi = b.ginfo()
    sizes = (b.ptr[1:] - b.ptr[:-1]).cpu().tolist()
    bad = (nodes < 0).any(dim=1).view(max(b.G, 0), -1).sum(dim=1).cpu().tolist() if nodes.numel() else [0] * b.G
    lines = []
    n_sampled = 0
    for g in range(b.G):
        if int(gi["cls"][g]) == 0:  # n < k: the batch driver never preprocesses it (:132-143)
            continue
        n_sampled += 1
        Z, lvl = float(gi["Z"][g]), int(gi["level"][g])
        viable = int(gi["n_eligible"][g]) if lvl == 0 else 0
        lines.append(_fmt_preproc(sizes[g], k, Z, viable))
        lines.append(_fmt_sample(sizes[g], k, m, Z, lvl, int(bad[g])))
    if with_cache_line:
        lines.append(_fmt_cache(n_sampled))
    sys.stderr.write("\n".join(lines) + "\n")


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("ugs_sampler (B200): no CUDA device available and there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _check_inputs(edge_index, ptr=None, allow_cuda=True):
    if not isinstance(edge_index, torch.Tensor):
        raise TypeError("edge_index must be a torch.Tensor")
    if not allow_cuda and edge_index.is_cuda:
        raise RuntimeError("edge_index must be on CPU")
    if edge_index.dtype != torch.int64:
        raise RuntimeError("edge_index must be int64")
    if ptr is not None:
        if ptr.dtype != torch.int64:
            raise RuntimeError("ptr must be int64")


def _to_host(tensors):
    out = []
    for t in tensors:
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        out.append(h)
    torch.cuda.current_stream().synchronize()
    return out


_PACKED_MIN_BYTES = 32 << 20  # below this a plain copy is as fast


def _host_threads() -> int:
    """Host threads that widen packed results: UGS_B200_HOST_THREADS, else the cores this rank can claim (the box's
    cores divided by the ranks torchrun started on it), at most 32."""
    try:
        if os.environ.get("UGS_B200_HOST_THREADS"):
            return max(1, int(os.environ["UGS_B200_HOST_THREADS"]))
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    except ValueError:
        return 1
    return max(1, min(32, (os.cpu_count() or 1) // local_world))


def _dma_fraction(threads: int) -> float:
    """Share of every output array that travels as plain int64 by DMA (the rest is packed and widened by host threads):
    UGS_B200_DMA_FRACTION, default 0. Measured on this pool's hosts (profiles/r02_e2e_delivery_{1,8}gpu.txt): the int64
    result is bound by the HOST memory system either way, and a byte delivered by DMA costs it more than a byte written
    by a widening thread -- one rank: 54 ms all packed, 69 ms at 0.2, 131 ms all DMA (PCIe-bound); eight ranks: 132 GB/s
    of int64 all packed against 92 GB/s all DMA. The split is kept for hosts whose cores are the scarcer resource."""
    env = os.environ.get("UGS_B200_DMA_FRACTION")
    if env:
        try:
            return min(1.0, max(0.0, float(env)))
        except ValueError:
            pass
    return 0.0


# measured on the pool's hosts (profiles/r02_pageable_probe.txt): 105 MB of columns 62.6 -> 59.2 ms per call, 69 MB 10.4 -> 9.3 ms;
# 10-13 MB are 1.4 ms SLOWER than the driver's own staging, hence the threshold
_STAGE_MIN_BYTES = 32 << 20
_stage_tls = threading.local()


def _to_device(t: torch.Tensor, dev) -> torch.Tensor:
    """Host -> device copy of an input tensor. Large PAGEABLE tensors (what the trainer passes) go through a cached pinned
    staging buffer filled by the delivery's host threads (ugs_host_copy): the driver's own staging of pageable memory
    runs on one thread at a fraction of the link rate. UGS_B200_STAGE_INPUTS=0 leaves every copy to the driver."""
    if (t.is_cuda or t.is_pinned() or not t.is_contiguous() or t.numel() * t.element_size() < _STAGE_MIN_BYTES
            or os.environ.get("UGS_B200_STAGE_INPUTS", "1") == "0"):
        return t.to(dev, non_blocking=True)
    from ugs_b200 import _lib
    nbytes = t.numel() * t.element_size()
    st = getattr(_stage_tls, "bufs", None)
    if st is None:
        st = _stage_tls.bufs = {}
    slot = st.get(dev)
    if slot is not None and slot[1] is not None:
        slot[1].synchronize()  # the H2D copy that last read this staging buffer
    if slot is None or slot[0].numel() < nbytes:
        slot = [torch.empty(int(nbytes * 1.25), dtype=torch.uint8, pin_memory=True), None]
        st[dev] = slot
    _lib.check(_lib.lib().ugs_host_copy(slot[0].data_ptr(), t.data_ptr(), nbytes, _host_threads()))
    d = torch.empty(t.shape, dtype=t.dtype, device=dev)
    d.copy_(slot[0][:nbytes].view(t.dtype).view(t.shape), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    slot[1] = ev
    return d


def _wire_emit_wanted(G: int, m: int, k: int, E: int) -> bool:
    """Large results for a CPU caller: emit the edges in the transfer format (no int64 copy of them in HBM, no packing
    pass). UGS_B200_WIRE_EMIT=0 keeps the int64 emission + packing; the size test mirrors _PACKED_MIN_BYTES with the
    input's column count as the estimate of the output's entries per sample."""
    if os.environ.get("UGS_B200_WIRE_EMIT", "1") == "0" or os.environ.get("UGS_B200_PACKED_D2H", "1") == "0":
        return False
    if _dma_fraction(_host_threads()) > 0.0 or m <= 16 or k > 8:
        return False
    return 8 * (G * m * k + 3 * m * E) >= _PACKED_MIN_BYTES


def _deliver_wire(b, outs, stats, k: int):
    """Emission into the wire format + copy / widen pipeline for the rows `b.sample(..., emit=False, alloc_edges=False)`
    left behind. Returns the five pinned CPU tensors, or None when a value does not fit the format (the caller then takes
    the int64 path)."""
    import ctypes
    from ugs_b200 import _lib
    from ugs_b200.engine import _ptr, _stream
    nodes, _ei, eptr, sptr, _es = outs
    L = _lib.lib()
    B, kk, total = int(nodes.shape[0]), int(nodes.shape[1]), int(stats["n_edge_entries"])
    dev = nodes.device
    threads = _host_threads()
    with torch.cuda.device(dev):
        wb = int(L.ugs_deliver_wire_bytes(B, kk, total, 1))
        d_wire = torch.empty(wb, dtype=torch.uint8, device=dev)
        b.emit_last_wire(d_wire)
        h_wire = torch.empty(wb, dtype=torch.uint8, pin_memory=True)
        h_nodes = torch.empty(nodes.shape, dtype=torch.int64, pin_memory=True)
        h_eidx = torch.empty((2, total), dtype=torch.int64, pin_memory=True)
        h_esrc = torch.empty(total, dtype=torch.int64, pin_memory=True)
        h_eptr = torch.empty(eptr.shape, dtype=torch.int64, pin_memory=True)
        h_sptr = torch.empty(sptr.shape, dtype=torch.int64, pin_memory=True)
        h_eptr.copy_(eptr, non_blocking=True)
        h_sptr.copy_(sptr, non_blocking=True)
        st = (ctypes.c_int64 * 4)()
        rc = L.ugs_deliver_outputs_wire(_ptr(nodes), B, kk, total, _ptr(d_wire), ctypes.c_void_p(h_wire.data_ptr()), wb,
                                        ctypes.c_void_p(h_nodes.data_ptr()), ctypes.c_void_p(h_eidx.data_ptr()),
                                        ctypes.c_void_p(h_esrc.data_ptr()), threads, st, _stream())
        torch.cuda.current_stream().synchronize()
        if rc == -3:  # UGS_ELIMIT
            return None
        _lib.check(rc)
    _to_host_batch.last_d2h_bytes = int(st[0]) + 8 * (eptr.numel() + sptr.numel())
    _to_host_batch.last_delivery = (f"dma_fraction 0.00: 0 B as int64 by DMA + {int(st[0])} B packed (node ids: 16-bit deltas; "
                                    f"edges emitted in the wire format: u8 pairs + 8-bit source-column deltas), widened by "
                                    f"{threads} host threads")
    return [h_nodes, h_eidx, h_eptr, h_sptr, h_esrc]


def _to_host_batch(outs, mode: str, k: int):
    """The 5-tuple (nodes, edge_index, edge_ptr, sample_ptr, edge_src) as CPU pinned int64 tensors (the reference's
    output contract). Large results are split between a plain int64 DMA and a packed transfer (16-bit deltas / one byte
    per edge entry) that host threads widen into the final tensors while the copies run (csrc/ugs_hostio.cu,
    csrc/ugs_hostdeliver.cpp); UGS_B200_PACKED_D2H=0 forces the plain copy, UGS_B200_DMA_FRACTION the split."""
    import ctypes
    from ugs_b200 import _lib
    from ugs_b200.engine import _ptr, _stream
    nodes, eidx, eptr, sptr, esrc = outs
    nbytes = 8 * (nodes.numel() + eidx.numel() + esrc.numel())
    _to_host_batch.last_d2h_bytes = nbytes + 8 * (eptr.numel() + sptr.numel())  # what crossed PCIe for this call
    _to_host_batch.last_delivery = "plain int64 copy"
    if nbytes < _PACKED_MIN_BYTES or os.environ.get("UGS_B200_PACKED_D2H", "1") == "0":
        return _to_host(outs)
    L = _lib.lib()
    B, kk, total = int(nodes.shape[0]), int(nodes.shape[1]), int(esrc.numel())
    pairs = 1 if (mode == "sample" and int(k) <= 16) else 0
    dev = nodes.device
    threads = _host_threads()
    frac = _dma_fraction(threads)
    with torch.cuda.device(dev):
        wb = int(L.ugs_deliver_wire_bytes(B, kk, total, pairs)) if frac < 1.0 else 256  # nothing is packed at 1.0
        d_wire = torch.empty(wb, dtype=torch.uint8, device=dev)
        h_wire = torch.empty(wb, dtype=torch.uint8, pin_memory=True)
        h_nodes = torch.empty(nodes.shape, dtype=torch.int64, pin_memory=True)
        h_eidx = torch.empty(eidx.shape, dtype=torch.int64, pin_memory=True)
        h_esrc = torch.empty(esrc.shape, dtype=torch.int64, pin_memory=True)
        h_eptr = torch.empty(eptr.shape, dtype=torch.int64, pin_memory=True)
        h_sptr = torch.empty(sptr.shape, dtype=torch.int64, pin_memory=True)
        eidx_c = eidx.contiguous()
        h_eptr.copy_(eptr, non_blocking=True)
        h_sptr.copy_(sptr, non_blocking=True)
        stats = (ctypes.c_int64 * 4)()
        rc = L.ugs_deliver_outputs(_ptr(nodes), B, kk, _ptr(eidx_c), _ptr(esrc), total, pairs, frac, _ptr(d_wire),
                                   ctypes.c_void_p(h_wire.data_ptr()), wb,
                                   ctypes.c_void_p(h_nodes.data_ptr()), ctypes.c_void_p(h_eidx.data_ptr()),
                                   ctypes.c_void_p(h_esrc.data_ptr()), threads, stats, _stream())
        torch.cuda.current_stream().synchronize()
        if rc == -3:  # UGS_ELIMIT: a value outside the packed ranges (never for valid sampler output): plain copy
            return _to_host(outs)
        _lib.check(rc)
    _to_host_batch.last_d2h_bytes = int(stats[0]) + int(stats[1]) + 8 * (eptr.numel() + sptr.numel())
    _to_host_batch.last_delivery = (f"dma_fraction {frac:.2f}: {int(stats[1])} B as int64 by DMA + {int(stats[0])} B packed "
                                    f"({'16-bit deltas' if stats[2] else 'int32'} / {'u8 pairs' if pairs else 'int32 pairs'}), "
                                    f"widened by {threads} host threads")
    return [h_nodes, h_eidx, h_eptr, h_sptr, h_esrc]


_to_host_batch.last_d2h_bytes = 0
_to_host_batch.last_delivery = None


def _want_cuda_out(inp):
    return inp.is_cuda or os.environ.get("UGS_B200_OUTPUT", "").lower() == "cuda"


def create_preproc(edge_index: torch.Tensor, num_nodes: int, k: int) -> int:
    """Create preprocessing for a graph and return a handle (int). Reference: preproc.cpp:262-284."""
    global _next_handle
    _check_inputs(edge_index)
    dev = edge_index.device if edge_index.is_cuda else _device()
    n = int(num_nodes)
    ei = edge_index.to(dev, non_blocking=True)
    ptr = torch.tensor([0, max(n, 0)], dtype=torch.int64).to(dev, non_blocking=True)
    b = _engine.DeviceBatch(ei, ptr, int(k), 0, max(n, 0), skip_lt_k=False)
    b._cuda_out = _want_cuda_out(edge_index)
    with _registry_lock:
        h = _next_handle
        _next_handle += 1
        _registry[h] = b
    return h


def destroy_preproc(handle: int) -> None:
    """Destroy a preprocessing handle (unknown handles are ignored, preproc.cpp:286-292)."""
    with _registry_lock:
        _registry.pop(int(handle), None)


def has_graphlets(handle: int) -> bool:
    """True if the preprocessed graph contains k-graphlets, i.e. Z > 0 (preproc.cpp:294-299)."""
    with _registry_lock:
        b = _registry.get(int(handle))
    if b is None or b.G < 1:
        return False
    return bool(float(b.ginfo()["Z"][0]) > 0.0)


def get_preproc_info(handle: int) -> dict:
    """Small metadata dict (preproc.cpp:301-314); empty for an unknown handle."""
    with _registry_lock:
        b = _registry.get(int(handle))
    if b is None:
        return {}
    gi = b.ginfo()
    nz = int((b.view("bucket_b") > 0).sum().item())
    return {"num_nodes": b.NT, "num_edges_stored": int(gi["nnz"][0]), "Z": float(gi["Z"][0]),
            "bucket_count_nonzero": nz}


def sample(handle: int, m_per_graph: int, k: int, edge_mode: str = "local", base_offset: int = 0, seed: int = 42):
    """Sample m_per_graph subgraphs of size k from a preprocessed graph; edge_mode in {'local','flat','global'}.

    Returns (nodes[m,k], edge_index[2,E], edge_ptr[m+1], edge_src[E]) like sampler.cpp:91-290."""
    with _registry_lock:
        b = _registry.get(int(handle))
    if b is None:
        raise RuntimeError("Invalid preproc handle")
    if b.NT <= 0:
        raise RuntimeError("No viable roots available")
    # Any edge_mode other than 'local'/'flat' behaves as 'global' in the reference (sampler.cpp:262-273).
    mode = _EDGE_MODES.get(edge_mode, "global")
    if int(k) != b.k:
        # The reference takes k again here and mixes it with the tables of the preprocessing k (rows of -1 when it is
        # larger). The tables are cheap to rebuild on the device, so the handle is re-preprocessed for the new k.
        nb = _engine.DeviceBatch(b.edge_index, b.ptr, int(k), 0, b.NT, skip_lt_k=False)
        nb._cuda_out = b._cuda_out
        with _registry_lock:
            if int(handle) in _registry:
                _registry[int(handle)] = nb
        b = nb
    (nodes, ei, eptr, _sptr, esrc), _ = b.sample(int(m_per_graph), int(k), mode, int(seed))
    if _debug_on():
        _debug_report(b, nodes, int(m_per_graph), int(k), with_cache_line=False)
    if mode == "global" and base_offset:
        nodes = torch.where(nodes >= 0, nodes + int(base_offset), nodes)
        ei = ei + int(base_offset)
    outs = [nodes, ei, eptr, esrc]
    return tuple(outs if b._cuda_out else _to_host(outs))


def sample_batch(edge_index: torch.Tensor, ptr: torch.Tensor, m_per_graph: int, k: int, mode: str = "sample",
                 seed: int = 42):
    """Sample m_per_graph k-subgraphs per graph from a batched PyG edge_index + ptr.

    Returns (nodes[G*m,k], edge_index[2,Ebar], edge_ptr[G*m+1], sample_ptr[G+1], edge_src_global[Ebar]),
    reference ugs_sampler_batch_extension.cpp:77-300."""
    _check_inputs(edge_index, ptr)
    if mode not in ("sample", "graph", "global"):
        raise RuntimeError("mode must be one of: 'sample', 'graph', 'global'")
    dev = edge_index.device if edge_index.is_cuda else _device()
    if ptr.is_cuda:
        pb = ptr[[0, -1]].tolist() if ptr.numel() > 1 else [0, 0]
    else:
        pb = [int(ptr[0]), int(ptr[-1])] if ptr.numel() > 1 else [0, 0]
    with torch.cuda.device(dev):
        ei = _to_device(edge_index, dev)
        p = _to_device(ptr, dev)
        # UGS_B200_VARIANT=uniform selects the exact-uniform accept/reject mode; default = reference law
        variant = os.environ.get("UGS_B200_VARIANT", "ref").lower()
        r = None
        G, E, NT = int(p.numel()) - 1, int(ei.size(1)) if ei.dim() == 2 else 0, pb[1] - pb[0]
        if (variant == "ref" and not _debug_on() and ei.dim() == 2 and ei.size(0) == 2
                and _engine.fused_eligible(E, G, NT, int(m_per_graph), int(k))):
            # DataLoader-sized call: the whole pipeline as one library call with one host round trip
            r = _engine.sample_batch_fused(ei, p, int(m_per_graph), int(k), mode, int(seed), pb[0], NT)
        bufs = None
        if r is not None:  # fused path: the two device allocations behind the five outputs
            bufs = getattr(_engine._fused_tls, "last_buffers", None)
            _engine._fused_tls.last_buffers = None
        else:
            b = _engine.DeviceBatch(ei, p, int(k), pb[0], NT, variant=variant)
            if (variant == "ref" and mode == "sample" and not _want_cuda_out(edge_index) and not _debug_on()
                    and _wire_emit_wanted(G, int(m_per_graph), int(k), E)):
                # CPU caller, large result: sample, then emit the edges in the transfer format and deliver them
                r = b.sample(int(m_per_graph), int(k), mode, int(seed), emit=False, alloc_edges=False)
                sample_batch.last_stats = r[1]
                if b.wire_eligible(mode):
                    host = _deliver_wire(b, r[0], r[1], int(k))
                    if host is not None:
                        return tuple(host)
                b.emit_last(mode)  # the int64 path after all (general-path graphs, unsorted columns, a delta out of range)
                r = ((r[0][0], b._emit_out[0], r[0][2], r[0][3], b._emit_out[1]), r[1])
            else:
                r = b.sample(int(m_per_graph), int(k), mode, int(seed))
            if _debug_on():
                _debug_report(b, r[0][0], int(m_per_graph), int(k), with_cache_line=True)
        outs, _stats = r
        sample_batch.last_stats = _stats
        if (bufs is not None and not _want_cuda_out(edge_index)
                and 8 * (bufs[0].numel() + bufs[1].numel()) < _PACKED_MIN_BYTES):
            # fused small-batch path: the five outputs are views of two device allocations -> two copies, five host views
            fixed, edges = bufs
            h_fixed = torch.empty(fixed.shape, dtype=torch.int64, pin_memory=True)
            h_edges = torch.empty(edges.shape, dtype=torch.int64, pin_memory=True)
            h_fixed.copy_(fixed, non_blocking=True)
            h_edges.copy_(edges, non_blocking=True)
            B, kk, total, Gn = outs[0].shape[0], outs[0].shape[1], outs[4].shape[0], outs[3].shape[0]
            torch.cuda.current_stream().synchronize()
            _to_host_batch.last_delivery = "plain int64 copy (two buffers)"
            _to_host_batch.last_d2h_bytes = 8 * (fixed.numel() + edges.numel())
            return (h_fixed[:B * kk].view(B, kk), h_edges[:2 * total].view(2, total), h_fixed[B * kk:B * kk + B + 1],
                    h_fixed[B * kk + B + 1:B * kk + B + 1 + Gn], h_edges[2 * total:])
        return tuple(outs if _want_cuda_out(edge_index) else _to_host_batch(outs, mode, int(k)))


sample_batch.last_stats = None
