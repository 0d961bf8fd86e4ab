"""Device pipeline for one PyG batch: segment -> preprocess -> sample -> scan -> emit.

Mirrors the reference batch driver (ugs_sampler_batch_extension.cpp:77-300) but runs every stage for the
whole batch at once through libugs_b200.so. PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

from . import _lib

MODES = {"sample": 0, "graph": 1, "global": 2, "encoder": 3}  # "encoder": UGS_MODE_ENCODER (not a reference mode)
VARIANTS = {"ref": 0, "uniform": 1}
CTR = dict(unsorted=0, n_general=1, max_n_general=2, n_skipped=3, total_edges=4, incomplete=5, relaxed=6, trials=7,
           n_large=9, emit_overflow=15)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceBatch:
    """Device-resident preprocessing of a batch of graphs: the equivalent of one reference `Preproc`
    (sampler.hpp:81-93) per graph, built by three kernels for the whole batch.

    edge_index: cuda int64 [2,E]; ptr: cuda int64 [G+1] (non-decreasing); node_base = ptr[0] and
    total_nodes = ptr[G]-ptr[0] must be passed by the caller (known on the host), so that no device->host
    read of ptr is needed.
    """

    def __init__(self, edge_index: torch.Tensor, ptr: torch.Tensor, k: int, node_base: int, total_nodes: int,
                 skip_lt_k: bool = True, variant: str = "ref"):
        L = _lib.lib()
        if not (edge_index.is_cuda and ptr.is_cuda):
            raise RuntimeError("DeviceBatch needs CUDA tensors (there is no CPU path)")
        if edge_index.dtype != torch.int64:
            raise RuntimeError("edge_index must be int64")
        if ptr.dtype != torch.int64:
            raise RuntimeError("ptr must be int64")
        if edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise RuntimeError("edge_index must have shape [2, E]")
        self.edge_index = edge_index.contiguous()
        self.ptr = ptr.contiguous()
        self.k = int(k)
        self.E = int(edge_index.size(1))
        self.G = int(ptr.numel()) - 1
        self.node_base = int(node_base)
        self.NT = int(total_nodes)
        self.device = edge_index.device
        self.layout = _lib.Layout()
        _lib.check(L.ugs_workspace_layout(self.E, max(self.G, 0), self.NT, ctypes.byref(self.layout)))
        self.ws = torch.empty(self.layout.total_bytes, dtype=torch.uint8, device=self.device)
        self.perm = None
        self._info_host = torch.empty(64, dtype=torch.int64).pin_memory()  # UGS_CTR_COUNT words
        self._skip_lt_k = skip_lt_k
        self.variant = variant
        if variant not in ("ref", "uniform"):
            raise RuntimeError(f"unknown variant {variant!r}")
        self._build(fast=True)

    def _build(self, fast: bool):
        """segment -> (sort columns by graph if needed) -> preprocess. `fast` assumes graph-contiguous columns
        (every PyG batch); the kernels verify it and `_late_check` redoes the batch on the slow path otherwise."""
        L = _lib.lib()
        self.perm = None
        self._enum = None
        self._fast = fast
        with torch.cuda.device(self.device):
            self._segment(fast)
            speculative = False
            if fast:
                # The bitmask-path preprocessing needs nothing from the host: queue it now and read the segmentation
                # flags / class counts through the side stream while it runs (no GPU idle time for this round trip).
                _lib.check(L.ugs_batch_info_begin(_ptr(self.ws), ctypes.byref(self.layout),
                                                  ctypes.c_void_p(self._info_host.data_ptr()), _stream()))
                self._preproc(0, 0, 1)
                _lib.check(L.ugs_batch_info_end())
                info = [int(x) for x in self._info_host[:16].tolist()]
                speculative = True
            else:
                info = self.info()
            if info[CTR["unsorted"]] & 2:
                raise RuntimeError("ptr must be non-decreasing")
            if fast and (info[CTR["unsorted"]] & 4):
                self._fast = fast = False
                speculative = False
                self._segment(False)
                info = self.info()
            if info[CTR["unsorted"]] & 1:
                # Columns are not graph-contiguous (never the case for PyG batches): order them stably by
                # graph id once and read them through `perm` from here on.
                speculative = False
                sb = int(L.ugs_sort_scratch_bytes(self.E))
                scratch = torch.empty(sb, dtype=torch.uint8, device=self.device)
                self.perm = torch.empty(self.E, dtype=torch.int64, device=self.device)
                _lib.check(L.ugs_sort_columns(self.E, self.G, _ptr(self.ws), ctypes.byref(self.layout), _ptr(scratch),
                                              sb, _ptr(self.perm), _stream()))
                del scratch
                self._segment(False)
                info = self.info()
            self.n_general = int(info[CTR["n_general"]])
            self.max_n_general = int(info[CTR["max_n_general"]])
            self.n_skipped = int(info[CTR["n_skipped"]])
            self.n_large = int(info[CTR["n_large"]])
            self._preproc(self.n_general, self.n_large, 2 if speculative else 3)
            if self.variant == "uniform":  # degree-dominating-order tables of the exact-uniform mode
                _lib.check(L.ugs_preproc_uniform(_ptr(self.ptr), self.G, self.node_base, self.k, self.n_general,
                                                 self.max_n_general, _ptr(self.ws), ctypes.byref(self.layout),
                                                 _stream()))

    def _preproc(self, n_general: int, n_large: int, parts: int):
        _lib.check(_lib.lib().ugs_preproc_batch_parts(_ptr(self.edge_index), self.E, _ptr(self.ptr), self.G,
                                                      self.node_base, self.k, _ptr(self.perm), n_general, n_large,
                                                      parts, _ptr(self.ws), ctypes.byref(self.layout), _stream()))

    def _check_emit_overflow(self, total):
        """General-path column emission (k_eps_emit_general) skips a sample it cannot hold (more than 1024 induced
        columns); its slice of the output would be uninitialised memory, so that is an error, not a result."""
        if self.n_general > 0 and total > 0:
            n_bad = self.info()[CTR["emit_overflow"]]
            if n_bad:
                raise RuntimeError(f"{n_bad} samples have more induced columns than the general-path emission "
                                   "holds (1024)")

    def _check_k(self, k) -> int:
        """The tables (path class, skip rule, reachability, bucket weights) were built for self.k: sampling another k on
        them would mix two laws (and overrun the bitmask path's 16-entry rows), so it is refused. The reference
        `sample(handle, m, k)` (sampler.cpp:91) takes k again but silently returns -1 rows when it differs from the
        preprocessing k (SURVEY section 7 "reference quirks"); `ugs_sampler.sample` re-preprocesses instead."""
        k = self.k if k is None else int(k)
        if k != self.k:
            raise RuntimeError(f"this batch was preprocessed for k={self.k}, not k={k}")
        return k

    def _late_check(self, info) -> bool:
        """True if the preprocessing kernels found a column outside its graph's range (the fast segmentation was
        fooled by a batch that is not graph-contiguous): the batch has been rebuilt on the slow path, redo the op."""
        if self._fast and (info[CTR["unsorted"]] & 8):
            self._build(fast=False)
            return True
        return False

    def _segment(self, fast):
        _lib.check(_lib.lib().ugs_segment_batch(_ptr(self.edge_index), self.E, _ptr(self.ptr), self.G, self.k,
                                                1 if self._skip_lt_k else 0, 1 if fast else 0, _ptr(self.perm),
                                                _ptr(self.ws),
                                                ctypes.byref(self.layout), _stream()))

    def info(self):
        """counters[0..15] as a list of python ints (synchronises the current stream)."""
        _lib.check(_lib.lib().ugs_batch_info(_ptr(self.ws), ctypes.byref(self.layout),
                                             ctypes.c_void_p(self._info_host.data_ptr()), _stream()))
        return [int(x) for x in self._info_host[:16].tolist()]

    # -- typed views into the workspace (parity tests read the deterministic stages through these) --
    _VIEWS = {
        "cstart": (torch.int64, lambda s: s.G + 2), "glist": (torch.int32, lambda s: s.G),
        "counters": (torch.int64, lambda s: 16), "colkey": (torch.int32, lambda s: s.E),
        "rowptr": (torch.int32, lambda s: s.NT + s.G), "order": (torch.int32, lambda s: s.NT),
        "index_of": (torch.int32, lambda s: s.NT), "suffix_deg": (torch.int32, lambda s: s.NT),
        "bucket_b": (torch.float64, lambda s: s.NT), "cum": (torch.float64, lambda s: s.NT),
        "adj": (torch.int64, lambda s: s.NT), "suf": (torch.int64, lambda s: s.NT),
        "csr_idx": (torch.int32, lambda s: 2 * s.E), "csr_col": (torch.int32, lambda s: 2 * s.E),
        "gaux": (torch.float64, lambda s: s.G),
    }

    def _ensure_valid(self):
        if self._fast:
            self._late_check(self.info())

    def view(self, name: str) -> torch.Tensor:
        if name != "colkey":
            self._ensure_valid()
        dt, nfn = self._VIEWS[name]
        n = nfn(self)
        off = getattr(self.layout, name)
        nbytes = n * torch.empty(0, dtype=dt).element_size()
        return self.ws[off:off + nbytes].view(dt)

    def ginfo(self) -> dict:
        """Per-graph records as CPU tensors: Z, nnz, level, n_eligible, last_viable, mult, cls."""
        self._ensure_valid()
        raw = self.ws[self.layout.ginfo:self.layout.ginfo + 32 * self.G].cpu()
        z = raw.view(torch.float64).view(self.G, 4)[:, 0].clone()
        ints = raw.view(torch.int32).view(self.G, 8)
        return dict(Z=z, nnz=ints[:, 2].clone(), level=ints[:, 3].clone(), n_eligible=ints[:, 4].clone(),
                    last_viable=ints[:, 5].clone(), mult=ints[:, 6].clone(), cls=ints[:, 7].clone())

    # ------------------------------------------------------------------------------------------------
    def sample(self, m: int, k: int | None = None, mode: str = "sample", seed: int = 42,
               graph_index_base: int = 0, variant: str | None = None, emit: bool = True,
               sample_index_base: int = 0, per_graph_seed: bool = False, alloc_edges: bool = True):
        """Returns device tensors (nodes[G*m,k], edge_index[2,Ebar], edge_ptr[G*m+1], sample_ptr[G+1],
        edge_src[Ebar]) and the stats dict. One host sync (reading Ebar) between sampling and emission."""
        L = _lib.lib()
        if mode not in MODES:
            raise RuntimeError("mode must be one of: 'sample', 'graph', 'global'")
        k = self._check_k(k)
        variant = self.variant if variant is None else variant
        if variant != self.variant:
            raise RuntimeError(f"this batch was preprocessed for variant {self.variant!r}")
        m = int(m)
        G = max(self.G, 0)
        dev = self.device
        with torch.cuda.device(dev):
            nodes = torch.empty((G * m, k), dtype=torch.int64, device=dev)
            edge_ptr = torch.empty(G * m + 1, dtype=torch.int64, device=dev)
            sample_ptr = torch.empty(G + 1, dtype=torch.int64, device=dev)
            sbytes = int(L.ugs_sample_scratch_bytes(G, m))
            scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
            early = variant == "ref"
            # reference-law path: the counters travel to the host as soon as the sampling kernels are done, i.e. while the
            # scan kernel runs, so allocating the data-dependent outputs does not leave the GPU idle
            _lib.check(L.ugs_sample_batch_early(_ptr(self.ptr), G, self.node_base, m, k, seed & (2 ** 64 - 1),
                                                graph_index_base, int(sample_index_base), 1 if per_graph_seed else 0,
                                                VARIANTS[variant], self.n_general, self.max_n_general,
                                                _ptr(self.ws), ctypes.byref(self.layout), _ptr(scratch), sbytes,
                                                _ptr(nodes), _ptr(edge_ptr), _ptr(sample_ptr),
                                                ctypes.c_void_p(self._info_host.data_ptr()) if early else None,
                                                _stream()))
            if early:
                _lib.check(L.ugs_batch_info_end())
                raw = self._info_host.tolist()
                info = [int(x) for x in raw[:16]]
                info[CTR["total_edges"]] = int(sum(raw[16:48]))  # UGS_CTR_EARLY_TOTAL slots
            else:
                info = self.info()
            if self._late_check(info):
                return self.sample(m, k, mode, seed, graph_index_base, variant, emit, sample_index_base,
                                   per_graph_seed, alloc_edges)
            if info[CTR["unsorted"]] & 16:
                raise RuntimeError(f"k={k} exceeds the bitmask path's limit of 16 for a batch preprocessed for k={self.k}")
            total = info[CTR["total_edges"]]
            # the reference law grows every row exactly once: trials == sampled rows (the counter is only written by the
            # rejection variants)
            n_trials = info[CTR["trials"]] if variant != "ref" else (G - self.n_skipped) * m
            stats = dict(n_graphlets=G * m, n_edge_entries=total, n_incomplete=info[CTR["incomplete"]],
                         n_relaxed_graphs=info[CTR["relaxed"]], n_skipped_graphs=self.n_skipped,
                         n_general_graphs=self.n_general, n_trials=n_trials)
            # alloc_edges=False (with emit=False): the caller decides between emit_last and emit_last_wire first
            edge_index = torch.empty((2, total), dtype=torch.int64, device=dev) if (alloc_edges or emit) else None
            edge_src = torch.empty(total, dtype=torch.int64, device=dev) if (alloc_edges or emit) else None
            # only the reference-law kernels leave packed rows behind the scan words of `scratch`
            packed = scratch if variant == "ref" else None
            self._emit_args = (nodes, edge_ptr, total, m, k, edge_index, edge_src, packed, sbytes)
            if emit:
                self.emit_last(mode)
        return (nodes, edge_index, edge_ptr, sample_ptr, edge_src), stats

    def wire_eligible(self, mode: str) -> bool:
        """True if the rows of the last `sample(..., emit=False)` call can be emitted in the host-transfer format
        (`emit_last_wire`): mode "sample", k <= 8, graph-contiguous columns, no general-path graph, reference law."""
        a = self._emit_args
        if not (a is not None and mode == "sample" and a[4] <= 8 and self.perm is None and self.n_general == 0
                and a[7] is not None and a[2] > 0):
            return False
        # the 8-bit source-column deltas need every 1024-entry block of the output to span fewer than 255 columns: with
        # at least 1024 entries per sampled graph a block covers two graphs (the delivery reports an overflow and the
        # caller takes the int64 emission, so this is only about not wasting an emission pass)
        return a[2] >= 1024 * max(self.G - self.n_skipped, 1)

    def emit_last_wire(self, d_wire: torch.Tensor):
        """Induced-edge emission of the last `sample(..., emit=False)` call straight into the transfer format of the host
        delivery (csrc/ugs_hostio.h): d_wire holds >= ugs_deliver_wire_bytes(G*m, k, total, 1) bytes. The int64 edge
        arrays are never materialised in HBM; `ugs_deliver_outputs_wire` finishes the job."""
        nodes, edge_ptr, total, m, k, _ei, _es, scratch, sbytes = self._emit_args
        G = max(self.G, 0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ugs_emit_edges_wire(
                _ptr(self.ptr), G, self.node_base, m, k, self.n_general, _ptr(self.ws), ctypes.byref(self.layout), _ptr(nodes),
                _ptr(edge_ptr), total, _ptr(d_wire), int(d_wire.numel()), _ptr(scratch), sbytes, _stream()))

    def emit_last(self, mode: str = "sample"):
        """Induced-edge emission for the rows of the last `sample(..., emit=False)` call (bench.py times the two halves
        separately). Uses the packed rows that call left in its scratch buffer."""
        nodes, edge_ptr, total, m, k, edge_index, edge_src, scratch, sbytes = self._emit_args
        G = max(self.G, 0)
        if edge_index is None:  # sample(..., alloc_edges=False)
            edge_index = torch.empty((2, total), dtype=torch.int64, device=self.device)
            edge_src = torch.empty(total, dtype=torch.int64, device=self.device)
            self._emit_out = (edge_index, edge_src)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ugs_emit_edges_packed(
                _ptr(self.ptr), G, self.node_base, m, k, MODES[mode], 1 if self.perm is not None else 0, self.n_general,
                _ptr(self.ws), ctypes.byref(self.layout), _ptr(nodes), _ptr(edge_ptr), total, _ptr(edge_index),
                _ptr(edge_src), _ptr(scratch), sbytes, _stream()))
        self._emit_args = None

    def sample_epsilon(self, m: int, k: int | None = None, mode: str = "sample", seed: int = 42,
                       epsilon: float = 0.1, graph_index_base: int = 0, given_nodes=None):
        """epsilon_uniform_sampler.sample_batch on the device (same 5-tuple layout, SORTED node rows).
        given_nodes: skip sampling and extract edges for these rows (pins the extraction stage)."""
        L = _lib.lib()
        if not (epsilon > 0.0 and epsilon <= 1.0):
            raise RuntimeError("epsilon must be in (0, 1]")
        k = self._check_k(k)
        m = int(m)
        G = max(self.G, 0)
        dev = self.device
        md = MODES.get(mode, 2)  # any mode other than "sample" emits batch-global ids (:279-285)
        with torch.cuda.device(dev):
            edge_ptr = torch.empty(G * m + 1, dtype=torch.int64, device=dev)
            sample_ptr = torch.empty(G + 1, dtype=torch.int64, device=dev)
            sbytes = int(L.ugs_sample_scratch_bytes(G, m))
            scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
            if given_nodes is None:
                nodes = torch.empty((G * m, k), dtype=torch.int64, device=dev)
                _lib.check(L.ugs_eps_sample_batch(_ptr(self.ptr), G, self.node_base, m, k, seed & (2 ** 64 - 1),
                                                  graph_index_base, float(epsilon), self.n_general, _ptr(self.ws),
                                                  ctypes.byref(self.layout), _ptr(scratch), sbytes, _ptr(nodes),
                                                  _ptr(edge_ptr), _ptr(sample_ptr), _stream()))
            else:
                nodes = given_nodes.to(dev).contiguous()
                # counts for caller-provided rows: CSR entries between members / 2 == columns inside the sample
                _lib.check(L.ugs_count_edges(_ptr(self.ptr), G, self.node_base, m, k, self.n_general, _ptr(self.ws),
                                             ctypes.byref(self.layout), _ptr(scratch), sbytes, _ptr(nodes),
                                             _ptr(edge_ptr), _stream()))
                edge_ptr //= 2
                sample_ptr = torch.arange(G + 1, dtype=torch.int64, device=dev) * m
            info = self.info()
            if self._late_check(info):
                return self.sample_epsilon(m, k, mode, seed, epsilon, graph_index_base, given_nodes)
            total = int(edge_ptr[-1].item()) if given_nodes is not None else info[CTR["total_edges"]]
            stats = dict(n_graphlets=G * m, n_edge_entries=total, n_incomplete=info[CTR["incomplete"]],
                         n_trials=info[CTR["trials"]], n_skipped_graphs=self.n_skipped,
                         n_general_graphs=self.n_general)
            edge_index = torch.empty((2, total), dtype=torch.int64, device=dev)
            edge_src = torch.empty(total, dtype=torch.int64, device=dev)
            _lib.check(L.ugs_eps_emit_edges(_ptr(self.edge_index), self.E, _ptr(self.ptr), G, self.node_base, m, k, md,
                                            _ptr(self.perm), self.n_general, _ptr(self.ws), ctypes.byref(self.layout),
                                            _ptr(nodes), _ptr(edge_ptr), total, _ptr(edge_index), _ptr(edge_src),
                                            _stream()))
            self._check_emit_overflow(total)
        return (nodes, edge_index, edge_ptr, sample_ptr, edge_src), stats

    def sample_rwr(self, m: int, k: int | None = None, mode: str = "sample", seed: int = 42, p_restart: float = 0.2,
                   graph_index_base: int = 0):
        """rwr_sampler.sample_batch on the device (rwr_sampler.cpp:72-295): same 5-tuple layout, node rows in
        visiting order, UGS-style entries (both directions, duplicates kept), edge_src = -1."""
        L = _lib.lib()
        k = self._check_k(k)
        if k < 1:
            raise RuntimeError("k must be >= 1")
        if not (p_restart >= 0.0 and p_restart <= 1.0):
            raise RuntimeError("p_restart in [0,1]")
        m = int(m)
        G = max(self.G, 0)
        dev = self.device
        md = MODES["sample"] if mode == "sample" else MODES["global"]
        with torch.cuda.device(dev):
            nodes = torch.empty((G * m, k), dtype=torch.int64, device=dev)
            edge_ptr = torch.empty(G * m + 1, dtype=torch.int64, device=dev)
            sample_ptr = torch.empty(G + 1, dtype=torch.int64, device=dev)
            sbytes = int(L.ugs_sample_scratch_bytes(G, m))
            scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
            _lib.check(L.ugs_rwr_sample_batch(_ptr(self.ptr), G, self.node_base, m, k, seed & (2 ** 64 - 1),
                                              graph_index_base, float(p_restart), _ptr(self.ws),
                                              ctypes.byref(self.layout), _ptr(scratch), sbytes, _ptr(nodes),
                                              _ptr(edge_ptr), _ptr(sample_ptr), _stream()))
            info = self.info()
            if self._late_check(info):
                return self.sample_rwr(m, k, mode, seed, p_restart, graph_index_base)
            total = info[CTR["total_edges"]]
            stats = dict(n_graphlets=G * m, n_edge_entries=total, n_incomplete=info[CTR["incomplete"]],
                         n_walk_steps=info[CTR["trials"]], n_skipped_graphs=self.n_skipped,
                         n_general_graphs=self.n_general)
            edge_index = torch.empty((2, total), dtype=torch.int64, device=dev)
            edge_src = torch.empty(total, dtype=torch.int64, device=dev)
            _lib.check(L.ugs_rwr_emit_edges(_ptr(self.ptr), G, self.node_base, m, k, md,
                                            1 if self.perm is not None else 0, self.n_general, _ptr(self.ws),
                                            ctypes.byref(self.layout), _ptr(nodes), _ptr(edge_ptr), total,
                                            _ptr(edge_index), _ptr(edge_src), _stream()))
        return (nodes, edge_index, edge_ptr, sample_ptr, edge_src), stats

    def enumerate_connected(self, k: int | None = None):
        """All connected k-subsets of every graph (uniform_sampler.cpp:47-81): (list[T, rw] int64, enum scratch).
        rw = 1: one 64-bit vertex mask per subset (every graph has <= 64 vertices); rw = 2: 16-byte records holding
        the mask (small graphs) or k 16-bit vertex ids (graphs above 64 vertices). Cached per batch."""
        L = _lib.lib()
        k = self.k if k is None else int(k)
        if k != self.k:
            raise RuntimeError(f"this batch was preprocessed for k={self.k}")
        cache = getattr(self, "_enum", None)
        if cache is not None and cache[0] == k:
            return cache[1], cache[2]
        self._ensure_valid()
        G = max(self.G, 0)
        dev = self.device
        with torch.cuda.device(dev):
            sb = int(L.ugs_enum_scratch_bytes(G, self.NT, self.n_general, self.max_n_general))
            es = torch.empty(sb, dtype=torch.uint8, device=dev)
            _lib.check(L.ugs_enum_count(_ptr(self.ptr), G, self.NT, self.node_base, k, self.n_general,
                                        self.max_n_general, _ptr(self.ws), ctypes.byref(self.layout), _ptr(es), sb,
                                        _stream()))
            total = self.info()[CTR["total_edges"]]
            rw = int(L.ugs_enum_record_words(self.n_general, self.max_n_general))
            lst = torch.empty((total, rw), dtype=torch.int64, device=dev)
            _lib.check(L.ugs_enum_fill(_ptr(self.ptr), G, self.NT, self.node_base, k, self.n_general,
                                       self.max_n_general, _ptr(self.ws), ctypes.byref(self.layout), _ptr(es),
                                       _ptr(lst), total, _stream()))
        self._enum = (k, lst, es)
        return lst, es

    def enum_graph_offsets(self, es: torch.Tensor) -> torch.Tensor:
        G = max(self.G, 0)
        return es[8 * (self.NT + 1): 8 * (self.NT + 1 + G + 1)].view(torch.int64)

    def sample_uniform_enum(self, m: int, k: int | None = None, mode: str = "sample", seed: int = 42,
                            graph_index_base: int = 0):
        """uniform_sampler.sample_batch on the device (uniform_sampler.cpp:86-282): one enumerated connected k-subset
        per sample, uniformly; ascending node rows; one edge entry per batch column inside the sample."""
        L = _lib.lib()
        k = self._check_k(k)
        m = int(m)
        G = max(self.G, 0)
        dev = self.device
        md = MODES.get(mode, 2) if mode == "sample" else 2
        lst, es = self.enumerate_connected(k)
        with torch.cuda.device(dev):
            nodes = torch.empty((G * m, k), dtype=torch.int64, device=dev)
            edge_ptr = torch.empty(G * m + 1, dtype=torch.int64, device=dev)
            sample_ptr = torch.empty(G + 1, dtype=torch.int64, device=dev)
            sbytes = int(L.ugs_sample_scratch_bytes(G, m))
            scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
            _lib.check(L.ugs_enum_sample_batch(_ptr(self.ptr), G, self.NT, self.node_base, m, k, seed & (2 ** 64 - 1),
                                               graph_index_base, self.n_general, self.max_n_general, _ptr(self.ws),
                                               ctypes.byref(self.layout), _ptr(es),
                                               _ptr(lst), _ptr(scratch), sbytes, _ptr(nodes), _ptr(edge_ptr),
                                               _ptr(sample_ptr), _stream()))
            info = self.info()
            total = info[CTR["total_edges"]]
            stats = dict(n_graphlets=G * m, n_edge_entries=total, n_incomplete=info[CTR["incomplete"]],
                         n_enumerated=int(lst.shape[0]), n_skipped_graphs=self.n_skipped,
                         n_general_graphs=self.n_general)
            edge_index = torch.empty((2, total), dtype=torch.int64, device=dev)
            edge_src = torch.empty(total, dtype=torch.int64, device=dev)
            _lib.check(L.ugs_eps_emit_edges(_ptr(self.edge_index), self.E, _ptr(self.ptr), G, self.node_base, m, k, md,
                                            _ptr(self.perm), self.n_general, _ptr(self.ws), ctypes.byref(self.layout),
                                            _ptr(nodes), _ptr(edge_ptr), total, _ptr(edge_index), _ptr(edge_src),
                                            _stream()))
            self._check_emit_overflow(total)
        return (nodes, edge_index, edge_ptr, sample_ptr, edge_src), stats

    def extract(self, nodes: torch.Tensor, m: int, k: int | None = None, mode: str = "sample"):
        """Induced-edge extraction (A.7) for caller-provided node rows [G*m,k] (batch-global ids, -1 pad)."""
        L = _lib.lib()
        k = self._check_k(k)
        G = max(self.G, 0)
        dev = self.device
        nodes = nodes.to(dev).contiguous()
        with torch.cuda.device(dev):
            edge_ptr = torch.empty(G * m + 1, dtype=torch.int64, device=dev)
            sbytes = int(L.ugs_sample_scratch_bytes(G, m))
            scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
            _lib.check(L.ugs_count_edges(_ptr(self.ptr), G, self.node_base, m, k, self.n_general, _ptr(self.ws),
                                         ctypes.byref(self.layout), _ptr(scratch), sbytes, _ptr(nodes),
                                         _ptr(edge_ptr), _stream()))
            info = self.info()
            if self._late_check(info):
                return self.extract(nodes, m, k, mode)
            total = info[CTR["total_edges"]]
            edge_index = torch.empty((2, total), dtype=torch.int64, device=dev)
            edge_src = torch.empty(total, dtype=torch.int64, device=dev)
            _lib.check(L.ugs_emit_edges(_ptr(self.ptr), G, self.node_base, m, k, MODES[mode], 1 if self.perm is not None else 0, self.n_general,
                                        _ptr(self.ws), ctypes.byref(self.layout), _ptr(nodes), _ptr(edge_ptr),
                                        total, _ptr(edge_index), _ptr(edge_src), _stream()))
        return edge_index, edge_ptr, edge_src


# ---- small batches: one library call, one host round trip (csrc/ugs_fused.cu) --------------------------------------
FUSED_MAX_BYTES = 1 << 30
# bound-sized edge arrays up to this size are handed out as the results themselves (views; no narrowing pass)
FUSED_COMPACT_MAX_BYTES = int(os.environ.get("UGS_B200_FUSED_COMPACT_BYTES", 64 << 20))     # worst-case output (24 bytes x 2E x m) up to which the fused path is taken
_fused_tls = threading.local()  # per thread and device: (pinned int64[64] tensor for the counters, its numpy view)


def fused_eligible(E: int, G: int, NT: int, m: int, k: int) -> bool:
    """The fused path serves DataLoader-sized calls: its output buffers are sized for the upper bound 2E*m entries."""
    if os.environ.get("UGS_B200_FUSED", "1") == "0":
        return False
    return G > 0 and m > 0 and 1 <= k <= 32 and NT < (1 << 20) and 24 * 2 * max(E, 1) * m <= FUSED_MAX_BYTES


def sample_batch_fused(edge_index: torch.Tensor, ptr: torch.Tensor, m: int, k: int, mode: str = "sample", seed: int = 42,
                       node_base: int = 0, total_nodes: int = 0, graph_index_base: int = 0,
                       sample_index_base: int = 0, per_graph_seed: bool = False):
    """`DeviceBatch(...).sample(...)` for a small batch as ONE library call (segment -> preprocess -> sample -> scan ->
    emit queued back to back, one host synchronisation). Returns (outs, stats) exactly like DeviceBatch.sample, or None
    when the batch needs the staged path (columns not graph-contiguous, decreasing ptr): the caller then uses DeviceBatch."""
    L = _lib.lib()
    if mode not in MODES:
        raise RuntimeError("mode must be one of: 'sample', 'graph', 'global'")
    if not (edge_index.is_cuda and ptr.is_cuda):
        raise RuntimeError("sample_batch_fused needs CUDA tensors (there is no CPU path)")
    if edge_index.dtype != torch.int64 or ptr.dtype != torch.int64:
        raise RuntimeError("edge_index and ptr must be int64")
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise RuntimeError("edge_index must have shape [2, E]")
    dev = edge_index.device
    if not edge_index.is_contiguous():
        edge_index = edge_index.contiguous()
    if not ptr.is_contiguous():
        ptr = ptr.contiguous()
    E, G, NT = int(edge_index.size(1)), int(ptr.numel()) - 1, int(total_nodes)
    m, k = int(m), int(k)
    B = G * m
    # Host time is what a call of this size costs: plain ints instead of ctypes objects, separate torch.empty calls
    # instead of views (2.5 us against 3 us per slice), no device context switch when the device is already current.
    switch = torch.cuda.current_device() != dev.index
    if switch:
        prev = torch.cuda.current_device()
        torch.cuda.set_device(dev)
    try:
        lay = _lib.Layout()
        _lib.check(L.ugs_workspace_layout(E, G, NT, ctypes.byref(lay)))
        cap = int(L.ugs_fused_capacity(E, m))
        sbytes = int(L.ugs_sample_scratch_bytes(G, m))
        # one temporary allocation: workspace | scan + packed-row scratch | bound-sized edge arrays. Small bounds get the
        # edge arrays as an allocation of their own that the emission kernels fill compactly (UGS_FUSED_COMPACT): the
        # results are views of it and the narrowing pass (a launch behind the host round trip) is not needed.
        off_s = (lay.total_bytes + 255) & ~255
        off_e = off_s + ((sbytes + 255) & ~255)
        compact = 24 * cap <= FUSED_COMPACT_MAX_BYTES
        if compact:
            tmp = torch.empty(off_e, dtype=torch.uint8, device=dev)
            ebuf = torch.empty(3 * cap, dtype=torch.int64, device=dev)
            e_base = ebuf.data_ptr()
        else:
            tmp = torch.empty(off_e + 24 * cap, dtype=torch.uint8, device=dev)
            e_base = tmp.data_ptr() + off_e
        # the three outputs whose sizes do not depend on the samples share one allocation (so that a caller who wants
        # them on the host needs one copy): nodes | edge_ptr | sample_ptr
        fixed = torch.empty(B * k + (B + 1) + (G + 1), dtype=torch.int64, device=dev)
        nodes = fixed[:B * k].view(B, k)
        edge_ptr = fixed[B * k:B * k + B + 1]
        sample_ptr = fixed[B * k + B + 1:]
        slots = _fused_tls.__dict__.setdefault("info", {})
        hi = slots.get(dev.index)
        if hi is None:
            t = torch.empty(64, dtype=torch.int64).pin_memory()
            hi = slots[dev.index] = (t, t.numpy())
        base = tmp.data_ptr()
        stream = torch.cuda.current_stream().cuda_stream
        # Batches of one data set look alike: after a batch without general-path graphs (every MUTAG / QM9 / ZINC batch)
        # the next one is queued without the general-path kernels, and repeated with them if the device count says the
        # guess was wrong (the outputs of the first attempt are void then).
        guess = 1 if (slots.get(("no_general", dev.index)) and k <= 16) else 0
        cflag = 2 if compact else 0
        while True:
            rc = L.ugs_sample_batch_fused(
                edge_index.data_ptr(), E, ptr.data_ptr(), G, int(node_base), NT, m, k, MODES[mode], seed & (2 ** 64 - 1),
                int(graph_index_base), int(sample_index_base), 1 if per_graph_seed else 0, base, ctypes.byref(lay),
                base + off_s, sbytes, nodes.data_ptr(), edge_ptr.data_ptr(), sample_ptr.data_ptr(), e_base,
                e_base + 8 * cap, e_base + 16 * cap, cap, hi[0].data_ptr(), guess | cflag, stream)
            if rc != 0:
                _lib.check(rc)
            info = hi[1]
            if info[CTR["unsorted"]]:
                return None
            no_general = slots[("no_general", dev.index)] = int(info[CTR["n_general"]]) == 0
            if not guess or no_general:
                break
            guess = 0
        total = int(info[CTR["total_edges"]])
        n_skipped = int(info[CTR["n_skipped"]])
        stats = dict(n_graphlets=B, n_edge_entries=total, n_incomplete=int(info[CTR["incomplete"]]),
                     n_relaxed_graphs=int(info[CTR["relaxed"]]), n_skipped_graphs=n_skipped,
                     n_general_graphs=int(info[CTR["n_general"]]), n_trials=(G - n_skipped) * m)
        if compact:
            edges = ebuf[:3 * total]  # edge_index rows | edge_src as the kernels wrote them
        else:
            edges = torch.empty(3 * total, dtype=torch.int64, device=dev)  # edge_index rows | edge_src, one allocation
        edge_index_out = edges[:2 * total].view(2, total)
        edge_src = edges[2 * total:]
        if total > 0 and not compact:
            rc = L.ugs_fused_finish(e_base, e_base + 8 * cap, e_base + 16 * cap, total,
                                    edge_index_out.data_ptr(), edge_src.data_ptr(), stream)
            if rc != 0:
                _lib.check(rc)
        _fused_tls.last_buffers = (fixed, edges)  # for ugs_sampler (same thread): two device->host copies instead of five
    finally:
        if switch:
            torch.cuda.set_device(prev)
    return (nodes, edge_index_out, edge_ptr, sample_ptr, edge_src), stats


def sample_batch_device(edge_index, ptr, m_per_graph, k, mode="sample", seed=42, node_base=0, total_nodes=None,
                        graph_index_base=0, variant="ref"):
    """One-shot device path: inputs and outputs are CUDA tensors. total_nodes = ptr[-1]-ptr[0] (host int)."""
    if total_nodes is None:
        pb = ptr[[0, -1]].tolist() if ptr.numel() > 0 else [0, 0]
        node_base, total_nodes = pb[0], pb[1] - pb[0]
    if variant == "ref" and fused_eligible(int(edge_index.size(1)), int(ptr.numel()) - 1, total_nodes, m_per_graph, k):
        r = sample_batch_fused(edge_index, ptr, m_per_graph, k, mode, seed, node_base, total_nodes, graph_index_base)
        if r is not None:
            return r
    b = DeviceBatch(edge_index, ptr, k, node_base, total_nodes, variant=variant)
    return b.sample(m_per_graph, k, mode, seed, graph_index_base)
