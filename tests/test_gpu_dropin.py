"""GPU: the drop-in `ugs_sampler` module through its public API -- the reference's own structural tests
(tests/test_sampler.py:7-81, tests/test_sample_batch.py:18-41) re-expressed, the golden vectors generated from
the real reference, and the sampling law (chi-square against the exact reference law P_ref)."""
import os
from collections import Counter

import numpy as np
import pytest
import torch
from scipy.stats import chi2

from helpers import bidir, per_graph_columns, random_batch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ugs_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_reference_test_sampler_global_mode_invariants(cuda_lib):
    """reference tests/test_sampler.py:7-81 verbatim semantics (n=4 with out-of-range columns)."""
    import ugs_sampler
    edge_index_local = torch.tensor([[0, 1, 2, 3, 4, 5, 5, 6, 7, 8, 9], [1, 2, 3, 4, 5, 0, 6, 7, 8, 9, 0]], dtype=torch.long)
    n, k, m, base_offset = 4, 3, 4, 10
    handle = ugs_sampler.create_preproc(edge_index_local, n, k)
    try:
        nodes_t, edge_index_t, edge_ptr_t, edge_src_idx_t = ugs_sampler.sample(handle, m, k, "global", base_offset)
        info = ugs_sampler.get_preproc_info(handle)
        assert ugs_sampler.has_graphlets(handle)
    finally:
        ugs_sampler.destroy_preproc(handle)
    assert info["num_nodes"] == 4 and info["num_edges_stored"] == 6  # columns touching ids >= n are skipped
    assert not nodes_t.is_cuda and nodes_t.is_pinned() and nodes_t.dtype == torch.int64
    assert nodes_t.shape == (m, k) and edge_ptr_t.shape == (m + 1,)
    assert edge_index_t.shape[0] == 2 and edge_src_idx_t.numel() == edge_index_t.shape[1]
    for b in range(m):
        node_set = {int(x) for x in nodes_t[b].tolist() if x >= 0}
        for u, v in edge_index_t[:, int(edge_ptr_t[b]):int(edge_ptr_t[b + 1])].t().tolist():
            assert base_offset <= u < base_offset + n and base_offset <= v < base_offset + n
            assert u in node_set and v in node_set
    gathered = edge_index_local[:, edge_src_idx_t]
    for ug, vg, ul, vl in zip(edge_index_t[0].tolist(), edge_index_t[1].tolist(), gathered[0].tolist(), gathered[1].tolist()):
        assert sorted((ug - base_offset, vg - base_offset)) == sorted((ul, vl))


def test_reference_test_sample_batch(cuda_lib):
    """reference tests/test_sample_batch.py:18-41 (two graphs, single-direction columns)."""
    import ugs_sampler
    ptr = torch.tensor([0, 4, 8], dtype=torch.long)
    edge_index = torch.tensor([[0, 1, 2, 4, 5, 6, 4], [1, 2, 3, 5, 6, 7, 7]], dtype=torch.long)
    m, k = 2, 3
    nodes_t, ei_t, eptr, sptr, esrc = ugs_sampler.sample_batch(edge_index, ptr, m, k, mode="global")
    assert ei_t.shape[1] > 0
    for i, (u, v) in enumerate(ei_t.t().tolist()):
        iu, iv = edge_index.t()[esrc[i]].tolist()
        assert (u, v) == (iu, iv) or (u, v) == (iv, iu)
    nodes_t, ei_t, eptr, sptr, esrc = ugs_sampler.sample_batch(edge_index, ptr, m, k, mode="sample")
    sid = np.searchsorted(eptr.numpy()[1:], np.arange(ei_t.shape[1]), side="right")
    for i, (u, v) in enumerate(ei_t.t().tolist()):
        ug, vg = int(nodes_t[sid[i], u]), int(nodes_t[sid[i], v])
        iu, iv = edge_index.t()[esrc[i]].tolist()
        assert (ug, vg) == (iu, iv) or (ug, vg) == (iv, iu)
    assert sptr.tolist() == [0, m, 2 * m]
    # keyword call exactly as gps/experiment.py:882-883 / wl_vocab.py:144-151 do it
    out = ugs_sampler.sample_batch(edge_index=edge_index, ptr=ptr, m_per_graph=m, k=k, mode="sample", seed=7)
    assert len(out) == 5 and all(not t.is_cuda and t.dtype == torch.int64 and t.is_contiguous() for t in out)
    # CUDA inputs are accepted and keep the outputs on the device
    out = ugs_sampler.sample_batch(edge_index.cuda(), ptr.cuda(), m, k)
    assert all(t.is_cuda for t in out)


@pytest.mark.parametrize("mode", ["sample", "graph", "global"])
def test_cuda_extraction_matches_reference_golden(cuda_lib, gold, mode):
    """Feed the REAL reference's node rows (golden) to the CUDA extractor: every output array must equal what
    the reference emitted (deterministic stage 'induced subgraphs for a given node set')."""
    from ugs_b200.engine import DeviceBatch
    for name in [str(x) for x in gold["names"]]:
        ei, ptr = gold[f"{name}/edge_index"], gold[f"{name}/ptr"]
        k, m = (int(x) for x in gold[f"{name}/km"])
        b = DeviceBatch(torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda(), k, int(ptr[0]), int(ptr[-1] - ptr[0]))
        nodes = torch.from_numpy(gold[f"{name}/{mode}/nodes"]).cuda()
        e_idx, e_ptr, e_src = b.extract(nodes, m, k, mode)
        assert np.array_equal(e_idx.cpu().numpy(), gold[f"{name}/{mode}/edge_index_out"]), name
        assert np.array_equal(e_ptr.cpu().numpy(), gold[f"{name}/{mode}/edge_ptr"]), name
        assert np.array_equal(e_src.cpu().numpy(), gold[f"{name}/{mode}/edge_src"]), name


def test_cuda_tables_match_reference_golden(cuda_lib, gold):
    from ugs_b200.engine import DeviceBatch
    for name in [str(x) for x in gold["names"]]:
        ei, ptr = gold[f"{name}/edge_index"], gold[f"{name}/ptr"]
        k = int(gold[f"{name}/km"][0])
        b = DeviceBatch(torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda(), k, int(ptr[0]), int(ptr[-1] - ptr[0]),
                        skip_lt_k=False)
        gi = b.ginfo()
        V = {nm: b.view(nm).cpu().numpy() for nm in ("cstart", "rowptr", "order", "index_of", "suffix_deg", "bucket_b",
                                                      "csr_idx", "csr_col")}
        for g in range(len(ptr) - 1):
            lo, n = int(ptr[g]), int(ptr[g + 1] - ptr[g])
            if n <= 0:
                continue
            _, gcol = per_graph_columns(ei, ptr, g)
            nb, cb = lo - int(ptr[0]), 2 * int(V["cstart"][g])
            nnz = int(gold[f"{name}/g{g}/indptr"][-1])
            assert np.array_equal(V["rowptr"][nb + g: nb + g + n + 1], gold[f"{name}/g{g}/indptr"]), (name, g)
            assert np.array_equal(V["csr_idx"][cb: cb + nnz], gold[f"{name}/g{g}/indices"]), (name, g)
            assert np.array_equal(V["csr_col"][cb: cb + nnz], gcol[gold[f"{name}/g{g}/edge_col"]]), (name, g)
            for key in ("order", "index_of", "suffix_deg", "bucket_b"):
                assert np.array_equal(V[key][nb: nb + n], gold[f"{name}/g{g}/{key}"]), (name, g, key)
            assert float(gi["Z"][g]) == float(gold[f"{name}/g{g}/Z"][0]), (name, g)


LAW_CASES = [
    ("triangle_tail", [(0, 1), (1, 2), (0, 2), (2, 3)], 4, 3),
    ("ring6", [(i, (i + 1) % 6) for i in range(6)], 6, 3),
    ("toluene_like", [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 0), (0, 6), (6, 7), (7, 8), (3, 9)], 10, 4),
]


@pytest.mark.parametrize("name,und,n,k", LAW_CASES)
def test_cuda_sampling_law_is_the_reference_law(cuda_lib, name, und, n, k):
    """Chi-square of 200k CUDA samples against the exact reference law P_ref at alpha = 0.001 (fixed seed)."""
    import ugs_sampler
    from oracle import exact_law as xl, ugs_oracle as uo
    ei = bidir(und)
    law = xl.p_ref(uo.Preproc(ei, n, k).dump(), k)
    M = 200000
    nodes = ugs_sampler.sample_batch(torch.from_numpy(ei).cuda(), torch.tensor([0, n]).cuda(), M, k, seed=2024)[0]
    cnt = Counter(frozenset(r) for r in nodes.cpu().numpy().tolist())
    stat, df, extra = xl.chi_square(cnt, law, M)
    assert extra == 0
    assert stat < chi2.ppf(0.999, df), (name, stat, df)


def test_completeness_on_qm9_shaped_graphs(cuda_lib):
    """reference tests/test_ugs_qm9_completeness.py:22-49: k in {4,5}, m=10 -- no row may contain -1 when the
    graph has a connected k-subgraph (all QM9-shaped graphs with n >= k are connected here)."""
    import ugs_sampler
    from ugs_b200 import synth
    ei, ptr = synth.make_batch("qm9", 2000, seed=21)
    n = np.diff(ptr)
    for k in (4, 5):
        nodes = ugs_sampler.sample_batch(torch.from_numpy(ei), torch.from_numpy(ptr), 10, k, mode="sample", seed=1)[0]
        bad = (nodes < 0).any(dim=1).view(-1, 10).any(dim=1).numpy()
        assert not bad[n >= k].any()
        assert bad[n < k].all()
        st = ugs_sampler.sample_batch.last_stats
        # n_incomplete counts partial rows of SAMPLED graphs; the all -1 rows of skipped graphs (n < k) are separate
        assert st["n_incomplete"] == 0
        assert st["n_skipped_graphs"] == int((n < k).sum()) == int(bad.sum())


def test_handle_api_k_quirks_and_errors(cuda_lib):
    import ugs_sampler
    ei = torch.from_numpy(bidir([(0, 1)]))
    h = ugs_sampler.create_preproc(ei, 3, 3)  # n >= k but no 3-graphlet: Z = 0 -> relaxation, partial rows
    assert not ugs_sampler.has_graphlets(h)
    nodes, e, ep, es = ugs_sampler.sample(h, 5, 3)
    assert (nodes[:, 2] == -1).all() and (nodes[:, :2] >= 0).all() and e.shape == (2, 0) and ep.tolist() == [0] * 6
    ugs_sampler.destroy_preproc(h)
    with pytest.raises(RuntimeError, match="Invalid preproc handle"):
        ugs_sampler.sample(h, 1, 3)
    with pytest.raises(RuntimeError):
        ugs_sampler.sample_batch(ei, torch.tensor([0, 3]), 2, 33)  # k above the implementation limit


def test_k_other_than_the_preprocessing_k(cuda_lib):
    """The tables are built for one k: DeviceBatch refuses another k; the handle API re-preprocesses, so
    that sample(handle, m, k2) equals a fresh create_preproc(k2) + sample."""
    import ugs_sampler
    from ugs_b200 import synth
    from ugs_b200.engine import DeviceBatch
    ei, ptr = synth.make_batch("zinc", 40, seed=3)
    b = DeviceBatch(torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda(), 4, 0, int(ptr[-1]))
    for call in (lambda: b.sample(8, 20), lambda: b.sample(8, 5), lambda: b.sample_epsilon(8, 5),
                 lambda: b.sample_rwr(8, 5), lambda: b.extract(torch.zeros((40 * 8, 5), dtype=torch.int64), 8, 5)):
        with pytest.raises(RuntimeError, match="preprocessed for k=4"):
            call()
    e0 = torch.from_numpy(ei[:, ei[0] < ptr[1]])
    n0 = int(ptr[1])
    h = ugs_sampler.create_preproc(e0, n0, 4)
    got = ugs_sampler.sample(h, 16, 20 if n0 >= 20 else 6, "local", 0, 7)
    h2 = ugs_sampler.create_preproc(e0, n0, 20 if n0 >= 20 else 6)
    want = ugs_sampler.sample(h2, 16, 20 if n0 >= 20 else 6, "local", 0, 7)
    assert all(torch.equal(a, b_) for a, b_ in zip(got, want))
    assert (got[0] >= 0).all()
    ugs_sampler.destroy_preproc(h)
    ugs_sampler.destroy_preproc(h2)


def test_ugs_debug_prints_the_reference_lines(cuda_lib, monkeypatch, capfd):
    """UGS_DEBUG=1: one PREPROC + one SAMPLE line per sampled graph and the cache line."""
    import ugs_sampler
    ei = bidir([(0, 1), (1, 2), (2, 0), (3, 4)])  # triangle (n=3) + single edge in a graph of 3 nodes
    ptr = torch.tensor([0, 3, 6, 8])  # third graph: n=2 < k -> skipped, no lines
    monkeypatch.setenv("UGS_DEBUG", "1")
    capfd.readouterr()
    ugs_sampler.sample_batch(torch.from_numpy(ei), ptr, 5, 3, mode="sample", seed=1)
    err = capfd.readouterr().err.strip().splitlines()
    assert len(err) == 5
    assert err[0].startswith("[UGS PREPROC] n=3 k=3 Z=") and err[0].endswith("/3")
    assert err[1].startswith("[UGS SAMPLE] n=3 k=3 m=5 Z=") and "relax=0 weighted=5 uniform=0 incomplete=0" in err[1]
    assert err[2] == "[UGS PREPROC] n=3 k=3 Z=0.00e+00 viable=0/3"
    assert err[3] == "[UGS SAMPLE] n=3 k=3 m=5 Z=0.00e+00 | relax=1 weighted=0 uniform=5 incomplete=5"
    assert err[4].startswith("[UGS CACHE] hits=0 misses=2 ")
    monkeypatch.delenv("UGS_DEBUG")
    ugs_sampler.sample_batch(torch.from_numpy(ei), ptr, 5, 3, mode="sample", seed=1)
    assert capfd.readouterr().err == ""


@pytest.mark.parametrize("mode", ["sample", "graph", "global"])
def test_packed_host_delivery_equals_plain_copy(cuda_lib, mode, monkeypatch):
    """CPU outputs are split between a plain int64 DMA and a packed transfer (16-bit deltas or int32 / one byte per
    edge entry) widened on the host (csrc/ugs_hostio.cu, csrc/ugs_hostdeliver.cpp): bit-identical to the plain int64
    copy, still pinned, in every mode, at every split, for every sampler that returns the 5-tuple."""
    import epsilon_uniform_sampler
    import rwr_sampler
    import ugs_sampler
    from ugs_b200 import synth
    ei, ptr = synth.make_batch("zinc", 3000, seed=2)
    tei, tptr = torch.from_numpy(ei), torch.from_numpy(ptr)
    monkeypatch.setattr(ugs_sampler, "_PACKED_MIN_BYTES", 0)
    for mod, kw in ((ugs_sampler, {}), (epsilon_uniform_sampler, dict(epsilon=0.5)), (rwr_sampler, {})):
        monkeypatch.setenv("UGS_B200_PACKED_D2H", "0")
        plain = mod.sample_batch(tei, tptr, 37, 8, mode, 5, **kw)
        monkeypatch.delenv("UGS_B200_PACKED_D2H")
        for frac in (None, "0", "0.37", "0.999", "1"):  # model default / all packed / mixed / almost all DMA / all DMA
            if frac is None:
                monkeypatch.delenv("UGS_B200_DMA_FRACTION", raising=False)
            else:
                monkeypatch.setenv("UGS_B200_DMA_FRACTION", frac)
            packed = mod.sample_batch(tei, tptr, 37, 8, mode, 5, **kw)
            assert ugs_sampler._to_host_batch.last_delivery.startswith("dma_fraction")
            for a, b in zip(packed, plain):
                assert a.dtype == torch.int64 and a.is_pinned() and not a.is_cuda
                assert torch.equal(a, b), (mod.__name__, frac)
        monkeypatch.delenv("UGS_B200_DMA_FRACTION", raising=False)
    # one graph with 120k columns: source columns of neighbouring entries are far apart -> the 16-bit frame-of-reference
    # packing overflows and the int32 wire format takes over
    from ugs_b200 import synth as _synth
    bei, bptr = _synth.barabasi_albert(20000, 3, 7)
    monkeypatch.setenv("UGS_B200_DMA_FRACTION", "0.25")
    a = ugs_sampler.sample_batch(torch.from_numpy(bei), torch.from_numpy(bptr), 3000, 5, mode, 2)
    assert "(int32 /" in ugs_sampler._to_host_batch.last_delivery  # int32 source columns, not 16-bit deltas
    monkeypatch.setenv("UGS_B200_PACKED_D2H", "0")
    b = ugs_sampler.sample_batch(torch.from_numpy(bei), torch.from_numpy(bptr), 3000, 5, mode, 2)
    monkeypatch.delenv("UGS_B200_PACKED_D2H")
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    # odd sizes / a single host thread / k > 16 (two int32 per entry) / empty edge list
    monkeypatch.setenv("UGS_B200_HOST_THREADS", "1")
    ei2, ptr2 = random_batch(np.random.default_rng(9), 23, 1, 60, deg=2.5, big=[(3, 200)])
    for k, m in ((17, 3), (3, 1)):
        a = ugs_sampler.sample_batch(torch.from_numpy(ei2), torch.from_numpy(ptr2), m, k, mode, 1)
        monkeypatch.setenv("UGS_B200_PACKED_D2H", "0")
        b = ugs_sampler.sample_batch(torch.from_numpy(ei2), torch.from_numpy(ptr2), m, k, mode, 1)
        monkeypatch.delenv("UGS_B200_PACKED_D2H")
        assert all(torch.equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("k,m,wire_expected", [(8, 37, True), (6, 64, True), (5, 17, False)])
def test_wire_emission_equals_plain_copy(cuda_lib, k, m, wire_expected, monkeypatch):
    """Large CPU results in mode "sample" with k <= 8: the emission pass writes the host-transfer format itself
    (ugs_emit_edges_wire: one byte per entry, 8-bit source-column deltas against per-block graph bases) and
    ugs_deliver_outputs_wire widens it -- bit-identical to the int64 emission + plain copy. Batches with skipped graphs,
    multi-edges / self loops and graphs above 32 vertices in between; a batch with a general-path graph and
    UGS_B200_WIRE_EMIT=0 take the int64 emission."""
    import ugs_sampler
    from ugs_b200 import synth
    monkeypatch.setattr(ugs_sampler, "_PACKED_MIN_BYTES", 0)
    monkeypatch.setattr(ugs_sampler._engine, "fused_eligible", lambda *a: False)  # the staged path, as for large batches
    rng = np.random.default_rng(3 * k + m)
    # molecules, graphs smaller than k in between (no entries: the next graph's columns start further away)
    parts = [synth.make_batch("zinc", 1500, seed=k), random_batch(rng, 25, 1, 3, deg=1.0),
             synth.make_batch("qm9", 700, seed=m)]
    ei, ptr = synth.concat_batches(parts)
    tei, tptr = torch.from_numpy(ei), torch.from_numpy(ptr)
    monkeypatch.setenv("UGS_B200_PACKED_D2H", "0")
    plain = ugs_sampler.sample_batch(tei, tptr, m, k, "sample", 11)
    monkeypatch.delenv("UGS_B200_PACKED_D2H")
    wire = ugs_sampler.sample_batch(tei, tptr, m, k, "sample", 11)
    # few entries per graph (k = 5, m = 17): a 1024-entry block would span many graphs' columns, the int64 emission is taken
    assert ("emitted in the wire format" in ugs_sampler._to_host_batch.last_delivery) == wire_expected
    for a, b in zip(wire, plain):
        assert a.dtype == torch.int64 and a.is_pinned() and torch.equal(a, b)
    assert ugs_sampler.sample_batch.last_stats["n_edge_entries"] == plain[4].numel()
    # multigraphs with hundreds of columns between the molecules: wherever a 1024-entry block spans more than 254 columns
    # the delivery reports it and the call takes the int64 emission; either way the result is the plain copy's
    parts.insert(1, random_batch(rng, 40, 1, 45, deg=2.5, weird=True))
    ei_w, ptr_w = synth.concat_batches(parts)
    monkeypatch.setenv("UGS_B200_PACKED_D2H", "0")
    plain_w = ugs_sampler.sample_batch(torch.from_numpy(ei_w), torch.from_numpy(ptr_w), m, k, "sample", 13)
    monkeypatch.delenv("UGS_B200_PACKED_D2H")
    got_w = ugs_sampler.sample_batch(torch.from_numpy(ei_w), torch.from_numpy(ptr_w), m, k, "sample", 13)
    assert all(torch.equal(a, b) for a, b in zip(got_w, plain_w))
    monkeypatch.setenv("UGS_B200_WIRE_EMIT", "0")
    classic = ugs_sampler.sample_batch(tei, tptr, m, k, "sample", 11)
    assert "emitted in the wire format" not in ugs_sampler._to_host_batch.last_delivery
    monkeypatch.delenv("UGS_B200_WIRE_EMIT")
    assert all(torch.equal(a, b) for a, b in zip(classic, plain))
    # graphs of 55-64 vertices with up to 320 columns each and thousands of entries per graph: eligible by the entry count,
    # but a 1024-entry block spans more than 254 columns -> the delivery reports it, the int64 emission takes over
    ei_d, ptr_d = random_batch(rng, 200, 55, 64, deg=2.5, weird=False)
    monkeypatch.setenv("UGS_B200_PACKED_D2H", "0")
    plain_d = ugs_sampler.sample_batch(torch.from_numpy(ei_d), torch.from_numpy(ptr_d), 64, k, "sample", 14)
    monkeypatch.delenv("UGS_B200_PACKED_D2H")
    got_d = ugs_sampler.sample_batch(torch.from_numpy(ei_d), torch.from_numpy(ptr_d), 64, k, "sample", 14)
    assert "emitted in the wire format" not in ugs_sampler._to_host_batch.last_delivery
    assert all(torch.equal(a, b) for a, b in zip(got_d, plain_d))
    # one general-path graph (70 vertices) in the batch: not eligible, the int64 path delivers
    ei2, ptr2 = synth.concat_batches(parts + [random_batch(rng, 3, 1, 30, deg=2.5, big=[(1, 70)])])
    monkeypatch.setenv("UGS_B200_PACKED_D2H", "0")
    plain2 = ugs_sampler.sample_batch(torch.from_numpy(ei2), torch.from_numpy(ptr2), m, k, "sample", 12)
    monkeypatch.delenv("UGS_B200_PACKED_D2H")
    mixed = ugs_sampler.sample_batch(torch.from_numpy(ei2), torch.from_numpy(ptr2), m, k, "sample", 12)
    assert "emitted in the wire format" not in ugs_sampler._to_host_batch.last_delivery
    assert all(torch.equal(a, b) for a, b in zip(mixed, plain2))


def test_large_pageable_inputs_are_staged_by_host_threads(cuda_lib, monkeypatch):
    """The trainer hands over pageable `.cpu()` tensors (experiment.py:877-878): above 32 MB they reach the device through a
    cached pinned staging buffer filled by the delivery's host threads (ugs_host_copy). Same results as with pinned
    inputs, call after call (the staging buffer is reused), also for a non-contiguous edge_index (plain copy)."""
    import ugs_sampler
    from ugs_b200 import synth
    monkeypatch.setattr(ugs_sampler, "_STAGE_MIN_BYTES", 1 << 20)
    ei, ptr = synth.make_batch("zinc", 6000, seed=3)
    assert ei.nbytes > ugs_sampler._STAGE_MIN_BYTES
    pinned = ugs_sampler.sample_batch(torch.from_numpy(ei).pin_memory(), torch.from_numpy(ptr).pin_memory(), 4, 5, "sample", 9)
    for rep in range(3):
        got = ugs_sampler.sample_batch(torch.from_numpy(ei.copy()), torch.from_numpy(ptr.copy()), 4, 5, "sample", 9)
        assert all(torch.equal(a, b) for a, b in zip(got, pinned))
    ei2, ptr2 = synth.make_batch("zinc", 7000, seed=4)   # a larger one: the staging buffer grows
    a = ugs_sampler.sample_batch(torch.from_numpy(ei2), torch.from_numpy(ptr2), 3, 4, "sample", 1)
    b = ugs_sampler.sample_batch(torch.from_numpy(ei2).pin_memory(), torch.from_numpy(ptr2).pin_memory(), 3, 4, "sample", 1)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    strided = torch.from_numpy(np.ascontiguousarray(ei.T)).t()   # [2, E] view of an [E, 2] array
    assert not strided.is_contiguous()
    c = ugs_sampler.sample_batch(strided, torch.from_numpy(ptr), 4, 5, "sample", 9)
    assert all(torch.equal(x, y) for x, y in zip(c, pinned))


def test_concurrent_calls_from_python_threads(cuda_lib, monkeypatch):
    """ctypes releases the GIL, so two Python threads can be inside libugs_b200 at the same time (the reference holds
    the GIL for the whole call): results must be the ones of the serial calls."""
    import threading
    import ugs_sampler
    from ugs_b200 import synth
    monkeypatch.setattr(ugs_sampler, "_PACKED_MIN_BYTES", 0)
    jobs = []
    for i in range(4):
        ei, ptr = synth.make_batch("zinc", 800 + 100 * i, seed=20 + i)
        jobs.append((torch.from_numpy(ei), torch.from_numpy(ptr), 16 + i, 5 + (i % 3), 100 + i))
    serial = [ugs_sampler.sample_batch(e, p, m, k, "sample", s) for e, p, m, k, s in jobs]
    got = [None] * len(jobs)

    def run(i):
        e, p, m, k, s = jobs[i]
        for _ in range(3):
            got[i] = ugs_sampler.sample_batch(e, p, m, k, "sample", s)

    th = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
    [t.start() for t in th]
    [t.join() for t in th]
    for a, b in zip(serial, got):
        assert all(torch.equal(x, y) for x, y in zip(a, b))
