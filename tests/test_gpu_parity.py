"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle, bit-exact.

Deterministic stages are compared array by array; sampled node rows are compared
bit-exactly too because the oracle and the kernels share the Philox draw rule (oracle/ugs_oracle.c header).
"""
import numpy as np
import pytest
import torch

from helpers import bidir, per_graph_columns, random_batch

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _batch(ei, ptr, k, **kw):
    from ugs_b200.engine import DeviceBatch
    return DeviceBatch(_dev(ei), _dev(ptr), k, int(ptr[0]), int(ptr[-1] - ptr[0]), **kw)


def _check_tables(b, ei, ptr, k):
    from oracle import ugs_oracle as uo
    gi = b.ginfo()
    V = {name: b.view(name).cpu().numpy() for name in
         ("cstart", "rowptr", "order", "index_of", "suffix_deg", "bucket_b", "cum", "csr_idx", "csr_col")}
    base = int(ptr[0])
    for g in range(len(ptr) - 1):
        lo, hi = int(ptr[g]), int(ptr[g + 1])
        n = hi - lo
        if n <= 0 or n < k:
            assert int(gi["cls"][g]) == 0
            continue
        assert int(gi["cls"][g]) in (1, 2)
        lcols, gcol = per_graph_columns(ei, ptr, g)
        P = uo.Preproc(lcols, n, k)
        d = P.dump()
        nb = lo - base
        rp = V["rowptr"][nb + g: nb + g + n + 1].astype(np.int64)
        cb = 2 * int(V["cstart"][g])
        assert np.array_equal(rp, d["indptr"]), f"graph {g}: rowptr"
        nnz = int(d["indptr"][-1])
        assert int(gi["nnz"][g]) == nnz
        assert np.array_equal(V["csr_idx"][cb: cb + nnz], d["indices"]), f"graph {g}: csr indices"
        assert np.array_equal(V["csr_col"][cb: cb + nnz], gcol[d["edge_col"]]), f"graph {g}: csr columns"
        assert np.array_equal(V["order"][nb: nb + n], d["order"]), f"graph {g}: order"
        assert np.array_equal(V["index_of"][nb: nb + n], d["index_of"]), f"graph {g}: index_of"
        assert np.array_equal(V["suffix_deg"][nb: nb + n], d["suffix_deg"]), f"graph {g}: suffix_deg"
        assert np.array_equal(V["bucket_b"][nb: nb + n], d["bucket_b"]), f"graph {g}: bucket_b"
        assert float(gi["Z"][g]) == d["Z"], f"graph {g}: Z"
        assert int(gi["level"][g]) == d["level"]
        assert np.array_equal(V["cum"][nb: nb + n], d["cum"]), f"graph {g}: root CDF"


@pytest.mark.parametrize("seed,k", [(0, 3), (1, 4), (2, 6), (3, 1), (4, 8)])
def test_preproc_tables_bit_exact_small_graphs(cuda_lib, seed, k):
    rng = np.random.default_rng(seed)
    ei, ptr = random_batch(rng, 60, 1, 40, deg=2.5)
    _check_tables(_batch(ei, ptr, k), ei, ptr, k)


@pytest.mark.parametrize("seed,k", [(10, 3), (11, 6), (12, 20)])
def test_preproc_tables_bit_exact_mixed_paths(cuda_lib, seed, k):
    """Graphs above 64 nodes / 384 columns (and any graph when k > 16) take the block-per-graph path."""
    rng = np.random.default_rng(seed)
    ei, ptr = random_batch(rng, 24, 1, 90, deg=3.0, big=[(3, 300), (7, 1500), (11, 65)])
    b = _batch(ei, ptr, k)
    assert b.n_general > 0
    _check_tables(b, ei, ptr, k)


def test_preproc_hub_rows_sorted_by_column(cuda_lib):
    """A star + clique: rows far longer than a warp (block-level sort of hub rows)."""
    n = 3000
    und = [(0, i) for i in range(1, n)] + [(i, i + 1) for i in range(1, 200)]
    ei = bidir(und)
    ptr = np.array([0, n], np.int64)
    _check_tables(_batch(ei, ptr, 5), ei, ptr, 5)


def _oracle_batch(ei, ptr, m, k, mode, seed, **kw):
    from oracle import ugs_oracle as uo
    return uo.sample_batch(ei, ptr, m, k, mode, seed, **kw)


@pytest.mark.parametrize("mode", ["sample", "graph", "global"])
@pytest.mark.parametrize("seed,k,m", [(20, 3, 7), (21, 6, 64), (22, 8, 33), (23, 1, 5)])
def test_sample_batch_bit_exact_vs_oracle(cuda_lib, mode, seed, k, m):
    rng = np.random.default_rng(seed)
    ei, ptr = random_batch(rng, 50, 1, 45, deg=2.5, big=[(5, 120), (9, 70)])
    b = _batch(ei, ptr, k)
    outs, stats = b.sample(m, k, mode, seed=seed)
    ref = _oracle_batch(ei, ptr, m, k, mode, seed)
    names = ["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"]
    for name, a, r in zip(names, outs, ref):
        a = a.cpu().numpy()
        assert a.shape == r.shape, name
        if name == "nodes" and not np.array_equal(a, r):
            bad = np.nonzero((a != r).any(axis=1))[0]
            gids = sorted(set(int(x) // m for x in bad))
            sizes = [int(ptr[g + 1] - ptr[g]) for g in gids]
            raise AssertionError(f"nodes differ in {len(bad)} rows, graphs {gids[:10]} sizes {sizes[:10]}: "
                                 f"gpu {a[bad[0]]} oracle {r[bad[0]]}")
        assert np.array_equal(a, r), f"{name} differs (mode={mode})"
    assert stats["n_edge_entries"] == ref[1].shape[1]
    # n_incomplete = partial rows of sampled graphs; skipped graphs (n < k) contribute m all -1 rows each
    assert stats["n_incomplete"] == int((ref[0] < 0).any(axis=1).sum()) - m * stats["n_skipped_graphs"]
    assert stats["n_trials"] == (len(ptr) - 1 - stats["n_skipped_graphs"]) * m


def test_both_paths_give_identical_samples(cuda_lib):
    """k > 16 forces the general path; the same graphs with k <= 16 take the bitmask path. Compare each with
    the oracle (which does not know about paths) on the same inputs."""
    rng = np.random.default_rng(5)
    ei, ptr = random_batch(rng, 30, 20, 60, deg=3.0, weird=False)
    for k in (5, 17):
        b = _batch(ei, ptr, k)
        assert (b.n_general == 0) == (k <= 16)
        outs, _ = b.sample(16, k, "sample", seed=9)
        ref = _oracle_batch(ei, ptr, 16, k, "sample", 9)
        for a, r in zip(outs, ref):
            assert np.array_equal(a.cpu().numpy(), r)


def test_unsorted_columns_take_the_perm_path(cuda_lib):
    rng = np.random.default_rng(7)
    ei, ptr = random_batch(rng, 20, 3, 30, deg=2.0)
    perm = rng.permutation(ei.shape[1])
    ei2 = np.ascontiguousarray(ei[:, perm])
    b = _batch(ei2, ptr, 4)
    _check_tables(b, ei2, ptr, 4)  # the first access notices the failed fast segmentation and rebuilds
    assert b.perm is not None
    outs, _ = b.sample(9, 4, "global", seed=1)
    ref = _oracle_batch(ei2, ptr, 9, 4, "global", 1)
    for a, r in zip(outs, ref):
        assert np.array_equal(a.cpu().numpy(), r)


def test_graph_index_base_shards_like_the_whole(cuda_lib):
    """Philox is keyed on the GLOBAL graph index: sampling a shard with graph_index_base equals the
    corresponding slice of sampling the whole batch (what makes results independent of the rank count)."""
    rng = np.random.default_rng(8)
    ei, ptr = random_batch(rng, 16, 5, 30, deg=2.2, weird=False)
    k, m = 4, 8
    whole, _ = _batch(ei, ptr, k).sample(m, k, "sample", seed=3)
    g0 = 6
    lo = int(ptr[g0])
    mask = ei[0] >= lo
    ei_s = ei[:, mask]
    ptr_s = ptr[g0:]
    shard, _ = _batch(ei_s, ptr_s, k).sample(m, k, "sample", seed=3, graph_index_base=g0)
    assert np.array_equal(shard[0].cpu().numpy(), whole[0].cpu().numpy()[g0 * m:])
    e0 = int(whole[2][g0 * m])
    assert np.array_equal(shard[1].cpu().numpy(), whole[1].cpu().numpy()[:, e0:])
    assert np.array_equal(shard[4].cpu().numpy() + int(np.nonzero(mask)[0][0]), whole[4].cpu().numpy()[e0:])


def test_sample_index_base_shards_like_the_whole(cuda_lib):
    """Philox counter word 0 is the sample's index in the whole job: drawing samples [i0, i0+m') with
    sample_index_base=i0 gives rows [i0, i0+m') of every graph of the single call (one huge graph
    replicated on all ranks, samples split). Both paths (bitmask and general) and the oracle."""
    rng = np.random.default_rng(21)
    ei, ptr = random_batch(rng, 9, 5, 40, deg=2.5, weird=False, big=[(2, 130)])
    k, m, i0, m1 = 4, 50, 17, 20
    b = _batch(ei, ptr, k)
    whole, _ = b.sample(m, k, "sample", seed=3)
    part, _ = b.sample(m1, k, "sample", seed=3, sample_index_base=i0)
    G = len(ptr) - 1
    wn = whole[0].cpu().numpy().reshape(G, m, k)
    assert np.array_equal(part[0].cpu().numpy().reshape(G, m1, k), wn[:, i0:i0 + m1])
    ref = _oracle_batch(ei, ptr, m1, k, "sample", 3, sample_index_base=i0)
    for a, r in zip(part, ref):
        assert np.array_equal(a.cpu().numpy(), r)
    with pytest.raises(RuntimeError, match="2\\^32"):
        b.sample(8, k, "sample", seed=3, sample_index_base=2 ** 32 - 4)


def test_empty_and_degenerate_inputs(cuda_lib):
    # no graphs
    b = _batch(np.zeros((2, 0), np.int64), np.array([0], np.int64), 3)
    outs, _ = b.sample(4, 3)
    assert outs[0].shape == (0, 3) and outs[2].tolist() == [0] and outs[3].tolist() == [0]
    # graphs without edges / smaller than k, m = 0
    ei = bidir([(4, 5), (5, 6), (6, 7), (4, 7)])
    ptr = np.array([0, 2, 4, 8], np.int64)
    for m in (0, 3):
        outs, _ = _batch(ei, ptr, 3).sample(m, 3)
        ref = _oracle_batch(ei, ptr, m, 3, "sample", 42)
        for a, r in zip(outs, ref):
            assert np.array_equal(a.cpu().numpy(), r)


def test_big_m_single_graph_and_large_scale_properties(cuda_lib):
    """Full-size property checks (sizes the oracle does not need to replay): every complete row is a connected
    set of distinct vertices of its own graph, edge_ptr is the exclusive scan of per-sample counts, every
    emitted entry maps back to its source column."""
    from ugs_b200 import synth
    ei, ptr = synth.make_batch("zinc", 4096, seed=3)
    k, m = 8, 64
    b = _batch(ei, ptr, k)
    (nodes, eidx, eptr, sptr, esrc), stats = b.sample(m, k, "sample", seed=11)
    G = len(ptr) - 1
    assert nodes.shape == (G * m, k) and int(eptr[-1]) == eidx.shape[1] == esrc.numel()
    assert torch.equal(sptr, torch.arange(G + 1, device=sptr.device) * m)
    assert bool((eptr[1:] >= eptr[:-1]).all())
    dptr = _dev(ptr)
    gid = torch.arange(G, device=nodes.device).repeat_interleave(m)
    lo, hi = dptr[gid], dptr[gid + 1]
    assert bool(((nodes >= lo[:, None]) & (nodes < hi[:, None])).all())  # zinc-shaped: connected, n >= 9 > k
    srt = nodes.sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())  # distinct
    # edge entries: local ids map to the source column's endpoints up to orientation
    sid = torch.bucketize(torch.arange(eidx.shape[1], device=eidx.device), eptr[1:], right=True)
    u = nodes[sid, eidx[0]]
    v = nodes[sid, eidx[1]]
    d_ei = _dev(ei)
    su, sv = d_ei[0, esrc], d_ei[1, esrc]
    assert bool((((u == su) & (v == sv)) | ((u == sv) & (v == su))).all())
    # every undirected induced edge appears 4x: counts are multiples of 4, >= 4(k-1)
    cnt = eptr[1:] - eptr[:-1]
    assert bool((cnt % 4 == 0).all()) and bool((cnt >= 4 * (k - 1)).all())
    assert stats["n_incomplete"] == 0


def test_unsorted_columns_detected_late_by_sample(cuda_lib):
    """Columns grouped by graph except for two swapped columns deep inside: the per-graph binary search may be
    fooled, the preprocessing kernels notice a foreign column, and sample() redoes the batch on the slow path."""
    rng = np.random.default_rng(17)
    ei, ptr = random_batch(rng, 30, 5, 30, deg=2.0, weird=False)
    g = np.searchsorted(ptr, ei[0], side="right") - 1
    a = int(np.nonzero(g == 3)[0][1])
    b_ = int(np.nonzero(g == 20)[0][1])
    ei2 = ei.copy()
    ei2[:, [a, b_]] = ei2[:, [b_, a]]
    for m, k in ((5, 3),):
        outs, _ = _batch(ei2, ptr, k).sample(m, k, "global", seed=2)
        ref = _oracle_batch(ei2, ptr, m, k, "global", 2)
        for x, r in zip(outs, ref):
            assert np.array_equal(x.cpu().numpy(), r)


@pytest.mark.parametrize("k", [3, 6])
def test_preproc_large_graphs_cooperative_kernel(cuda_lib, k):
    """Graphs above UGS_LARGE_N = 4096 nodes are preprocessed by the whole GPU (cooperative kernel): a Barabasi-
    Albert graph with hubs (rows far longer than a warp / the CTA sort buffer), a large random multigraph with
    self loops and dropped columns, and small graphs in the same batch."""
    from ugs_b200 import synth
    rng = np.random.default_rng(41)
    ba, pba = synth.barabasi_albert(6000, 3, seed=2)
    star = bidir([(0, i) for i in range(1, 5000)] + [(i, i + 1) for i in range(1, 300)])  # hub row of 9998 entries
    wei, wptr = random_batch(rng, 4, 5, 30, deg=2.0, big=[(1, 5000)])
    ei, ptr = synth.concat_batches([(ba, pba), (star, np.array([0, 5000])), (wei, wptr)])
    b = _batch(ei, ptr, k)
    assert b.n_large == 3 and b.n_general >= 3
    _check_tables(b, ei, ptr, k)
    outs, _ = b.sample(50, k, "sample", seed=4)
    ref = _oracle_batch(ei, ptr, 50, k, "sample", 4)
    for a, r in zip(outs, ref):
        assert np.array_equal(a.cpu().numpy(), r)


def test_column_sort_is_stable_on_many_graphs(cuda_lib):
    """The slow path's LSD radix sort (several 8-bit passes: > 256 graphs, > 4096 columns per CTA tile)."""
    rng = np.random.default_rng(23)
    ei, ptr = random_batch(rng, 700, 2, 12, deg=2.0, weird=False)
    perm = rng.permutation(ei.shape[1])
    ei2 = np.ascontiguousarray(ei[:, perm])
    b = _batch(ei2, ptr, 3)
    b.ginfo()
    assert b.perm is not None
    gid = np.searchsorted(ptr, ei2[0], side="right") - 1
    want = np.argsort(gid, kind="stable")
    assert np.array_equal(b.perm.cpu().numpy(), want)
    outs, _ = b.sample(4, 3, "sample", seed=6)
    ref = _oracle_batch(ei2, ptr, 4, 3, "sample", 6)
    for a, r in zip(outs, ref):
        assert np.array_equal(a.cpu().numpy(), r)


@pytest.mark.parametrize("n,k,m", [(100000, 6, 1000), (30000, 6, 3000), (100000, 3, 500)])
def test_powerlaw_graph_at_baseline_size_bit_exact(cuda_lib, n, k, m):
    """BASELINE configs[4] at its full size: ONE Barabasi-Albert graph with 10^5 nodes (hub rows of thousands of
    entries, L2-resident CSR, two-level cut bitmap of the general-path sampler, cooperative preprocessing kernel):
    every table and all five outputs bit-exact against the oracle."""
    from ugs_b200 import synth
    ei, ptr = synth.barabasi_albert(n, 3, seed=7)
    b = _batch(ei, ptr, k)
    assert b.n_large == 1 and b.n_general == 1
    _check_tables(b, ei, ptr, k)
    outs, stats = b.sample(m, k, "sample", seed=4)
    ref = _oracle_batch(ei, ptr, m, k, "sample", 4)
    for name, a, r in zip(["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"], outs, ref):
        assert np.array_equal(a.cpu().numpy(), r), name
    assert stats["n_incomplete"] == 0 and stats["n_edge_entries"] == ref[1].shape[1]
    # the same rows when the samples are split over two "ranks" (one huge graph is sharded by sample)
    half = m // 2
    lo_, _ = b.sample(half, k, "sample", seed=4)
    hi_, _ = b.sample(m - half, k, "sample", seed=4, sample_index_base=half)
    assert np.array_equal(torch.cat([lo_[0], hi_[0]]).cpu().numpy(), ref[0])


def test_general_path_sampler_block_boundaries(cuda_lib):
    """General-path graphs whose sizes straddle the two-level cut bitmap's block size (1024 vertices = 32 words) and
    multi-pass block scans (> 32 blocks), dense enough that the cut spans many blocks."""
    rng = np.random.default_rng(77)
    sizes = [65, 1023, 1024, 1025, 2048, 3000, 33 * 1024 + 5]
    ei, ptr = random_batch(rng, len(sizes), 5, 30, deg=6.0, weird=False, big=list(enumerate(sizes)))
    for k in (4, 9):
        b = _batch(ei, ptr, k)
        outs, _ = b.sample(200, k, "global", seed=11)
        ref = _oracle_batch(ei, ptr, 200, k, "global", 11)
        for name, a, r in zip(["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"], outs, ref):
            assert np.array_equal(a.cpu().numpy(), r), (name, k)


@pytest.mark.parametrize("m", [1, 5, 8, 9, 10, 16])
@pytest.mark.parametrize("k", [3, 4, 5, 7])
def test_small_m_packs_several_graphs_per_warp(cuda_lib, m, k):
    """m <= 16: a warp samples 2 (m <= 16) or 4 (m <= 8) graphs at once (k_sample_small_packed). Batches with every kind
    of graph in one warp -- uniform multiplicities, multi-edges / self loops (sequential fallback), graphs smaller than
    k, graphs above 32 and above 64 vertices (general path) -- bit-exact against the oracle in every mode."""
    rng = np.random.default_rng(100 * m + k)
    e1, p1 = random_batch(rng, 37, 1, 45, deg=2.5, weird=True, big=[(5, 120), (9, 70)])
    from ugs_b200 import synth
    e2, p2 = synth.make_batch("qm9", 45, seed=m + k)        # uniform multiplicity 2, n <= 29: the packed path proper
    e3, p3 = synth.make_batch("zinc", 11, seed=m)            # some graphs above 32 vertices: 64-bit masks
    ei, ptr = synth.concat_batches([(e2, p2), (e1, p1), (e3, p3)])
    for mode in ("sample", "global"):
        b = _batch(ei, ptr, k)
        outs, stats = b.sample(m, k, mode, seed=7)
        ref = _oracle_batch(ei, ptr, m, k, mode, 7)
        for name, a, r in zip(["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"], outs, ref):
            assert np.array_equal(a.cpu().numpy(), r), (name, mode)
        assert stats["n_incomplete"] == int((ref[0] < 0).any(axis=1).sum()) - m * stats["n_skipped_graphs"]
    if k <= 8:  # encoder mode (batch-level ids s*k + local, ss_gnn.py:461-462) through the packed emission kernel
        b = _batch(ei, ptr, k)
        (n_s, e_s, p_s, _, s_s), _ = b.sample(m, k, "sample", seed=7)
        (n_e, e_e, p_e, _, s_e), _ = b.sample(m, k, "encoder", seed=7)
        counts = (p_s[1:] - p_s[:-1])
        off = torch.repeat_interleave(torch.arange(n_s.shape[0], device=e_s.device), counts) * k
        assert torch.equal(e_e, e_s + off.unsqueeze(0)) and torch.equal(s_e, s_s) and torch.equal(p_e, p_s)


@pytest.mark.parametrize("k,m", [(6, 640), (3, 700), (12, 800)])
def test_many_general_path_samples_take_the_thread_per_sample_kernel(cuda_lib, k, m):
    """A batch with >= 10 000 general-path samples hands its PROTEINS-sized graphs (<= 1024 vertices, <= 3072 columns) to
    k_sample_medium (one thread per sample, tables in shared memory, induced entries counted by symmetry during the
    growth) and leaves the larger ones with the warp-per-sample kernel: same bits as the oracle either way. The graphs
    carry self loops, duplicate and one-direction columns, and sizes at the kernel's limits."""
    rng = np.random.default_rng(1000 + k)
    sizes = [65, 100, 333, 700, 1023, 1024, 1025, 1500, 90, 256]
    ei, ptr = random_batch(rng, 64, 1, 60, deg=1.4, weird=True, big=[(3 * i + 1, s_) for i, s_ in enumerate(sizes)])
    ei2, ptr2 = random_batch(rng, 3, 800, 1000, deg=3.5, weird=True)  # more than 3072 columns: not "medium"
    from ugs_b200 import synth
    ei, ptr = synth.concat_batches([(ei, ptr), (ei2, ptr2), synth.make_batch("proteins", 30, seed=k)])
    b = _batch(ei, ptr, k)
    assert b.n_general * m >= 10000
    for mode in ("sample", "global"):
        outs, stats = b.sample(m, k, mode, seed=3)
        ref = _oracle_batch(ei, ptr, m, k, mode, 3)
        for name, a, r in zip(["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"], outs, ref):
            assert np.array_equal(a.cpu().numpy(), r), (name, mode)
        assert stats["n_incomplete"] == int((ref[0] < 0).any(axis=1).sum()) - m * stats["n_skipped_graphs"]
        assert stats["n_edge_entries"] == ref[1].shape[1]


@pytest.mark.parametrize("k,m", [(8, 64), (4, 40), (5, 33)])
def test_large_batch_takes_the_pipelined_sampling_kernel(cuda_lib, k, m):
    """Batches with at least two graphs per resident warp (>= 2 x 148 SMs x 4 CTAs x 8 warps = 9472 graphs) and every tile of
    a graph in one work item are sampled by k_sample_small_pipe (persistent warps, cp.async fetch-ahead of the next graph's
    tables). Every kind of graph between the molecules -- multi-edges / self loops (CSR rows staged when the graph's turn
    comes), graphs smaller than k, graphs above 32 vertices (64-bit masks) and above 64 (general path, skipped by this
    kernel) -- all five outputs bit-exact against the oracle."""
    from ugs_b200 import synth
    rng = np.random.default_rng(1000 + k)
    parts = []
    for i in range(6):
        parts.append(synth.make_batch("zinc" if i % 2 else "qm9", 1990, seed=10 * k + i))
        parts.append(random_batch(rng, 30, 1, 45, deg=2.5, weird=True, big=[(3, 120), (17, 70)]))
    ei, ptr = synth.concat_batches(parts)
    assert len(ptr) - 1 >= 12000
    b = _batch(ei, ptr, k)
    outs, stats = b.sample(m, k, "sample", seed=5)
    ref = _oracle_batch(ei, ptr, m, k, "sample", 5)
    for name, a, r in zip(["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"], outs, ref):
        assert np.array_equal(a.cpu().numpy(), r), name
    assert stats["n_incomplete"] == int((ref[0] < 0).any(axis=1).sum()) - m * stats["n_skipped_graphs"]
