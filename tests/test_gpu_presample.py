"""GPU: device presample store against a line-by-line restatement of the reference's Python
cache -- Experiment._setup_presampling / _load_from_presample_cache (gps/experiment.py:379-440,895-995) -- driven by
the drop-in `ugs_sampler.sample_batch` exactly as the reference drives its sampler (one call per dataset graph)."""
import numpy as np
import pytest
import torch

from helpers import random_batch

pytestmark = pytest.mark.gpu


def _dataset(seed, G=40):
    """A 'dataset': list of per-graph local edge_index tensors + node counts, plus its concatenation as one batch."""
    rng = np.random.default_rng(seed)
    ei, ptr = random_batch(rng, G, 1, 30, deg=2.2, weird=False, big=[(5, 90)])
    graphs = []
    for g in range(G):
        lo, hi = int(ptr[g]), int(ptr[g + 1])
        msk = (ei[0] >= lo) & (ei[0] < hi)
        graphs.append((torch.from_numpy(ei[:, msk] - lo), hi - lo))
    return graphs, torch.from_numpy(ei), torch.from_numpy(ptr)


def reference_setup_presampling(sampler, graphs, m, k, seed):
    """experiment.py:400-432 for one split."""
    cache, total_presampled = {}, 0
    for i, (edge_index, num_nodes) in enumerate(graphs):
        ptr = torch.tensor([0, num_nodes], dtype=torch.long)
        graph_seed = seed + total_presampled
        try:
            nodes_t, edge_index_t, edge_ptr_t, sample_ptr_t, edge_src_t = sampler(edge_index, ptr, m, k, mode="sample",
                                                                                  seed=graph_seed)
            cache[i] = {'nodes': nodes_t.clone(), 'edge_index': edge_index_t.clone(), 'edge_ptr': edge_ptr_t.clone(),
                        'edge_src': edge_src_t.clone()}
        except Exception:
            cache[i] = None
        total_presampled += 1
    return cache


def reference_load_from_presample_cache(cache, graph_indices, ptr, batch_edge_index, m, k):
    """experiment.py:920-990 verbatim in behaviour."""
    num_graphs = len(graph_indices)
    all_nodes, all_edges, all_edge_src = [], [], []
    edge_ptr_list, sample_ptr_list = [0], [0]
    cumulative_edge_offset = cumulative_sample_offset = 0
    orig_edge_offsets = [0]
    for g in range(num_graphs - 1):
        node_start, node_end = ptr[g].item(), ptr[g + 1].item()
        edge_mask = (batch_edge_index[0] >= node_start) & (batch_edge_index[0] < node_end)
        orig_edge_offsets.append(orig_edge_offsets[-1] + edge_mask.sum().item())
    for g_idx, orig_idx in enumerate(graph_indices):
        ps = cache.get(orig_idx)
        nodes_g, edges_g, edge_ptr_g, edge_src_g = ps['nodes'], ps['edge_index'], ps['edge_ptr'], ps['edge_src']
        n_samples, n_edges = nodes_g.shape[0], edges_g.shape[1]
        all_nodes.append(nodes_g + ptr[g_idx].item())
        all_edges.append(edges_g)
        all_edge_src.append(edge_src_g + orig_edge_offsets[g_idx])
        for i in range(n_samples):
            edge_ptr_list.append(cumulative_edge_offset + edge_ptr_g[i + 1].item())
        cumulative_edge_offset += n_edges
        cumulative_sample_offset += n_samples
        sample_ptr_list.append(cumulative_sample_offset)
    return (torch.cat(all_nodes, dim=0), torch.cat(all_edges, dim=1), torch.tensor(edge_ptr_list, dtype=torch.long),
            torch.tensor(sample_ptr_list, dtype=torch.long), torch.cat(all_edge_src, dim=0))


@pytest.mark.parametrize("k,m", [(3, 5), (4, 16), (6, 3)])
def test_store_equals_the_reference_python_cache(cuda_lib, k, m):
    import ugs_sampler
    from ugs_b200.presample import PresampleStore
    graphs, ei, ptr = _dataset(7)
    seed = 1000
    cache = reference_setup_presampling(ugs_sampler.sample_batch, graphs, m, k, seed)
    store = PresampleStore.build(ei.cuda(), ptr.cuda(), m, k, seed)
    assert store.num_graphs == len(graphs)
    rng = np.random.default_rng(1)
    for trial in range(4):
        idx = rng.permutation(len(graphs))[: int(rng.integers(1, 17))].tolist()
        if trial == 3:
            idx = idx + idx[:2]  # the same dataset graph twice in one mini-batch
        # the mini-batch as PyG would collate it
        sizes = [graphs[i][1] for i in idx]
        bptr = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.long)
        bei = torch.cat([graphs[i][0] + int(bptr[b]) for b, i in enumerate(idx)], dim=1)
        want = reference_load_from_presample_cache(cache, idx, bptr, bei, m, k)
        got = store.load(torch.tensor(idx), bptr.cuda(), bei.cuda())
        for w, g_, nm in zip(want, got, ["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"]):
            assert torch.equal(g_.cpu(), w), (trial, nm)
        # edge_src really indexes the mini-batch's columns: the column joins the two sampled endpoints
        nodes, eidx, eptr, sptr, esrc = [t.cpu() for t in store.load(torch.tensor(idx), bptr.cuda(), bei.cuda(),
                                                                      keep_padding=True)]
        owner = torch.repeat_interleave(torch.arange(nodes.shape[0]), eptr[1:] - eptr[:-1])
        u = nodes[owner, eidx[0]]
        v = nodes[owner, eidx[1]]
        cols = bei[:, esrc]
        assert (((cols[0] == u) & (cols[1] == v)) | ((cols[0] == v) & (cols[1] == u))).all()
        assert ((nodes == -1) == (want[0] - bptr[:-1].repeat_interleave(m).unsqueeze(1) == -1)).all()


def test_store_edge_cases(cuda_lib):
    from ugs_b200.presample import PresampleStore
    graphs, ei, ptr = _dataset(3, G=10)
    store = PresampleStore.build(ei.cuda(), ptr.cuda(), 4, 3, 5)
    out = store.load(torch.zeros(0, dtype=torch.long), torch.zeros(1, dtype=torch.long).cuda(),
                     torch.zeros((2, 0), dtype=torch.long).cuda())
    assert out[0].shape == (0, 3) and out[1].shape == (2, 0) and out[2].tolist() == [0] and out[3].tolist() == [0]
    with pytest.raises(RuntimeError, match="out of range"):
        store.load(torch.tensor([10]), torch.tensor([0, 5]).cuda(), torch.zeros((2, 0), dtype=torch.long).cuda())
    with pytest.raises(RuntimeError, match="len\\(graph_idx\\)"):
        store.load(torch.tensor([1, 2]), torch.tensor([0, 5]).cuda(), torch.zeros((2, 0), dtype=torch.long).cuda())
    assert store.nbytes() > 0
