"""CPU, world_size 2, gloo: the N>1 host logic -- partition by graph, per-rank Philox graph offsets, and the
one collective (all-gather of the statistics vector). The per-rank 'sampler' here is the CPU oracle standing in
for the GPU (tests may use oracle/), so the test also proves that sharded sampling == unsharded sampling."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ei, ptr, m, k, seed, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "ss-gnn_b200"))
    from oracle import ugs_oracle as uo
    from ugs_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ei_s, ptr_s, gbase, c0 = sharding.shard(ptr, ei, rank, world, m, k)
    nodes, eidx, eptr, sptr, esrc = uo.sample_batch(ei_s, ptr_s, m, k, "sample", seed, graph_index_base=gbase)
    stats = {"n_graphlets": nodes.shape[0], "n_edge_entries": eidx.shape[1],
             "n_incomplete": int((nodes < 0).any(axis=1).sum())}
    g = sharding.gather_stats(stats)
    offs = sharding.global_edge_offsets(g["per_rank"])
    ret[rank] = dict(nodes=nodes, eidx=eidx, eptr=eptr + offs[rank], esrc=esrc + c0, total=g["total"], gbase=gbase)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_rank():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "ss-gnn_b200"))
    from oracle import ugs_oracle as uo
    from ugs_b200 import sharding, synth
    ei, ptr = synth.make_batch("mutag", 37, seed=4)
    m, k, seed, world = 6, 4, 11, 2
    parts = sharding.partition(ptr, ei, world, m, k)
    assert parts[0][0] == 0 and parts[-1][1] == 37 and parts[0][1] == parts[1][0]
    assert parts[0][3] == parts[1][2] and parts[-1][3] == ei.shape[1]
    costs = [sharding.graph_cost(ptr, np.searchsorted(np.searchsorted(ptr, ei[0], side="right") - 1, np.arange(38)),
                                 m, k)[a:b].sum() for a, b, _, _ in parts]
    assert abs(costs[0] - costs[1]) / sum(costs) < 0.15  # balanced by the cost proxy

    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ei, ptr, m, k, seed, ret), nprocs=world, join=True)
    whole = uo.sample_batch(ei, ptr, m, k, "sample", seed)
    nodes = np.concatenate([ret[r]["nodes"] for r in range(world)])
    eidx = np.concatenate([ret[r]["eidx"] for r in range(world)], axis=1)
    esrc = np.concatenate([ret[r]["esrc"] for r in range(world)])
    eptr = np.concatenate([ret[0]["eptr"]] + [ret[r]["eptr"][1:] for r in range(1, world)])
    assert np.array_equal(nodes, whole[0]) and np.array_equal(eidx, whole[1])
    assert np.array_equal(eptr, whole[2]) and np.array_equal(esrc, whole[4])
    assert ret[0]["total"] == ret[1]["total"]
    assert ret[0]["total"]["n_graphlets"] == 37 * m and ret[0]["total"]["n_edge_entries"] == whole[1].shape[1]


def _sample_worker(rank, world, port, ei, ptr, m, k, seed, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "ss-gnn_b200"))
    from oracle import ugs_oracle as uo
    from ugs_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, m_local = sharding.shard_samples(m, rank, world)
    nodes, eidx, eptr, sptr, esrc = uo.sample_batch(ei, ptr, m_local, k, "sample", seed, sample_index_base=base)
    g = sharding.gather_stats({"n_graphlets": nodes.shape[0], "n_edge_entries": eidx.shape[1]})
    ret[rank] = dict(nodes=nodes, eidx=eidx, eptr=eptr, esrc=esrc, base=base, m_local=m_local, total=g["total"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sample_sharding_of_one_replicated_graph():
    """Section 8e, power-law configs: ONE graph replicated on every rank, the m samples split across ranks.
    Rank r's rows are rows [base_r, base_r + m_r) of the single-rank call (Philox counter = global sample index)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "ss-gnn_b200"))
    from oracle import ugs_oracle as uo
    from ugs_b200 import sharding, synth
    assert [sharding.shard_samples(10, r, 4) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert sharding.shard_samples(0, 1, 2) == (0, 0)
    ei, ptr = synth.barabasi_albert(400, 3, 5)
    m, k, seed, world = 37, 5, 3, 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sample_worker, args=(world, port, ei, ptr, m, k, seed, ret), nprocs=world, join=True)
    whole = uo.sample_batch(ei, ptr, m, k, "sample", seed)
    assert ret[0]["base"] == 0 and ret[1]["base"] == ret[0]["m_local"] and ret[0]["m_local"] + ret[1]["m_local"] == m
    assert np.array_equal(np.concatenate([ret[0]["nodes"], ret[1]["nodes"]]), whole[0])
    assert np.array_equal(np.concatenate([ret[0]["eidx"], ret[1]["eidx"]], axis=1), whole[1])
    assert np.array_equal(np.concatenate([ret[0]["esrc"], ret[1]["esrc"]]), whole[4])
    assert np.array_equal(np.concatenate([ret[0]["eptr"], ret[1]["eptr"][1:] + ret[0]["eptr"][-1]]), whole[2])
    assert ret[0]["total"]["n_graphlets"] == m


def test_gather_stats_without_process_group():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "ss-gnn_b200"))
    from ugs_b200 import sharding
    g = sharding.gather_stats({"n_graphlets": 5, "n_edge_entries": 7})
    assert g["total"]["n_graphlets"] == 5 and g["per_rank"] == [[5, 7, 0, 0, 0, 0]]
    assert sharding.global_edge_offsets([[1, 10, 0, 0, 0, 0], [1, 4, 0, 0, 0, 0]]) == [0, 10]
