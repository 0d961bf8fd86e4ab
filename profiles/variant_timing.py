"""Device-resident timing of the sampler variants on the named shapes (CUDA events, 3 warm-up + 5 timed passes)."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import numpy as np
import torch
from ugs_b200 import synth
from ugs_b200.engine import DeviceBatch
import apx_ugs_sampler


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    out = None
    for _ in range(reps):
        out = None  # release the previous outputs first: the caching allocator then reuses their blocks
        out = fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps, out


d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
for kind, G, k, m in (("zinc", 131072, 8, 64), ("proteins", 1113, 6, 100), ("qm9", 131072, 4, 10)):
    ei, ptr = synth.make_batch(kind, G, seed=1234)
    dei, dptr = d(ei), d(ptr)
    NT = int(ptr[-1])
    ms, (o, st) = timed(lambda: DeviceBatch(dei, dptr, k, 0, NT).sample(m, k, "sample", seed=1))
    print(f"{kind:9s} G={G} k={k} m={m} ugs(ref law)      {ms:8.3f} ms  {G*m/ms/1e3:10.1f} M graphlets/s  entries/graphlet {st['n_edge_entries']/(G*m):.2f}")
    for eps in (0.1,):
        ms, (o, st) = timed(lambda: DeviceBatch(dei, dptr, k, 0, NT).sample_epsilon(m, k, "sample", seed=1, epsilon=eps))
        print(f"{kind:9s} G={G} k={k} m={m} epsilon({eps})       {ms:8.3f} ms  {G*m/ms/1e3:10.1f} M graphlets/s  trials/sample {st['n_trials']/(G*m):.2f} failed {st['n_incomplete']}")
    if kind != "proteins" and k <= 5:
        ms, (o, st) = timed(lambda: DeviceBatch(dei, dptr, k, 0, NT, variant='uniform').sample(m, k, "sample", seed=1))
        print(f"{kind:9s} G={G} k={k} m={m} uniform (A.8)      {ms:8.3f} ms  {G*m/ms/1e3:10.1f} M graphlets/s  trials/sample {st['n_trials']/(G*m):.2f} failed {st['n_incomplete']}")
    # the other reference samplers behind the same workspace
    ms, (o, st) = timed(lambda: DeviceBatch(dei, dptr, k, 0, NT).sample_rwr(m, k, "sample", seed=1, p_restart=0.2))
    print(f"{kind:9s} G={G} k={k} m={m} rwr(p=0.2)          {ms:8.3f} ms  {G*m/ms/1e3:10.1f} M graphlets/s  walk steps/sample {st['n_walk_steps']/(G*m):.1f} failed {st['n_incomplete']}")
    if True:
        ms, (o, st) = timed(lambda: DeviceBatch(dei, dptr, k, 0, NT).sample_uniform_enum(m, k, "sample", seed=1), reps=3, warm=1)
        print(f"{kind:9s} G={G} k={k} m={m} uniform(enumerate)  {ms:8.3f} ms  {G*m/ms/1e3:10.1f} M graphlets/s  connected k-subsets enumerated {st['n_enumerated']} ({st['n_enumerated']/G:.0f}/graph)")
    # WL graphlet classification of the UGS samples and the presample store (f2)
    from ugs_b200 import wl
    from ugs_b200.presample import PresampleStore
    (nodes, eidx, eptr, sptr, esrc), _ = DeviceBatch(dei, dptr, k, 0, NT).sample(m, k, "sample", seed=1)
    ms, _o = timed(lambda: wl.wl_digests(nodes, eidx, eptr, 3))
    print(f"{kind:9s} G={G} k={k} m={m} WL hash (3 iter)    {ms:8.3f} ms  {G*m/ms/1e3:10.1f} M graphlets/s")
    del nodes, eidx, eptr, sptr, esrc, _o
    Gs = min(G, 32768)
    store = PresampleStore.build(dei[:, : int(np.searchsorted(ei[0], ptr[Gs]))], dptr[: Gs + 1], m, k, 7)
    gi = torch.randperm(Gs, generator=torch.Generator().manual_seed(1))[:1024]
    sizes = np.diff(ptr)[gi.numpy()]
    bptr = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)])).cuda()
    cs = np.searchsorted(ei[0], ptr[:-1])
    ce = np.searchsorted(ei[0], ptr[1:])
    bei = torch.from_numpy(np.concatenate([ei[:, cs[g]:ce[g]] - ptr[g] + int(bptr[b]) for b, g in enumerate(gi.tolist())], axis=1)).cuda()
    ms, o = timed(lambda: store.load(gi.cuda(), bptr, bei))
    print(f"{kind:9s} presample store: {Gs} graphs = {store.nbytes()/1e6:.0f} MB in HBM; stitching a 1024-graph mini-batch ({o[1].shape[1]} entries): {ms:.3f} ms")
    del store, o
edges = np.array([[0, 1], [1, 0], [1, 2], [2, 1], [2, 0], [0, 2], [2, 3], [3, 2], [3, 4], [4, 3]], np.int64).T.copy()
ms, (s, p) = timed(lambda: apx_ugs_sampler.sample_batch(d(edges), torch.tensor([0, 10]).cuda(), 100, 3, "sample", 7, 0.5), reps=2, warm=1)
print(f"apx_ugs 5-node graph k=3 eps=0.5 m=100: {ms:.1f} ms per call, {s.shape[1]} returned (reference: 745 s, SURVEY section 6)")
