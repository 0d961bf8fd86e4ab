"""One invocation of every kernel behind the SURVEY section 8(f) rows on a ZINC-shaped shard (for ncu):
encoder-input gathers, rwr / uniform-enumeration samplers, presample stitching, WL hashing, packed host delivery."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import numpy as np
import torch
from ugs_b200 import encode, synth, wl
from ugs_b200.engine import DeviceBatch
from ugs_b200.presample import PresampleStore
import ugs_sampler

G = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
k, m = 8, 64
ei, ptr = synth.make_batch("zinc", G, seed=1234)
dei, dptr = torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda()
N, E = int(ptr[-1]), ei.shape[1]
b = DeviceBatch(dei, dptr, k, 0, N)
x, ea = torch.randn(N, 64, device="cuda"), torch.randn(E, 16, device="cuda")
for rep in range(2):
    (nodes, eidx, eptr, sptr, esrc), _ = b.sample(m, k, "sample", seed=rep)
    encode.encoder_inputs(x, ea, nodes, eidx, eptr, esrc)
    encode.sample_encoder_inputs(b, x, ea, m, k, seed=rep)
    b.sample_rwr(m, k, "sample", seed=rep)
    bq = DeviceBatch(dei, dptr, 4, 0, N)
    bq.sample_uniform_enum(16, 4, "sample", seed=rep)
    wl.wl_digests(nodes, eidx, eptr, 3)
    store = PresampleStore.build(dei, dptr, 16, 4, 7)
    gi = torch.arange(0, G, 7)[:2048]
    sizes = np.diff(ptr)[gi.numpy()]
    bptr = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)])).cuda()
    cs, ce = np.searchsorted(ei[0], ptr[:-1]), np.searchsorted(ei[0], ptr[1:])
    bei = torch.from_numpy(np.concatenate([ei[:, cs[g]:ce[g]] - ptr[g] + int(bptr[i]) for i, g in enumerate(gi.tolist())], axis=1)).cuda()
    store.load(gi.cuda(), bptr, bei)
    ugs_sampler._to_host_batch((nodes, eidx, eptr, sptr, esrc), "sample", k)
    torch.cuda.synchronize()
print("ok")
