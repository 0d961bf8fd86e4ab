"""WL hashing throughput on sampler output: python profiles/wl_timing.py"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import torch
from ugs_b200 import synth, wl
from ugs_b200.engine import DeviceBatch
for name, G, k, m in (("zinc", 131072, 8, 64), ("qm9", 131072, 4, 10), ("proteins", 1113, 6, 100)):
    ei, ptr = synth.make_batch(name, G, 1234)
    b = DeviceBatch(torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda(), k, 0, int(ptr[-1]))
    (nodes, eidx, eptr, _, _), st = b.sample(m, k, "sample", seed=1)
    for it in range(2):
        wl.wl_digests(nodes, eidx, eptr, 3)
    ts = []
    for it in range(5):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); d, nn = wl.wl_digests(nodes, eidx, eptr, 3); c.record(); c.synchronize()
        ts.append(a.elapsed_time(c))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name} G={G} k={k} m={m}: {nodes.shape[0]} samples, WL 3 iterations {ms:.2f} ms = {nodes.shape[0] / ms / 1e3:.1f} M samples/s", flush=True)
