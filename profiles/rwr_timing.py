"""rwr_sampler throughput on the named shapes: python profiles/rwr_timing.py"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import torch
from ugs_b200 import synth
from ugs_b200.engine import DeviceBatch
for name, G, k, m in (("zinc", 131072, 8, 64), ("qm9", 131072, 4, 10), ("proteins", 1113, 6, 100)):
    ei, ptr = synth.make_batch(name, G, 1234)
    b = DeviceBatch(torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda(), k, 0, int(ptr[-1]))
    for it in range(2):
        outs, st = b.sample_rwr(m, k, "sample", seed=it)
    ts = []
    for it in range(5):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); outs, st = b.sample_rwr(m, k, "sample", seed=10 + it); c.record(); c.synchronize()
        ts.append(a.elapsed_time(c))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name} G={G} k={k} m={m}: rwr sample+emit {ms:.2f} ms = {G * m / ms / 1e3:.1f} M samples/s, "
          f"{st['n_walk_steps'] / (G * m):.1f} walk steps per sample, incomplete {st['n_incomplete']}", flush=True)
