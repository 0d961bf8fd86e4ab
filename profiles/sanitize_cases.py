"""Small run that touches every kernel of libugs_b200 once; meant to be run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python profiles/sanitize_cases.py
    compute-sanitizer --tool racecheck python profiles/sanitize_cases.py
"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200")); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import torch
from helpers import random_batch, bidir
from ugs_b200 import synth
from ugs_b200.engine import DeviceBatch
import apx_ugs_sampler

d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
rng = np.random.default_rng(0)
# bitmask + general paths, weird columns, all modes
ei, ptr = random_batch(rng, 24, 1, 40, deg=2.5, big=[(3, 150), (9, 70)])
for mode in ("sample", "graph", "global"):
    b = DeviceBatch(d(ei), d(ptr), 4, 0, int(ptr[-1]))
    b.sample(9, 4, mode, seed=1)
    b.sample_epsilon(9, 4, mode, seed=1, epsilon=0.3)
# k > 8 (32-bit staging), k > 16 (general path only)
for k in (10, 18):
    DeviceBatch(d(ei), d(ptr), k, 0, int(ptr[-1])).sample(5, k, "sample", seed=2)
# unsorted columns -> radix sort + perm
perm = rng.permutation(ei.shape[1])
DeviceBatch(d(ei[:, perm]), d(ptr), 3, 0, int(ptr[-1])).sample(5, 3, "global", seed=3)
# large graph (cooperative kernel) with a hub
star = bidir([(0, i) for i in range(1, 4500)] + [(i, i + 1) for i in range(1, 100)])
DeviceBatch(d(star), d(np.array([0, 4500])), 4, 0, 4500).sample(64, 4, "sample", seed=4)
# uniform variant
ez, pz = synth.make_batch("mutag", 12, seed=3)
DeviceBatch(d(ez), d(pz), 4, 0, int(pz[-1]), variant="uniform").sample(8, 4, "sample", seed=5)
# APX-UGS with few trials
edges = np.array([[0, 1], [1, 0], [1, 2], [2, 1], [2, 0], [0, 2], [2, 3], [3, 2], [3, 4], [4, 3]], np.int64).T.copy()
apx_ugs_sampler.sample_batch(d(edges), torch.tensor([0, 10]).cuda(), 4, 3, "sample", 5, 0.5, _max_trials=2048)
# PROTEINS-sized graphs: k_preproc_medium; enough general-path samples for the thread-per-sample kernel (k_sample_medium)
ep, pp = synth.make_batch("proteins", 60, seed=2)
DeviceBatch(d(ep), d(pp), 5, 0, int(pp[-1])).sample(1400, 5, "sample", seed=6)
# general-path exact-uniform tables and sampler, random-walk sampler
DeviceBatch(d(ep), d(pp), 4, 0, int(pp[-1]), variant="uniform").sample(4, 4, "sample", seed=7)
DeviceBatch(d(ep), d(pp), 4, 0, int(pp[-1])).sample_rwr(6, 4, "sample", seed=8)
# fused one-call path: guess / compact / narrowed, small and mixed batches, host delivery
from ugs_b200 import engine
from ugs_b200.engine import sample_batch_fused
for e_, p_ in ((ez, pz), (ez, pz), (ep, pp), (ei, ptr)):
    sample_batch_fused(d(e_), d(p_), 7, 4, "sample", 9, 0, int(p_[-1]))
engine.FUSED_COMPACT_MAX_BYTES = 0
sample_batch_fused(d(ep), d(pp), 7, 4, "global", 9, 0, int(pp[-1]))
import ugs_sampler
ugs_sampler.sample_batch(torch.from_numpy(ez), torch.from_numpy(pz), 8, 4, mode="sample", seed=3)
torch.cuda.synchronize()
print("sanitize_cases: done")
