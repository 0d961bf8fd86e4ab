"""Where the step's non-kernel time goes: CUDA events after every host call of one ZINC-shaped step (the deltas that
contain no kernel are GPU idle time waiting for the host). python profiles/gap_probe.py"""
import ctypes
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import numpy as np
import torch
from ugs_b200 import _lib, synth
from ugs_b200 import engine
from ugs_b200.engine import DeviceBatch

L = _lib.lib()
ei, ptr = synth.make_batch("zinc", 131072, seed=1234)
d_ei, d_ptr = torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda()
NT = int(ptr[-1]); k, m = 8, 64
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
marks = []


def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((name, e))


orig_info = DeviceBatch.info
orig_segment = DeviceBatch._segment


def info(self):
    mark("before info (kernels queued)")
    r = orig_info(self)
    mark("after info (host has the counters)")
    return r


DeviceBatch.info = info
acc = {}
for it in range(8):
    flush.zero_()
    marks.clear()
    mark("start")
    b = DeviceBatch(d_ei, d_ptr, k, 0, NT)
    mark("preproc queued")
    outs, st = b.sample(m, k, "sample", seed=it, emit=True)
    mark("emit queued")
    torch.cuda.synchronize()
    if it >= 3:
        for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
            acc.setdefault(f"{n0} -> {n1}", []).append(e0.elapsed_time(e1) * 1e3)
        acc.setdefault("TOTAL", []).append(marks[0][1].elapsed_time(marks[-1][1]) * 1e3)
    del outs, b
for kx, v in acc.items():
    print(f"{np.mean(v):9.1f} us  {kx}")
