"""Cross-call overlap probe: emission of batch i (HBM-bound) on one stream while batch i+1 is preprocessed and sampled
(issue-bound) on another.  python profiles/pipeline_probe.py [graphs] [steps]
Prints ms per step for: serial (one stream), two streams without priorities, two streams with the sampling stream at
high priority, and with the emission stream at high priority."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import torch
from ugs_b200 import synth
from ugs_b200.engine import DeviceBatch

G = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
k, m = 8, 64
ei, ptr = synth.make_batch("zinc", G, 1234)
d_ei, d_ptr = torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda()
NT = int(ptr[-1])


def serial(n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        b = DeviceBatch(d_ei, d_ptr, k, 0, NT)
        outs, st = b.sample(m, k, "sample", seed=i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def piped(n, prio_sample, prio_emit):
    sa = torch.cuda.Stream(priority=prio_sample)
    sb = torch.cuda.Stream(priority=prio_emit)
    keep = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        with torch.cuda.stream(sa):
            b = DeviceBatch(d_ei, d_ptr, k, 0, NT)
            outs, st = b.sample(m, k, "sample", seed=i, emit=False)
            ev = torch.cuda.Event()
            ev.record(sa)
        with torch.cuda.stream(sb):
            sb.wait_event(ev)
            b.emit_last("sample")
            done = torch.cuda.Event()
            done.record(sb)
        keep.append((b, outs, done))
        if len(keep) > 2:  # the third-last batch's emission must be over before its memory goes back to the allocator
            keep[0][2].synchronize()
            keep.pop(0)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for name, fn in (("serial", lambda n: serial(n)),
                 ("two streams, equal priority", lambda n: piped(n, 0, 0)),
                 ("two streams, sampling high priority", lambda n: piped(n, -1, 0)),
                 ("two streams, emission high priority", lambda n: piped(n, 0, -1))):
    fn(30)
    a = fn(steps)
    b_ = fn(steps)
    print(f"{name:42s} {a:.3f} {b_:.3f} ms per step (wall, {steps} steps, no L2 flush: every step writes 6.9 GB)", flush=True)
