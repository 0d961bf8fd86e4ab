"""End-to-end ugs_sampler.sample_batch (pinned inputs) for the wire chunk size of this process:
for c in 1 2 4 8 16 32 64; do UGS_B200_WIRE_CHUNK_MB=$c python profiles/chunk_sweep.py; done"""
import os, sys, time, statistics
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import torch
import ugs_sampler
from ugs_b200 import synth

ei, ptr = synth.make_batch("zinc", 131072, 1234)
tei, tptr = torch.from_numpy(ei).pin_memory(), torch.from_numpy(ptr).pin_memory()
ts = []
for rep in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = ugs_sampler.sample_batch(tei, tptr, 64, 8, "sample", rep)
    torch.cuda.synchronize()
    ts.append(1e3 * (time.perf_counter() - t0))
    del outs
print(f"chunk {os.environ.get('UGS_B200_WIRE_CHUNK_MB', '32 (default)')} MB, threads {ugs_sampler._host_threads()}: "
      f"median {statistics.median(ts[2:]):.2f} ms, min {min(ts[2:]):.2f} ms", flush=True)
