"""End-to-end calls with PAGEABLE inputs (what the trainer passes), staging by host threads on / off, alternating in one
process: python profiles/pageable_probe.py"""
import os, sys, time, statistics
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import torch
import ugs_sampler
from ugs_b200 import synth

cases = [("zinc 131072 graphs k=8 m=64", synth.make_batch("zinc", 131072, 1234), 64, 8),
         ("qm9 131072 graphs k=4 m=10", synth.make_batch("qm9", 131072, 1234), 10, 4),
         ("BA n=1e5 k=6 m=20000", synth.barabasi_albert(100000, 3, 7), 20000, 6),
         ("zinc 16384 graphs k=8 m=64", synth.make_batch("zinc", 16384, 5), 64, 8)]
for name, (ei, ptr), m, k in cases:
    tei, tptr = torch.from_numpy(ei), torch.from_numpy(ptr)
    res = {"1": [], "0": []}
    for rep in range(7):
        for flag in ("1", "0"):
            os.environ["UGS_B200_STAGE_INPUTS"] = flag
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            outs = ugs_sampler.sample_batch(tei, tptr, m, k, "sample", rep)
            torch.cuda.synchronize()
            res[flag].append(1e3 * (time.perf_counter() - t0))
            del outs
    print(f"{name}: input {ei.nbytes / 1e6:.1f} MB; staged by host threads {statistics.median(res['1'][1:]):.2f} ms, "
          f"driver-staged pageable copy {statistics.median(res['0'][1:]):.2f} ms (median of 6 calls after one warm-up)", flush=True)
