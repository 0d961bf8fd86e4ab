"""Per-call latency of DataLoader-sized batches, device-resident (CUDA events around the whole call) and through the
drop-in API with CPU tensors: fused path vs staged path (UGS_B200_FUSED=0).  python profiles/latency_probe.py"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import numpy as np, torch
from ugs_b200 import synth
from ugs_b200.engine import sample_batch_device
import ugs_sampler

CASES = [("mutag", 188, 6, 64), ("proteins", 32, 6, 100), ("proteins", 32, 6, 50), ("proteins", 128, 6, 100),
         ("proteins", 1113, 6, 100), ("zinc", 1024, 8, 64), ("qm9", 1024, 4, 10)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, G, k, m in CASES:
    ei, ptr = synth.make_batch(name, G, 1234)
    d_ei, d_ptr = torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda()
    h_ei, h_ptr = torch.from_numpy(ei), torch.from_numpy(ptr)
    NT = int(ptr[-1])
    row = [f"{name} G={G} k={k} m={m}"]
    for fused in ("1", "0"):
        os.environ["UGS_B200_FUSED"] = fused
        for it in range(5):
            sample_batch_device(d_ei, d_ptr, m, k, "sample", it, 0, NT)
        ts, ws = [], []
        for it in range(30):
            flush.zero_()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            outs, st = sample_batch_device(d_ei, d_ptr, m, k, "sample", 100 + it, 0, NT)
            b.record(); b.synchronize()
            ws.append(1e3 * (time.perf_counter() - t0)); ts.append(a.elapsed_time(b))
        for it in range(3):
            ugs_sampler.sample_batch(h_ei, h_ptr, m, k, mode="sample", seed=it)
        es = []
        for it in range(20):
            t0 = time.perf_counter()
            ugs_sampler.sample_batch(h_ei, h_ptr, m, k, mode="sample", seed=50 + it)
            es.append(1e3 * (time.perf_counter() - t0))
        row.append(f"{'fused' if fused == '1' else 'staged'}: device {np.median(ts):.3f} ms (wall {np.median(ws):.3f}) "
                   f"e2e-cpu-tensors {np.median(es):.3f} ms")
    print(" | ".join(row), f"| graphlets {G*m} entries {st['n_edge_entries']}", flush=True)
