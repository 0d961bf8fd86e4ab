import sys, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'ss-gnn_b200'))
import numpy as np, torch
from ugs_b200 import synth
from ugs_b200.engine import DeviceBatch
for n in (10000, 30000, 100000):
    ei, ptr = synth.barabasi_albert(n, 3, seed=7)
    d_ei = torch.from_numpy(ei).cuda(); d_ptr = torch.from_numpy(ptr).cuda()
    for rep in range(3):
        torch.cuda.synchronize(); t=time.perf_counter()
        b = DeviceBatch(d_ei, d_ptr, 6, 0, n)
        torch.cuda.synchronize(); t1=time.perf_counter()
        outs, st = b.sample(20000, 6, "sample", seed=rep)
        torch.cuda.synchronize(); t2=time.perf_counter()
    deg = np.bincount(ei[0], minlength=n)
    print(f"n={n} maxdeg={deg.max()} rows>16:{(deg>16).sum()} rows>64:{(deg>64).sum()} preproc {1e3*(t1-t):.2f} ms sample+emit(20000) {1e3*(t2-t1):.2f} ms incomplete {st['n_incomplete']}", flush=True)
