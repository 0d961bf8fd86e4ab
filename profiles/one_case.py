"""One workload, a few fused/staged calls (for ncu launch lists): python profiles/one_case.py proteins 32 6 100"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import torch
from ugs_b200 import synth
from ugs_b200.engine import sample_batch_device
name, G, k, m = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ei, ptr = synth.make_batch(name, G, 1234)
d_ei, d_ptr = torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda()
for it in range(4):
    outs, st = sample_batch_device(d_ei, d_ptr, m, k, "sample", it, 0, int(ptr[-1]))
torch.cuda.synchronize()
print(st)
