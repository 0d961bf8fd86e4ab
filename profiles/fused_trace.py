"""Stage timeline of the fused small-batch call (UGS_B200_FUSED_TRACE=1): python profiles/fused_trace.py mutag 188 6 64"""
import os, sys
os.environ["UGS_B200_FUSED_TRACE"] = "1"
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import torch
from ugs_b200 import synth
from ugs_b200.engine import sample_batch_fused
name, G, k, m = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ei, ptr = synth.make_batch(name, G, 1234)
d_ei, d_ptr = torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda()
print(f"== {name} G={G} k={k} m={m}", file=sys.stderr)
for it in range(6):
    sample_batch_fused(d_ei, d_ptr, m, k, "sample", it, 0, int(ptr[-1]))
    torch.cuda.synchronize()
