"""Device portion of the host delivery: int64 emission + packing pass against the emission in the wire format.
python profiles/wire_emit_probe.py [graphs]"""
import ctypes, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import torch
from ugs_b200 import _lib, synth
from ugs_b200.engine import DeviceBatch, _ptr, _stream

G = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
k, m = 8, 64
ei, ptr = synth.make_batch("zinc", G, 1234)
d_ei, d_ptr = torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda()
NT = int(ptr[-1])
L = _lib.lib()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)


def run_wire():
    ts = []
    for it in range(12):
        b = DeviceBatch(d_ei, d_ptr, k, 0, NT)
        outs, st = b.sample(m, k, "sample", seed=it, emit=False, alloc_edges=False)
        nodes, total = outs[0], st["n_edge_entries"]
        B = nodes.shape[0]
        wb = int(L.ugs_deliver_wire_bytes(B, k, total, 1))
        d_wire = torch.empty(wb, dtype=torch.uint8, device="cuda")
        flush.zero_()
        a, c = ev(), ev()
        a.record()
        b.emit_last_wire(d_wire)
        c.record(); c.synchronize()
        ts.append(a.elapsed_time(c))
    return sorted(ts)[len(ts) // 2]


def run_int64_and_pack():
    """emit_last + the packing kernels of ugs_deliver_outputs, timed through a delivery into pinned memory minus nothing:
    the packing pass alone is measured as `pack` below with torch ops of the same traffic as a lower bound."""
    ts, tp = [], []
    for it in range(12):
        b = DeviceBatch(d_ei, d_ptr, k, 0, NT)
        outs, st = b.sample(m, k, "sample", seed=it, emit=False)
        flush.zero_()
        a, c, d = ev(), ev(), ev()
        a.record()
        b.emit_last("sample")
        c.record()
        eidx, esrc = outs[1], outs[4]
        # the traffic of k_pack_pairs_u8 + k_pack_delta16: read every int64 once, write 1 + 2 bytes per entry
        w8 = (eidx[0] | (eidx[1] << 4)).to(torch.uint8)
        w16 = (esrc - esrc[0]).to(torch.int16)
        d.record(); d.synchronize()
        ts.append(a.elapsed_time(c)); tp.append(c.elapsed_time(d))
    return sorted(ts)[len(ts) // 2], sorted(tp)[len(tp) // 2]


t_wire = run_wire()
t_emit, t_pack = run_int64_and_pack()
print(f"emission in the wire format (k_wire_src_bases + k_emit_small<1,1>): {t_wire:.3f} ms")
print(f"int64 emission (k_emit_small<1>): {t_emit:.3f} ms; packing traffic as torch ops (upper bound of the pack kernels): {t_pack:.3f} ms")
