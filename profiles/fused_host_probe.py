"""Where the host time of one fused small-batch call goes (perf_counter around every step, MUTAG-shaped G=188)."""
import ctypes, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import numpy as np, torch
from ugs_b200 import synth, _lib, engine
from ugs_b200.engine import _ptr, _stream, MODES
L = _lib.lib()
ei, ptr = synth.make_batch(sys.argv[1] if len(sys.argv) > 1 else "mutag", int(sys.argv[2]) if len(sys.argv) > 2 else 188, 1234)
k, m = 6, 64
d_ei, d_ptr = torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda()
E, G, NT = ei.shape[1], len(ptr) - 1, int(ptr[-1])
dev = d_ei.device
acc = {}
def T(name, t0):
    t = time.perf_counter(); acc.setdefault(name, []).append(1e6 * (t - t0)); return t
for it in range(40):
    torch.cuda.synchronize()
    t = time.perf_counter()
    with torch.cuda.device(dev):
        t = T("ctx", t)
        lay = _lib.Layout(); _lib.check(L.ugs_workspace_layout(E, G, NT, ctypes.byref(lay))); t = T("layout", t)
        cap = int(L.ugs_fused_capacity(E, m)); sbytes = int(L.ugs_sample_scratch_bytes(G, m)); t = T("sizes", t)
        a256 = lambda x: (x + 255) & ~255
        off_s = a256(lay.total_bytes); off_e = off_s + a256(sbytes)
        tmp = torch.empty(off_e + 24 * cap, dtype=torch.uint8, device=dev); t = T("alloc tmp", t)
        B = G * m
        fixed = torch.empty(B * k + (B + 1) + (G + 1), dtype=torch.int64, device=dev); t = T("alloc fixed", t)
        nodes = fixed[:B * k].view(B, k); edge_ptr = fixed[B * k:B * k + B + 1]; sample_ptr = fixed[B * k + B + 1:]; t = T("views", t)
        hi = None
        hi = globals().setdefault("_HI", torch.empty(64, dtype=torch.int64).pin_memory())
        base = tmp.data_ptr(); vp = ctypes.c_void_p
        args = (_ptr(d_ei), E, _ptr(d_ptr), G, 0, NT, m, k, 0, 42, 0, 0, 0, vp(base), ctypes.byref(lay), vp(base + off_s), sbytes,
                _ptr(nodes), _ptr(edge_ptr), _ptr(sample_ptr), vp(base + off_e), vp(base + off_e + 8 * cap), vp(base + off_e + 16 * cap),
                cap, vp(hi.data_ptr()), int(os.environ.get("ASSUME_NO_GENERAL", "0")), _stream()); t = T("marshal", t)
        rc = L.ugs_sample_batch_fused(*args); t = T("C call (launches+sync)", t)
        info = hi[:16].tolist(); t = T("tolist", t)
        total = int(info[4])
        eo = torch.empty(3 * total, dtype=torch.int64, device=dev); eidx = eo[:2 * total].view(2, total); esrc = eo[2 * total:]; t = T("alloc out", t)
        L.ugs_fused_finish(vp(base + off_e), vp(base + off_e + 8 * cap), vp(base + off_e + 16 * cap), total, _ptr(eidx), _ptr(esrc), _stream()); t = T("finish", t)
    t = T("ctx exit", t)
for k_, v in acc.items():
    print(f"{k_:28s} {np.median(v[5:]):8.1f} us")
print("sum", sum(np.median(v[5:]) for v in acc.values()))
# C call split: time the staged C entry points without sync
import subprocess
