"""Timing of the encoder-input step on one B200: the reference's own torch ops
(encode_subgraphs, gps/models/ss_gnn.py:437-466, on CUDA tensors -- what the reference runs on a GPU) against
ugs_b200.encode (materialised and fused paths). ZINC-shaped shard, k=8, m=64. CUDA events, L2 flushed between runs.
Usage: python profiles/encode_timing.py [graphs] > profiles/rNN_encode_timing.txt"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ss-gnn_b200"))
import torch  # noqa: E402

from ugs_b200 import encode, synth  # noqa: E402
from ugs_b200.engine import DeviceBatch  # noqa: E402


def torch_reference(x, edge_attr, nodes_t, edge_index_t, edge_ptr_t, edge_src_global_t):
    device = x.device
    num_subgraphs, k = nodes_t.shape
    stacked_nodes = nodes_t.flatten()
    valid_mask = stacked_nodes >= 0
    global_x = x[stacked_nodes.clamp(min=0)]
    if not valid_mask.all():
        global_x = global_x * valid_mask.unsqueeze(-1).float()
    global_edge_attr = edge_attr[edge_src_global_t] if edge_attr is not None else None
    gei = torch.repeat_interleave(torch.arange(0, num_subgraphs, device=device), edge_ptr_t[1:] - edge_ptr_t[:-1]) * k
    gei = gei + edge_index_t
    gb = torch.repeat_interleave(torch.arange(0, num_subgraphs, device=device), k)
    return global_x, global_edge_attr, gei, gb


def timed(fn, flush, reps=5):
    ts = []
    for _ in range(reps + 2):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
        del out
    return sorted(ts[2:])[len(ts[2:]) // 2]


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    k, m = 8, 64
    ei, ptr = synth.make_batch("zinc", G, seed=1237)
    dev = torch.device("cuda", 0)
    d_ei, d_ptr = torch.from_numpy(ei).to(dev), torch.from_numpy(ptr).to(dev)
    N, E = int(ptr[-1]), ei.shape[1]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    b = DeviceBatch(d_ei, d_ptr, k, 0, N)
    outs, st = b.sample(m, k, "sample", seed=1)
    nodes, eidx, eptr, sptr, esrc = outs
    Eb = int(eidx.shape[1])
    rows = []
    for name, x, ea in [("zinc-raw: x int64[N,1], edge_attr int64[E]", torch.randint(0, 28, (N, 1), device=dev),
                         torch.randint(0, 4, (E,), device=dev)),
                        ("dense: x f32[N,64], edge_attr f32[E,16]", torch.randn(N, 64, device=dev),
                         torch.randn(E, 16, device=dev))]:
        xb, eb = x[0].numel() * x.element_size(), ea[0].numel() * ea.element_size()
        alg_gather = nodes.numel() * (8 + xb) + Eb * (8 + eb)            # index read + row written (source rows hit L2)
        alg_mat = alg_gather + Eb * 32 + nodes.numel() * 8               # + edge_index read/write + batch vector
        t_ref = timed(lambda: torch_reference(x, ea, nodes, eidx, eptr, esrc), flush)
        t_mat = timed(lambda: encode.encoder_inputs(x, ea, nodes, eidx, eptr, esrc), flush)
        t_samp = timed(lambda: b.sample(m, k, "sample", seed=2), flush)
        t_ref_e2e = timed(lambda: torch_reference(x, ea, *[(o := b.sample(m, k, "sample", seed=2)[0])[i] for i in (0, 1, 2, 4)]), flush)
        t_fused = timed(lambda: encode.sample_encoder_inputs(b, x, ea, m, k, seed=2), flush)
        rows.append({"features": name, "graphlets": int(nodes.shape[0]), "edge_entries": Eb,
                     "torch_reference_ops_ms": t_ref, "encoder_inputs_ms": t_mat,
                     "encoder_inputs_GBps": alg_mat / t_mat / 1e6,
                     "sample_only_ms": t_samp, "sample_then_torch_reference_ms": t_ref_e2e,
                     "sample_encoder_inputs_fused_ms": t_fused})
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "graphs": G, "k": k, "m": m, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
