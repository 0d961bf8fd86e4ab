"""End-to-end ugs_sampler.sample_batch (CPU tensors in, pinned CPU int64 tensors out) on the ZINC-shaped shard for a range
of DMA fractions, all ranks at once (they share the box's host memory system):
    python profiles/e2e_delivery.py                        # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 profiles/e2e_delivery.py
Prints one line per fraction: whole-box graphlets/s (max over ranks of the mean step), delivered GB/s, host peak GB/s."""
import ctypes, json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "ss-gnn_b200"))
import numpy as np, torch, torch.distributed as dist
from ugs_b200 import synth, _lib
import ugs_sampler

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
G = int(os.environ.get("GRAPHS", 131072)); k, m = 8, 64
ei, ptr = synth.make_batch("zinc", G, 1234 + rank)
h_ei, h_ptr = torch.from_numpy(ei).pin_memory(), torch.from_numpy(ptr).pin_memory()
L = _lib.lib()

def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

fracs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["auto", "0", "0.2", "0.4", "0.6", "0.8", "1"]
thr = ugs_sampler._host_threads()
probe = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
barrier()
pk = torch.tensor([float(L.ugs_host_store_bandwidth(ctypes.c_void_p(probe.data_ptr()), probe.numel(), thr, 3))], dtype=torch.float64, device=dev)
barrier()
if world > 1: dist.all_reduce(pk)
del probe
if rank == 0:
    print(json.dumps({"world": world, "cpu_count": os.cpu_count(), "threads_per_rank": thr, "isa": L.ugs_host_isa().decode(),
                      "host_store_peak_gbs_all_ranks": float(pk.item())}), flush=True)
for fr in fracs:
    if fr == "auto": os.environ.pop("UGS_B200_DMA_FRACTION", None)
    else: os.environ["UGS_B200_DMA_FRACTION"] = fr
    for w in range(2):
        outs = ugs_sampler.sample_batch(h_ei, h_ptr, m, k, mode="sample", seed=w)
    barrier()
    ts = []
    for s in range(3):
        t0 = time.perf_counter()
        outs = ugs_sampler.sample_batch(h_ei, h_ptr, m, k, mode="sample", seed=10 + s)
        ts.append(time.perf_counter() - t0)
    delivered = sum(o.numel() * 8 for o in outs)
    desc = ugs_sampler._to_host_batch.last_delivery
    del outs
    barrier()
    t = torch.tensor([sum(ts) / len(ts)], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        tm = float(t.item())
        print(json.dumps({"dma_fraction": fr, "ms_per_step": 1e3 * tm, "graphlets_per_s_all_ranks": world * G * m / tm,
                          "delivered_gbs_all_ranks": world * delivered / tm / 1e9, "delivery": desc}), flush=True)
if world > 1: dist.destroy_process_group()
