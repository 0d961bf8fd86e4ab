#!/usr/bin/env python
"""bench.py -- sampled k-graphlets/s of the UGS hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload zinc|proteins|mutag|qm9|powerlaw]
                    [--graphs G] [--impl reference]

A "step" is one pass of the whole hot path over one synthetic batch resident in HBM:
segment -> preprocess -> sample -> scan -> emit (what ugs_sampler.sample_batch does per mini-batch).
Default workload: BASELINE.json configs[3] (ZINC/molpcba-shaped, k=8, m=64, "1M graphs sharded by graph over
8xB200"): every rank samples its own shard of 131072 graphs (weak scaling; 8 ranks = the 1M-graph batch).
Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for the definitions of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ss-gnn_b200"))

import numpy as np  # noqa: E402

METRIC = "sampled k-graphlets/sec"
UNIT = "graphlets/s"
DEFAULT_G = {"zinc": 131072, "proteins": 1113, "mutag": 188, "qm9": 131072, "powerlaw": 1}
KM = {"zinc": (8, 64), "proteins": (6, 100), "mutag": (6, 64), "qm9": (5, 10), "powerlaw": (6, 100000)}


def make_workload(name: str, G: int, seed: int):
    from ugs_b200 import synth
    if name == "powerlaw":
        return synth.barabasi_albert(30000, 3, seed)
    return synth.make_batch(name, G, seed)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi SM clocks and throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.stop = threading.Event()
        self.th = None

    def _run(self):
        # NVML in process (nvidia_ml_py): forking an nvidia-smi every 100 ms from a rank with gigabytes mapped costs that
        # rank ~0.5 ms hiccups inside the timed region; the subprocess is only the fallback.
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.idx)
            mx = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            getr = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            while not self.stop.is_set():
                sm = float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                r = int(getr(h))
                row = [str(self.idx), str(sm), str(mx), "", hex(r)]
                row += ["Active" if r & bits[nm] else "Not Active"
                        for nm in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")]
                self.rows.append(row)
                self.stop.wait(0.05)
            return
        except Exception:
            pass
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.idx)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nme, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU sampler (oracle/_ref) on the host cores
# ---------------------------------------------------------------------------------------------------------
def _ref_worker(args):
    """Sample a contiguous shard of graphs with the unmodified reference, in DataLoader-sized chunks."""
    ei, ptr, m, k, seed, chunk, use_port = args
    import torch
    torch.set_num_threads(1)
    if use_port:
        from oracle import ugs_oracle as uo
        fn = lambda e, p: uo.sample_batch(e, p, m, k, "sample", seed)  # noqa: E731
    else:
        from oracle import ref_loader
        mod = ref_loader.load("ugs_sampler", "nopin")  # pageable outputs: no CUDA context per worker
        fn = lambda e, p: mod.sample_batch(torch.from_numpy(e), torch.from_numpy(p), m, k, "sample", seed)  # noqa: E731
    G = len(ptr) - 1
    col_graph = np.searchsorted(ptr, ei[0], side="right") - 1
    cstart = np.searchsorted(col_graph, np.arange(G + 1))
    chunks = []
    for g0 in range(0, G, chunk):
        g1 = min(G, g0 + chunk)
        chunks.append((np.ascontiguousarray(ei[:, cstart[g0]:cstart[g1]] - ptr[g0]),
                       np.ascontiguousarray(ptr[g0:g1 + 1] - ptr[g0]), (g1 - g0) * m))
    # pass 1 = cold (every graph misses the reference's LRU preproc cache, cache.hpp:17-79); pass 2 = warm (the pool is
    # started with UGS_CACHE_SIZE=0 = unbounded, so every graph of the second pass hits and skips preprocessing)
    t0 = time.perf_counter()
    done = 0
    for e, p, cnt in chunks:
        fn(e, p)
        done += cnt
    t_cold = time.perf_counter() - t0
    t0 = time.perf_counter()
    for e, p, cnt in chunks:
        fn(e, p)
    t_warm = time.perf_counter() - t0
    return done, t_cold, t_warm


def cpu_reference_throughput(ei, ptr, m, k, seed, n_graphs, procs, chunk=256):
    """graphlets/s of the reference CPU sampler over the first n_graphs graphs using `procs` processes."""
    import multiprocessing as mp
    from oracle import ref_loader
    use_port = not ref_loader.available("ugs_sampler", "nopin")
    G = min(n_graphs, len(ptr) - 1)
    col_graph = np.searchsorted(ptr, ei[0], side="right") - 1
    cstart = np.searchsorted(col_graph, np.arange(len(ptr)))
    procs = max(1, min(procs, G))
    bounds = np.linspace(0, G, procs + 1).astype(int)
    jobs = []
    for r in range(procs):
        g0, g1 = int(bounds[r]), int(bounds[r + 1])
        if g1 <= g0:
            continue
        e = np.ascontiguousarray(ei[:, cstart[g0]:cstart[g1]] - ptr[g0])
        p = np.ascontiguousarray(ptr[g0:g1 + 1] - ptr[g0])
        jobs.append((e, p, m, k, seed, chunk, use_port))
    os.environ["UGS_CACHE_SIZE"] = "0"  # read once when the reference module loads in a worker: unbounded LRU
    if len(jobs) == 1:
        res = [_ref_worker(jobs[0])]
    else:
        ctx = mp.get_context("spawn")
        with ctx.Pool(len(jobs)) as pool:
            # warm the pool (imports, library load) outside the timed region
            pool.map(_ref_worker, [(j[0][:, :0], np.zeros(1, np.int64), m, k, seed, chunk, use_port) for j in jobs])
            res = pool.map(_ref_worker, jobs)
    total = sum(r[0] for r in res)
    wall = max(r[1] for r in res)       # the slowest worker's cold pass: all workers start together
    wall_warm = max(r[2] for r in res)
    return {"value": total / wall, "unit": UNIT, "cores": len(jobs), "kind": "port" if use_port else "reference",
            "warm_value": total / wall_warm,
            "sample": f"reference ugs_sampler.sample_batch (oracle/_ref, outputs pageable instead of pinned), "
                      f"first {G} graphs of the workload, {chunk} graphs per call, {len(jobs)} worker processes x 1 "
                      f"thread; value = cold LRU cache (first pass, {wall:.2f}s), warm_value = second pass over the "
                      f"same graphs with every preprocessing cached ({wall_warm:.2f}s)"}


def workload_name(workload: str, graphs_per_rank: int) -> str:
    """The `config.workload` string, shared by both arms."""
    if workload == "zinc":
        return (f"zinc-shaped batch (BASELINE configs[3]: 1M graphs k=8 m=64 sharded by graph over 8 GPUs -> "
                f"{graphs_per_rank} graphs per rank)")
    return f"{workload}-shaped batch"


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    k, m = KM[args.workload]
    procs = os.cpu_count() or 1
    G = args.graphs or DEFAULT_G[args.workload]
    n_sample = min(G, max(8192, procs * 2048))
    ei, ptr = make_workload(args.workload, n_sample, 1234)
    vals = []
    last = None
    for it in range(args.warmup + args.steps):
        last = cpu_reference_throughput(ei, ptr, m, k, 42 + it, n_sample, procs)
        if it >= args.warmup:
            vals.append(last["value"])
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n_sample * m / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, G), "k": k, "m": m, "graphs_per_rank": G,
                   "graphs_per_step": n_sample,
                   "sample": f"each step = the first {n_sample} graphs of that workload on all host cores"},
        "cpu_baseline": dict(last, value=v),
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ---------------------------------------------------------------------------------------------------------
# Every BASELINE.json config (and the trainer's mini-batch sizes) next to the headline workload
# ---------------------------------------------------------------------------------------------------------
CONFIG_CASES = [
    # name, generator, graphs, k, m, BASELINE config it belongs to
    ("mutag", ("batch", "mutag", 188), 6, 64, "configs[0] MUTAG-shaped, G=188"),
    ("proteins", ("batch", "proteins", 1113), 6, 100, "configs[1] PROTEINS-shaped, G=1113"),
    ("minibatch_32", ("batch", "proteins", 32), 6, 100, "configs[1] as the trainer calls it: one DataLoader mini-batch of 32 graphs"),
    ("minibatch_1024", ("batch", "proteins", 1024), 6, 100, "configs[1], mini-batch of 1024 graphs"),
    ("qm9_k4", ("batch", "qm9", 131072), 4, 10, "configs[2] QM9-shaped, k=4, m=10"),
    ("qm9_k5", ("batch", "qm9", 131072), 5, 10, "configs[2] QM9-shaped, k=5, m=10"),
    ("powerlaw_1e4", ("ba", 10000), 6, 20000, "configs[4] Barabasi-Albert n=10^4, one graph"),
    ("powerlaw_1e5", ("ba", 100000), 6, 20000, "configs[4] Barabasi-Albert n=10^5, one graph"),
]


def _gen_case(gen, seed):
    from ugs_b200 import synth
    if gen[0] == "ba":
        return synth.barabasi_albert(gen[1], 3, seed)
    return synth.make_batch(gen[1], gen[2], seed)


def l2_copy_peak_gbs(torch, dev):
    """Measured L2 bandwidth of this GPU: device copy of a 24 MB buffer (source + destination stay resident in the 126 MB
    L2), read + write bytes, best of 20 -- the peak the general-path ("l2"-bound) kernels are reported against."""
    a = torch.empty(24 << 20, dtype=torch.uint8, device=dev)
    b = torch.empty_like(a)
    for _ in range(5):
        b.copy_(a)
    best = 0.0
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        e1.record()
        e1.synchronize()
        best = max(best, 2 * a.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def run_config_cases(torch, dev, flush_buf, peak_hbm, with_reference=True, steps=10):
    """device-resident whole-call time (CUDA events, L2 flushed between calls), end-to-end time through
    ugs_sampler.sample_batch with pageable CPU tensors (what gps/experiment.py:877-883 passes), and the reference's own
    sample_batch on the same batch on one host core (it is single-threaded), for every BASELINE config."""
    import ugs_sampler
    from ugs_b200.engine import sample_batch_device, fused_eligible
    out = {}
    ref_mod = None
    if with_reference:
        try:
            from oracle import ref_loader
            if ref_loader.available("ugs_sampler", "nopin"):
                ref_mod = ref_loader.load("ugs_sampler", "nopin")
        except Exception:
            ref_mod = None
    for idx, (name, gen, k, m, what) in enumerate(CONFIG_CASES):
        ei, ptr = _gen_case(gen, 1234 + idx)
        G, E, NT = len(ptr) - 1, ei.shape[1], int(ptr[-1] - ptr[0])
        d_ei, d_ptr = torch.from_numpy(ei).to(dev), torch.from_numpy(ptr).to(dev)
        for w in range(3):
            outs, st = sample_batch_device(d_ei, d_ptr, m, k, "sample", 5 + w, int(ptr[0]), NT)
        ts = []
        for s_ in range(steps):
            flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            outs, st = sample_batch_device(d_ei, d_ptr, m, k, "sample", 100 + s_, int(ptr[0]), NT)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        ms = statistics.median(ts)
        ebar = st["n_edge_entries"] / max(st["n_graphlets"], 1)
        b_alg = 8 * k + 8 + 24 * ebar + (16 * E / max(G, 1) + 16) / m
        del outs
        # end to end: pageable CPU int64 tensors in, pinned CPU int64 tensors out
        h_ei, h_ptr = torch.from_numpy(ei), torch.from_numpy(ptr)
        for w in range(3):
            ugs_sampler.sample_batch(h_ei, h_ptr, m, k, mode="sample", seed=w)
        te = []
        for s_ in range(max(5, steps // 2)):
            t0 = time.perf_counter()
            r = ugs_sampler.sample_batch(h_ei, h_ptr, m, k, mode="sample", seed=50 + s_)
            te.append(time.perf_counter() - t0)
        del r
        e2e_ms = 1e3 * statistics.median(te)  # median: the pinned-memory allocator grows on a new output size
        row = {"what": what, "graphs": G, "nodes": NT, "edge_columns": E, "k": k, "m": m,
               "device_ms": ms, "graphlets_per_s": G * m / (ms * 1e-3),
               "path": "fused (one call, one host round trip)" if fused_eligible(E, G, NT, m, k) else "staged",
               "general_path_graphs": st["n_general_graphs"], "mean_edge_entries_per_graphlet": ebar,
               "roofline_step": {"bytes_per_graphlet": b_alg, "frac": b_alg * G * m / (ms * 1e-3) / 1e9 / peak_hbm},
               "e2e_ms": e2e_ms, "e2e_graphlets_per_s": G * m / (e2e_ms * 1e-3)}
        if ref_mod is not None:
            # the reference, one core, cold preprocessing cache; big configs: a bounded sample of the same batch
            g_ref = min(G, 4096)
            m_ref = min(m, 1000)
            col_graph = np.searchsorted(ptr, ei[0], side="right") - 1
            c1 = int(np.searchsorted(col_graph, g_ref))
            e_ref = torch.from_numpy(np.ascontiguousarray(ei[:, :c1]))
            p_ref = torch.from_numpy(np.ascontiguousarray(ptr[:g_ref + 1]))
            t0 = time.perf_counter()
            for g0 in range(0, g_ref, 256):  # DataLoader-sized calls: the reference's slice step is O(G*E) per call
                g1 = min(g_ref, g0 + 256)
                ca, cb = int(np.searchsorted(col_graph, g0)), int(np.searchsorted(col_graph, g1))
                ref_mod.sample_batch(e_ref[:, ca:cb] - int(ptr[g0]), p_ref[g0:g1 + 1] - int(ptr[g0]), m_ref, k, "sample", 42)
            t_ref = time.perf_counter() - t0
            row["reference_cpu"] = {"ms": 1e3 * t_ref, "graphlets_per_s": g_ref * m_ref / t_ref, "cores": 1,
                                    "sample": f"unmodified reference sample_batch, first {g_ref} graphs, m={m_ref}, "
                                              f"256 graphs per call, cold cache"}
            row["speedup_e2e_vs_reference_1core"] = row["e2e_graphlets_per_s"] / row["reference_cpu"]["graphlets_per_s"]
        out[name] = row
        del d_ei, d_ptr
    return out


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=10)  # 2 ms steps: a few of them do not bring the clocks up from idle
    ap.add_argument("--workload", default="zinc", choices=sorted(KM))
    ap.add_argument("--graphs", type=int, default=0, help="graphs per rank (default: workload-specific)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config block (other BASELINE configs)")
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="length of the sustained loop (0 = skip)")
    ap.add_argument("--prewarm-seconds", type=float, default=1.0, help="untimed load before the warm-up steps (clock ramp)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from ugs_b200 import _lib
    from ugs_b200.engine import DeviceBatch
    import ugs_sampler  # the drop-in public API (e2e leg)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 sampler has no CPU fallback")
    L = _lib.lib()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    k, m = KM[args.workload]
    G = args.graphs or DEFAULT_G[args.workload]
    # power-law config: ONE huge graph replicated on every rank, its m samples split across ranks;
    # every other workload: each rank owns its own shard of graphs (weak scaling).
    sample_sharded = args.workload == "powerlaw" and world > 1
    ei_np, ptr_np = make_workload(args.workload, G, 1234 + (0 if sample_sharded else rank))
    ibase, m_job = 0, m
    if sample_sharded:
        from ugs_b200 import sharding
        ibase, m = sharding.shard_samples(m_job, rank, world)
    G = len(ptr_np) - 1
    E = ei_np.shape[1]
    NT = int(ptr_np[-1] - ptr_np[0])
    d_ei = torch.from_numpy(ei_np).to(dev)
    d_ptr = torch.from_numpy(ptr_np).to(dev)
    gbase = 0 if sample_sharded else rank * G  # Philox is keyed on the global graph index: results do not depend on the sharding
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    stage_ms = {"preproc": [], "sample": [], "emit": []}
    last = {}

    def step(seed, timed):
        """One pass of the hot path; returns device-event milliseconds of the whole pass."""
        flush_buf.zero_()  # L2 flush between iterations (not timed)
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        b = DeviceBatch(d_ei, d_ptr, k, int(ptr_np[0]), NT)
        e1.record()
        outs, stats = b.sample(m, k, "sample", seed=seed, graph_index_base=gbase, emit=False,
                               sample_index_base=ibase)
        e2.record()
        nodes, eidx, eptr, sptr, esrc = outs
        total = stats["n_edge_entries"]
        b.emit_last("sample")
        e3.record()
        e3.synchronize()
        if timed:
            stage_ms["preproc"].append(e0.elapsed_time(e1))
            stage_ms["sample"].append(e1.elapsed_time(e2))
            stage_ms["emit"].append(e2.elapsed_time(e3))
        last.update(stats=stats, total=total, n_general=b.n_general, nodes=nodes, eptr=eptr, batch=b)
        return e0.elapsed_time(e3)

    clk = ClockSampler(local_rank)
    if rank == 0:  # one sampler per job: nvidia-smi takes a driver lock, N of them perturb the other ranks' launches
        clk.__enter__()  # sampled from the warm-up to the end of the e2e leg (the timed regions are milliseconds long)
    # Pre-warm: a fresh process on an idle GPU needs a few hundred milliseconds of load before SM and memory clocks and
    # the allocator's block sizes are those of a running job (measured: the first 20 steps after 10 warm-up steps ran at
    # 2.02 ms, the same steps after 2 s of load at 1.79 ms). Untimed, before the W warm-up steps, stated in `config`.
    t_pw = time.perf_counter()
    n_pw = 0
    while time.perf_counter() - t_pw < args.prewarm_seconds:
        step(50000 + n_pw, False)
        n_pw += 1
    for w in range(args.warmup):
        step(100 + w, False)
    # Allocator warm-up: the size of edge_index / edge_src is data dependent (it changes by ~0.01 % with the seed), and
    # a step whose output is a few KB larger than every earlier one would pay a cudaMalloc (2-3 ms, seen as a one-off
    # 4 ms step on one rank of a multi-GPU run). Park blocks with 3 % headroom in torch's caching allocator once.
    if last.get("total"):
        pad = [torch.empty(int(2 * last["total"] * 1.03), dtype=torch.int64, device=dev),
               torch.empty(int(last["total"] * 1.03), dtype=torch.int64, device=dev)]
        del pad
    import gc
    gc.collect()
    gc.disable()  # no cyclic-GC pause inside the timed region
    launches0 = int(L.ugs_launch_count())
    barrier()
    times = [step(1000 + s, True) for s in range(args.steps)]
    barrier()
    launches = int(L.ugs_launch_count()) - launches0
    # ---- sustained: the same step back to back for >= --sustained-seconds (the headline region above is milliseconds
    # long, i.e. a burst number); same L2 flush between steps, device-event time of the steps only
    sustained = None
    if args.sustained_seconds > 0:
        n_before = len(stage_ms["emit"])
        t_wall0 = time.perf_counter()
        st_times = []
        while time.perf_counter() - t_wall0 < args.sustained_seconds:
            st_times.append(step(5000 + len(st_times), True))
        em = stage_ms["emit"][n_before:]
        sustained = {"wall_seconds": time.perf_counter() - t_wall0, "steps": len(st_times),
                     "ms_per_step": sum(st_times) / len(st_times),
                     "emit_ms": sum(em) / len(em)}
        for kk in stage_ms:  # keep the headline's stage times those of the K timed steps
            del stage_ms[kk][n_before:]
    gc.enable()
    t_local = sum(times) / 1e3
    t = torch.tensor([t_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_max = float(t.item())
    # per-rank diagnostics: step time, stage times, the slowest single step, and this rank's SM clock right after the
    # timed loop (NVML, in process -- no nvidia-smi subprocess per rank)
    def _nvml_clock(idx):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            return float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return -1.0
    diag = [1e3 * t_local / args.steps] + [sum(stage_ms[kk]) / len(stage_ms[kk]) for kk in ("preproc", "sample", "emit")] \
        + [max(times), _nvml_clock(local_rank)]
    per_rank_diag = [diag]
    if world > 1:
        tl = torch.tensor(diag, dtype=torch.float64, device=dev)
        tg = [torch.empty_like(tl) for _ in range(world)]
        dist.all_gather(tg, tl)
        per_rank_diag = [x.tolist() for x in tg]
    per_rank_ms = [d[0] for d in per_rank_diag]

    # the one collective of the design: all-gather of per-rank counts/statistics (never on the sampling path)
    st = last["stats"]
    svec = torch.tensor([st["n_graphlets"], st["n_edge_entries"], st["n_incomplete"], st["n_relaxed_graphs"],
                         st["n_skipped_graphs"], st["n_general_graphs"]], dtype=torch.int64, device=dev)
    if world > 1:
        gathered = [torch.empty_like(svec) for _ in range(world)]
        dist.all_gather(gathered, svec)
        allstats = torch.stack(gathered).sum(0).tolist()
    else:
        allstats = svec.tolist()
    graphlets_step_all = allstats[0]
    value = graphlets_step_all * args.steps / t_max

    # ---- roofline: algorithmic bytes -----------------------------------------------
    peak, peak_src = peaks()
    ebar = st["n_edge_entries"] / max(st["n_graphlets"], 1)
    b_alg = 8 * k + 8 + 24 * ebar + (16 * E / max(G, 1) + 16) / m          # bytes per graphlet, whole path
    emit_bytes = 24.0 * st["n_edge_entries"]                                # the emit kernel's share per launch
    emit_ms = sum(stage_ms["emit"]) / len(stage_ms["emit"])
    achieved = emit_bytes / (emit_ms * 1e-3) / 1e9
    step_ms_local = sum(times) / len(times)
    if sustained is not None:
        sustained["value"] = st["n_graphlets"] / (sustained["ms_per_step"] * 1e-3)
        sustained["emit_gbs"] = emit_bytes / (sustained["emit_ms"] * 1e-3) / 1e9
        sustained["emit_frac_of_peak"] = sustained["emit_gbs"] / peak
        sustained["what"] = "this rank's steps back to back (L2 flushed between steps, device-event time of the steps)"
    general = last["n_general"] > 0 and last["n_general"] * 2 >= G - st["n_skipped_graphs"]  # general path dominates
    stage_avg = {kk: sum(v) / len(v) for kk, v in stage_ms.items()}
    top_stage = max(stage_avg, key=stage_avg.get)
    KERNELS = {False: {"preproc": "k_preproc_small (CSR + order + weights, warp per graph)",
                       "sample": "k_sample_small (+ k_scan_counts)",
                       "emit": "k_emit_small (induced-edge emission + compaction)"},
               True: {"preproc": "k_preproc_large / k_preproc_general (cooperative, L2-resident tables)",
                      "sample": "k_sample_general (+ k_scan_counts)", "emit": "k_emit_general"}}
    if general:
        # General path: the graph's tables (CSR, order, CDF: a few MB) live in L2; what the sampling and emission kernels
        # must move is every member row once per pass (4 B per CSR entry; growth + entry count + emission = 3 passes) and
        # the outputs. Sum of member-row lengths from the produced node rows (outside the timed region).
        nodes_l, eptr_l = last["nodes"], last["eptr"]
        rp = last["batch"].view("rowptr").to(torch.int64)
        deg = rp[1:NT + 1] - rp[:NT] if G == 1 else None
        rows_entries = float(deg[(nodes_l.clamp(min=0) - int(ptr_np[0])).reshape(-1)].sum().item()) if deg is not None else 0.0
        l2_peak = l2_copy_peak_gbs(torch, dev)
        stage_bytes = {"sample": 2 * 4.0 * rows_entries + (8 * k + 8) * st["n_graphlets"],
                       "emit": 4.0 * rows_entries + 24.0 * st["n_edge_entries"] + 8 * k * st["n_graphlets"],
                       "preproc": 16.0 * E + 28.0 * 2 * E + 40.0 * NT}
        roof = {"bound": "l2", "kernel": KERNELS[True][top_stage], "achieved": stage_bytes[top_stage] / (stage_avg[top_stage] * 1e-3) / 1e9,
                "peak": l2_peak, "unit": "GB/s", "traffic": None,
                "peak_source": "measured in this run: device copy of a 24 MB L2-resident buffer, read + write bytes",
                "algorithmic_bytes_per_launch": stage_bytes[top_stage], "launch_ms": stage_avg[top_stage],
                "bytes_model": "member rows read once per pass at 4 B per CSR entry (sample: growth + count, emit: 1) + "
                               "output bytes; preproc: 16 B/column read + CSR/tables written"}
        roof["frac"] = roof["achieved"] / roof["peak"]
    else:
        roof = None
    step_achieved = b_alg * st["n_graphlets"] / (step_ms_local * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get(args.workload)
        except Exception:
            traffic = None

    # ---- e2e: the drop-in public API with HOST buffers (H2D + D2H inside the timed region) ----------------
    e2e = None
    if not args.no_e2e:
        h_ei = torch.from_numpy(ei_np).pin_memory()
        h_ptr = torch.from_numpy(ptr_np).pin_memory()
        d2h = delivered = 0
        for w in range(2):
            outs = ugs_sampler.sample_batch(h_ei, h_ptr, m, k, mode="sample", seed=7 + w)
        barrier()
        tt = []
        for s in range(args.e2e_steps):
            a, bq = ev(), ev()
            a.record()
            outs = ugs_sampler.sample_batch(h_ei, h_ptr, m, k, mode="sample", seed=50 + s)
            bq.record()
            bq.synchronize()
            tt.append(a.elapsed_time(bq))
            delivered = sum(o.numel() * o.element_size() for o in outs)  # int64 bytes handed to the caller
            d2h = int(getattr(ugs_sampler._to_host_batch, "last_d2h_bytes", 0)) or delivered  # bytes that crossed PCIe
            assert not outs[0].is_cuda and outs[0].is_pinned()
        barrier()
        # the same call with PAGEABLE inputs, which is what the trainer hands over (experiment.py:877-878: .cpu() tensors)
        p_ei, p_ptr = torch.from_numpy(ei_np), torch.from_numpy(ptr_np)
        tp_ = []
        for s in range(3):  # the first call is the warm-up (it allocates the pinned staging buffer of the inputs)
            a, bq = ev(), ev()
            a.record()
            outs = ugs_sampler.sample_batch(p_ei, p_ptr, m, k, mode="sample", seed=70 + s)
            bq.record()
            bq.synchronize()
            if s > 0:
                tp_.append(a.elapsed_time(bq))
        barrier()
        # host memory system: what the box's cores can WRITE with streaming stores, measured in this run with the
        # threads the delivery uses (all ranks at once: they share one memory system), against what the delivery achieved
        host_peak = None
        n_thr = ugs_sampler._host_threads()
        if hasattr(L, "ugs_host_store_bandwidth"):
            import ctypes
            L.ugs_host_store_bandwidth.restype = ctypes.c_double
            L.ugs_host_store_bandwidth.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
            probe = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
            barrier()
            host_peak = float(L.ugs_host_store_bandwidth(ctypes.c_void_p(probe.data_ptr()), probe.numel(), n_thr, 3))
            barrier()
            del probe
        te = torch.tensor([sum(tt) / 1e3, sum(tp_) / 1e3, host_peak or 0.0], dtype=torch.float64, device=dev)
        tsum = te.clone()
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_e2e = float(te[0].item())
        delivered_all = delivered * world  # every rank delivers its own shard (same shapes up to 0.01 %)
        e2e = {"value": graphlets_step_all * args.e2e_steps / t_e2e, "unit": UNIT,
               "h2d_bytes_per_step": int(h_ei.numel() * 8 + h_ptr.numel() * 8), "d2h_bytes_per_step": int(d2h),
               "host_bytes_delivered_per_step": int(delivered), "steps": args.e2e_steps,
               "ms_per_step": 1e3 * t_e2e / args.e2e_steps,
               "pageable_inputs_value": graphlets_step_all * 2 / float(te[1].item()),
               "delivery": getattr(ugs_sampler._to_host_batch, "last_delivery", None),
               "host_threads_per_rank": n_thr,
               "host_gbs": delivered_all * args.e2e_steps / t_e2e / 1e9,
               "host_peak_gbs": (float(tsum[2].item()) if host_peak else None),
               "host_peak_source": "ugs_host_store_bandwidth: 1 GiB of non-temporal stores per rank with the delivery's "
                                   "thread count, all ranks at the same time, sum over ranks",
               "api": "ugs_sampler.sample_batch(cpu int64 tensors) -> pinned cpu int64 tensors; results cross PCIe packed "
                      "(node ids as 16-bit deltas per 4096-value block; edge entries emitted by the emission kernel itself as "
                      "one byte of index pair + an 8-bit source-column delta per 1024-entry block) and are widened by a "
                      "persistent pool of host threads (AVX-512 / AVX2 non-temporal stores) while the copy runs; bound by the "
                      "host memory system: host_gbs vs host_peak_gbs"}
        del outs

    if rank == 0:
        clk.__exit__(None, None, None)
    configs = None
    if rank == 0 and world == 1 and not args.no_configs and args.workload == "zinc":
        del d_ei, d_ptr
        last.clear()
        torch.cuda.empty_cache()
        try:
            configs = run_config_cases(torch, dev, flush_buf, peak, with_reference=not args.no_cpu_baseline)
        except Exception as ex:  # a report next to the headline, never a reason to lose the headline
            configs = {"error": repr(ex)[:300]}
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        procs = os.cpu_count() or 1
        n_sample = min(G, max(8192, procs * 2048))
        try:
            cpu_base = cpu_reference_throughput(ei_np, ptr_np, m, k, 42, n_sample, procs)
        except Exception as ex:  # the baseline is a report, never the product
            cpu_base = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if sample_sharded else "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {
                "workload": workload_name(args.workload, G),
                "k": k, "m": m, "graphs_per_rank": G, "nodes_per_rank": NT, "edge_columns_per_rank": E,
                "mean_edge_entries_per_graphlet": ebar, "l2": "512 MiB flush between timed iterations",
                "prewarm": f"{args.prewarm_seconds:g} s of untimed steps before the {args.warmup} warm-up steps",
                "sharding": (f"dp{world} by sample of one replicated graph (m={m_job} split), no data-path collective"
                             if sample_sharded else f"dp{world} by graph, no data-path collective"),
            },
            "stage_ms": {kk: sum(v) / len(v) for kk, v in stage_ms.items()},
            "ms_per_step_per_rank": per_rank_ms,
            "per_rank": [dict(zip(["ms_per_step", "preproc_ms", "sample_ms", "emit_ms", "slowest_step_ms", "sm_mhz_after"],
                                  [round(v, 4) for v in d])) for d in per_rank_diag],
            "roofline": roof if roof is not None else {
                "bound": "hbm", "kernel": KERNELS[False]["emit"],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": emit_bytes, "launch_ms": emit_ms,
                "share_of_step": emit_ms / step_ms_local, "dominant_stage": top_stage},
            "roofline_step": {"bytes_per_graphlet": b_alg, "achieved": step_achieved, "peak": peak, "unit": "GB/s",
                              "frac": step_achieved / peak},
            "sustained": sustained,
            "cpu_baseline": cpu_base,
            "e2e": e2e,
            "configs": configs,
            "gpu_launches": launches,
            "clocks": clk.summary(),
            "stats_all_ranks": dict(zip(["n_graphlets", "n_edge_entries", "n_incomplete", "n_relaxed_graphs",
                                         "n_skipped_graphs", "n_general_graphs"], allstats)),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
