"""Long randomized cross-check of the device paths against the oracles (same generators as tests/test_gpu_fuzz.py, many
more seeds): python profiles/fuzz_long.py [seconds]. Prints the number of cases per path; raises on the first mismatch."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (R, os.path.join(R, "ss-gnn_b200"), os.path.join(R, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from helpers import random_batch
from oracle import ugs_oracle as uo, eps_oracle as eo, walk_oracle as wo
from ugs_b200.engine import DeviceBatch, sample_batch_fused

MODES = ["sample", "graph", "global"]
NAMES = ["nodes", "edge_index", "edge_ptr", "sample_ptr", "edge_src"]
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
t0 = time.time()
counts = dict(ugs=0, fused=0, eps=0, rwr=0, uniform=0)
seed0 = int(os.environ.get("FUZZ_SEED", "9000"))


def eq(outs, ref, tag):
    for a, r, nm in zip(outs, ref, NAMES):
        if not np.array_equal(a.cpu().numpy(), r):
            raise SystemExit(f"MISMATCH {tag} {nm}")


it = 0
while time.time() - t0 < budget:
    rng = np.random.default_rng(seed0 + it)
    it += 1
    G = int(rng.integers(1, 40))
    big = [(int(rng.integers(0, G)), int(rng.integers(65, 400)))] if rng.random() < 0.4 else []
    ei, ptr = random_batch(rng, G, 1, 60, deg=float(rng.uniform(1.0, 3.5)), weird=bool(rng.random() < 0.7), big=big)
    contiguous = True
    if rng.random() < 0.15:
        ei = np.ascontiguousarray(ei[:, rng.permutation(ei.shape[1])]); contiguous = False
    if rng.random() < 0.3:
        off = int(rng.integers(1, 50)); ei, ptr = ei + off, ptr + off
    k = int(rng.integers(1, 13)); m = int(rng.integers(1, 70)); mode = MODES[int(rng.integers(0, 3))]
    seed = int(rng.integers(0, 2 ** 40))
    d_ei, d_ptr = torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda()
    tag = (it, G, k, m, mode)
    ref = uo.sample_batch(ei, ptr, m, k, mode, seed)
    b = DeviceBatch(d_ei, d_ptr, k, int(ptr[0]), int(ptr[-1] - ptr[0]))
    outs, st = b.sample(m, k, mode, seed=seed)
    eq(outs, ref, ("ugs",) + tag); counts["ugs"] += 1
    f = sample_batch_fused(d_ei, d_ptr, m, k, mode, seed, int(ptr[0]), int(ptr[-1] - ptr[0]))
    if contiguous:
        assert f is not None, tag
    if f is not None:
        eq(f[0], ref, ("fused",) + tag); counts["fused"] += 1
        fs = dict(f[1]); fs.pop("_buffers", None)
        assert fs == st, (fs, st)
    if it % 3 == 0 and k <= 8:
        eps = float(rng.choice([0.05, 0.3, 1.0]))
        o, _ = b.sample_epsilon(m, k, mode, seed=seed, epsilon=eps)
        eq(o, eo.sample_batch(ei, ptr, m, k, mode, seed, eps), ("eps",) + tag); counts["eps"] += 1
    if it % 3 == 1 and k <= 10:
        p = float(rng.choice([0.0, 0.2, 0.9]))
        o, _ = b.sample_rwr(m, k, mode, seed=seed, p_restart=p)
        eq(o, wo.rwr_sample_batch(ei, ptr, m, k, mode, seed, p), ("rwr",) + tag); counts["rwr"] += 1
    if it % 3 == 2 and k <= 4 and m <= 12:
        bu = DeviceBatch(d_ei, d_ptr, k, int(ptr[0]), int(ptr[-1] - ptr[0]), variant="uniform")
        o, _ = bu.sample(m, k, mode, seed=seed)
        eq(o, uo.sample_batch_uniform(ei, ptr, m, k, mode, seed), ("uniform",) + tag); counts["uniform"] += 1
print("fuzz ok", counts, f"{time.time() - t0:.0f}s", flush=True)
