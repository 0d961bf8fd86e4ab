"""Generate tests/golden/walk_golden.npz from the REAL reference rwr_sampler and uniform_sampler (oracle/_ref), run with
a single OpenMP thread so that their output order is the graph order (rwr_sampler.cpp:122-266 writes rows through a
shared counter).     OMP_NUM_THREADS=1 python tests/golden/make_golden_walk.py
The reference's node rows are INPUTS to the extraction stage; edges / pointers are the pinned outputs. Also records,
for the law tests, the reference's sample frequencies on two small graphs (`freq/...`)."""
import os
import sys

os.environ["OMP_NUM_THREADS"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "ss-gnn_b200")):
    sys.path.insert(0, p)
from helpers import bidir  # noqa: E402
from make_golden import cases  # noqa: E402
from oracle import ref_loader  # noqa: E402

FREQ_CASES = [
    ("triangle_tail", [(0, 1), (1, 2), (0, 2), (2, 3)], 4, 3),
    ("ring6_chord", [(i, (i + 1) % 6) for i in range(6)] + [(0, 3)], 6, 4),
]
FREQ_M = 60000


def main():
    torch.set_num_threads(1)
    rwr = ref_loader.load("rwr_sampler", "nopin")
    uni = ref_loader.load("uniform_sampler", "nopin")
    blob, names = {}, []
    for name, ei, ptr, k, m in cases():
        ei = np.ascontiguousarray(ei.astype(np.int64))
        ptr = np.ascontiguousarray(ptr.astype(np.int64))
        sizes = np.diff(ptr)
        if sizes.size and sizes.max() > 40:   # the reference's uniform sampler tests all C(n,k) combinations
            continue
        names.append(name)
        blob[f"{name}/edge_index"] = ei
        blob[f"{name}/ptr"] = ptr
        blob[f"{name}/km"] = np.array([k, m], np.int64)
        for mode in ("sample", "global"):
            for tag, mod in (("rwr", rwr), ("uniform", uni)):
                outs = mod.sample_batch(torch.from_numpy(ei), torch.from_numpy(ptr), m, k, mode, 42)
                for nm, t in zip(("nodes", "edge_index_out", "edge_ptr", "sample_ptr", "edge_src"), outs):
                    blob[f"{name}/{tag}/{mode}/{nm}"] = t.numpy().copy()
    blob["names"] = np.array(names)
    for name, und, n, k in FREQ_CASES:
        ei = bidir(und)
        for tag, mod, kw in (("rwr", rwr, dict(p_restart=0.2)), ("uniform", uni, {})):
            nodes = mod.sample_batch(torch.from_numpy(ei), torch.tensor([0, n]), FREQ_M, k, "sample", 7, **kw)[0].numpy()
            rows, counts = np.unique(nodes, axis=0, return_counts=True)
            blob[f"freq/{name}/{tag}/rows"] = rows
            blob[f"freq/{name}/{tag}/counts"] = counts
    out = os.path.join(HERE, "walk_golden.npz")
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out), "bytes;", len(names), "cases")


if __name__ == "__main__":
    main()
