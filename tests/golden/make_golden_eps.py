"""Generate tests/golden/eps_golden.npz from the REAL reference epsilon_uniform_sampler (oracle/_ref), run with a
single OpenMP thread so that its output order is the graph order (the reference's order otherwise depends on thread
scheduling, epsilon_uniform_sampler.cpp:295-319).   OMP_NUM_THREADS=1 python tests/golden/make_golden_eps.py
The reference's node rows are INPUTS to the extraction stage; edges / pointers are the pinned outputs."""
import os
import sys

os.environ["OMP_NUM_THREADS"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "ss-gnn_b200")):
    sys.path.insert(0, p)
from make_golden import cases  # noqa: E402
from oracle import ref_loader  # noqa: E402


def main():
    mod = ref_loader.load("epsilon_uniform_sampler", "nopin")
    blob, names = {}, []
    for name, ei, ptr, k, m in cases():
        names.append(name)
        ei = np.ascontiguousarray(ei.astype(np.int64))
        ptr = np.ascontiguousarray(ptr.astype(np.int64))
        blob[f"{name}/edge_index"] = ei
        blob[f"{name}/ptr"] = ptr
        blob[f"{name}/km"] = np.array([k, m], np.int64)
        for mode in ("sample", "global"):
            outs = mod.sample_batch(torch.from_numpy(ei), torch.from_numpy(ptr), m, k, mode, 42, 0.1)
            for nm, t in zip(("nodes", "edge_index_out", "edge_ptr", "sample_ptr", "edge_src"), outs):
                blob[f"{name}/{mode}/{nm}"] = t.numpy().copy()
    blob["names"] = np.array(names)
    out = os.path.join(HERE, "eps_golden.npz")
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
