"""Generate tests/golden/ugs_golden.npz from the REAL reference (oracle/_ref, built by oracle/build_ref.py).

Run in the container where /root/reference exists:   python tests/golden/make_golden.py
The fixtures pin, for every case: the reference's internal tables (through oracle/ref_dump_shim.cpp),
get_preproc_info, and sample_batch outputs (its own node rows + the edges it extracts for them) in all three
modes. Sample VALUES are the reference's (xorshift + unordered_set order) and are only used as INPUTS to the
extraction stage; the deterministic stages are compared bit-exactly.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "ss-gnn_b200"))

from helpers import bidir, random_batch  # noqa: E402
from oracle import ref_loader  # noqa: E402
from ugs_b200 import synth  # noqa: E402


def cases():
    rng = np.random.default_rng(2024)
    out = []
    # SURVEY section 0 / section 4 golden edge cases (reference tests/test_debug_sampling.py:7-10, test_uniformity.py:57-60,
    # test_diagnostics.py:25-63, test_sampler.py:9-12, test_sample_batch.py:5-11)
    out.append(("triangle_tail", bidir([(0, 1), (1, 2), (0, 2), (2, 3)]), np.array([0, 4]), 3, 5))
    out.append(("ring6", bidir([(i, (i + 1) % 6) for i in range(6)]), np.array([0, 6]), 3, 5))
    out.append(("triangle_isolated", bidir([(0, 1), (1, 2), (0, 2)]), np.array([0, 4]), 3, 5))
    out.append(("chain5_onedir", np.array([[0, 1, 2, 3], [1, 2, 3, 4]], np.int64), np.array([0, 5]), 4, 5))
    out.append(("single_edge_n3_k3", bidir([(0, 1)]), np.array([0, 3]), 3, 5))
    out.append(("small_then_cycle", bidir([(2, 3), (3, 4), (4, 5), (5, 2)]), np.array([0, 2, 6]), 3, 4))
    out.append(("ref_test_sampler", np.array([[0, 1, 2, 3, 4, 5, 5, 6, 7, 8, 9], [1, 2, 3, 4, 5, 0, 6, 7, 8, 9, 0]], np.int64),
                np.array([0, 4]), 3, 4))
    out.append(("ref_test_sample_batch", np.array([[0, 1, 2, 4, 5, 6, 4], [1, 2, 3, 5, 6, 7, 7]], np.int64),
                np.array([0, 4, 8]), 3, 2))
    ei, ptr = synth.make_batch("mutag", 6, seed=11)
    out.append(("mutag6", ei, ptr, 6, 8))
    ei, ptr = synth.make_batch("proteins", 8, seed=12)
    out.append(("proteins8", ei, ptr, 6, 5))
    ei, ptr = synth.make_batch("qm9", 10, seed=13)
    out.append(("qm9_10", ei, ptr, 4, 10))
    ei, ptr = random_batch(rng, 12, 1, 30, deg=2.5)
    out.append(("weird12", ei, ptr, 4, 6))
    ei, ptr = random_batch(rng, 6, 40, 100, deg=3.0)
    out.append(("weird_big6", ei, ptr, 5, 4))
    return out


def main():
    mod = ref_loader.load("ugs_sampler", "nopin")
    blob = {}
    names = []
    for name, ei, ptr, k, m in cases():
        names.append(name)
        ei = np.ascontiguousarray(ei.astype(np.int64))
        ptr = np.ascontiguousarray(ptr.astype(np.int64))
        blob[f"{name}/edge_index"] = ei
        blob[f"{name}/ptr"] = ptr
        blob[f"{name}/km"] = np.array([k, m], np.int64)
        G = len(ptr) - 1
        for g in range(G):
            lo, hi = int(ptr[g]), int(ptr[g + 1])
            n = hi - lo
            if n <= 0:
                continue
            msk = (ei[0] >= lo) & (ei[0] < hi) & (ei[1] >= lo) & (ei[1] < hi)
            lei = np.ascontiguousarray(ei[:, msk] - lo)
            h = mod.create_preproc(torch.from_numpy(lei), n, k)
            d = ref_loader.dump_preproc(mod, h)
            info = mod.get_preproc_info(h)
            mod.destroy_preproc(h)
            for key in ("indptr", "indices", "edge_col", "order", "index_of", "suffix_deg", "bucket_b",
                        "alias_prob", "alias_alias"):
                blob[f"{name}/g{g}/{key}"] = d[key]
            blob[f"{name}/g{g}/Z"] = np.array([d["Z"]])
            blob[f"{name}/g{g}/info"] = np.array([info["num_nodes"], info["num_edges_stored"],
                                                  info["bucket_count_nonzero"]], np.int64)
        for mode in ("sample", "graph", "global"):
            outs = mod.sample_batch(torch.from_numpy(ei), torch.from_numpy(ptr), m, k, mode, 42)
            for nm, t in zip(("nodes", "edge_index_out", "edge_ptr", "sample_ptr", "edge_src"), outs):
                blob[f"{name}/{mode}/{nm}"] = t.numpy().copy()
    blob["names"] = np.array(names)
    out = os.path.join(HERE, "ugs_golden.npz")
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out), "bytes,", len(names), "cases")


if __name__ == "__main__":
    main()
