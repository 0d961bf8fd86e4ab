"""Record the REAL reference's sampling frequencies on PROTEINS-shaped graphs -> tests/golden/law_ref_counts.npz.

    python tests/golden/make_golden_law.py          (needs oracle/_ref, i.e. /root/reference + oracle/build_ref.py)

BASELINE configs[1] asks for a "uniformity chi-square vs reference" on a PROTEINS-shaped batch at k=6; the reference's own
script for this picks the first PROTEINS graph with 20-30 nodes and draws 5000 samples at k=8
(tests/test_ugs_uniformity_proteins.py:34-46), its chi-square is tests/test_uniformity_analysis.py:36-45, and its
"graphlet isomorphism classes" are WL-hash classes (gps/utils/wl_vocab.py:21-67). Recorded here, per case:
  * the graph (first graph with 20-30 nodes of ugs_b200.synth.make_batch("proteins", 64, seed)), k, m, seed;
  * the unmodified reference `ugs_sampler.sample_batch` called ONCE with m = 500 000 (its RNG stream depends on the sample
    index only, sampler.cpp:160-161): counts per node set (vertex bitmask, n <= 30) and per WL class, the class of a
    node set computed by the reference's own `compute_wl_hash` on the induced simple graph.
tests/test_gpu_law_proteins.py compares 5*10^5 CUDA samples of the same graph with these counts (two-sample chi-square
over node sets and over WL classes) and, for k=6, with the exact law of oracle/exact_law.py.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ss-gnn_b200"))
sys.path.insert(0, os.environ.get("SSGNN_REFERENCE", "/root/reference") + "/src/gps")
warnings.filterwarnings("ignore")

CASES = [  # (name, synth seed, k, m, sampler seed)
    ("proteins_k6", 11, 6, 500000, 42),
    ("proteins_k8", 12, 8, 500000, 43),
]


def pick_graph(seed):
    """First graph with 20..30 nodes of a PROTEINS-shaped batch (test_ugs_uniformity_proteins.py:34), as local columns."""
    from ugs_b200 import synth
    ei, ptr = synth.make_batch("proteins", 64, seed)
    n = np.diff(ptr)
    g = int(np.nonzero((n >= 20) & (n <= 30))[0][0])
    lo, hi = int(ptr[g]), int(ptr[g + 1])
    sel = (ei[0] >= lo) & (ei[0] < hi)
    return np.ascontiguousarray(ei[:, sel] - lo), hi - lo


def masks_of(nodes):
    """uint32 vertex bitmask per complete row; rows with a -1 are dropped (returned separately as a count)."""
    ok = (nodes >= 0).all(axis=1)
    m = np.zeros(int(ok.sum()), np.uint32)
    for j in range(nodes.shape[1]):
        m |= (np.uint32(1) << nodes[ok, j].astype(np.uint32))
    return m, int((~ok).sum())


def wl_of_mask(mask, ei, compute_wl_hash):
    """WL class of the induced subgraph of a node set, by the reference's compute_wl_hash (nx.Graph: simple graph)."""
    verts = [v for v in range(32) if (int(mask) >> v) & 1]
    loc = {v: i for i, v in enumerate(verts)}
    cols = [(loc[int(a)], loc[int(b)]) for a, b in ei.T if int(a) in loc and int(b) in loc]
    t = torch.tensor(cols, dtype=torch.long).t().reshape(2, -1) if cols else torch.zeros((2, 0), dtype=torch.long)
    return compute_wl_hash(t, len(verts), None, 3)


def main():
    from gps.utils.wl_vocab import compute_wl_hash
    from oracle import ref_loader
    ref = ref_loader.load("ugs_sampler", "nopin")
    out = {}
    for name, gseed, k, m, seed in CASES:
        ei, n = pick_graph(gseed)
        nodes = ref.sample_batch(torch.from_numpy(ei), torch.tensor([0, n]), m, k, "sample", seed)[0].numpy()
        masks, n_bad = masks_of(nodes)
        keys, counts = np.unique(masks, return_counts=True)
        wl = [wl_of_mask(x, ei, compute_wl_hash) for x in keys]
        classes = sorted(set(wl))
        cls_of_key = np.array([classes.index(h) for h in wl], np.int32)
        out[name + "_ei"] = ei
        out[name + "_meta"] = np.array([n, k, m, seed, n_bad, gseed], np.int64)
        out[name + "_keys"] = keys
        out[name + "_counts"] = counts.astype(np.int64)
        out[name + "_wl_of_key"] = cls_of_key
        out[name + "_wl_classes"] = np.array(classes)
        print(f"{name}: n={n} columns={ei.shape[1]} distinct node sets={len(keys)} WL classes={len(classes)} "
              f"incomplete={n_bad} min/max count={counts.min()}/{counts.max()}")
    np.savez_compressed(os.path.join(HERE, "law_ref_counts.npz"), **out)


if __name__ == "__main__":
    main()
