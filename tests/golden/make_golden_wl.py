"""Generate tests/golden/wl_golden.json with the REAL reference function gps.utils.wl_vocab.compute_wl_hash
(imported from /root/reference/src/gps; needs networkx, installed here: 3.6.1).   python tests/golden/make_golden_wl.py
Cases: small hand-made graphs (incl. self loops, parallel entries, isolated nodes, single node) and random local edge
lists of 1..12 nodes, each hashed with 1, 2 and 3 WL iterations."""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("SSGNN_REFERENCE", "/root/reference") + "/src/gps")
warnings.filterwarnings("ignore")
from gps.utils.wl_vocab import compute_wl_hash  # noqa: E402


def cases():
    out = [("single", 1, []), ("pair", 2, [(0, 1)]), ("pair_both_dirs", 2, [(0, 1), (1, 0), (0, 1)]),
           ("isolated3", 3, []), ("path3", 3, [(0, 1), (1, 2)]), ("triangle", 3, [(0, 1), (1, 2), (2, 0)]),
           ("self_loop", 3, [(0, 0), (0, 1), (1, 2)]), ("star5", 5, [(0, i) for i in range(1, 5)]),
           ("ring6", 6, [(i, (i + 1) % 6) for i in range(6)]), ("k4", 4, [(a, b) for a in range(4) for b in range(a)]),
           ("two_components", 6, [(0, 1), (1, 2), (3, 4)]),
           ("deg11", 12, [(0, i) for i in range(1, 12)] + [(1, 2)])]
    rng = np.random.default_rng(77)
    for i in range(60):
        n = int(rng.integers(1, 13))
        e = int(rng.integers(0, 3 * n))
        edges = [(int(rng.integers(0, n)), int(rng.integers(0, n))) for _ in range(e)]
        out.append((f"rand{i}", n, edges))
    return out


def main():
    rows = []
    for name, n, edges in cases():
        ei = torch.tensor(edges, dtype=torch.long).t().reshape(2, -1) if edges else torch.zeros((2, 0), dtype=torch.long)
        row = {"name": name, "n": n, "edges": edges,
               "hash": {str(it): compute_wl_hash(ei, n, None, it) for it in (1, 2, 3)}}
        # use_node_features=True: float32 feature rows of 1, 7 (28 bytes) and 20 (80 bytes: two md5 blocks) columns and
        # int64 rows of 7 columns (56 bytes: the md5 padding spills into a second block)
        g = torch.Generator().manual_seed(len(rows))
        feats = {"f32x1": torch.randint(0, 5, (n, 1), generator=g).float(),
                 "f32x7": torch.randn(n, 7, generator=g), "f32x20": torch.randn(n, 20, generator=g),
                 "i64x7": torch.randint(0, 9, (n, 7), generator=g)}
        if len(rows) % 3 == 0:  # every third case carries features (keeps the fixture small)
            row["feat"] = {kk: {"dtype": str(v.dtype).replace("torch.", ""), "values": v.tolist(),
                                "hash": compute_wl_hash(ei, n, v, 3)} for kk, v in feats.items()}
        rows.append(row)
    with open(os.path.join(HERE, "wl_golden.json"), "w") as f:
        json.dump(rows, f)
    print("wrote", len(rows), "cases")


if __name__ == "__main__":
    main()
