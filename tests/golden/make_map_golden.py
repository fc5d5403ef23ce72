#!/usr/bin/env python
"""Generates tests/golden/map_golden.npz from the REFERENCE ITSELF: evaluate/eval_metric.py (MApMetric /
VOC07MApMetric) loaded from /root/reference behind a stand-in for the `mxnet` base class (oracle/ref_map.py).
Run in the build container:

    python tests/golden/make_map_golden.py

Stored per case: the seeded inputs, the reference's records / counts (flattened) and its mAP value.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = [  # (seed, B, L, M, C, label width, use_difficult, voc07)
    (11, 4, 12, 60, 5, 6, False, False),
    (12, 4, 12, 60, 5, 6, True, True),
    (13, 3, 58, 200, 20, 6, False, True),
    (14, 5, 8, 40, 3, 5, False, False),
]


def case(seed, B, L, M, C, width):
    """Labels (B, L, width) padded with -1 rows and predictions (B, M, 7): jittered copies of the gts (so that IoUs
    straddle the threshold and duplicates occur), random boxes, wrong classes, id -1 rows."""
    r = np.random.Generator(np.random.PCG64(seed))
    lab = np.full((B, L, width), -1, np.float32)
    pred = np.full((B, M, 7), -1, np.float32)
    for b in range(B):
        g = int(r.integers(0, L + 1))
        xy = r.uniform(0, 0.7, (g, 2))
        wh = r.uniform(0.05, 0.3, (g, 2))
        lab[b, :g, 0] = r.integers(0, C, g)
        lab[b, :g, 1:3] = xy
        lab[b, :g, 3:5] = xy + wh
        if width >= 6:
            lab[b, :g, 5] = (r.random(g) < 0.3) * r.random(g)
        m = int(r.integers(0, M + 1))
        src = r.integers(0, max(g, 1), m)
        for j in range(m):
            if g > 0 and r.random() < 0.7:
                box = lab[b, src[j], 1:5] + r.normal(0, 0.02, 4).astype(np.float32)
                cls = lab[b, src[j], 0] if r.random() < 0.8 else r.integers(0, C + 1)
            else:
                p = r.uniform(0, 0.7, 2)
                box = np.concatenate([p, p + r.uniform(0.05, 0.3, 2)])
                cls = r.integers(-1, C + 1)
            pred[b, j, 0] = cls
            pred[b, j, 1] = r.random()
            pred[b, j, 2:6] = box
            pred[b, j, 6] = r.random()
    return lab, pred


def flatten(records, counts):
    """{cid: (n, 2)} / {cid: int} in insertion order -> one (N, 3) array [cid, score, flag] and one (K, 2) array."""
    rec = np.concatenate([np.hstack((np.full((v.shape[0], 1), k, np.float64), v)) for k, v in records.items()]) if records \
        else np.zeros((0, 3))
    cnt = np.array([[k, v] for k, v in counts.items()], dtype=np.int64).reshape(-1, 2)
    return rec, cnt


def main():
    from oracle import ref_map
    assert ref_map.available(), "/root/reference/evaluate/eval_metric.py not found"
    out = {}
    for i, (seed, B, L, M, C, width, ud, voc) in enumerate(CASES):
        lab, pred = case(seed, B, L, M, C, width)
        records, counts, (_, value) = ref_map.run_metric(lab, pred, 0.5, ud, voc, batches=2)
        rec, cnt = flatten(records, counts)
        out["labels_%d" % i], out["preds_%d" % i] = lab, pred
        out["records_%d" % i], out["counts_%d" % i], out["map_%d" % i] = rec, cnt, np.float64(value)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "map_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
