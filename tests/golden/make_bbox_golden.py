#!/usr/bin/env python
"""Generates tests/golden/bbox_golden.npz from the REFERENCE ITSELF: cython/bbox.pyx (bbox_overlaps_cython) compiled
in place from /root/reference by oracle/build_ref.py (oracle/_ref/bbox_ref*.so).  Run in the build container:

    python oracle/build_ref.py && python tests/golden/make_bbox_golden.py

A small case is stored in full; a large one as the SHA-256 digest of the output bytes plus a digest of its inputs.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as R  # noqa: E402


def boxes(seed, n, degenerate=True):
    """Pixel-scale [x1, y1, x2, y2] float64 boxes; a few zero / negative extents and exact duplicates."""
    r = np.random.Generator(np.random.PCG64(seed))
    xy = r.uniform(0, 1000, (n, 2))
    wh = r.uniform(4, 300, (n, 2))
    b = np.concatenate([xy, xy + wh], axis=1)
    if degenerate and n >= 8:
        b[1, 2] = b[1, 0] - 1.0          # width exactly 0 under the +1 convention
        b[2, 3] = b[2, 1] - 5.0          # negative height
        b[3] = b[0]                      # duplicate
        b[4] = np.round(b[4])            # integer coordinates
    return b


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def cases():
    small_b, small_q = boxes(7001, 37), boxes(7002, 21)
    small_q[5] = small_b[6]              # an identical pair across the two sets -> overlap exactly 1
    big_b, big_q = boxes(7003, 3000), boxes(7004, 700)
    return small_b, small_q, big_b, big_q


def main():
    assert R.bbox_available(), "run oracle/build_ref.py first"
    small_b, small_q, big_b, big_q = cases()
    out = {"small_boxes": small_b, "small_query": small_q, "small_overlaps": R.bbox_overlaps_cython(small_b, small_q),
           "big_inputs_digest": np.array(digest(big_b, big_q)),
           "big_overlaps_digest": np.array(digest(R.bbox_overlaps_cython(big_b, big_q)))}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bbox_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
