"""bbox_overlaps_cython (cython/bbox.pyx) on the GPU vs the compiled reference / the oracle on one host thread.
Prints one JSON line; run under gpurun:  python scripts/bbox_bench.py > gpurun_out/bbox_r1.json"""
import json
import os
import sys
import time

sys.path.insert(0, '.')
import numpy as np
import torch

from dspnet_b200.bbox import bbox_overlaps_cython
from tests.golden.make_bbox_golden import boxes

N, K = 20000, 2000
dev = torch.device('cuda', 0)
b_h, q_h = boxes(1, N), boxes(2, K)
b, q = torch.from_numpy(b_h).to(dev), torch.from_numpy(q_h).to(dev)
for _ in range(5):
    out = bbox_overlaps_cython(b, q)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 50
e0.record()
for _ in range(reps):
    out = bbox_overlaps_cython(b, q)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
t0 = time.perf_counter()
host = bbox_overlaps_cython(b_h, q_h)  # numpy in -> numpy out: H2D + kernel + 320 MB D2H
e2e_ms = 1e3 * (time.perf_counter() - t0)

# CPU baseline on a bounded sample (2000 x 2000) of the same boxes, one thread
from oracle import ref, oracle as O
fn, kind = (ref.bbox_overlaps_cython, "reference") if ref.bbox_available() else (O.bbox_overlaps, "port")
sb = b_h[:2000]
fn(sb[:10], q_h)
t0 = time.perf_counter()
want = fn(sb, q_h)
cpu_s = time.perf_counter() - t0
assert np.array_equal(want.view(np.uint64), host[:2000].view(np.uint64))
peak = json.load(open('MEASURED_PEAKS.json')).get('hbm_gbs', 6464.3) if os.path.exists('MEASURED_PEAKS.json') else 6464.3
byts = 8.0 * N * K + 32.0 * (N + K)
print(json.dumps({"workload": "bbox_overlaps_cython N=%d K=%d float64" % (N, K), "pairs_per_s": N * K / (ms * 1e-3),
                  "kernel_ms": ms, "roofline": {"bound": "hbm", "achieved": byts / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                                "frac": byts / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": byts},
                  "e2e_host_buffers_ms": e2e_ms,
                  "cpu_baseline": {"kind": kind, "cores": 1, "pairs_per_s": 2000 * K / cpu_s, "sample": "2000 x %d pairs" % K},
                  "identical_to_cpu": True}))
