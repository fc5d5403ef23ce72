"""Evaluation consumers of the detection output on the bench workload: post-filter + TP/FP
matching on the GPU vs the reference's host path (asnumpy of the (B, A, 7) output, numpy post-filter,
MApMetric.update).  Prints one JSON line; run under gpurun:  python scripts/evalmap_bench.py > gpurun_out/evalmap_r1.json"""
import json
import sys
import time

sys.path.insert(0, '.')
import numpy as np
import torch

import bench
from dspnet_b200 import evalmap
from dspnet_b200.plan import DetectionPlan
from dspnet_b200.symbol import multibox_anchors

dev = torch.device('cuda', 0)
B = bench.BATCH
inputs, _ = bench.make_inputs(0, B)
A, C, L = inputs['A'], inputs['C'], inputs['L']
anchors = multibox_anchors(bench.PRESET, device=dev)
plan = DetectionPlan(B, A, C, dev, **bench.DET_PARAMS)
out = plan.run(torch.from_numpy(inputs['prob']).to(dev), torch.from_numpy(inputs['loc']).to(dev), anchors, plan.new_output())
labels = torch.from_numpy(inputs['lab']).to(dev)
torch.cuda.synchronize()


def gpu_path():
    rows, counts = evalmap.postfilter(out, 200, 0.25)
    flags = evalmap.match_flags(labels, rows, 0.5, False)
    return rows.cpu(), flags.cpu(), counts.cpu()


for _ in range(5):
    gpu_path()
torch.cuda.synchronize()
reps = 50
t0 = time.perf_counter()
for _ in range(reps):
    rows_h, flags_h, counts_h = gpu_path()
torch.cuda.synchronize()
gpu_ms = 1e3 * (time.perf_counter() - t0) / reps
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    rows, counts = evalmap.postfilter(out, 200, 0.25)
    flags = evalmap.match_flags(labels, rows, 0.5, False)
e1.record()
torch.cuda.synchronize()
kern_ms = e0.elapsed_time(e1) / reps

# host path of the reference: full output to the host, numpy filter, MApMetric.update
from oracle import map_oracle, ref_map
lab_h = inputs['lab']
t0 = time.perf_counter()
out_h = out.cpu().numpy()
t_copy = time.perf_counter() - t0
t0 = time.perf_counter()
wrows, wcounts = map_oracle.postfilter(out_h, 200, 0.25)
t_filter = time.perf_counter() - t0
t0 = time.perf_counter()
if ref_map.available():
    kind = "reference"
    ref_map.run_metric(lab_h, wrows, 0.5, False)
else:
    kind = "port"
    acc = map_oracle.MApAccumulator(False, False)
    acc.update(lab_h, wrows, map_oracle.match_flags(lab_h, wrows, 0.5, False))
t_match = time.perf_counter() - t0
assert np.array_equal(rows_h.numpy(), wrows) and np.array_equal(flags_h.numpy(), map_oracle.match_flags(lab_h, wrows, 0.5, False))
print(json.dumps({"workload": "post-filter (id >= 0, score > 0.25, pad 200) + mAP TP/FP matching, SSD-512 detections, batch %d" % B,
                  "gpu_images_per_s_with_readback": B / (gpu_ms * 1e-3), "gpu_ms_with_readback": gpu_ms, "gpu_kernels_ms": kern_ms,
                  "d2h_bytes": int(rows_h.numel() * 4 + flags_h.numel() * 4 + counts_h.numel() * 4),
                  "cpu_baseline": {"kind": kind, "cores": 1, "images_per_s": B / (t_copy + t_filter + t_match),
                                   "ms": {"asnumpy_of_full_output": 1e3 * t_copy, "numpy_filter": 1e3 * t_filter, "metric_update": 1e3 * t_match},
                                   "d2h_bytes": int(out_h.nbytes)},
                  "surviving_rows_per_image": float(wcounts.mean()), "identical_to_cpu": True}))
