"""Standalone NMS sweep (BASELINE.json configs[4]): N = 1k..200k boxes, IoU 0.45, single class (force_suppress on)
and per-class, device time vs the reference's cpu_nms (oracle/_ref when present, else the oracle port)."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from dspnet_b200 import synth
from dspnet_b200.nms import nms_device
from oracle import oracle as O, ref as R

dev = torch.device('cuda', 0)
cpu_nms = R.cpu_nms if R.nms_available() else O.cpu_nms
rows = []
for n in (1000, 2000, 5000, 10000, 20000, 50000, 100000, 200000):
    dets = synth.nms_boxes(1000 + n, n)
    d = torch.from_numpy(dets).to(dev)
    keep, num = nms_device(d, 0.45, rule='ge')
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5 if n <= 50000 else 2
    ev0.record()
    for _ in range(reps):
        keep, num = nms_device(d, 0.45, rule='ge')
    ev1.record()
    torch.cuda.synchronize()
    gpu_ms = ev0.elapsed_time(ev1) / reps
    k = int(num.item())
    got = keep[:k].cpu().tolist()
    row = {"n": n, "kept": k, "gpu_ms": gpu_ms, "pairs_per_s": n * (n - 1) / 2 / (gpu_ms * 1e-3)}
    if n <= 100000:
        t0 = time.perf_counter()
        want = cpu_nms(dets, 0.45)
        row["cpu_ms"] = 1e3 * (time.perf_counter() - t0)
        row["equal"] = got == want
        row["speedup"] = row["cpu_ms"] / gpu_ms
    # per-class (force_suppress off): 20 classes
    dc = synth.nms_boxes(2000 + n, n, with_class=True, num_classes=20)
    dcd = torch.from_numpy(dc).to(dev)
    nms_device(dcd, 0.45, rule='ge', class_col=5)
    torch.cuda.synchronize()
    ev0.record()
    kc, nc = nms_device(dcd, 0.45, rule='ge', class_col=5)
    ev1.record()
    torch.cuda.synchronize()
    row["gpu_ms_per_class"] = ev0.elapsed_time(ev1)
    if n <= 20000:
        kept = []
        for c in range(20):
            idx = np.nonzero(dc[:, 5] == c)[0]
            kept += [int(idx[i]) for i in cpu_nms(dc[idx, :5], 0.45)]
        kept.sort(key=lambda i: -dc[i, 4])
        row["per_class_equal"] = kc[: int(nc.item())].cpu().tolist() == kept
    rows.append(row)
    print(json.dumps(row), flush=True)
