#!/usr/bin/env python
"""Benchmark of the multibox hot path (BASELINE.json: "SSD-512 multibox target+detect images/s at 1/2/4/8 B200;
% of HBM peak").

    python bench.py --gpus N --steps K --warmup W [--workload NAME]   # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...           # the reference's CPU path on the host cores

Workloads (one per BASELINE.json config; the default is configs[1], the one the metric is quoted on):

    detection   SSD-512 VOC MultiBoxDetection + NMS, 32 images per GPU (weak scaling)           configs[1]
    target      SSD-512 MultiBoxTarget, mining ratio 3, 64 images split over the GPUs (strong)  configs[2]
    ssd300      SSD-300 prior + target + detection, batch 1                                     configs[0]
    dspnet_cs   DSPNet Cityscapes 1024x512 head, prior + target + detection, 16 images split    configs[3]
    nms         standalone NMS sweep 1k-200k boxes, force_suppress on / off                     configs[4]
    detection_heads  SURVEY 8f row f1: SSD-512 detection fed by the per-scale conv heads (layout shuffles + channel
                     softmax fused into the stream kernel), 32 images per GPU
    train_tail       SURVEY 8f row f2: SSD-512 MultiBoxTarget + training-graph forward (softmax, masked smooth-L1) +
                     MultiBoxMetric statistics, 64 images split over the GPUs

A step is one pass of the workload's operators over one batch of synthetic head tensors.  Inputs rotate over resident
copies whose total footprint exceeds the 126 MB L2 (or, for the small configs, L2 is flushed between event-timed
steps); the timed region is K steps between barriers, repeated `timed_blocks` times -- the line reports the median
block (every block is listed).  One JSON line is printed by rank 0 (keys: see the task contract; `roofline`,
`cpu_baseline`, `e2e`, `parity_check`, `clocks` are described in DESIGN.md section 5).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DET_PARAMS = dict(threshold=0.01, clip=True, nms_threshold=0.45, force_suppress=False, nms_topk=400,
                  variances=(0.1, 0.1, 0.2, 0.2))
TGT_PARAMS = dict(overlap_threshold=0.5, ignore_label=-1.0, negative_mining_ratio=3.0, negative_mining_thresh=0.5,
                  minimum_negative_samples=0, variances=(0.1, 0.1, 0.2, 0.2))
L2_BYTES = 126 * 2 ** 20
NMS_SIZES = (1000, 2000, 5000, 10000, 20000, 50000, 100000, 200000)
NMS_THRESH = 0.45

WORKLOADS = {
    "detection": dict(preset="ssd512", batch=32, scaling="weak", ops=("detection",), config_id=2, max_gt=8,
                      metric="ssd512_multibox_detection_images_per_s",
                      name="ssd512_voc21_multibox_detection_nms_batch32"),
    "target": dict(preset="ssd512", batch=64, scaling="strong", ops=("target",), config_id=2, max_gt=8,
                   metric="ssd512_multibox_target_images_per_s",
                   name="ssd512_voc21_multibox_target_mining3_batch64"),
    "ssd300": dict(preset="ssd300", batch=1, scaling="weak", ops=("prior", "target", "detection"), config_id=0,
                   max_gt=8, metric="ssd300_multibox_prior_target_detection_images_per_s",
                   name="ssd300_voc21_prior_target_detection_batch1"),
    "dspnet_cs": dict(preset="dspnet_cs", batch=16, scaling="strong", ops=("prior", "target", "detection"),
                      config_id=3, max_gt=50, metric="dspnet_cs_multibox_prior_target_detection_images_per_s",
                      name="dspnet_cityscapes_1024x512_prior_target_detection_batch16"),
    "nms": dict(metric="nms_sweep_boxes_per_s", name="standalone_nms_sweep_1k_200k_iou045_force_on_off",
                scaling="weak"),
    "detection_heads": dict(preset="ssd512", batch=32, scaling="weak", ops=("detection_heads",), config_id=2, max_gt=8,
                            metric="ssd512_multibox_detection_from_heads_images_per_s",
                            name="ssd512_voc21_layout_softmax_multibox_detection_nms_batch32"),
    "train_tail": dict(preset="ssd512", batch=64, scaling="strong", ops=("target", "loss"), config_id=2, max_gt=8,
                       metric="ssd512_multibox_target_loss_metric_images_per_s",
                       name="ssd512_voc21_multibox_target_softmax_smoothl1_metric_batch64"),
}
# the names older scripts use for the default workload
PRESET, BATCH = WORKLOADS["detection"]["preset"], WORKLOADS["detection"]["batch"]


def det_algorithmic_bytes(B, A, C):
    """4*B*C*A + 4*B*A*5 + 16*A read, 28*B*A written."""
    return 4 * B * C * A + 20 * B * A + 16 * A + 28 * B * A


def tgt_algorithmic_bytes(B, A, C, L):
    """16*A + 24*B*L + 4*B*C*A read, 4*B*A*(2*5+1) written."""
    return 16 * A + 24 * B * L + 4 * B * C * A + 44 * B * A


def prior_algorithmic_bytes(A):
    return 16 * A


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(name)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled in a side process while the GPU is under load."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.proc.wait()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.tmp.close()
        os.unlink(self.tmp.name)
        sm.sort()
        # the median of the upper half approximates "under load" when idle samples are mixed in
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# synthetic inputs (numpy, host) and the CPU arm
# ---------------------------------------------------------------------------------------------------------------
def make_inputs(first_image, B, workload="detection"):
    """Seeded synthetic head tensors of the workload's preset (numpy, host): SURVEY.md section 8d."""
    import numpy as np
    from dspnet_b200 import presets, synth
    w = WORKLOADS[workload]
    p = presets.PRESETS[w["preset"]]
    A = presets.num_anchors(p)
    cid = w["config_id"]
    d = dict(A=A, C=p.num_classes, L=p.label_slots, preset=w["preset"], B=B)
    if "detection" in w["ops"]:
        d["prob"] = synth.cls_prob(cid, B, p.num_classes, A, first_image=first_image)
        d["loc"] = synth.loc_pred(cid, B, A, first_image=first_image)
    if "target" in w["ops"] or workload == "detection":  # (detection keeps them for the profiling scripts)
        d["lab"] = synth.labels(cid, B, p.label_slots, p.num_classes, max_gt=w["max_gt"], first_image=first_image)
        d["logits"] = synth.cls_preds(cid, B, p.num_classes, A, first_image=first_image)
    return d, np


def cpu_backend():
    """The reference's own operator .cc files compiled in place (oracle/_ref) when present, else the restatement
    (bit-identical, tests/test_oracle_golden.py)."""
    from oracle import ref as R
    if R.available():
        return R, "reference"
    from oracle import oracle as O
    return O, "port"


def oracle_anchors(preset=PRESET):
    import numpy as np
    from dspnet_b200 import presets
    M, _ = cpu_backend()
    p = presets.PRESETS[preset]
    return np.concatenate([M.multibox_prior(fm.height, fm.width, fm.sizes, fm.ratios, False, (fm.step, fm.step))
                           for fm in p.maps], axis=1)


def cpu_pass(workload, inputs, anchors, threads, collect=False):
    """One step of the workload on the host: the reference loops are single-threaded per image, so the batch is cut
    into contiguous image slices, one per host thread (the best the CPU path can do without changing it).
    Returns the per-slice outputs when collect is set."""
    from concurrent.futures import ThreadPoolExecutor
    M, _ = cpu_backend()
    w = WORKLOADS[workload]
    B = inputs["B"]
    n = max(1, min(threads, B))
    bounds = [(i * B // n, (i + 1) * B // n) for i in range(n)]

    def run(be):
        b, e = be
        res = {}
        an = anchors
        if "prior" in w["ops"] and b == 0:
            an = oracle_anchors(w["preset"])
        if "target" in w["ops"]:
            res["target"] = M.multibox_target(an, inputs["lab"][b:e], inputs["logits"][b:e], **TGT_PARAMS)
        if "detection" in w["ops"]:
            res["detection"] = M.multibox_detection(inputs["prob"][b:e], inputs["loc"][b:e], an, **DET_PARAMS)
        return res
    if n == 1:
        parts = [run(bounds[0])]
    else:
        with ThreadPoolExecutor(n) as ex:
            parts = list(ex.map(run, bounds))
    return parts if collect else None


def cpu_time(workload, inputs, anchors, threads, steps, warmup, budget_s=25.0):
    """images/s of the CPU arm: `steps` timed passes after `warmup`, cut short once `budget_s` of CPU work is spent
    (the sample is reported).  ONE method for the cpu_baseline object and for --impl reference."""
    for _ in range(max(warmup, 1)):
        cpu_pass(workload, inputs, anchors, threads)
    done, t0 = 0, time.perf_counter()
    while done < steps:
        cpu_pass(workload, inputs, anchors, threads)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return inputs["B"] * done / dt, done, dt


def nms_cpu_time(sizes, budget_s=25.0):
    """boxes/s of the reference's Cython cpu_nms (one thread, as shipped) over the sweep sizes that fit the budget."""
    from dspnet_b200 import synth
    from oracle import ref as R
    from oracle import oracle as O
    use_ref = R.nms_available()
    boxes, t_total, done = 0, 0.0, []
    for n in sizes:
        dets = synth.nms_boxes(100 + n, n)
        t0 = time.perf_counter()
        (R.cpu_nms if use_ref else O.cpu_nms)(dets, NMS_THRESH)
        dt = time.perf_counter() - t0
        boxes += n
        t_total += dt
        done.append(n)
        if t_total + 4.5 * dt > budget_s:  # the next size costs >= 4x (O(N^2))
            break
    return boxes / t_total, done, "reference" if use_ref else "port"


def workload_config(workload, inputs, batch_per_gpu):
    w = WORKLOADS[workload]
    cfg = {"workload": w["name"], "preset": w["preset"], "batch_per_gpu": batch_per_gpu, "anchors": inputs["A"],
           "classes": inputs["C"]}
    if "target" in w["ops"]:
        cfg["label_slots"] = inputs["L"]
        cfg.update({"target_" + k: v for k, v in TGT_PARAMS.items() if k != "variances"})
    if "detection" in w["ops"]:
        cfg.update({k: v for k, v in DET_PARAMS.items()})
    return cfg


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host threads; rank 0 only.
    (MXNet itself cannot be installed here -- DESIGN.md section 9 -- so the operator bodies are driven directly.)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    if args.workload == "nms":
        v, done, kind = nms_cpu_time(NMS_SIZES)
        line = {"impl": "reference", "metric": w["metric"], "value": v, "unit": "boxes/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": w["name"], "sizes": list(NMS_SIZES), "thresh": NMS_THRESH},
                "cpu_baseline": {"value": v, "unit": "boxes/s", "cores": 1, "kind": kind,
                                 "sample": "cpu_nms (single-threaded as shipped) on N = %s, one pass each" % done},
                "e2e": {"value": v, "unit": "boxes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return
    if args.workload in ("detection_heads", "train_tail"):
        B = w["batch"]
        inputs = extra_inputs(args.workload, 0, B)
        anchors = oracle_anchors(w["preset"])
        value, done, dt = extra_cpu_time(args.workload, inputs, anchors, threads, args.steps, budget_s=120.0)
        line = {"impl": "reference", "metric": w["metric"], "value": value, "unit": "images/s", "n_gpus": args.gpus,
                "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True,
                "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": w["name"], "preset": w["preset"], "batch_per_gpu": B, "anchors": inputs["A"],
                           "classes": inputs["C"]},
                "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                                 "sample": "%d steps x one B=%d batch, image slices over %d host threads" % (done, B, threads)},
                "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return
    # weak workloads: one GPU's batch (what one step of the GPU arm processes per GPU); strong: the whole batch
    B = w["batch"]
    inputs, _ = make_inputs(0, B, args.workload)
    anchors = oracle_anchors(w["preset"])
    _, kind = cpu_backend()
    value, done, dt = cpu_time(args.workload, inputs, anchors, threads, args.steps, args.warmup, budget_s=120.0)
    line = {
        "impl": "reference", "metric": w["metric"], "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done,
        "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, inputs, B),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": kind,
                         "sample": "%d steps x one B=%d %s batch, image slices over %d host threads"
                                   % (done, B, w["preset"], threads)},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
class E2EPipeline:
    """The reference-facing call with HOST buffers: per step the inputs travel pinned host -> device, the public
    operators run, and every output is copied back to pinned host memory.  Three streams with double-buffered device
    tensors let the copy-in of step i+1, the operators of step i and the copy-out of step i-1 overlap (PCIe is full
    duplex); every step still moves its own inputs in and its own results out."""

    def __init__(self, torch, dev, host_inputs, run):
        self.torch, self.run = torch, run
        self.pin_in = [torch.from_numpy(x).pin_memory() for x in host_inputs]
        self.d_in = [[torch.empty(p.shape, dtype=p.dtype, device=dev) for p in self.pin_in] for _ in range(2)]
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_run = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.d_out = [None, None]
        self.pin_out = None
        self.n = 0
        self.h2d = sum(p.numel() * p.element_size() for p in self.pin_in)
        self.d2h = 0

    def step(self):
        torch, i = self.torch, self.n
        k = i & 1
        with torch.cuda.stream(self.s_in):
            if i >= 2:
                self.s_in.wait_event(self.ev_run[k])  # the operators that read this buffer two steps ago are done
            for d, p in zip(self.d_in[k], self.pin_in):
                d.copy_(p, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.ev_in[k])
            if i >= 2:
                self.s_run.wait_event(self.ev_out[k])  # the previous results in this slot have been copied out
            self.d_out[k] = self.run(*self.d_in[k])
            self.ev_run[k].record(self.s_run)
        if self.pin_out is None:
            self.pin_out = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in self.d_out[k]]
            self.d2h = sum(p.numel() * p.element_size() for p in self.pin_out)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_run[k])
            for p, o in zip(self.pin_out, self.d_out[k]):
                p.copy_(o, non_blocking=True)
            self.ev_out[k].record(self.s_out)
        self.n = i + 1


# ---------------------------------------------------------------------------------------------------------------
# SURVEY.md section 8f rows f1 / f2 (the callers either side of the hot path)
# ---------------------------------------------------------------------------------------------------------------
def extra_inputs(workload, first_image, B):
    from dspnet_b200 import presets, synth
    w = WORKLOADS[workload]
    p = presets.PRESETS[w["preset"]]
    A, C = presets.num_anchors(p), p.num_classes
    d = dict(A=A, C=C, L=p.label_slots, B=B, preset=w["preset"])
    if workload == "detection_heads":
        d["logits"] = synth.det_logits(w["config_id"], B, C, A, first_image=first_image)
        d["loc"] = synth.loc_pred(w["config_id"], B, A, first_image=first_image)
        d["cls_heads"], d["loc_heads"] = synth.heads_from_logits(p, d["logits"], d["loc"])
        d["shapes"] = [(fm.height, fm.width, len(fm.sizes) + len(fm.ratios) - 1) for fm in p.maps]
    else:
        d["lab"] = synth.labels(w["config_id"], B, p.label_slots, C, max_gt=w["max_gt"], first_image=first_image)
        d["logits"] = synth.cls_preds(w["config_id"], B, C, A, first_image=first_image)
        d["loc"] = synth.loc_pred(w["config_id"], B, A, first_image=first_image)
    return d


def extra_cpu_pass(workload, inputs, anchors, threads, collect=False):
    """The reference graph of the row on the host, image slices over the host threads: f1 = layout shuffles (numpy) ->
    channel softmax -> MultiBoxDetection; f2 = MultiBoxTarget -> softmax / smooth-L1 / metric sums."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    M, _ = cpu_backend()
    B = inputs["B"]
    n = max(1, min(threads, B))
    bounds = [(i * B // n, (i + 1) * B // n) for i in range(n)]

    def run(be):
        b, e = be
        if workload == "detection_heads":
            cp, lp = O.head_layout([h[b:e] for h in inputs["cls_heads"]], [h[b:e] for h in inputs["loc_heads"]], inputs["C"])
            return M.multibox_detection(O.softmax_channel(cp), lp, anchors, **DET_PARAMS)
        tgt = M.multibox_target(anchors, inputs["lab"][b:e], inputs["logits"][b:e], **TGT_PARAMS)
        return tgt, O.multibox_training_outputs(inputs["logits"][b:e], inputs["loc"][b:e], *tgt)
    if n == 1:
        parts = [run(bounds[0])]
    else:
        with ThreadPoolExecutor(n) as ex:
            parts = list(ex.map(run, bounds))
    return parts if collect else None


def extra_cpu_time(workload, inputs, anchors, threads, steps, budget_s):
    extra_cpu_pass(workload, inputs, anchors, threads)
    done, t0 = 0, time.perf_counter()
    while done < steps:
        extra_cpu_pass(workload, inputs, anchors, threads)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return inputs["B"] * done / dt, done, dt


def run_extra(args, torch, dist, dev, rank, world):
    import ctypes
    import numpy as np
    from dspnet_b200 import _lib
    from dspnet_b200.plan import DetectionHeadsPlan, TargetPlan
    from dspnet_b200.symbol import multibox_anchors
    from dspnet_b200.loss import multibox_training_outputs
    wl = args.workload
    w = WORKLOADS[wl]
    lib = _lib.lib()
    Bg = w["batch"] if w["scaling"] == "weak" else max(1, w["batch"] // world)
    inputs = extra_inputs(wl, rank * Bg, Bg)
    A, C, L = inputs["A"], inputs["C"], inputs["L"]
    anchors = multibox_anchors(w["preset"], device=dev)
    rotate = 4
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    if wl == "detection_heads":
        plan = DetectionHeadsPlan(Bg, A, C, inputs["shapes"], dev, **DET_PARAMS)
        sets = []
        for _ in range(rotate):
            ch, lh = [t(h) for h in inputs["cls_heads"]], [t(h) for h in inputs["loc_heads"]]
            sets.append((ch, lh, plan.bind(ch, lh), plan.new_output()))
        abytes = 4 * Bg * C * A + 20 * Bg * A + 16 * A + 28 * Bg * A  # every class row is read now (softmax), 28 B/anchor out
        avoided = 2 * 4 * Bg * C * A                                   # cls_prob written by the softmax, read by the operator
        avoided_shuffles = 3 * 2 * 4 * Bg * C * A                      # transpose, concat, transpose of the class tensor

        def step(i):
            s = sets[i % rotate]
            plan.run(s[2], anchors, s[3])
        launches = lambda: plan.launches_per_run
        names = [(1, "det_stream_heads_kernel"), (2, "det_sort_kernel"), (4, "det_pair_kernel")]
    else:
        tplan = TargetPlan(Bg, A, L, C, dev, **TGT_PARAMS)
        lab_d = t(inputs["lab"])
        sets = [(t(inputs["logits"]), t(inputs["loc"]), tplan.new_outputs(), tplan.new_stats()) for _ in range(rotate)]
        ws = torch.empty(max(int(lib.dspmb_multibox_loss_workspace_bytes(Bg, A)), 256), dtype=torch.uint8, device=dev)
        stats = [torch.empty((Bg, 4), dtype=torch.float64, device=dev) for _ in range(rotate)]
        p = lambda x: ctypes.c_void_p(x.data_ptr())
        # statistics-only pass behind the target: three loc tensors + labels are read (64 B / anchor), the logits only
        # for the labelled anchors (a few percent; not counted), 32 B of sums per image are written
        loss_bytes = 64 * Bg * A + 32 * Bg
        abytes = tgt_algorithmic_bytes(Bg, A, C, L) + loss_bytes
        # the reference graph writes cls_prob and loc_loss and the metric reads them back (plus SoftmaxOutput's own read
        # of cls_preds, which the statistics kernel skips for unlabelled anchors)
        avoided = 2 * 4 * Bg * C * A + 2 * 20 * Bg * A + 4 * Bg * C * A

        def step(i):
            s = sets[i % rotate]
            tplan.run(anchors, lab_d, s[0], s[2], stats=s[3])
            rc = lib.dspmb_multibox_loss_f32(p(s[0]), p(s[1]), p(s[2][0]), p(s[2][1]), p(s[2][2]), None, None,
                                             p(stats[i % rotate]), Bg, A, C, ctypes.c_float(1e-8), p(ws), ws.numel(),
                                             ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            if rc:
                _lib.check(rc)
        launches = lambda: (tplan.launches_per_run or 0) + 2
        names = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(dev.index) if rank == 0 else None
    for i in range(max(args.warmup, 3)):
        step(i)
    soak = 0 if args.no_soak else 4000
    for i in range(soak):
        step(i)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    block_ms = []
    for _ in range(1 if args.no_soak else 9):
        barrier()
        ev0.record()
        for k in range(args.steps):
            step(k)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        block_ms.append(ms)
    ms = sorted(block_ms)[len(block_ms) // 2]
    # per-kernel times (profile events around every launch, direct launches)
    cache_was = lib.dspmb_set_tuning(_lib.TUNE_GRAPH_CACHE, 0)
    lib.dspmb_profile_enable(1)
    for k in range(20):
        step(k)
    torch.cuda.synchronize()
    ms_k, ln_k = (ctypes.c_float * 32)(), (ctypes.c_int * 32)()
    nslots = lib.dspmb_profile_read(ms_k, ln_k, 32)
    lib.dspmb_profile_enable(0)
    lib.dspmb_set_tuning(_lib.TUNE_GRAPH_CACHE, cache_was)
    lib.dspmb_profile_kernel_name.restype = ctypes.c_char_p
    kernels = {lib.dspmb_profile_kernel_name(i).decode(): ms_k[i] / 20 for i in range(nslots) if ln_k[i]}
    if wl == "detection_heads" and "det_stream_kernel" in kernels:
        kernels["det_stream_heads_kernel"] = kernels.pop("det_stream_kernel")
    # parity of the last outputs against the oracle's walk through the reference graph
    M, kind = cpu_backend()
    ref_anchors = oracle_anchors(w["preset"])
    want = extra_cpu_pass(wl, inputs, ref_anchors, os.cpu_count() or 1, collect=True)
    if wl == "detection_heads":
        got = sets[(args.steps - 1) % rotate][3].cpu().numpy()
        exp = np.concatenate(want, axis=0)
        ok = bool(np.array_equal(got.view(np.uint32), exp.view(np.uint32)))
        checked = ["detection (B,A,7) bit-exact against head_layout -> softmax_channel -> multibox_detection"]
    else:
        s = sets[(args.steps - 1) % rotate]
        ok = True
        for k in range(3):
            g = s[2][k].cpu().numpy()
            e = np.concatenate([pp[0][k] for pp in want], axis=0).reshape(g.shape)
            ok = ok and bool(((g.view(np.uint32) == e.view(np.uint32)) | ((g == 0) & (e == 0))).all())
        st = stats[(args.steps - 1) % rotate].cpu().numpy()
        est = np.concatenate([pp[1][2] for pp in want], axis=0)
        ok = ok and bool(np.array_equal(st[:, 0], est[:, 0])) and bool(np.allclose(st[:, 1:3], est[:, 1:3], rtol=1e-6))
        checked = ["loc_target / loc_mask / cls_target bit-exact", "valid count exact, cross-entropy and smooth-L1 sums to 1e-6"]
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = int(flag.item()) == 1
    # end to end with host buffers through the public operators
    e2e = None
    if not args.no_e2e:
        from dspnet_b200 import MultiBoxDetectionFromHeads, MultiBoxTarget
        if wl == "detection_heads":
            host_in = list(inputs["cls_heads"]) + list(inputs["loc_heads"])
            k = len(inputs["cls_heads"])

            def e2e_run(*d):
                return [MultiBoxDetectionFromHeads(list(d[:k]), list(d[k:]), anchors, C, **DET_PARAMS)]
        else:
            host_in = [inputs["lab"], inputs["logits"], inputs["loc"]]

            def e2e_run(lab, logits, loc):
                lt, lm, ct = MultiBoxTarget(anchors, lab, logits, **TGT_PARAMS)
                _, _, st = multibox_training_outputs(logits, loc, lt, lm, ct, want_cls_prob=False, want_loc_loss=False)
                return [lt, lm, ct, st]
        pipe = E2EPipeline(torch, dev, host_in, e2e_run)
        e2e_steps = max(100, args.steps)  # the 3-stream pipeline needs a few dozen steps to show its steady state
        for _ in range(4):
            pipe.step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pipe.step()
        torch.cuda.synchronize()
        e2e_ms = 1e3 * (time.perf_counter() - t0)
        barrier()
        if world > 1:
            tt = torch.tensor([e2e_ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_ms = float(tt.item())
        e2e = {"value": Bg * world * e2e_steps / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": int(pipe.h2d),
               "d2h_bytes_per_step": int(pipe.d2h), "steps": e2e_steps,
               "api": ("dspnet_b200.MultiBoxDetectionFromHeads" if wl == "detection_heads" else
                       "dspnet_b200.MultiBoxTarget + dspnet_b200.loss.multibox_training_outputs (statistics only)") +
                      "; pinned host -> device copy of every input, the operators, every output copied back"}
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    step_ms = ms / args.steps
    dominant = max(kernels, key=kernels.get)
    rk = "det_stream_heads_kernel" if wl == "detection_heads" else "multibox_loss_kernel"
    rbytes = abytes if wl == "detection_heads" else loss_bytes
    achieved = rbytes / (kernels[rk] * 1e-3) / 1e9
    line = {"metric": w["metric"], "value": Bg * world * args.steps / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "preset": w["preset"], "batch_per_gpu": Bg, "anchors": A, "classes": C,
                       "l2": "inputs/outputs rotate over %d resident sets (> 126 MB L2)" % rotate,
                       "parallelism": "one GPU" if world == 1 else "images sharded, %d per GPU, no exchange" % Bg,
                       "survey_row": "8f " + ("f1" if wl == "detection_heads" else "f2")},
            "soak_steps": soak, "timed_blocks": len(block_ms), "block_ms": [round(x, 4) for x in block_ms],
            "roofline": {"bound": "hbm", "kernel": rk, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": kernel_traffic(rk), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": rbytes, "kernel_ms": kernels[rk], "all_kernels_ms": kernels,
                         "dominant_kernel": dominant, "whole_op_frac": abytes / (step_ms * 1e-3) / 1e9 / peak,
                         "whole_op_algorithmic_bytes": abytes,
                         "hbm_bytes_the_fusion_removes": avoided,
                         "note": ("cls_prob (B,C,A) is neither written by a softmax pass nor read by the operator; the "
                                  "reference graph additionally moves the class tensor through transpose / Concat / "
                                  "transpose (%d more bytes)" % avoided_shuffles) if wl == "detection_heads" else
                                 "the metric's inputs (cls_prob, loc_loss) stay in registers; only (B,4) sums are written"},
            "e2e": e2e, "gpu_launches": args.steps * launches(), "launches_per_step": launches(),
            "parity_check": {"result": "ok" if ok else "MISMATCH", "checked": checked,
                             "against": "oracle restatement of the reference graph (MXNet's softmax / smooth_l1 are not "
                                        "in the reference tree: parity pinned on multibox_target.cc:220-231's softmax)"},
            "clocks": clocks}
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        v, done, dt = extra_cpu_time(wl, inputs, ref_anchors, threads, steps=10, budget_s=20.0)
        line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                                "sample": "%d steps x one B=%d batch: layout shuffles / softmax (oracle) + the reference "
                                          "operator, image slices over %d host threads" % (done, Bg, threads)}
    print(json.dumps(line))


def run_nms(args, torch, dist, dev, rank, world):
    """configs[4]: standalone NMS sweep, force_suppress on (single class) and off (20 classes), device-resident and
    through the host-buffer helpers (the cpu_nms / gpu_nms drop-ins).  One box set does not shard (greedy
    dependency): with N GPUs every rank runs the sweep on its own box sets (replicas)."""
    from dspnet_b200 import _lib, synth
    from dspnet_b200 import nms as N
    from oracle import ref as R
    from oracle import oracle as O
    sizes = NMS_SIZES
    sets = {}
    for n in sizes:
        d1 = synth.nms_boxes(100 + n + 7919 * rank, n)
        d2 = synth.nms_boxes(300 + n + 7919 * rank, n, with_class=True, num_classes=20)
        sets[n] = (d1, d2, torch.from_numpy(d1).to(dev), torch.from_numpy(d2).to(dev))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib = _lib.lib()
    launches = [0]

    def sweep(count=False):
        for n in sizes:
            _, _, g1, g2 = sets[n]
            N.nms_device(g1, NMS_THRESH, rule="ge")
            if count:
                launches[0] += lib.dspmb_last_launch_count()
            N.nms_device(g2, NMS_THRESH, rule="ge", class_col=5)
            if count:
                launches[0] += lib.dspmb_last_launch_count()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(dev.index) if rank == 0 else None
    warm = max(args.warmup, 3)
    for k in range(warm):
        sweep(count=(k == 0))
    barrier()
    ev0.record()
    for _ in range(args.steps):
        sweep()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    boxes_per_step = 2 * sum(sizes)
    # per-size figures (device-resident), pairs/s and the kept-list check against cpu_nms
    cpu_nms_ref = R.cpu_nms if R.nms_available() else O.cpu_nms
    per_size, parity_ok, checked = [], True, []
    limit = 200000 if args.full_parity else 20000
    for n in sizes:
        d1, d2, g1, g2 = sets[n]
        row = {"n": n}
        for tag, g, kw in (("force", g1, {}), ("per_class", g2, {"class_col": 5})):
            for _ in range(2):
                N.nms_device(g, NMS_THRESH, rule="ge", **kw)
            torch.cuda.synchronize()
            reps = 20 if n <= 20000 else 5
            ev0.record()
            for _ in range(reps):
                keep, num = N.nms_device(g, NMS_THRESH, rule="ge", **kw)
            ev1.record()
            torch.cuda.synchronize()
            row[tag + "_ms"] = ev0.elapsed_time(ev1) / reps
            row[tag + "_kept"] = int(num.item())
        row["force_pairs_per_s"] = n * (n - 1) / 2 / (row["force_ms"] * 1e-3)
        if rank == 0 and n <= limit:
            want = cpu_nms_ref(d1, NMS_THRESH)
            keep, num = N.nms_device(g1, NMS_THRESH, rule="ge")
            parity_ok &= keep[: int(num.item())].cpu().tolist() == want
            checked.append(n)
        if rank == 0 and R.gpu_nms_available() and n <= 100000:
            # same-box comparator: the reference's own GPU NMS (cython/nms_kernel.cu compiled unmodified for sm_100a)
            # against this library through the SAME host-pointer ABI (_nms / dspmb_nms_host: rows presorted, host in,
            # host out, synchronous, allocation and copies included on both sides), iou > thresh rule on both
            import ctypes
            import numpy as np_
            srt = np_.ascontiguousarray(d1[np_.argsort(-d1[:, 4], kind="stable")])
            kp = np_.empty(n, np_.int32)
            num_c = ctypes.c_int(0)
            reps_h = 5 if n <= 20000 else 2
            R.gpu_nms_sorted(srt, NMS_THRESH)
            t0 = time.perf_counter()
            for _ in range(reps_h):
                ref_keep = R.gpu_nms_sorted(srt, NMS_THRESH)
            row["ref_gpu_nms_host_ms"] = (time.perf_counter() - t0) / reps_h * 1e3
            call = lambda: lib.dspmb_nms_host(kp.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), ctypes.byref(num_c),
                                              srt.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n, 5,
                                              ctypes.c_float(NMS_THRESH), dev.index)
            call()
            t0 = time.perf_counter()
            for _ in range(reps_h):
                call()
            row["ours_nms_host_ms"] = (time.perf_counter() - t0) / reps_h * 1e3
            row["same_keep_as_ref_gpu_nms"] = bool(np_.array_equal(kp[: num_c.value], ref_keep))
        per_size.append(row)
    # per-kernel device times at the largest size (profile events around every launch, direct launches)
    kernels_ms = {}
    if rank == 0:
        import ctypes
        lib.dspmb_profile_enable(1)
        N.nms_device(sets[sizes[-1]][2], NMS_THRESH, rule="ge")
        torch.cuda.synchronize()
        ms_k = (ctypes.c_float * 32)()
        ln_k = (ctypes.c_int * 32)()
        nslots = lib.dspmb_profile_read(ms_k, ln_k, 32)
        lib.dspmb_profile_enable(0)
        lib.dspmb_profile_kernel_name.restype = ctypes.c_char_p
        kernels_ms = {lib.dspmb_profile_kernel_name(i).decode(): {"ms": ms_k[i], "launches": ln_k[i]}
                      for i in range(nslots) if ln_k[i]}
    # end to end through the host helpers
    e2e_sizes = [n for n in sizes if n <= 50000]
    for n in e2e_sizes[:2]:
        N.cpu_nms(sets[n][0], NMS_THRESH)
    t0 = time.perf_counter()
    for n in e2e_sizes:
        N.cpu_nms(sets[n][0], NMS_THRESH)
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        return
    w = WORKLOADS["nms"]
    top = per_size[-1]
    peak, peak_src = measured_peaks()
    gbs = 24.0 * top["n"] / (top["force_ms"] * 1e-3) / 1e9
    dominant = max(kernels_ms, key=lambda k: kernels_ms[k]["ms"]) if kernels_ms else "nms_cull_kernel"
    line = {"metric": w["metric"], "value": boxes_per_step * world * args.steps / (ms * 1e-3), "unit": "boxes/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "sizes": list(sizes), "thresh": NMS_THRESH, "rule": "cpu_nms (>=, double)",
                       "modes": ["force_suppress (single class)", "per class (20 classes)"],
                       "parallelism": "replicas only: one box set does not shard",
                       "l2": "box sets are KB-MB sized; the work is O(N^2) pair tests, not HBM traffic"},
            "roofline": {"bound": "hbm", "kernel": "%s at N=%d (dominant)" % (dominant, top["n"]), "achieved": gbs,
                         "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": kernel_traffic(dominant),
                         "peak_source": peak_src, "all_kernels_ms_at_top_size": kernels_ms,
                         "note": "formality (20 B read + 4 B written per box over the whole call): the sweep is "
                                 "ALU / latency bound -- O(N x kept) pair tests -- see pairs_per_s (all N(N-1)/2 pairs "
                                 "the greedy rule ranges over, per second)", "pairs_per_s": top["force_pairs_per_s"]},
            "sweep": per_size,
            "e2e": {"value": sum(e2e_sizes) / e2e_s, "unit": "boxes/s",
                    "h2d_bytes_per_step": 20 * sum(e2e_sizes), "d2h_bytes_per_step": 4 * sum(e2e_sizes),
                    "api": "dspnet_b200.nms.cpu_nms(dets numpy) for N <= 50000: host array in, kept-index list out"},
            "gpu_launches": launches[0] * args.steps, "launches_per_step": launches[0],
            "parity_check": {"result": "ok" if parity_ok else "MISMATCH", "against": "cython cpu_nms (oracle/_ref)",
                             "what": "kept-index lists at N = %s" % checked},
            "clocks": clocks}
    if not args.no_cpu_baseline:
        v, done, kind = nms_cpu_time(sizes)
        line["cpu_baseline"] = {"value": v, "unit": "boxes/s", "cores": 1, "kind": kind,
                                "sample": "cpu_nms (single-threaded as shipped) on N = %s, one pass each" % done}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="detection", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-soak", action="store_true", help="skip the clock soak and repeat blocks (runs under ncu)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--full-parity", action="store_true", help="nms: compare every sweep size with cpu_nms")
    ap.add_argument("--consume", action="store_true",
                    help="multi-GPU: read the gathered detections back on every step, inside the timed region")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from dspnet_b200 import _lib
    from dspnet_b200.plan import DetectionPlan, TargetPlan
    from dspnet_b200.symbol import multibox_anchors

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.workload == "nms":
        run_nms(args, torch, dist, dev, rank, world)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload in ("detection_heads", "train_tail"):
        run_extra(args, torch, dist, dev, rank, world)
        if world > 1:
            dist.destroy_process_group()
        return

    w = WORKLOADS[args.workload]
    ops = w["ops"]
    Bg = w["batch"] if w["scaling"] == "weak" else max(1, w["batch"] // world)  # images per GPU
    inputs, _ = make_inputs(rank * Bg, Bg, args.workload)
    A, C, L = inputs["A"], inputs["C"], inputs["L"]
    anchors = multibox_anchors(w["preset"], device=dev)
    lib = _lib.lib()

    # ---- resident rotating sets (same values, distinct addresses): consecutive steps must not hit in L2 ----
    abytes = 0
    if "prior" in ops:
        abytes += prior_algorithmic_bytes(A)
    if "target" in ops:
        abytes += tgt_algorithmic_bytes(Bg, A, C, L)
    if "detection" in ops:
        abytes += det_algorithmic_bytes(Bg, A, C)
    rotate = 2
    while rotate * abytes <= 1.5 * L2_BYTES and rotate < 8:
        rotate += 1
    flush_mode = rotate * abytes <= L2_BYTES  # small configs: explicit L2 flush between event-timed steps
    flush_buf = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev) if flush_mode else None
    dplan = tplan = None
    if "detection" in ops:
        dplan = DetectionPlan(Bg, A, C, dev, **DET_PARAMS)
        prob_sets = [torch.from_numpy(inputs["prob"]).to(dev) for _ in range(rotate)]
        loc_sets = [torch.from_numpy(inputs["loc"]).to(dev) for _ in range(rotate)]
        out_sets = [dplan.new_output() for _ in range(rotate)]
        # multi-GPU: the gather kernel only scans the rows below the operator's valid count (one per resident set: it
        # reads them beside the next step); without it every step re-read the whole 22 MB output on the side stream
        valid_sets = [dplan.new_valid_count() if world > 1 else None for _ in range(rotate)]
    if "target" in ops:
        tplan = TargetPlan(Bg, A, L, C, dev, **TGT_PARAMS)
        lab_d = torch.from_numpy(inputs["lab"]).to(dev)
        logit_sets = [torch.from_numpy(inputs["logits"]).to(dev) for _ in range(rotate)]
        tout_sets = [tplan.new_outputs() for _ in range(rotate)]
        stat_sets = [tplan.new_stats() for _ in range(rotate)]
    gatherer = None
    if world > 1:
        # the one exchange step of the path: surviving detections and / or per-image target statistics,
        # stored straight into every peer's buffer over NVLink
        from dspnet_b200.dist import P2PDetectionGatherer
        gatherer = P2PDetectionGatherer(Bg, A, DET_PARAMS["nms_topk"] if dplan is not None else 0, dev, world, rank,
                                        stats_width=4 if tplan is not None else 0)
    last_anchors = [anchors]

    def step(i, consume=False):
        s = i % rotate
        an = anchors
        if "prior" in ops:
            an = multibox_anchors(w["preset"], device=dev)
            last_anchors[0] = an
        if tplan is not None:
            tplan.run(an, lab_d, logit_sets[s], tout_sets[s], stats=stat_sets[s])
        if dplan is not None:
            dplan.run(prob_sets[s], loc_sets[s], an, out_sets[s], valid=valid_sets[s])
        if gatherer is not None:
            gatherer.submit(out_sets[s] if dplan is not None else None, i,
                            valid_count=valid_sets[s] if dplan is not None else None,
                            stats=stat_sets[s] if tplan is not None else None)
            if consume:
                gatherer.gathered(i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    # clock soak: a FIXED number of untimed steps (about 1 s), identical on every rank (the exchange counts steps)
    soak_steps = 0 if args.no_soak else (8000 if not flush_mode else 400)
    i = args.warmup
    for _ in range(soak_steps // 50):
        for _ in range(50):
            step(i)
            i += 1
        torch.cuda.synchronize()
    if gatherer is not None:
        gatherer.drain()

    # ---- timed region: exactly K steps between barriers, device-timed, max over ranks; repeated `blocks` times ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    blocks = 1 if args.no_soak else 15
    block_ms = []
    for _ in range(blocks):
        if flush_mode:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            barrier()
            for k in range(args.steps):
                flush_buf.fill_(k & 0xff)  # 252 MB written: nothing of the previous step is left in L2
                evs[k][0].record()
                step(i + k, consume=args.consume)
                evs[k][1].record()
            if gatherer is not None:
                gatherer.drain()
            barrier()
            ms = sum(a.elapsed_time(b) for a, b in evs)
        else:
            barrier()
            ev0.record()
            for k in range(args.steps):
                step(i + k, consume=args.consume)
            if gatherer is not None:
                gatherer.drain()
            ev1.record()
            barrier()
            ms = ev0.elapsed_time(ev1)
        i += args.steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        block_ms.append(ms)
    ms = sorted(block_ms)[len(block_ms) // 2]
    launches_per_step = sum(p.launches_per_run or 0 for p in (dplan, tplan) if p is not None)
    launches_per_step += 1 if "prior" in ops else 0
    launches_per_step += gatherer.launches_per_submit if gatherer else 0
    last_step = i - 1
    s_last = last_step % rotate

    # ---- parity of a timed output: the last step's results against the reference CPU operator ----
    M, kind = cpu_backend()
    ref_anchors = oracle_anchors(w["preset"])
    parity = {"checked": []}
    ok = True
    if "prior" in ops:
        ok = bool(np.array_equal(last_anchors[0].cpu().numpy(), ref_anchors))
        parity["checked"].append("anchors")
    want = cpu_pass(args.workload, inputs, ref_anchors, os.cpu_count() or 1, collect=True)
    if dplan is not None:
        got = out_sets[s_last].cpu().numpy()
        exp = np.concatenate([p["detection"] for p in want], axis=0)
        ok = ok and bool(np.array_equal(got.view(np.uint32), exp.view(np.uint32)))
        parity["checked"].append("detection (B,A,7) bit-exact")
    if tplan is not None:
        tplan.status()
        for k in range(3):
            got = tout_sets[s_last][k].cpu().numpy()
            exp = np.concatenate([p["target"][k] for p in want], axis=0).reshape(got.shape)
            same = (got.view(np.uint32) == exp.view(np.uint32)) | ((got == 0) & (exp == 0))
            ok = ok and bool(same.all())
        parity["checked"].append("loc_target / loc_mask / cls_target bit-exact")
    gather_ok = True
    if gatherer is not None:
        # sanity of the exchange step: every rank holds every rank's compacted detections / target statistics
        gok = True
        if dplan is not None:
            from dspnet_b200.dist import compact_rows
            rows, counts = gatherer.gathered(last_step)
            mine, mine_n = compact_rows(out_sets[s_last], DET_PARAMS["nms_topk"])
            gok = torch.equal(rows[rank * Bg:(rank + 1) * Bg], mine) and torch.equal(counts[rank * Bg:(rank + 1) * Bg], mine_n)
            gok = gok and bool((counts > 0).all())
        if tplan is not None:
            st = gatherer.gathered_stats(last_step)
            gok = gok and torch.equal(st[rank * Bg:(rank + 1) * Bg], stat_sets[s_last]) and bool((st[:, 0] >= 0).all())
        gok = gok and gatherer.check()
        gather_ok = bool(gok)
    if world > 1:
        flag = torch.tensor([1 if ok else 0, 1 if gather_ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok, gather_ok = int(flag[0].item()) == 1, int(flag[1].item()) == 1
    parity["result"] = "ok" if ok else "MISMATCH"
    parity["against"] = ("oracle/_ref (the reference's own operator .cc compiled in place)" if kind == "reference"
                         else "oracle port (bit-identical restatement)")
    parity["what"] = "outputs of the last timed step, every rank's shard"

    # ---- per-kernel durations: after the timed region each launch of the step is timed on its own -- K back-to-back
    #      launches of ONE phase (DSPMB_TUNE_PHASES) between a single cudaEvent pair on the launching stream, so the
    #      average carries no per-launch event overhead.  The other phases' inputs are still in the workspace.
    def time_phases(run, names):
        res = {}
        # plain launches here: a single-kernel CUDA graph would add its own launch overhead to every sample
        cache_was = lib.dspmb_set_tuning(_lib.TUNE_GRAPH_CACHE, 0)
        for bit, name in names:
            lib.dspmb_set_tuning(_lib.TUNE_PHASES, bit)
            for q in range(3):
                run(q)
            torch.cuda.synchronize()
            reps = max(args.steps, 200)  # a 20-step driver run would put the ramp of the first launch on every sample
            ev0.record()
            for q in range(reps):
                run(q)
            ev1.record()
            torch.cuda.synchronize()
            res[name] = ev0.elapsed_time(ev1) / reps
        lib.dspmb_set_tuning(_lib.TUNE_PHASES, _lib.PHASES_ALL)
        lib.dspmb_set_tuning(_lib.TUNE_GRAPH_CACHE, cache_was)
        return res

    kernels = {}
    if dplan is not None:
        def det_run(q):
            s = q % rotate
            dplan.run(prob_sets[s], loc_sets[s], anchors, out_sets[s])
        for s in range(rotate):  # every workspace-dependent phase input exists for every rotating set
            det_run(s)
        pipeline = lib.dspmb_set_tuning(_lib.TUNE_DET_PIPELINE, 1)
        lib.dspmb_set_tuning(_lib.TUNE_DET_PIPELINE, pipeline)
        if pipeline and C in (21, 9) and not DET_PARAMS["force_suppress"]:
            names = [(1, "det_stream_kernel"), (2, "det_sort_kernel"), (4, "det_pair_kernel")]
        else:
            names = [(1, "det_stream_kernel"), (2, "det_sort_kernel"), (4, "det_nms_kernel")]
        kernels.update(time_phases(det_run, names))
        for s in range(rotate):
            det_run(s)
    if tplan is not None:
        def tgt_run(q):
            s = q % rotate
            tplan.run(anchors, lab_d, logit_sets[s], tout_sets[s], stats=stat_sets[s])
        for s in range(rotate):
            tgt_run(s)
        names = [(1, "target_stream_kernel"), (2, "target_match_kernel")]
        if (tplan.launches_per_run or 0) >= 4:
            names.append((4, "target_select_kernel"))
        kernels.update(time_phases(tgt_run, names))
        tplan.status()

    # ---- end to end through the public operators with HOST buffers (pinned in, results read back) ----
    e2e = None
    if not args.no_e2e:
        from dspnet_b200 import MultiBoxDetection, MultiBoxTarget
        host_in, order = [], []
        if "target" in ops:
            host_in += [inputs["lab"], inputs["logits"]]
            order += ["lab", "logits"]
        if "detection" in ops:
            host_in += [inputs["prob"], inputs["loc"]]
            order += ["prob", "loc"]

        def e2e_run(*d):
            t = dict(zip(order, d))
            an = multibox_anchors(w["preset"], device=dev) if "prior" in ops else anchors
            outs = []
            if "target" in ops:
                outs += list(MultiBoxTarget(an, t["lab"], t["logits"], **TGT_PARAMS))
            if "detection" in ops:
                outs.append(MultiBoxDetection(t["prob"], t["loc"], an, **DET_PARAMS))
            return outs
        pipe = E2EPipeline(torch, dev, host_in, e2e_run)
        e2e_steps = max(100, args.steps)  # the 3-stream pipeline needs a few dozen steps to show its steady state
        for _ in range(4):
            pipe.step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pipe.step()
        torch.cuda.synchronize()
        e2e_ms = 1e3 * (time.perf_counter() - t0)  # wall clock around fully synchronised work on three streams
        barrier()
        if world > 1:
            t = torch.tensor([e2e_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        names = " + ".join(("dspnet_b200.symbol.multibox_anchors" if o == "prior" else
                            "dspnet_b200.MultiBoxTarget" if o == "target" else "dspnet_b200.MultiBoxDetection") for o in ops)
        e2e = {"value": Bg * world * e2e_steps / (e2e_ms * 1e-3), "unit": "images/s",
               "h2d_bytes_per_step": int(pipe.h2d), "d2h_bytes_per_step": int(pipe.d2h), "steps": e2e_steps,
               "api": names + "; per step: pinned host -> device copy of every input tensor, the operators, every output "
                      "tensor copied back to pinned host; copy-in / operators / copy-out of consecutive steps overlap on "
                      "three streams"}
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peak, peak_src = measured_peaks()
        # the roofline kernel is the one that streams the operator's tensors; the dominant kernel (by time) is named too
        if dplan is not None and (tplan is None or kernels.get("det_stream_kernel", 0) >= kernels.get("target_stream_kernel", 0)):
            rk, rbytes = "det_stream_kernel", det_algorithmic_bytes(Bg, A, C)
        else:
            rk, rbytes = "target_stream_kernel", tgt_algorithmic_bytes(Bg, A, C, L)
        k_ms = kernels.get(rk, float("nan"))
        achieved = rbytes / (k_ms * 1e-3) / 1e9
        dominant = max(kernels, key=kernels.get)
        step_ms = ms / args.steps
        if world == 1:
            par = "one GPU, %d images" % Bg
        else:
            what = " + ".join(x for x, on in (("surviving detections", dplan is not None),
                                              ("per-image target statistics", tplan is not None)) if on)
            par = "images sharded, %d per GPU; exchange: %s stored into every peer over NVLink" % (Bg, what)
        line = {
            "metric": w["metric"], "value": Bg * world * args.steps / (ms * 1e-3),
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.workload, inputs, Bg),
                           l2=("L2 flushed (252 MB written) between steps, every step timed by its own event pair"
                               if flush_mode else "inputs/outputs rotate over %d resident sets (%d MB total) > 126 MB L2"
                               % (rotate, rotate * abytes // 2 ** 20)),
                           parallelism=par),
            "soak_steps": soak_steps, "timed_blocks": len(block_ms),
            "block_ms": [round(x, 4) for x in block_ms],
            "timing": "value = median of %d timed blocks of exactly %d steps each (barrier + synchronize on both sides, "
                      "CUDA events, max over ranks) after %d warm-up and %d soak steps" % (len(block_ms), args.steps,
                                                                                          args.warmup, soak_steps),
            "roofline": {"bound": "hbm", "kernel": rk, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": kernel_traffic(rk), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": rbytes, "kernel_ms": k_ms,
                         "timing": "max(K, 200) back-to-back launches of the kernel alone between one cudaEvent pair, right after the timed region",
                         "all_kernels_ms": kernels, "dominant_kernel": dominant,
                         "dominant_kernel_ms": kernels[dominant],
                         "dominant_kernel_bound": "hbm" if dominant == rk else "instruction issue / latency (moves a few MB)",
                         "whole_op_frac": abytes / (step_ms * 1e-3) / 1e9 / peak,
                         "whole_op_algorithmic_bytes": abytes},
            "e2e": e2e,
            "gpu_launches": args.steps * launches_per_step,
            "launches_per_step": launches_per_step,
            "parity_check": parity,
            "gather_check": ("ok" if gather_ok else "MISMATCH") if world > 1 else None,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            v, done, dt = cpu_time(args.workload, inputs, ref_anchors, threads, steps=20, warmup=1, budget_s=20.0)
            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": threads, "kind": kind,
                                    "sample": "%d steps x one B=%d %s batch, image slices over %d host threads (the method "
                                              "of --impl reference)" % (done, Bg, w["preset"], threads)}
        print(json.dumps(line))
    if gatherer is not None:
        gatherer.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
