#!/usr/bin/env python
"""Benchmark of the multibox hot path (BASELINE.json: "SSD-512 multibox target+detect images/s at 1/2/4/8 B200;
% of HBM peak").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

A step is one pass of MultiBoxDetection (decode + threshold + top-k sort + per-class NMS) over one batch of 32
synthetic SSD-512 VOC head tensors per GPU (BASELINE.json configs[1]); with N > 1 every rank processes its own 32
images (weak scaling, images are independent) and the compacted detections are all-gathered over NCCL, overlapped
with the next step's kernels.  Inputs rotate over several resident copies whose total footprint exceeds the 126 MB
L2, so every step streams from HBM.  One JSON line is printed by rank 0 (see the task contract for the keys).
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

PRESET = "ssd512"
BATCH = 32
CONFIG_ID = 2
DET_PARAMS = dict(threshold=0.01, clip=True, nms_threshold=0.45, force_suppress=False, nms_topk=400,
                  variances=(0.1, 0.1, 0.2, 0.2))
TGT_PARAMS = dict(overlap_threshold=0.5, ignore_label=-1.0, negative_mining_ratio=3.0, negative_mining_thresh=0.5,
                  minimum_negative_samples=0, variances=(0.1, 0.1, 0.2, 0.2))
ROTATE = 4  # resident input/output sets cycled through: 4 x ~104 MB > 126 MB L2


def det_algorithmic_bytes(B, A, C):
    """4*B*C*A + 4*B*A*5 + 16*A read, 28*B*A written."""
    return 4 * B * C * A + 20 * B * A + 16 * A + 28 * B * A


def tgt_algorithmic_bytes(B, A, C, L):
    return 16 * A + 24 * B * L + 4 * B * C * A + 44 * B * A


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


def make_inputs(first_image, B):
    """Seeded synthetic SSD-512 head tensors (numpy, host)."""
    import numpy as np
    from dspnet_b200 import presets, synth
    p = presets.PRESETS[PRESET]
    A = presets.num_anchors(p)
    prob = synth.cls_prob(CONFIG_ID, B, p.num_classes, A, first_image=first_image)
    loc = synth.loc_pred(CONFIG_ID, B, A, first_image=first_image)
    lab = synth.labels(CONFIG_ID, B, p.label_slots, p.num_classes, first_image=first_image)
    logits = synth.cls_preds(CONFIG_ID, B, p.num_classes, A, first_image=first_image)
    return dict(prob=prob, loc=loc, lab=lab, logits=logits, A=A, C=p.num_classes, L=p.label_slots), np


def oracle_anchors():
    import numpy as np
    from dspnet_b200 import presets
    from oracle import oracle as O
    p = presets.PRESETS[PRESET]
    return np.concatenate([O.multibox_prior(fm.height, fm.width, fm.sizes, fm.ratios, False, (fm.step, fm.step))
                           for fm in p.maps], axis=1)


def cpu_detection(inputs, anchors, threads):
    """One pass of the reference CPU operator over the batch, one image slice per host thread.  Uses oracle/_ref
    (the reference's own multibox_detection.cc compiled in place) when it is present, else the oracle port; the two
    are bit-identical (tests/test_oracle_golden.py).  Returns the kind that ran."""
    from oracle import ref as R
    B = inputs["prob"].shape[0]
    if R.available():
        from concurrent.futures import ThreadPoolExecutor
        n = max(1, min(threads, B))
        bounds = [(i * B // n, (i + 1) * B // n) for i in range(n)]

        def run(be):
            b, e = be
            return R.multibox_detection(inputs["prob"][b:e], inputs["loc"][b:e], anchors, **DET_PARAMS)
        if n == 1:
            run(bounds[0])
        else:
            with ThreadPoolExecutor(n) as ex:
                list(ex.map(run, bounds))
        return "reference"
    from oracle import oracle as O
    O.multibox_detection(inputs["prob"], inputs["loc"], anchors, nthreads=threads, **DET_PARAMS)
    return "port"


def cpu_baseline(inputs, threads, repeats=3):
    """images/s of the CPU operator on one B=32 detection batch (median of `repeats` after one warm-up)."""
    anchors = oracle_anchors()
    times, kind = [], "port"
    for _ in range(repeats + 1):
        t0 = time.perf_counter()
        kind = cpu_detection(inputs, anchors, threads)
        times.append(time.perf_counter() - t0)
    times = sorted(times[1:])
    return BATCH / times[len(times) // 2], kind


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host threads; rank 0 only.
    (MXNet itself cannot be installed here -- DESIGN.md section 9 -- so the operator body is driven directly.)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    inputs, _ = make_inputs(0, BATCH)
    anchors = oracle_anchors()
    kind = "port"
    for _ in range(max(args.warmup, 1)):
        kind = cpu_detection(inputs, anchors, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_detection(inputs, anchors, threads)
    dt = time.perf_counter() - t0
    value = BATCH * args.steps / dt
    line = {
        "impl": "reference", "metric": "ssd512_multibox_detection_images_per_s", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ssd512_voc21_multibox_detection_nms_batch32", "preset": PRESET, "batch_per_gpu": BATCH,
                   "anchors": inputs["A"], "classes": inputs["C"], **{k: v for k, v in DET_PARAMS.items()}},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": kind,
                         "sample": "%d steps x one B=32 SSD-512 detection batch, image slices over %d host threads"
                                   % (args.steps, threads)},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-soak", action="store_true", help="skip the 1.5 s clock soak (for runs under ncu)")
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

    inputs, _ = make_inputs(rank * BATCH, BATCH)
    A, C, L = inputs["A"], inputs["C"], inputs["L"]
    anchors = multibox_anchors(PRESET, device=dev)
    plan = DetectionPlan(BATCH, A, C, dev, **DET_PARAMS)
    # resident rotating sets (same values, distinct addresses) so that consecutive steps do not hit in L2
    prob_sets = [torch.from_numpy(inputs["prob"]).to(dev) for _ in range(ROTATE)]
    loc_sets = [torch.from_numpy(inputs["loc"]).to(dev) for _ in range(ROTATE)]
    out_sets = [plan.new_output() for _ in range(ROTATE)]
    gatherer = None
    if world > 1:
        from dspnet_b200.dist import P2PDetectionGatherer
        gatherer = P2PDetectionGatherer(BATCH, A, DET_PARAMS["nms_topk"], dev, world, rank)

    def step(i):
        s = i % ROTATE
        plan.run(prob_sets[s], loc_sets[s], anchors, out_sets[s])
        if gatherer is not None:
            gatherer.submit(out_sets[s], i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # warm-up: W steps as asked, then keep the GPU busy for ~1.5 s so clocks and the sampler reach steady state
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    # the soak is a FIXED number of steps (about 1.5 s), identical on every rank: the exchange step counts arrivals
    soak_steps = 0 if args.no_soak else 12000
    i = args.warmup
    for _ in range(soak_steps // 50):
        for _ in range(50):
            step(i)
            i += 1
        torch.cuda.synchronize()
    if gatherer is not None:
        gatherer.drain()

    # ---- timed region: exactly K steps, device-timed, max over ranks ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for k in range(args.steps):
        step(i + k)
    if gatherer is not None:
        gatherer.drain()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = args.steps * (plan.launches_per_run + (gatherer.launches_per_submit if gatherer else 0))
    gather_ok = True
    if world > 1:
        # sanity of the exchange step: every rank holds every rank's compacted detections
        last_step = i + args.steps - 1
        rows, counts = gatherer.gathered(last_step)
        from dspnet_b200.dist import compact_rows
        mine, mine_n = compact_rows(out_sets[last_step % ROTATE], DET_PARAMS["nms_topk"])
        ok = torch.equal(rows[rank * BATCH:(rank + 1) * BATCH], mine) and torch.equal(counts[rank * BATCH:(rank + 1) * BATCH], mine_n)
        ok = ok and bool((counts > 0).all())
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_ok = int(flag.item()) == 1

    # ---- per-kernel durations: after the timed region each launch of the step is timed on its own -- K back-to-back
    #      launches of ONE phase (DSPMB_TUNE_PHASES) between a single cudaEvent pair on the launching stream, so the
    #      average carries no per-launch event overhead.  The other phases' inputs are still in the workspace.
    lib = _lib.lib()
    import ctypes

    def time_phases(run, names):
        res = {}
        # plain launches here: a single-kernel CUDA graph would add its own launch overhead to every sample
        cache_was = lib.dspmb_set_tuning(_lib.TUNE_GRAPH_CACHE, 0)
        for bit, name in names:
            lib.dspmb_set_tuning(_lib.TUNE_PHASES, bit)
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            ev0.record()
            for i in range(args.steps):
                run(i)
            ev1.record()
            torch.cuda.synchronize()
            res[name] = ev0.elapsed_time(ev1) / args.steps
        lib.dspmb_set_tuning(_lib.TUNE_PHASES, _lib.PHASES_ALL)
        lib.dspmb_set_tuning(_lib.TUNE_GRAPH_CACHE, cache_was)
        return res

    def det_run(i):
        s = i % ROTATE
        plan.run(prob_sets[s], loc_sets[s], anchors, out_sets[s])
    for s in range(ROTATE):  # every workspace-dependent phase input exists for every rotating set
        det_run(s)
    kernels = time_phases(det_run, [(1, "det_stream_kernel"), (2, "det_sort_kernel"), (4, "det_pair_kernel"),
                                    (8, "det_resolve_kernel")])
    nslots = 16

    # ---- end to end through the public operator with HOST buffers (pinned in, result read back) ----
    from dspnet_b200 import MultiBoxDetection
    pin_prob = torch.from_numpy(inputs["prob"]).pin_memory()
    pin_loc = torch.from_numpy(inputs["loc"]).pin_memory()
    pin_out = torch.empty((BATCH, A, 7), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    # Three streams, double-buffered device tensors: the H2D copy of step i+1, the operator of step i and the D2H
    # copy of step i-1 overlap (PCIe is full duplex); every step still moves its own inputs in and its result out.
    s_in, s_run, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    d_prob = [torch.empty_like(prob_sets[0]) for _ in range(2)]
    d_loc = [torch.empty_like(loc_sets[0]) for _ in range(2)]
    d_out = [None, None]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_run = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    state = {"n": 0}

    def e2e_step():
        i = state["n"]
        k = i & 1
        with torch.cuda.stream(s_in):
            if i >= 2:
                s_in.wait_event(ev_run[k])      # the operator that read this input buffer two steps ago is done
            d_prob[k].copy_(pin_prob, non_blocking=True)
            d_loc[k].copy_(pin_loc, non_blocking=True)
            ev_in[k].record(s_in)
        with torch.cuda.stream(s_run):
            s_run.wait_event(ev_in[k])
            if i >= 2:
                s_run.wait_event(ev_out[k])     # the previous result in this slot has been copied out
            d_out[k] = MultiBoxDetection(d_prob[k], d_loc[k], anchors, **DET_PARAMS)
            ev_run[k].record(s_run)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_run[k])
            pin_out.copy_(d_out[k], non_blocking=True)
            ev_out[k].record(s_out)
        state["n"] = i + 1

    for _ in range(4):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)  # wall clock around fully synchronised work on three streams
    barrier()
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    clocks = sampler.stop() if sampler else None

    # ---- secondary figure: MultiBoxTarget with hard-negative mining (BASELINE.json configs[2]) ----
    tgt = None
    try:
        TB = 64 // world if world > 1 else 64
        tin, _ = make_inputs(rank * TB, TB)
        tplan = TargetPlan(TB, A, L, C, dev, **TGT_PARAMS)
        lab_d = torch.from_numpy(tin["lab"]).to(dev)
        logit_sets = [torch.from_numpy(tin["logits"]).to(dev) for _ in range(2)]
        touts = [tplan.new_outputs() for _ in range(2)]
        for i in range(5):
            tplan.run(anchors, lab_d, logit_sets[i % 2], touts[i % 2])
        barrier()
        ev0.record()
        tsteps = max(10, min(args.steps, 100))
        for i in range(tsteps):
            tplan.run(anchors, lab_d, logit_sets[i % 2], touts[i % 2])
        ev1.record()
        barrier()
        tms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([tms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tms = float(t.item())
        tplan.status()
        def tgt_run(i):
            tplan.run(anchors, lab_d, logit_sets[i % 2], touts[i % 2])
        tk = time_phases(tgt_run, [(1, "target_stream_kernel"), (2, "target_match_kernel")])
        peak, _ = measured_peaks()
        tbytes = tgt_algorithmic_bytes(TB, A, C, L)
        tgt = {"workload": "ssd512_multibox_target_mining3_batch64_%s" % ("sharded" if world > 1 else "1gpu"),
               "images_per_s": TB * world * tsteps / (tms * 1e-3), "ms_per_step": tms / tsteps,
               "kernel_ms": tk,
               "stream_kernel_gbs": tbytes / (tk.get("target_stream_kernel", float("nan")) * 1e-3) / 1e9,
               "stream_kernel_frac_of_hbm": tbytes / (tk.get("target_stream_kernel", float("nan")) * 1e-3) / 1e9 / peak,
               "scaling": "strong"}
    except Exception as e:  # the headline must still print
        tgt = {"error": repr(e)}

    if rank == 0:
        peak, peak_src = measured_peaks()
        abytes = det_algorithmic_bytes(BATCH, A, C)
        k_ms = kernels.get("det_stream_kernel", float("nan"))
        achieved = abytes / (k_ms * 1e-3) / 1e9
        line = {
            "metric": "ssd512_multibox_detection_images_per_s", "value": BATCH * world * args.steps / (ms * 1e-3),
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ssd512_voc21_multibox_detection_nms_batch32", "preset": PRESET,
                       "batch_per_gpu": BATCH, "anchors": A, "classes": C,
                       "l2": "inputs/outputs rotate over %d resident sets (%d MB total) > 126 MB L2" % (
                           ROTATE, ROTATE * abytes // 2 ** 20),
                       "parallelism": "images sharded, %d per GPU" % BATCH, **DET_PARAMS},
            "roofline": {"bound": "hbm", "kernel": "det_stream_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": abytes, "kernel_ms": k_ms,
                         "timing": "K back-to-back launches of the kernel alone between one cudaEvent pair, right after the timed region",
                         "all_kernels_ms": kernels,
                         "whole_op_frac": abytes / (ms / args.steps * 1e-3) / 1e9 / peak},
            "e2e": {"value": BATCH * world * e2e_steps / (e2e_ms * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": int(pin_prob.numel() * 4 + pin_loc.numel() * 4),
                    "d2h_bytes_per_step": int(pin_out.numel() * 4), "steps": e2e_steps,
                    "api": "dspnet_b200.MultiBoxDetection; per step: pinned host -> device copy of cls_prob+loc_pred, operator, full (B,A,7) result copied back to pinned host; copy-in / operator / copy-out of consecutive steps overlap on three streams"},
            "gpu_launches": launches,
            "gather_check": ("ok" if gather_ok else "MISMATCH") if world > 1 else None,
            "clocks": clocks,
            "target": tgt,
        }
        traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_path):
            with open(traffic_path) as f:
                line["roofline"]["traffic"] = json.load(f).get("det_stream_kernel")
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            v, kind = cpu_baseline(inputs, threads)
            v1, _ = cpu_baseline(inputs, 1, repeats=1)
            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": threads, "kind": kind,
                                    "sample": "one B=32 SSD-512 detection batch, median of 3 after 1 warm-up, one image "
                                              "per thread; single thread: %.1f images/s" % v1}
        print(json.dumps(line))
    if world > 1:
        gatherer.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
