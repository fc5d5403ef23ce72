"""Device-resident step time of the three operators at every BASELINE.json config (one JSON line per config)."""
import json
import sys

sys.path.insert(0, '.')
import torch

from dspnet_b200 import presets, synth
from dspnet_b200.plan import DetectionPlan, TargetPlan
from dspnet_b200.symbol import multibox_anchors

dev = torch.device('cuda', 0)
CASES = [("ssd300", 1, 8), ("ssd512", 32, 8), ("ssd512", 64, 8), ("dspnet_cs", 16, 50), ("dspnet_cs", 2, 50)]
for preset, B, max_gt in CASES:
    p = presets.PRESETS[preset]
    A, C, L = presets.num_anchors(p), p.num_classes, p.label_slots
    anchors = multibox_anchors(preset, device=dev)
    prob = torch.from_numpy(synth.cls_prob(3, B, C, A)).to(dev)
    loc = torch.from_numpy(synth.loc_pred(3, B, A)).to(dev)
    lab = torch.from_numpy(synth.labels(3, B, L, C, max_gt=max_gt)).to(dev)
    logits = torch.from_numpy(synth.cls_preds(3, B, C, A)).to(dev)
    dplan = DetectionPlan(B, A, C, dev, threshold=0.01, clip=True, nms_threshold=0.45, nms_topk=400)
    tplan = TargetPlan(B, A, L, C, dev, negative_mining_ratio=3.0, negative_mining_thresh=0.5)
    out, touts = dplan.new_output(), tplan.new_outputs()
    res = {"preset": preset, "batch": B, "anchors": A, "classes": C, "label_slots": L}
    for name, fn in (("detection", lambda: dplan.run(prob, loc, anchors, out)), ("target", lambda: tplan.run(anchors, lab, logits, touts))):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 200
        res[name + "_ms"] = round(ms, 5)
        res[name + "_images_per_s"] = round(B / ms * 1e3)
    tplan.status()
    print(json.dumps(res), flush=True)
