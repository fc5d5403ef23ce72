"""Device-timed MultiBoxDetectionFromHeads step (graph replay, rotating sets) beside the tensor-fed operator."""
import sys, torch
sys.path.insert(0, '.')
import numpy as np
import bench
from dspnet_b200 import _lib, presets, synth
from dspnet_b200.plan import DetectionHeadsPlan, DetectionPlan
from dspnet_b200.symbol import multibox_anchors
dev = torch.device('cuda', 0)
for kv in sys.argv[1:]:
    k, v = kv.split('=')
    _lib.lib().dspmb_set_tuning(int(k), int(v))
B = 32
p = presets.PRESETS['ssd512']
A, C = presets.num_anchors(p), p.num_classes
logits = synth.det_logits(2, B, C, A)
loc = synth.loc_pred(2, B, A)
ch, lh = synth.heads_from_logits(p, logits, loc)
anchors = multibox_anchors('ssd512', device=dev)
shapes = [(fm.height, fm.width, len(fm.sizes) + len(fm.ratios) - 1) for fm in p.maps]
plan = DetectionHeadsPlan(B, A, C, shapes, dev, **bench.DET_PARAMS)
sets = []
for _ in range(4):
    c = [torch.from_numpy(h).to(dev) for h in ch]
    l = [torch.from_numpy(h).to(dev) for h in lh]
    sets.append((c, l, plan.bind(c, l), plan.new_output()))
for i in range(3000):
    s = sets[i % 4]
    plan.run(s[2], anchors, s[3])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = []
for rep in range(5):
    e0.record()
    for i in range(200):
        s = sets[i % 4]
        plan.run(s[2], anchors, s[3])
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 200 * 1e3)
print('heads step us: median %.2f  all %s  launches %s' % (sorted(res)[2], ['%.1f' % r for r in res], plan.launches_per_run))
