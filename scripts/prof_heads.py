import sys, torch
sys.path.insert(0, '.')
import bench
from dspnet_b200 import presets, synth
from dspnet_b200.plan import DetectionHeadsPlan
from dspnet_b200.symbol import multibox_anchors
dev = torch.device('cuda', 0)
B = 32
p = presets.PRESETS['ssd512']
A, C = presets.num_anchors(p), p.num_classes
ch, lh = synth.heads_from_logits(p, synth.det_logits(2, B, C, A), synth.loc_pred(2, B, A))
anchors = multibox_anchors('ssd512', device=dev)
shapes = [(fm.height, fm.width, len(fm.sizes) + len(fm.ratios) - 1) for fm in p.maps]
plan = DetectionHeadsPlan(B, A, C, shapes, dev, **bench.DET_PARAMS)
c = [torch.from_numpy(h).to(dev) for h in ch]
l = [torch.from_numpy(h).to(dev) for h in lh]
bind = plan.bind(c, l)
out = plan.new_output()
for _ in range(2):
    plan.run(bind, anchors, out)
torch.cuda.synchronize()
