"""Runs each op a few times on the bench workload (for ncu captures)."""
import sys, torch
sys.path.insert(0, '.')
import bench
from dspnet_b200.plan import DetectionPlan, TargetPlan
from dspnet_b200.symbol import multibox_anchors
dev = torch.device('cuda', 0)
which = sys.argv[1] if len(sys.argv) > 1 else 'both'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if which in ('det', 'both'):
    inputs, _ = bench.make_inputs(0, bench.BATCH)
    A, C, L = inputs['A'], inputs['C'], inputs['L']
    anchors = multibox_anchors(bench.PRESET, device=dev)
    plan = DetectionPlan(bench.BATCH, A, C, dev, **bench.DET_PARAMS)
    prob = torch.from_numpy(inputs['prob']).to(dev); loc = torch.from_numpy(inputs['loc']).to(dev); out = plan.new_output()
    for _ in range(reps):
        plan.run(prob, loc, anchors, out)
    torch.cuda.synchronize()
if which in ('tgt', 'both'):
    tin, _ = bench.make_inputs(0, 64)
    A, C, L = tin['A'], tin['C'], tin['L']
    anchors = multibox_anchors(bench.PRESET, device=dev)
    tplan = TargetPlan(64, A, L, C, dev, **bench.TGT_PARAMS)
    lab = torch.from_numpy(tin['lab']).to(dev); logits = torch.from_numpy(tin['logits']).to(dev); outs = tplan.new_outputs()
    for _ in range(reps):
        tplan.run(anchors, lab, logits, outs)
    torch.cuda.synchronize()
    tplan.status()
