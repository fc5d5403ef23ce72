"""Device-timed detection step (graph replay, rotating sets) on the bench batch: quick A/B measurements."""
import sys, torch
sys.path.insert(0, '.')
import bench
from dspnet_b200.plan import DetectionPlan
from dspnet_b200.symbol import multibox_anchors
dev = torch.device('cuda', 0)
from dspnet_b200 import _lib
for kv in sys.argv[1:]:
    k, v = kv.split('=')
    _lib.lib().dspmb_set_tuning(int(k), int(v))
inputs, _ = bench.make_inputs(0, bench.BATCH)
A, C = inputs['A'], inputs['C']
anchors = multibox_anchors(bench.PRESET, device=dev)
plan = DetectionPlan(bench.BATCH, A, C, dev, **bench.DET_PARAMS)
prob = [torch.from_numpy(inputs['prob']).to(dev) for _ in range(4)]
loc = [torch.from_numpy(inputs['loc']).to(dev) for _ in range(4)]
out = [plan.new_output() for _ in range(4)]
for i in range(6000):
    plan.run(prob[i % 4], loc[i % 4], anchors, out[i % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = []
for rep in range(7):
    e0.record()
    for i in range(200):
        plan.run(prob[i % 4], loc[i % 4], anchors, out[i % 4])
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 200 * 1e3)
print('step us: median %.2f  all %s  launches %s' % (sorted(res)[3], ['%.1f' % r for r in res], plan.launches_per_run))
