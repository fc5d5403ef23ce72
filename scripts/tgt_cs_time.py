"""Device-timed MultiBoxTarget on the DSPNet-Cityscapes head (B=16, L=200, one image with 200 gts): knobs as k=v args."""
import sys, torch
sys.path.insert(0, '.')
import bench
from dspnet_b200 import _lib
from dspnet_b200.plan import TargetPlan
from dspnet_b200.symbol import multibox_anchors
dev = torch.device('cuda', 0)
L_ = _lib.lib()
for kv in sys.argv[1:]:
    k, v = kv.split('=')
    L_.dspmb_set_tuning(int(k), int(v))
B = 16
tin, _ = bench.make_inputs(0, B, 'dspnet_cs')
A, C, L = tin['A'], tin['C'], tin['L']
anchors = multibox_anchors('dspnet_cs', device=dev)
plan = TargetPlan(B, A, L, C, dev, **bench.TGT_PARAMS)
lab = torch.from_numpy(tin['lab']).to(dev)
logits = [torch.from_numpy(tin['logits']).to(dev) for _ in range(4)]
outs = [plan.new_outputs() for _ in range(4)]
for i in range(500):
    plan.run(anchors, lab, logits[i % 4], outs[i % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = []
for rep in range(5):
    e0.record()
    for i in range(100):
        plan.run(anchors, lab, logits[i % 4], outs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 100 * 1e3)
plan.status()
print('dspnet_cs target B=%d step us: median %.2f %s' % (B, sorted(res)[2], sys.argv[1:]))
import ctypes
st = torch.zeros(B * 12, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
L_.dspmb_debug_target_stamps.argtypes = [ctypes.c_void_p]
L_.dspmb_debug_target_stamps(st.data_ptr())
plan.run(anchors, lab, logits[0], outs[0])
torch.cuda.synchronize()
L_.dspmb_debug_target_stamps(None)
t = st.cpu().view(B, 12).double() / 1e3
names = ['colbest', 'bipartite', 'fixup', 'stage keys', 'pivot', 'final pass', 'ambiguous']
for img in range(B):
    if t[img, 7] > 0:
        print('image %2d G=%3d ' % (img, int((tin['lab'][img, :, 0] != -1).sum())) + ' '.join('%s %.1f' % (n, (t[img, k + 1] - t[img, k]).item()) for k, n in enumerate(names)) + '  total %.1f' % (t[img, 7] - t[img, 0]).item())
