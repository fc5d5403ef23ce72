"""Device-timed target step on the bench batch (or a smaller one: argv[1] = images) with per-kernel times and the
per-image phase stamps of the match kernel (dspmb_debug_target_stamps)."""
import ctypes, sys, torch
sys.path.insert(0, '.')
import bench
from dspnet_b200 import _lib
from dspnet_b200.plan import TargetPlan
from dspnet_b200.symbol import multibox_anchors
dev = torch.device('cuda', 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L_ = _lib.lib()
for kv in sys.argv[2:]:
    k, v = kv.split('=')
    L_.dspmb_set_tuning(int(k), int(v))
tin, _ = bench.make_inputs(0, B, 'target')
A, C, L = tin['A'], tin['C'], tin['L']
anchors = multibox_anchors(bench.PRESET, device=dev)
plan = TargetPlan(B, A, L, C, dev, **bench.TGT_PARAMS)
lab = torch.from_numpy(tin['lab']).to(dev)
logits = [torch.from_numpy(tin['logits']).to(dev) for _ in range(4)]
outs = [plan.new_outputs() for _ in range(4)]
for i in range(3000):
    plan.run(anchors, lab, logits[i % 4], outs[i % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = []
for rep in range(5):
    e0.record()
    for i in range(200):
        plan.run(anchors, lab, logits[i % 4], outs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 200 * 1e3)
print('B=%d step us: median %.2f  all %s  launches %s' % (B, sorted(res)[2], ['%.1f' % r for r in res], plan.launches_per_run))
# per-kernel (direct launches, profile events)
L_.dspmb_profile_enable(1)
for i in range(50):
    plan.run(anchors, lab, logits[i % 4], outs[i % 4])
torch.cuda.synchronize()
ms = (ctypes.c_float * 32)(); ln = (ctypes.c_int * 32)()
n = L_.dspmb_profile_read(ms, ln, 32)
L_.dspmb_profile_enable(0)
L_.dspmb_profile_kernel_name.restype = ctypes.c_char_p
print({L_.dspmb_profile_kernel_name(i).decode(): round(ms[i] / ln[i] * 1e3, 2) for i in range(n) if ln[i]})
st = torch.zeros(B * 12, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
L_.dspmb_debug_target_stamps.argtypes = [ctypes.c_void_p]
L_.dspmb_debug_target_stamps(st.data_ptr())
plan.run(anchors, lab, logits[0], outs[0])
torch.cuda.synchronize()
L_.dspmb_debug_target_stamps(None)
t = st.cpu().view(B, 12).double() / 1e3
names = ['colbest', 'bipartite', 'fixup', 'stage keys', 'pivot', 'final pass', 'ambiguous']
ok = t[:, 7] > 0
print('images with mining:', int(ok.sum()), 'of', B)
tt = t[ok]
for k, nme in enumerate(names):
    d = tt[:, k + 1] - tt[:, k]
    print('  %-11s mean %6.2f  max %6.2f us' % (nme, d.mean().item(), d.max().item()))
if (tt[:, 8] > 0).any():   # shortlist sub-phases of 'pivot' (stamps 8, 9, 10)
    sub = tt[tt[:, 8] > 0]
    for a_, b_, nme in ((4, 8, 'sample'), (8, 9, 'compaction'), (9, 10, 'histogram'), (10, 5, 'pivot rank')):
        d = sub[:, b_] - sub[:, a_]
        print('    %-11s mean %6.2f  max %6.2f us' % (nme, d.mean().item(), d.max().item()))
d = tt[:, 7] - tt[:, 0]
print('  %-11s mean %6.2f  max %6.2f us (image %d)' % ('total', d.mean().item(), d.max().item(), int(d.argmax())))
print('  span of the kernel: %.2f us' % (tt[:, 7].max() - tt[:, 0].min()).item())

# stream kernel: per-CTA phase stamps (grid (T, B))
T = (A + 255) // 256
ss = torch.zeros(B * T * 8, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
L_.dspmb_debug_target_stream_stamps.argtypes = [ctypes.c_void_p]
L_.dspmb_debug_target_stream_stamps(ss.data_ptr())
plan.run(anchors, lab, logits[1], outs[1])
torch.cuda.synchronize()
L_.dspmb_debug_target_stream_stamps(None)
s8 = ss.cpu().view(B, T, 8).double() / 1e3
t0 = s8[:, :, 0].min()
ph = ['count G', 'stage gts', 'gt loop', 'keys', 'outputs', 'publish']
for img in range(min(B, 4)):
    x = s8[img]
    print('image %d: first start %.2f last end %.2f' % (img, (x[:, 0].min() - t0).item(), (x[:, 6].max() - t0).item()),
          ' '.join('%s %.2f/%.2f' % (ph[k], (x[:, k + 1] - x[:, k]).mean().item(), (x[:, k + 1] - x[:, k]).max().item()) for k in range(6)))
print('stream kernel span %.2f us' % (s8[:, :, 6].max() - t0).item())
