"""Device-side timeline of one detection step in graph-replay mode (dspmb_debug_trace): per kernel the earliest CTA
start and the latest CTA end, relative to the start of the stream kernel."""
import sys, torch
sys.path.insert(0, '.')
import bench
from dspnet_b200 import _lib
from dspnet_b200.plan import DetectionPlan
from dspnet_b200.symbol import multibox_anchors
dev = torch.device('cuda', 0)
inputs, _ = bench.make_inputs(0, bench.BATCH)
A, C = inputs['A'], inputs['C']
anchors = multibox_anchors(bench.PRESET, device=dev)
plan = DetectionPlan(bench.BATCH, A, C, dev, **bench.DET_PARAMS)
prob = [torch.from_numpy(inputs['prob']).to(dev) for _ in range(4)]
loc = [torch.from_numpy(inputs['loc']).to(dev) for _ in range(4)]
out = [plan.new_output() for _ in range(4)]
for i in range(2000):
    plan.run(prob[i % 4], loc[i % 4], anchors, out[i % 4])
torch.cuda.synchronize()
L = _lib.lib()
for kv in sys.argv[1:]:
    if kv == 'nograph':
        L.dspmb_set_tuning(_lib.TUNE_GRAPH_CACHE, 0)
    else:
        k, v = kv.split('=')
        L.dspmb_set_tuning(int(k), int(v))
for i in range(20):
    plan.run(prob[i % 4], loc[i % 4], anchors, out[i % 4])
torch.cuda.synchronize()
names = ['stream', 'sort', 'pair', 'tail', 'resolve', 'pair:scans', 'pair:gather', 'pair:units', 'pair:waited', 'sort:staged', 'sort:bucket-sorted', 'sort:fallback-sorted', 'res:headscan', 'res:ordered', 'res:rounds']
acc = [[0.0, 0.0] for _ in names]
reps = 20
for r in range(reps):
    buf = torch.zeros(32, dtype=torch.int64, device=dev)
    buf[0::2] = torch.iinfo(torch.int64).max
    torch.cuda.synchronize()
    L.dspmb_debug_trace(buf.data_ptr())
    for i in range(3):
        plan.run(prob[i % 4], loc[i % 4], anchors, out[i % 4])
    torch.cuda.synchronize()
    L.dspmb_debug_trace(None)
    # only the LAST of the three steps is unambiguous for max; min comes from the first: run single steps instead
    buf = torch.zeros(32, dtype=torch.int64, device=dev)
    buf[0::2] = torch.iinfo(torch.int64).max
    torch.cuda.synchronize()
    L.dspmb_debug_trace(buf.data_ptr())
    plan.run(prob[r % 4], loc[r % 4], anchors, out[r % 4])
    torch.cuda.synchronize()
    L.dspmb_debug_trace(None)
    t = buf.cpu().tolist()
    maxrounds = max(globals().get('maxrounds', 0), t[31])
    t0 = t[0]
    for k in range(len(names)):
        acc[k][0] += (t[2 * k] - t0) / 1e3 / reps
        acc[k][1] += (t[2 * k + 1] - t0) / 1e3 / reps
for k, n in enumerate(names):
    print('%-22s start %7.2f us   end %7.2f us' % (n, acc[k][0], acc[k][1]))
print('max fixed-point rounds', maxrounds)

# per-CTA phase stamps of the pair kernel (no contended atomics): mean / max duration of every phase
st = torch.zeros(640 * 10, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
L.dspmb_debug_stamps(st.data_ptr())
plan.run(prob[0], loc[0], anchors, out[0])
torch.cuda.synchronize()
L.dspmb_debug_stamps(None)
t = st.cpu().view(640, 10).double() / 1e3
lab = ['scans', 'gather', 'units', 'wait', 'headscan', 'order', 'rounds', 'ids']
ok = (t[:, 8] > 0)
t = t[ok]
print('pair kernel CTAs with a resolve:', int(ok.sum()))
for k, name in enumerate(lab):
    d = t[:, k + 1] - t[:, k]
    print('  %-9s mean %6.2f  max %6.2f us' % (name, d.mean().item(), d.max().item()))
print('  %-9s mean %6.2f  max %6.2f us' % ('total', (t[:, 8] - t[:, 0]).mean().item(), (t[:, 8] - t[:, 0]).max().item()))
print('  first start %.2f, last end %.2f (span)' % (0.0, (t[:, 8].max() - t[:, 0].min()).item()))
