"""Probe: does splitting the B=32 detection batch into S sub-batches on forked streams shorten the step?
Usage: python scripts/split_probe.py [steps]"""
import sys
sys.path.insert(0, '.')
import torch
import bench
from dspnet_b200.plan import DetectionPlan
from dspnet_b200.symbol import multibox_anchors

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
inputs, _ = bench.make_inputs(0, bench.BATCH)
A, C = inputs['A'], inputs['C']
B = bench.BATCH
anchors = multibox_anchors(bench.PRESET, device=dev)
R = bench.ROTATE
prob = [torch.from_numpy(inputs['prob']).to(dev) for _ in range(R)]
loc = [torch.from_numpy(inputs['loc']).to(dev) for _ in range(R)]
ref_plan = DetectionPlan(B, A, C, dev, **bench.DET_PARAMS)
ref_out = ref_plan.run(prob[0], loc[0], anchors, ref_plan.new_output()).clone()

for S in (1, 2, 4, 8):
    sub = B // S
    plans = [DetectionPlan(sub, A, C, dev, **bench.DET_PARAMS) for _ in range(S)]
    outs = [torch.empty((B, A, 7), dtype=torch.float32, device=dev) for _ in range(R)]
    side = [torch.cuda.Stream(dev) for _ in range(S - 1)]
    fork = torch.cuda.Event()
    joins = [torch.cuda.Event() for _ in range(S - 1)]

    def step(i):
        r = i % R
        cur = torch.cuda.current_stream()
        fork.record(cur)
        for k in range(S):
            sl = slice(k * sub, (k + 1) * sub)
            if k == 0:
                plans[0].run(prob[r][sl], loc[r][sl], anchors, outs[r][sl])
            else:
                st = side[k - 1]
                st.wait_event(fork)
                plans[k].run(prob[r][sl], loc[r][sl], anchors, outs[r][sl], stream=st.cuda_stream)
                joins[k - 1].record(st)
        for k in range(S - 1):
            cur.wait_event(joins[k])

    for mode in ('eager', 'graph'):
        if mode == 'graph':
            g = torch.cuda.CUDAGraph()
            cs = torch.cuda.Stream(dev)
            with torch.cuda.stream(cs):
                with torch.cuda.graph(g, stream=cs):
                    for r in range(R):
                        step(r)
            def run_n(n):
                for _ in range(n // R):
                    g.replay()
        else:
            def run_n(n):
                for i in range(n):
                    step(i)
        run_n(2000)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_n(steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ok = torch.equal(outs[0], ref_out)
        print('SPLIT', S, mode, 'ms/step %.4f' % ms, 'img/s %.0f' % (B / ms * 1e3), 'equal', ok, flush=True)
