import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import oracle as O
from tests import util
from dspnet_b200 import MultiBoxDetection
dev = torch.device('cuda', 0)
preset = sys.argv[1] if len(sys.argv) > 1 else 'ssd300'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
force = len(sys.argv) > 3 and sys.argv[3] == 'force'
from dspnet_b200 import _lib
import os
for k in range(4):
    v = os.environ.get('TUNE%d' % k)
    if v is not None:
        print('tune', k, v, _lib.lib().dspmb_set_tuning(k, int(v)))
anchors, prob, lp = util.detection_inputs(O, preset, B, config_id=11)
kw = dict(nms_threshold=0.45, force_suppress=force, nms_topk=400)
want, wv = O.multibox_detection(prob, lp, anchors, return_valid=True, **kw)
got, gv = MultiBoxDetection(torch.from_numpy(prob).to(dev), torch.from_numpy(lp).to(dev), torch.from_numpy(anchors).to(dev), return_valid_count=True, **kw)
got = got.cpu().numpy()
print('valid', wv, gv.cpu().numpy())
for b in range(B):
    V = wv[b]
    same_rows = np.array_equal(got[b, :V, 1:], want[b, :V, 1:])
    d = np.nonzero(got[b, :V, 0] != want[b, :V, 0])[0]
    print('image', b, 'V', V, 'payload equal', same_rows, 'id mismatches', len(d))
    # recompute: what ids did rows have before NMS (from payload match)
    for r in d[:10]:
        print('  row', r, 'got', got[b, r, 0], 'want', want[b, r, 0], 'score', want[b, r, 1])
    if len(d):
        cls = want[b, d, 0]
        print('  classes of wrongly suppressed rows:', np.unique(cls, return_counts=True))
        print('  rows < 400:', (d < 400).sum(), ' rows >= 400:', (d >= 400).sum())
