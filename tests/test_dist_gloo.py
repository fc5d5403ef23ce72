"""world_size-2 gloo test of the image-sharded driver (host logic of dspnet_b200/dist.py) on CPU: with the oracle
injected as the per-shard compute provider, the gathered detections / target statistics must be identical to the
unsharded run -- for a batch that does not divide evenly, too."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _OracleOps:
    def __init__(self):
        from oracle import oracle as O
        self.O = O

    def MultiBoxDetection(self, cls_prob, loc_pred, anchor, **kw):
        return torch.from_numpy(self.O.multibox_detection(cls_prob.numpy(), loc_pred.numpy(), anchor.numpy(), **kw))

    def MultiBoxTarget(self, anchor, label, cls_pred, **kw):
        return [torch.from_numpy(x) for x in self.O.multibox_target(anchor.numpy(), label.numpy(), cls_pred.numpy(), **kw)]


def _worker(rank, world, port, batch, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dspnet_b200 import synth
        from dspnet_b200.dist import ShardedMultiBox
        from tests import util
        ops = _OracleOps()
        anchors, prob, lp = util.detection_inputs(ops.O, "ssd300", batch, config_id=81)
        sm = ShardedMultiBox(ops=ops)
        rows, counts, _ = sm.detection(torch.from_numpy(prob), torch.from_numpy(lp), torch.from_numpy(anchors), 50,
                                       nms_threshold=0.45, nms_topk=400)
        anchors, lab, cp = util.target_inputs(ops.O, "ssd300", batch, config_id=82)
        _, stats = sm.target(torch.from_numpy(anchors), torch.from_numpy(lab), torch.from_numpy(cp), negative_mining_ratio=3)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), rows=rows.numpy(), counts=counts.numpy(), stats=stats.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 5])
def test_sharded_equals_unsharded(tmp_path, oracle, batch):
    from dspnet_b200.dist import compact_rows
    from tests import util
    port = 29500 + (os.getpid() % 2000) + batch
    mp.spawn(_worker, args=(2, port, batch, str(tmp_path)), nprocs=2, join=True)
    anchors, prob, lp = util.detection_inputs(oracle, "ssd300", batch, config_id=81)
    full = torch.from_numpy(oracle.multibox_detection(prob, lp, anchors, nms_threshold=0.45, nms_topk=400))
    rows, counts = compact_rows(full, 50)
    anchors, lab, cp = util.target_inputs(oracle, "ssd300", batch, config_id=82)
    ct = oracle.multibox_target(anchors, lab, cp, negative_mining_ratio=3)[2]
    stats = np.stack([(ct > 0).sum(1), (ct == 0).sum(1), (ct == -1).sum(1)], axis=1)
    for r in range(2):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(z["rows"], rows.numpy())
        assert np.array_equal(z["counts"], counts.numpy())
        assert np.array_equal(z["stats"], stats)
    assert (counts.numpy() > 0).all() and (rows.numpy()[:, 0, 0] >= 0).all()
