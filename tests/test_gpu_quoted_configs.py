"""GPU parity at exactly the shapes BASELINE.json's configs are quoted on (the bench's own inputs), compared with the
reference's own operator .cc / cpu_nms.pyx compiled in place (oracle/_ref) whenever that library travelled to the box,
else with the restatement (bit-identical to it, tests/test_oracle_golden.py).  Plus the code paths that only exist
behind a knob on these sizes: the cluster matcher of MultiBoxTarget and the batch-split detection pipeline."""
import numpy as np
import pytest
import torch

import bench
from dspnet_b200 import presets, synth
from tests import util

pytestmark = pytest.mark.gpu


def _t(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


@pytest.fixture(scope="module")
def checker(oracle):
    """(module, name): oracle/_ref when it was built, else the port."""
    from oracle import ref as R
    if R.available():
        return R, "oracle/_ref"
    return oracle, "oracle port"


def _anchors(M, preset):
    return util.oracle_anchors(M, preset)


def test_detection_ssd512_batch32(checker, cuda):
    """configs[1]: SSD-512 VOC MultiBoxDetection + NMS, batch 32 -- the tensors bench.py times."""
    from dspnet_b200 import MultiBoxDetection
    M, _ = checker
    inputs, _np = bench.make_inputs(0, 32, "detection")
    anchors = _anchors(M, "ssd512")
    want = M.multibox_detection(inputs["prob"], inputs["loc"], anchors, **{k: v for k, v in bench.DET_PARAMS.items()})
    got, valid = MultiBoxDetection(_t(inputs["prob"], cuda), _t(inputs["loc"], cuda), _t(anchors, cuda),
                                   return_valid_count=True, **bench.DET_PARAMS)
    got = got.cpu().numpy()
    util.assert_bit_equal(got[:, :, 0], want[:, :, 0], "ids / kept rows at B=32")
    util.assert_bit_equal(got, want, "detection B=32 (bit-exact)")
    v = valid.cpu().numpy()
    assert (v > bench.DET_PARAMS["nms_topk"]).all(), "every image exercises the nms_topk tail quirk"
    # size-independent property: rows >= valid_count are untouched (-1), kept ids are classes, scores of the head descend
    for b in range(32):
        assert (got[b, v[b]:] == -1).all()
        head = got[b, :bench.DET_PARAMS["nms_topk"], 1]
        assert (np.diff(head) <= 0).all()


def test_target_ssd512_batch64(checker, cuda):
    """configs[2]: SSD-512 MultiBoxTarget, mining ratio 3, batch 64, L = 58 (image 1 has no gt, image 2 fills all)."""
    from dspnet_b200 import MultiBoxTarget
    M, _ = checker
    inputs, _np = bench.make_inputs(0, 64, "target")
    assert inputs["L"] == 58
    anchors = _anchors(M, "ssd512")
    want = M.multibox_target(anchors, inputs["lab"], inputs["logits"], **bench.TGT_PARAMS)
    got = MultiBoxTarget(_t(anchors, cuda), _t(inputs["lab"], cuda), _t(inputs["logits"], cuda), **bench.TGT_PARAMS)
    for g, w, n in zip(got, want, ("loc_target", "loc_mask", "cls_target")):
        util.assert_bit_equal(g.cpu().numpy(), w, n + " B=64")
    ct = got[2].cpu().numpy()
    assert (ct[1] == -1).all(), "image without ground truth stays at ignore_label"
    # property: negatives = 3 x positives (clamped), per image
    pos, neg = (ct > 0).sum(1), (ct == 0).sum(1)
    A = ct.shape[1]
    assert (neg == np.minimum((pos * 3.0).astype(np.int64), A - pos)).all()


def test_dspnet_cs_batch16(checker, cuda):
    """configs[3]: DSPNet Cityscapes head, prior + target + detection, batch 16, L = 200 with the G = 200 image."""
    from dspnet_b200 import MultiBoxDetection, MultiBoxTarget
    from dspnet_b200.symbol import multibox_anchors
    M, _ = checker
    inputs, _np = bench.make_inputs(0, 16, "dspnet_cs")
    assert inputs["L"] == 200 and (inputs["lab"][2, :, 0] != -1).all()
    anchors = _anchors(M, "dspnet_cs")
    util.assert_bit_equal(multibox_anchors("dspnet_cs", device=cuda).cpu().numpy(), anchors, "prior")
    want = M.multibox_target(anchors, inputs["lab"], inputs["logits"], **bench.TGT_PARAMS)
    got = MultiBoxTarget(_t(anchors, cuda), _t(inputs["lab"], cuda), _t(inputs["logits"], cuda), **bench.TGT_PARAMS)
    for g, w, n in zip(got, want, ("loc_target", "loc_mask", "cls_target")):
        util.assert_bit_equal(g.cpu().numpy(), w, n + " dspnet_cs B=16")
    want = M.multibox_detection(inputs["prob"], inputs["loc"], anchors, **bench.DET_PARAMS)
    got = MultiBoxDetection(_t(inputs["prob"], cuda), _t(inputs["loc"], cuda), _t(anchors, cuda), **bench.DET_PARAMS)
    util.assert_bit_equal(got.cpu().numpy(), want, "detection dspnet_cs B=16")


def test_ssd300_batch1(checker, cuda):
    """configs[0]: SSD-300 prior + target + detection, batch 1."""
    from dspnet_b200 import MultiBoxDetection, MultiBoxTarget
    M, _ = checker
    inputs, _np = bench.make_inputs(0, 1, "ssd300")
    anchors = _anchors(M, "ssd300")
    want = M.multibox_target(anchors, inputs["lab"], inputs["logits"], **bench.TGT_PARAMS)
    got = MultiBoxTarget(_t(anchors, cuda), _t(inputs["lab"], cuda), _t(inputs["logits"], cuda), **bench.TGT_PARAMS)
    for g, w, n in zip(got, want, ("loc_target", "loc_mask", "cls_target")):
        util.assert_bit_equal(g.cpu().numpy(), w, n)
    want = M.multibox_detection(inputs["prob"], inputs["loc"], anchors, **bench.DET_PARAMS)
    got = MultiBoxDetection(_t(inputs["prob"], cuda), _t(inputs["loc"], cuda), _t(anchors, cuda), **bench.DET_PARAMS)
    util.assert_bit_equal(got.cpu().numpy(), want, "detection ssd300 B=1")


@pytest.mark.parametrize("n", [20000, 50000])
def test_nms_sweep_large(oracle, cuda, n):
    """configs[4] beyond the small sizes of test_nms_sweep: kept lists identical to cython cpu_nms, force on / off."""
    from oracle import ref as R
    from dspnet_b200.nms import nms_device
    cpu_nms = R.cpu_nms if R.nms_available() else oracle.cpu_nms
    dets = synth.nms_boxes(100 + n, n)
    keep, num = nms_device(_t(dets, cuda), 0.45, rule="ge")
    assert keep[: int(num.item())].cpu().tolist() == cpu_nms(dets, 0.45)
    d6 = synth.nms_boxes(200 + n, n, with_class=True)
    keep, num = nms_device(_t(d6, cuda), 0.45, rule="ge", class_col=5)
    got = keep[: int(num.item())].cpu().tolist()
    want = []
    for c in range(20):  # per class, merged in score order: the oracle SURVEY.md 8d defines for force_suppress=off
        idx = np.nonzero(d6[:, 5] == c)[0]
        want += [int(idx[i]) for i in cpu_nms(d6[idx, :5], 0.45)]
    want.sort(key=lambda i: -d6[i, 4])
    assert got == want


@pytest.mark.slow
def test_nms_sweep_200k(oracle, cuda):
    from oracle import ref as R
    from dspnet_b200.nms import nms_device
    cpu_nms = R.cpu_nms if R.nms_available() else oracle.cpu_nms
    dets = synth.nms_boxes(424242, 200000)
    keep, num = nms_device(_t(dets, cuda), 0.45, rule="ge")
    assert keep[: int(num.item())].cpu().tolist() == cpu_nms(dets, 0.45)


@pytest.mark.parametrize("preset,batch,max_gt", [("ssd300", 3, 8), ("ssd512", 4, 8), ("dspnet_cs", 3, 50)])
def test_target_cluster_matcher(oracle, cuda, preset, batch, max_gt):
    """DSPMB_TUNE_TARGET_PIPELINE = 1: the thread-block-cluster matcher (8 CTAs per image, DSMEM histograms, cooperative
    column recompute) must equal the oracle like the default single-CTA matcher -- presets (G = 0 / G = L images,
    stale column maxima in the 200-gt image), duplicate gts, massive ties in the mining keys, mining disabled."""
    from dspnet_b200 import MultiBoxTarget, _lib
    L = _lib.lib()
    old = L.dspmb_set_tuning(_lib.TUNE_TARGET_PIPELINE, 1)
    try:
        anchors, lab, cp = util.target_inputs(oracle, preset, batch, config_id=21, max_gt=max_gt)
        lab2 = lab.copy()
        lab2[0, 1] = lab2[0, 0]  # duplicate ground truth: shared best anchor -> the sequential bipartite path
        cp_ties = (np.round(cp * 2) / 2).astype(np.float32)
        for lb, logits, kw in ((lab, cp, dict(negative_mining_ratio=3)), (lab2, cp, dict(negative_mining_ratio=3)),
                               (lab, cp_ties, dict(negative_mining_ratio=3)), (lab, np.zeros_like(cp), dict(negative_mining_ratio=3)),
                               (lab, cp, dict()), (lab, cp, dict(negative_mining_ratio=1.5, negative_mining_thresh=0.3)),
                               (lab, cp, dict(negative_mining_ratio=200.0))):
            try:
                want = oracle.multibox_target(anchors, lb, logits, **kw)
            except oracle.OracleError as e:
                from dspnet_b200 import DspmbError
                with pytest.raises(DspmbError) as ei:
                    MultiBoxTarget(_t(anchors, cuda), _t(lb, cuda), _t(logits, cuda), **kw)
                assert ei.value.code == e.code
                continue
            got = MultiBoxTarget(_t(anchors, cuda), _t(lb, cuda), _t(logits, cuda), **kw)
            for g, w, n in zip(got, want, ("loc_target", "loc_mask", "cls_target")):
                util.assert_bit_equal(g.cpu().numpy(), w, "%s %s (cluster matcher)" % (n, kw))
    finally:
        L.dspmb_set_tuning(_lib.TUNE_TARGET_PIPELINE, old)


@pytest.mark.parametrize("groups,batch", [(2, 8), (3, 13), (4, 16)])
def test_detection_batch_split(oracle, cuda, groups, batch):
    """DSPMB_TUNE_DET_SPLIT: image groups whose post-processing runs on a side branch of the library's graph; direct
    launches, the capturing call and the replays all equal the oracle."""
    from dspnet_b200 import MultiBoxDetection, _lib
    L = _lib.lib()
    old = L.dspmb_set_tuning(_lib.TUNE_DET_SPLIT, groups)
    try:
        anchors, prob, lp = util.detection_inputs(oracle, "ssd300", batch, config_id=17)
        want = oracle.multibox_detection(prob, lp, anchors, nms_threshold=0.45, nms_topk=400)
        p, l, a = _t(prob, cuda), _t(lp, cuda), _t(anchors, cuda)
        from dspnet_b200.plan import DetectionPlan
        plan = DetectionPlan(batch, anchors.shape[1], prob.shape[1], cuda, nms_threshold=0.45, nms_topk=400)
        out = plan.new_output()
        for it in range(4):  # first sighting, capture, two replays
            out.fill_(7.0)
            plan.run(p, l, a, out)
            util.assert_bit_equal(out.cpu().numpy(), want, "split %d, call %d" % (groups, it))
        assert plan.launches_per_run == 3 * groups
    finally:
        L.dspmb_set_tuning(_lib.TUNE_DET_SPLIT, old)
