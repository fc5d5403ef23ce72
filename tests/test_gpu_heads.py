"""MultiBoxDetection fed by the per-scale prediction heads (layout shuffles + channel softmax
fused into the stream kernel) against the oracle's restatement of the reference graph
(symbol/common.py:399-432, symbol/symbol_builder.py:156-165): head_layout -> softmax_channel -> multibox_detection."""
import numpy as np
import pytest
import torch

from dspnet_b200 import presets, synth
from tests import util

pytestmark = pytest.mark.gpu


def _t(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def _case(oracle, cuda, preset, batch, logits, lp, **kw):
    from dspnet_b200 import MultiBoxDetectionFromHeads
    p = presets.PRESETS[preset]
    anchors = util.oracle_anchors(oracle, preset)
    ch, lh = synth.heads_from_logits(p, logits, lp)
    # the oracle walks the reference graph: layout shuffles (numpy), channel softmax, detection
    cp, lp2 = oracle.head_layout(ch, lh, p.num_classes)
    util.assert_bit_equal(cp, logits, "head_layout inverts heads_from_logits")
    util.assert_bit_equal(lp2, lp, "head_layout (loc)")
    want, want_valid = oracle.multibox_detection(oracle.softmax_channel(cp), lp2, anchors, return_valid=True, **kw)
    got, valid = MultiBoxDetectionFromHeads([_t(h, cuda) for h in ch], [_t(h, cuda) for h in lh], _t(anchors, cuda),
                                            p.num_classes, return_valid_count=True, **kw)
    util.assert_bit_equal(valid.cpu().numpy(), want_valid, "valid_count")
    util.assert_bit_equal(got.cpu().numpy(), want, "detection from heads (bit-exact)")
    return want_valid


@pytest.mark.parametrize("preset,batch", [("ssd300", 2), ("ssd512", 3), ("dspnet_cs", 2)])
def test_detection_from_heads_presets(oracle, cuda, preset, batch):
    p = presets.PRESETS[preset]
    A = presets.num_anchors(p)
    logits = synth.det_logits(31, batch, p.num_classes, A)
    lp = synth.loc_pred(31, batch, A)
    v = _case(oracle, cuda, preset, batch, logits, lp, nms_threshold=0.45, nms_topk=400)
    assert (v > 100).all()
    _case(oracle, cuda, preset, batch, logits, lp, nms_threshold=0.45, nms_topk=-1, force_suppress=True)


def test_detection_from_heads_dense_ties_and_thresholds(oracle, cuda):
    """Plain N(0,1) logits (every anchor near or above the threshold band), logits rounded to halves (equal
    probabilities: the first-maximum rule on the ROUNDED probabilities decides), constant logits, threshold 0 and a
    threshold no anchor reaches."""
    p = presets.PRESETS["ssd300"]
    A = presets.num_anchors(p)
    rng = np.random.default_rng(5)
    lp = synth.loc_pred(32, 2, A)
    dense = rng.standard_normal((2, p.num_classes, A)).astype(np.float32)
    _case(oracle, cuda, "ssd300", 2, dense, lp, nms_threshold=0.45, nms_topk=400)
    _case(oracle, cuda, "ssd300", 2, (np.round(dense * 2) / 2).astype(np.float32), lp, nms_threshold=0.45, nms_topk=400)
    _case(oracle, cuda, "ssd300", 2, np.zeros_like(dense), lp, nms_threshold=0.45, nms_topk=50)
    _case(oracle, cuda, "ssd300", 2, dense, lp, threshold=0.0, nms_threshold=0.5, nms_topk=100)
    _case(oracle, cuda, "ssd300", 2, dense, lp, threshold=0.9999, nms_threshold=0.5)
    _case(oracle, cuda, "ssd300", 2, (dense * 30).astype(np.float32), lp, threshold=0.5, nms_threshold=0.45, nms_topk=400)


def test_detection_from_heads_near_threshold(oracle, cuda):
    """Scores engineered to sit within a few ulps of the threshold: the approximate pre-filter must hand every one of
    them to the exact evaluation."""
    p = presets.PRESETS["ssd300"]
    A = presets.num_anchors(p)
    rng = np.random.default_rng(9)
    x = np.zeros((1, p.num_classes, A), np.float32)
    # one foreground logit t against 20 zeros: p = e^t / (e^t + 20); p = 0.01 at t = log(20/99)
    t0 = np.log(20.0 / 99.0)
    x[0, 1 + rng.integers(0, p.num_classes - 1, A), np.arange(A)] = (t0 + rng.uniform(-3e-6, 3e-6, A)).astype(np.float32)
    lp = synth.loc_pred(33, 1, A)
    _case(oracle, cuda, "ssd300", 1, x, lp, nms_threshold=0.45, nms_topk=400)
