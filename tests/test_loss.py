"""forward of the training graph behind MultiBoxTarget (channel softmax of SoftmaxOutput,
masked smooth-L1 of MakeLoss) and the MultiBoxMetric statistics.  CPU part: the numpy oracle against the reference's
own train/metric.py (loaded behind an mxnet stand-in, build container only).  GPU part: the fused kernel against the
oracle -- cls_prob and loc_loss bit for bit, the fp64 sums to 1e-12 relative (summation order)."""
import numpy as np
import pytest
import torch

from dspnet_b200 import presets, synth
from tests import util

SUM_RTOL = 1e-12   # fp64 sums of identical fp32 terms in a different order
CE_RTOL = 1e-6     # -log(p + eps): the reference calls numpy's float32 log, the kernel glibc's logf (<= 1 ulp per term);
                   # north_star allows 1e-5 relative on float outputs
REF_RTOL = 1e-5    # the reference's metric adds float32 terms with numpy's float32 pairwise sum


def _inputs(oracle, preset, batch, config_id=41):
    p = presets.PRESETS[preset]
    anchors, lab, cp = util.target_inputs(oracle, preset, batch, config_id=config_id)
    A = anchors.shape[1]
    lt, lm, ct = oracle.multibox_target(anchors, lab, cp, negative_mining_ratio=3)
    lp = synth.loc_pred(config_id, batch, A)
    return cp, lp, lt, lm, ct


def test_oracle_metric_equals_the_reference_class(oracle):
    from oracle import ref_metric as RM
    if not RM.available():
        pytest.skip("/root/reference not present (GPU box): the pin is checked in the build container")
    cp, lp, lt, lm, ct = _inputs(oracle, "ssd300", 3)
    prob, loss, stats = oracle.multibox_training_outputs(cp, lp, lt, lm, ct)
    names, values, sums, counts = RM.multibox_metric(prob, loss, ct)
    assert names == ['CrossEntropy', 'SmoothL1']
    assert counts[0] == counts[1] == int(stats[:, 0].sum())
    np.testing.assert_allclose(sums[0], stats[:, 1].sum(), rtol=REF_RTOL)
    np.testing.assert_allclose(sums[1], stats[:, 2].sum(), rtol=REF_RTOL)
    # the product's host-side class walks numpy inputs through the reference's arithmetic: identical numbers
    from dspnet_b200.loss import MultiBoxMetric
    m = MultiBoxMetric()
    m.update(None, [prob, loss, ct])
    assert m.get()[0] == names and m.get()[1] == values
    # smooth_l1 of the oracle on a few hand-checked values (mx.symbol.smooth_l1, scalar = 1)
    np.testing.assert_array_equal(oracle.smooth_l1(np.array([0.0, 0.5, -0.5, 1.0, -2.0, 3.5], np.float32)),
                                  np.array([0.0, 0.125, 0.125, 0.5, 1.5, 3.0], np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("preset,batch", [("ssd300", 3), ("ssd512", 2), ("dspnet_cs", 2)])
def test_training_outputs_on_gpu(oracle, cuda, preset, batch):
    from dspnet_b200.loss import MultiBoxMetric, multibox_training_outputs
    cp, lp, lt, lm, ct = _inputs(oracle, preset, batch)
    want_prob, want_loss, want_stats = oracle.multibox_training_outputs(cp, lp, lt, lm, ct)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda)
    prob, loss, stats = multibox_training_outputs(t(cp), t(lp), t(lt), t(lm), t(ct))
    util.assert_bit_equal(prob.cpu().numpy(), want_prob, "cls_prob")
    util.assert_bit_equal(loss.cpu().numpy(), want_loss, "loc_loss")
    s = stats.cpu().numpy()
    np.testing.assert_array_equal(s[:, 0], want_stats[:, 0])
    np.testing.assert_allclose(s[:, 1], want_stats[:, 1], rtol=CE_RTOL)
    np.testing.assert_allclose(s[:, 2], want_stats[:, 2], rtol=SUM_RTOL)
    np.testing.assert_array_equal(s[:, 3], (want_loss > 0).sum(axis=1))
    # statistics only (a different kernel: softmax of the labelled anchors alone): same numbers up to the order of the
    # fp64 additions, and bit-identical run to run
    _, _, s2 = multibox_training_outputs(t(cp), t(lp), t(lt), t(lm), t(ct), want_cls_prob=False, want_loc_loss=False)
    _, _, s3 = multibox_training_outputs(t(cp), t(lp), t(lt), t(lm), t(ct), want_cls_prob=False, want_loc_loss=False)
    assert torch.equal(s2, s3)
    np.testing.assert_allclose(s2.cpu().numpy(), s, rtol=SUM_RTOL)
    _, l4, s4 = multibox_training_outputs(t(cp), t(lp), t(lt), t(lm), t(ct), want_cls_prob=False, want_loc_loss=True)
    util.assert_bit_equal(l4.cpu().numpy(), want_loss, "loc_loss (statistics kernel)")
    m1, m2 = MultiBoxMetric(), MultiBoxMetric()
    m1.update_from_stats(stats)
    m2.update(None, [prob, loss, t(ct)])
    np.testing.assert_allclose(m1.get()[1], m2.get()[1], rtol=1e-6)


@pytest.mark.gpu
def test_training_outputs_odd_shapes_and_generic_class_count(oracle, cuda):
    from dspnet_b200.loss import multibox_training_outputs
    rng = np.random.default_rng(3)
    for (B, C, A) in ((2, 5, 1003), (1, 21, 380), (3, 9, 64)):
        cp = rng.normal(0, 3, (B, C, A)).astype(np.float32)
        lp = rng.normal(0, 1, (B, A * 5)).astype(np.float32)
        lt = rng.normal(0, 1, (B, A * 5)).astype(np.float32)
        lm = (rng.random((B, A, 1)) < 0.1).astype(np.float32).repeat(5, axis=2).reshape(B, A * 5)
        ct = rng.integers(-1, C, (B, A)).astype(np.float32)
        want_prob, want_loss, want_stats = oracle.multibox_training_outputs(cp, lp, lt, lm, ct)
        t = lambda x: torch.from_numpy(x).to(cuda)
        prob, loss, stats = multibox_training_outputs(t(cp), t(lp), t(lt), t(lm), t(ct))
        util.assert_bit_equal(prob.cpu().numpy(), want_prob, "cls_prob %s" % ((B, C, A),))
        util.assert_bit_equal(loss.cpu().numpy(), want_loss, "loc_loss")
        np.testing.assert_allclose(stats.cpu().numpy()[:, :3], want_stats, rtol=CE_RTOL)
        _, _, s2 = multibox_training_outputs(t(cp), t(lp), t(lt), t(lm), t(ct), want_cls_prob=False, want_loc_loss=False)
        np.testing.assert_allclose(s2.cpu().numpy()[:, :3], want_stats, rtol=CE_RTOL)
