"""CPU tests: the oracle restatement (oracle/multibox_oracle.cc) against the golden vectors produced by the
reference's own sources (tests/golden/make_golden.py), and -- where oracle/_ref/ is present -- against the
reference-compiled library directly."""
import numpy as np
import pytest

from dspnet_b200 import presets, synth
from tests import golden_util, util


@pytest.fixture(scope="module")
def golden():
    z, meta = golden_util.load()
    return z, meta, golden_util.generator()


def test_tiny_cases_match_reference(oracle, golden):
    z, meta, gen = golden
    anchors = np.concatenate([oracle.multibox_prior(fm.height, fm.width, fm.sizes, fm.ratios, False, (fm.step, fm.step))
                              for fm in gen.TINY.maps], axis=1)
    util.assert_bit_equal(anchors, z["tiny_anchors"], "tiny anchors")
    lt, lm, ct = oracle.multibox_target(z["tiny_anchors"], z["tiny_label"], z["tiny_logits"], **gen.TARGET_KW)
    util.assert_bit_equal(lt, z["tiny_loc_target"], "loc_target")
    util.assert_bit_equal(lm, z["tiny_loc_mask"], "loc_mask")
    util.assert_bit_equal(ct, z["tiny_cls_target"], "cls_target")
    det = oracle.multibox_detection(z["tiny_prob"], z["tiny_loc"], z["tiny_anchors"], **gen.DET_KW)
    util.assert_bit_equal(det, z["tiny_detection"], "detection")
    det = oracle.multibox_detection(z["tiny_prob"], z["tiny_loc"], z["tiny_anchors"],
                                    **dict(gen.DET_KW, force_suppress=True, nms_topk=20))
    util.assert_bit_equal(det, z["tiny_detection_force_top20"], "detection force/top20")
    if z["nms_keep_045"].size:
        assert oracle.cpu_nms(z["nms_dets"], 0.45) == z["nms_keep_045"].tolist()


def test_anchor_digests(oracle, golden):
    _, meta, gen = golden
    for name, p in presets.PRESETS.items():
        assert gen.digest(util.oracle_anchors(oracle, name)) == meta["digests"]["anchors_" + name], name
    assert util.oracle_anchors(oracle, "ssd300").shape[1] == 8732
    assert util.oracle_anchors(oracle, "ssd512").shape[1] == 24564
    assert util.oracle_anchors(oracle, "dspnet_cs").shape[1] == 12264  # utils.py:37 of the reference


@pytest.mark.parametrize("case", ["ssd300_target_b2", "ssd300_detection_b2", "ssd512_detection_b2",
                                  "ssd512_detection_b2_force", "dspnet_cs_detection_b2"])
def test_baseline_sized_digests(oracle, golden, case):
    _, meta, gen = golden
    name, preset, batch, op, cid, extra = next(c for c in gen.BIG_CASES if c[0] == case)
    anchors = util.oracle_anchors(oracle, preset)
    x, y = gen.big_case_inputs(preset, batch, op, cid, extra, anchors)
    if gen.digest(x, y) != meta["digests"][case]["inputs"]:
        pytest.skip("numpy generator stream differs from the one the golden inputs were made with")
    if op == "target":
        res = oracle.multibox_target(anchors, x, y, **gen.TARGET_KW)
    else:
        res = [oracle.multibox_detection(x, y, anchors, **dict(gen.DET_KW, **{k: v for k, v in extra.items() if k != "max_gt"}))]
    assert gen.digest(*res) == meta["digests"][case]["outputs"]


def test_restatement_equals_reference_build(oracle):
    """Direct comparison with the reference's .cc bodies compiled in place (only where oracle/_ref exists)."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    anchors, lab, cp = util.target_inputs(oracle, "ssd300", 3, config_id=71)
    lab[0, 1] = lab[0, 0]  # duplicate gt -> bipartite re-evaluation
    for kw in (dict(negative_mining_ratio=3), dict(), dict(negative_mining_ratio=2, overlap_threshold=0.0, ignore_label=-3)):
        for a, b in zip(oracle.multibox_target(anchors, lab, cp, **kw), R.multibox_target(anchors, lab, cp, **kw)):
            util.assert_bit_equal(a, b, "target %r" % (kw,))
    anchors, prob, lp = util.detection_inputs(oracle, "ssd300", 2, config_id=72)
    for kw in (dict(nms_threshold=0.45, nms_topk=400), dict(nms_threshold=0.5, force_suppress=True), dict(nms_threshold=0.0),
               dict(clip=False, threshold=0.3, nms_topk=7)):
        util.assert_bit_equal(oracle.multibox_detection(prob, lp, anchors, **kw), R.multibox_detection(prob, lp, anchors, **kw),
                              "detection %r" % (kw,))
    if R.nms_available():
        d = synth.nms_boxes(73, 1500)
        for thr in (0.45, float(np.float32(0.45)), 0.7):
            assert oracle.cpu_nms(d, thr) == R.cpu_nms(d, thr)


def test_oracle_error_codes(oracle):
    anchors, lab, cp = util.target_inputs(oracle, "ssd300", 2, config_id=74)
    g = int((lab[0, :, 0] != -1).sum())
    bad = lab.copy()
    bad[0, g] = (-1, 0.3, -1, -1, -1, -1)
    with pytest.raises(oracle.OracleError) as e:
        oracle.multibox_target(anchors, bad, cp, negative_mining_ratio=3)
    assert e.value.code == -2
    with pytest.raises(oracle.OracleError) as e:
        oracle.multibox_target(anchors, lab, cp, negative_mining_ratio=3, negative_mining_thresh=0.0)
    assert e.value.code == -4


def test_oracle_semantics_spot_checks(oracle):
    """Hand-checkable facts of the reference semantics."""
    # prior: ratios[0] is never read; count = S + R - 1 per cell; centre = (c + 0.5) / W
    a = oracle.multibox_prior(1, 2, (0.5,), (7.0, 4.0))[0]
    assert a.shape == (4, 4)
    np.testing.assert_allclose(a[0], [0.25 - 0.5 * 1 / 2 / 2, 0.5 - 0.25, 0.25 + 0.125, 0.75], rtol=1e-6)
    np.testing.assert_allclose(a[1], [0.25 - 0.5 * 0.5 * 2 / 2, 0.5 - 0.5 / 2 / 2, 0.25 + 0.25, 0.5 + 0.125], rtol=1e-6)
    # detection: rows beyond nms_topk keep their pass-1 (anchor order) content and still take part in NMS
    anchors = np.array([[[0.1, 0.1, 0.3, 0.3], [0.1, 0.1, 0.3, 0.3], [0.6, 0.6, 0.9, 0.9]]], np.float32)
    prob = np.array([[[0.1, 0.1, 0.1], [0.5, 0.9, 0.7]]], np.float32)  # (1, 2, 3): one foreground class
    loc = np.zeros((1, 15), np.float32)
    out = oracle.multibox_detection(prob, loc, anchors, nms_threshold=0.5, nms_topk=1)
    # sorted head: anchor 1 (0.9); tail rows 1, 2 keep anchors 1 and 2 in anchor order; row 1 duplicates the head
    assert out[0, 0, 1] == np.float32(0.9) and out[0, 0, 0] == 0
    assert out[0, 1, 1] == np.float32(0.9) and out[0, 1, 0] == -1  # suppressed by its own copy (IoU 1)
    assert out[0, 2, 1] == np.float32(0.7) and out[0, 2, 0] == 0
    # target: an image without ground truth leaves every output at its initial value
    lab = np.full((1, 3, 6), -1, np.float32)
    lt, lm, ct = oracle.multibox_target(anchors, lab, np.zeros((1, 2, 3), np.float32), negative_mining_ratio=3)
    assert not lt.any() and not lm.any() and (ct == -1).all()


def test_py_nms_equals_the_reference_file(oracle):
    """oracle.py_nms against the reference's own detect/nms.py::nms (its two Cython imports stubbed), SURVEY 8c."""
    from oracle import ref_pynms as RP
    if not RP.available():
        pytest.skip("/root/reference not present (GPU box): the pin is checked in the build container")
    from dspnet_b200 import synth
    for seed, n in ((1, 50), (2, 400), (3, 1500)):
        dets = synth.nms_boxes(seed, n)
        for thr in (0.3, 0.45, 0.7):
            assert oracle.py_nms(dets, thr) == RP.nms(dets, thr)
