"""Shared helpers for the parity tests: seeded inputs for the BASELINE configs and exact comparisons."""
import numpy as np

from dspnet_b200 import presets, synth


def oracle_anchors(O, preset):
    p = presets.PRESETS[preset]
    return np.concatenate([O.multibox_prior(fm.height, fm.width, fm.sizes, fm.ratios, False, (fm.step, fm.step))
                           for fm in p.maps], axis=1)


def target_inputs(O, preset, batch, config_id, max_gt=8, first_image=0):
    p = presets.PRESETS[preset]
    anchors = oracle_anchors(O, preset)
    a = anchors.shape[1]
    lab = synth.labels(config_id, batch, p.label_slots, p.num_classes, max_gt=max_gt, first_image=first_image)
    cp = synth.cls_preds(config_id, batch, p.num_classes, a, first_image=first_image)
    return anchors, lab, cp


def detection_inputs(O, preset, batch, config_id, dense=False, first_image=0):
    p = presets.PRESETS[preset]
    anchors = oracle_anchors(O, preset)
    a = anchors.shape[1]
    prob = synth.cls_prob(config_id, batch, p.num_classes, a, first_image=first_image, dense=dense)
    lp = synth.loc_pred(config_id, batch, a, first_image=first_image)
    return anchors, prob, lp


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def assert_bit_equal(got, want, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (what, got.shape, want.shape)
    if got.dtype.kind == "f":
        bad = bits(got) != bits(want)
        # +0 / -0 and NaN payloads are not distinguished by the reference's comparisons
        bad &= ~((got == 0) & (want == 0))
        bad &= ~(np.isnan(got) & np.isnan(want))
    else:
        bad = got != want
    if bad.any():
        idx = np.argwhere(bad)[:5]
        raise AssertionError("%s: %d of %d elements differ, first at %s: got %s want %s" % (
            what, bad.sum(), bad.size, idx.tolist(), got[tuple(idx[0])], want[tuple(idx[0])]))
