#!/usr/bin/env python
"""Generates tests/golden/multibox_golden.npz from the REFERENCE ITSELF: oracle/_ref/ = the reference's
operator/multibox_{prior,target,detection}.cc and cython/cpu_nms.pyx compiled in place from /root/reference by
oracle/build_ref.py.  Run in the build container (the reference tree does not exist on the GPU box):

    python oracle/build_ref.py && python tests/golden/make_golden.py

Small cases are stored in full (inputs and outputs); the BASELINE-sized cases are stored as SHA-256 digests of the
output bytes together with a digest of the seeded inputs (so a drift of numpy's generators is detected instead of
being reported as a parity failure).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dspnet_b200 import presets, synth  # noqa: E402
from oracle import ref as R  # noqa: E402

TINY = presets.Preset("tiny", presets._maps([(8, 8), (4, 4), (2, 2), (1, 1)], presets._SSD300_SIZES[:4],
                                           [presets._R3, presets._R5, presets._R5, presets._R3], [-1.0] * 4), 5, 8)
TARGET_KW = dict(overlap_threshold=0.5, ignore_label=-1.0, negative_mining_ratio=3.0, negative_mining_thresh=0.5,
                 minimum_negative_samples=0, variances=(0.1, 0.1, 0.2, 0.2))
DET_KW = dict(threshold=0.01, clip=True, nms_threshold=0.45, force_suppress=False, nms_topk=400, variances=(0.1, 0.1, 0.2, 0.2))


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def anchors_of(mod, p):
    return np.concatenate([mod.multibox_prior(fm.height, fm.width, fm.sizes, fm.ratios, False, (fm.step, fm.step))
                           for fm in p.maps], axis=1)


def tiny_inputs():
    a = TINY_ANCHORS
    A = a.shape[1]
    lab = synth.labels(41, 3, TINY.label_slots, TINY.num_classes, max_gt=6)
    logits = synth.cls_preds(41, 3, TINY.num_classes, A)
    prob = synth.cls_prob(41, 3, TINY.num_classes, A)
    loc = synth.loc_pred(41, 3, A)
    return lab, logits, prob, loc


BIG_CASES = [  # name, preset, batch, op, config_id, extra kwargs
    ("ssd300_target_b2", "ssd300", 2, "target", 51, {}),
    ("ssd300_detection_b2", "ssd300", 2, "detection", 52, {}),
    ("ssd512_detection_b2", "ssd512", 2, "detection", 53, {}),
    ("ssd512_detection_b2_force", "ssd512", 2, "detection", 53, {"force_suppress": True}),
    ("ssd512_target_b3", "ssd512", 3, "target", 54, {}),
    ("dspnet_cs_target_b3", "dspnet_cs", 3, "target", 55, {"max_gt": 50}),
    ("dspnet_cs_detection_b2", "dspnet_cs", 2, "detection", 56, {}),
]


def big_case_inputs(preset, batch, op, config_id, extra, anchors):
    p = presets.PRESETS[preset]
    A = anchors.shape[1]
    if op == "target":
        lab = synth.labels(config_id, batch, p.label_slots, p.num_classes, max_gt=extra.get("max_gt", 8))
        logits = synth.cls_preds(config_id, batch, p.num_classes, A)
        return lab, logits
    return synth.cls_prob(config_id, batch, p.num_classes, A), synth.loc_pred(config_id, batch, A)


if __name__ == "__main__":
    assert R.available(), "build oracle/_ref first (python oracle/build_ref.py)"
    out = {}
    meta = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref (reference .cc / .pyx compiled in place)",
            "numpy": np.__version__, "digests": {}}
    TINY_ANCHORS = anchors_of(R, TINY)
    out["tiny_anchors"] = TINY_ANCHORS
    lab, logits, prob, loc = tiny_inputs()
    out.update(tiny_label=lab, tiny_logits=logits, tiny_prob=prob, tiny_loc=loc)
    lt, lm, ct = R.multibox_target(TINY_ANCHORS, lab, logits, **TARGET_KW)
    out.update(tiny_loc_target=lt, tiny_loc_mask=lm, tiny_cls_target=ct)
    out["tiny_detection"] = R.multibox_detection(prob, loc, TINY_ANCHORS, **DET_KW)
    out["tiny_detection_force_top20"] = R.multibox_detection(prob, loc, TINY_ANCHORS, **dict(DET_KW, force_suppress=True, nms_topk=20))
    dets = synth.nms_boxes(61, 300)
    out["nms_dets"] = dets
    out["nms_keep_045"] = np.array(R.cpu_nms(dets, 0.45), np.int64) if R.nms_available() else np.zeros(0, np.int64)
    for name in presets.PRESETS:
        meta["digests"]["anchors_" + name] = digest(anchors_of(R, presets.PRESETS[name]))
    for name, preset, batch, op, cid, extra in BIG_CASES:
        anchors = anchors_of(R, presets.PRESETS[preset])
        x, y = big_case_inputs(preset, batch, op, cid, extra, anchors)
        if op == "target":
            res = R.multibox_target(anchors, x, y, **TARGET_KW)
        else:
            res = [R.multibox_detection(x, y, anchors, **dict(DET_KW, **{k: v for k, v in extra.items() if k != "max_gt"}))]
        meta["digests"][name] = {"inputs": digest(x, y), "outputs": digest(*res)}
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multibox_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    print(json.dumps(meta["digests"], indent=1)[:600])
