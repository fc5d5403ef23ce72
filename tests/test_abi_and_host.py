"""CPU tests of the drop-in boundary and the host logic: the C-ABI library loads and exports every symbol
include/dspmb.h declares, argument CHECKs map to error codes without touching a GPU, presets / synthetic
generators / sharding arithmetic behave."""
import ctypes
import os
import re

import numpy as np
import pytest

from dspnet_b200 import _lib, presets, synth
from dspnet_b200.dist import shard_slice

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "dspmb.h")).read()
    declared = sorted(set(re.findall(r"\b(dspmb_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 15
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), "libdspmb.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == declared
    assert L.dspmb_version() == 100


def test_argument_checks_return_codes_without_gpu():
    L = _lib.lib()
    sizes = _lib.float_array([0.1])
    # CHECK_GE(offsets, 0) of MultiBoxPriorOp's ctor (multibox_prior-inl.h:90-93)
    assert L.dspmb_prior_f32(None, 4, 4, sizes, 1, sizes, 1, -1.0, -1.0, 1.5, 0.5, 0, None) == _lib.ERR_BAD_ARG
    assert b"offsets" in L.dspmb_last_error()
    assert L.dspmb_prior_f32(None, 0, 4, sizes, 1, sizes, 1, -1.0, -1.0, 0.5, 0.5, 0, None) == _lib.ERR_BAD_ARG
    var = _lib.float_array([0.1, 0.1, 0.2, 0.2])
    assert L.dspmb_detection_f32(None, None, None, None, 1, 0, 3, 0.01, 1, var, 0.5, 0, -1, None, None, 0, None) == _lib.ERR_BAD_ARG
    assert L.dspmb_target_f32(None, None, None, None, None, None, 1, 8, 4, 5, 3, 0.5, -1.0, 3.0, 0.5, 0, var, None, None,
                              None, 0, None) == _lib.ERR_BAD_ARG  # label width must be >= 6
    assert L.dspmb_nms_f32(None, 5, 4, 0.5, 0, -1, 0, None, None, None, 0, None) == _lib.ERR_BAD_ARG
    # workspace queries are pure functions of the shape
    assert L.dspmb_detection_workspace_bytes(32, 24564, 21) > 32 * 24564 * 32
    assert L.dspmb_target_workspace_bytes(64, 24564, 58, 21) >= 64 * 24564 * 4
    assert L.dspmb_nms_workspace_bytes(1000) >= 1000 * 16 * 8
    assert L.dspmb_set_libm_mode(0) == 0 and L.dspmb_set_libm_mode(1) == 1 and L.dspmb_set_libm_mode(-1) in (0, 1)
    old = L.dspmb_set_tuning(_lib.TUNE_NMS_MASK_ROWS, 100000)
    assert L.dspmb_set_tuning(_lib.TUNE_NMS_MASK_ROWS, old) == 320  # clamped to the compiled maximum
    assert L.dspmb_set_tuning(99, 1) == -1


def test_tuning_knobs_match_the_header():
    """Every DSPMB_TUNE_* index of include/dspmb.h has the same value under its Python name, the library accepts exactly
    DSPMB_NUM_TUNING knobs, and the defaults the header documents are the ones the library starts with."""
    import os, re
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "dspmb.h")).read()
    knobs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define DSPMB_TUNE_(\w+)\s+(\d+)", hdr)}
    num = int(re.search(r"#define DSPMB_NUM_TUNING\s+(\d+)", hdr).group(1))
    assert sorted(knobs.values()) == list(range(num))
    for name, idx in knobs.items():
        assert getattr(_lib, "TUNE_" + name) == idx, name
    L = _lib.lib()
    assert L.dspmb_set_tuning(num, 0) == -1
    defaults = {"DET_STREAM_VARIANT": 2, "DET_PIPELINE": 1, "TARGET_PIPELINE": 0, "DET_PREFETCH": 600, "TARGET_PREFETCH": 0,
                "DET_LEAN": 2, "TARGET_SHORTLIST": 1, "TARGET_PDL": 1, "DET_SORT_PDL": 1, "NMS_PDL": 1, "GRAPH_CACHE": 1}
    for name, want in defaults.items():
        old = L.dspmb_set_tuning(knobs[name], want)
        assert old == want, (name, old)


def test_ops_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    from dspnet_b200 import DspmbError, MultiBoxDetection, MultiBoxPrior
    from dspnet_b200.nms import cpu_nms
    with pytest.raises(DspmbError):
        MultiBoxPrior((1, 3, 4, 4), sizes=(0.1,), ratios=(1,))
    with pytest.raises(DspmbError):
        MultiBoxDetection(np.zeros((1, 3, 8), np.float32), np.zeros((1, 40), np.float32), np.zeros((1, 8, 4), np.float32))
    with pytest.raises(DspmbError):
        cpu_nms(np.zeros((4, 5), np.float32), 0.5)


def test_presets_match_reference_shape_facts():
    assert presets.num_anchors("ssd300") == 8732
    assert presets.num_anchors("ssd512") == 24564
    assert presets.num_anchors("ssd512_generic") == 24576
    assert presets.num_anchors("dspnet_cs") == 12264                      # utils.py:37 (1, 12264, 7)
    assert presets.PRESETS["dspnet_cs"].label_slots == 200                  # multi_solver.py:196
    assert [presets.anchors_per_location(fm) for fm in presets.PRESETS["ssd512"].maps] == [4, 6, 6, 6, 6, 4, 4]


def test_synthetic_inputs_are_per_image_deterministic():
    a = synth.cls_prob(2, 4, 21, 1000)
    b = synth.cls_prob(2, 2, 21, 1000, first_image=2)
    assert np.array_equal(a[2:], b)          # an image does not depend on the batch it sits in
    np.testing.assert_allclose(a.sum(axis=1), 1.0, rtol=1e-5)
    lab = synth.labels(2, 4, 58, 21)
    assert (lab[1] == -1).all() and (lab[2, :, 0] >= 0).all()   # the G = 0 and G = L edge images
    for img in lab:
        valid = img[:, 0] != -1
        g = int(valid.sum())
        assert valid[:g].all() and (img[g:] == -1).all()       # valid rows first, padding rows all -1
    d = synth.nms_boxes(3, 5000)
    assert len(np.unique(d[:, 4])) == 5000                       # tie-free scores


@pytest.mark.parametrize("batch,world", [(64, 8), (16, 8), (10, 4), (3, 8), (32, 1)])
def test_shard_slices_partition_the_batch(batch, world):
    covered = []
    for r in range(world):
        b, e = shard_slice(batch, world, r)
        assert 0 <= b <= e <= batch
        covered += list(range(b, e))
    assert covered == list(range(batch))
    sizes = [shard_slice(batch, world, r)[1] - shard_slice(batch, world, r)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def test_parameter_tuple_parsing():
    from dspnet_b200.ops import _tuple
    assert _tuple("(0.1,0.141)", "sizes") == (float(np.float32(0.1)), float(np.float32(0.141)))
    assert _tuple([1, 2, .5], "ratios") == (1.0, 2.0, 0.5)
    assert _tuple(0.3, "x") == (float(np.float32(0.3)),)
    assert _tuple("(-1.0, -1.0)", "steps") == (-1.0, -1.0)
