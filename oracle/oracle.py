"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/multibox_oracle.cc).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  The product package (dspnet_b200/) never does.

Functions mirror the reference operator names and keyword arguments
(operator/multibox_{prior,target,detection}-inl.h Param structs) and work on numpy arrays.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_multibox.so")
_lib = None

ERRORS = {
    0: "ok",
    -1: "bad argument",
    -2: "label padding row is not all -1 (multibox_target.cc:98-101)",
    -3: "fewer mining candidates than num_negative (multibox_target.cc:236)",
    -4: "negative_mining_thresh must be > 0 (multibox_target.cc:184)",
}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__("oracle: %s (code %d)" % (ERRORS.get(code, "?"), code))
        self.code = code


def build(force=False):
    """Compile the restatement with the flags SURVEY.md section 8c prescribes."""
    src = os.path.join(_HERE, "multibox_oracle.cc")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_multibox.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.oracle_multibox_prior.argtypes = [fp, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int, fp, ctypes.c_int,
                                            ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                            ctypes.c_int]
        L.oracle_multibox_prior.restype = ctypes.c_int
        L.oracle_target_iou.argtypes = [fp, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int, fp]
        L.oracle_target_iou.restype = None
        L.oracle_multibox_target.argtypes = [fp, fp, fp, fp, fp, fp] + [ctypes.c_int] * 5 + [ctypes.c_float] * 4 + [
            ctypes.c_int, fp, ip, fp, ctypes.POINTER(ctypes.c_int8), ip, ctypes.c_int]
        L.oracle_multibox_target.restype = ctypes.c_int
        L.oracle_multibox_detection.argtypes = [fp, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_float, ctypes.c_int, fp, ctypes.c_float, ctypes.c_int,
                                                ctypes.c_int, ip, ctypes.c_int]
        L.oracle_multibox_detection.restype = ctypes.c_int
        L.oracle_cpu_nms.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                     ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
        L.oracle_cpu_nms.restype = ctypes.c_int
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_bbox_overlaps.argtypes = [dp, ctypes.c_int, dp, ctypes.c_int, dp]
        L.oracle_bbox_overlaps.restype = None
        L.oracle_softmax_channel.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_softmax_channel.restype = None
        L.oracle_expf_array.argtypes = [fp, fp, ctypes.c_long]
        L.oracle_logf_array.argtypes = [fp, fp, ctypes.c_long]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, ty=ctypes.c_float):
    return a.ctypes.data_as(ctypes.POINTER(ty)) if a is not None else None


def multibox_prior(in_height, in_width, sizes=(1.0,), ratios=(1.0,), clip=False, steps=(-1.0, -1.0),
                   offsets=(0.5, 0.5)):
    """operator/multibox_prior.cc:29-71; returns (1, H*W*(S+R-1), 4) like InferShape (-inl.h:171-193)."""
    sizes = _f32(sizes)
    ratios = _f32(ratios)
    n = in_height * in_width * (len(sizes) + len(ratios) - 1)
    out = np.empty((1, n, 4), np.float32)
    rc = lib().oracle_multibox_prior(_p(out), in_height, in_width, _p(sizes), len(sizes), _p(ratios), len(ratios),
                                     float(steps[0]), float(steps[1]), float(offsets[0]), float(offsets[1]),
                                     int(bool(clip)))
    if rc:
        raise OracleError(rc)
    return out


def target_iou(anchors, gts):
    anchors = _f32(anchors).reshape(-1, 4)
    gts = _f32(gts).reshape(-1, 4)
    out = np.empty((anchors.shape[0], gts.shape[0]), np.float32)
    lib().oracle_target_iou(_p(anchors), anchors.shape[0], _p(gts), gts.shape[0], 4, _p(out))
    return out


def multibox_target(anchor, label, cls_pred, overlap_threshold=0.5, ignore_label=-1.0, negative_mining_ratio=-1.0,
                    negative_mining_thresh=0.5, minimum_negative_samples=0, variances=(0.1, 0.1, 0.2, 0.2),
                    debug=False, nthreads=1):
    """operator/multibox_target-inl.h:89-171 + multibox_target.cc:72-284.

    Returns [loc_target (B, A*5), loc_mask (B, A*5), cls_target (B, A)]; with debug=True also a dict with
    match_gt, match_iou, anchor_flags, stats (num_valid_gt, num_positive, num_negative, num_bipartite).
    """
    anchor = _f32(anchor)
    label = _f32(label)
    cls_pred = _f32(cls_pred)
    assert anchor.ndim == 3 and anchor.shape[0] == 1 and anchor.shape[2] == 4
    assert label.ndim == 3 and cls_pred.ndim == 3 and cls_pred.shape[2] == anchor.shape[1]
    B, L, W = label.shape
    A = anchor.shape[1]
    C = cls_pred.shape[1]
    var = _f32(variances)
    loc_target = np.empty((B, A * 5), np.float32)
    loc_mask = np.empty((B, A * 5), np.float32)
    cls_target = np.empty((B, A), np.float32)
    match_gt = np.empty((B, A), np.int32) if debug else None
    match_iou = np.empty((B, A), np.float32) if debug else None
    flags = np.empty((B, A), np.int8) if debug else None
    stats = np.zeros((B, 4), np.int32)
    rc = lib().oracle_multibox_target(_p(anchor), _p(label), _p(cls_pred), _p(loc_target), _p(loc_mask),
                                      _p(cls_target), B, A, L, W, C, overlap_threshold, ignore_label,
                                      negative_mining_ratio, negative_mining_thresh, minimum_negative_samples,
                                      _p(var), _p(match_gt, ctypes.c_int32), _p(match_iou),
                                      _p(flags, ctypes.c_int8), _p(stats, ctypes.c_int32), nthreads)
    if rc:
        raise OracleError(rc)
    outs = [loc_target, loc_mask, cls_target]
    if debug:
        return outs, dict(match_gt=match_gt, match_iou=match_iou, anchor_flags=flags, stats=stats)
    return outs


def multibox_detection(cls_prob, loc_pred, anchor, clip=True, threshold=0.01, background_id=0, nms_threshold=0.5,
                       force_suppress=False, variances=(0.1, 0.1, 0.2, 0.2), nms_topk=-1, return_valid=False,
                       nthreads=1):
    """operator/multibox_detection-inl.h:81-107 + multibox_detection.cc:53-169.  Returns (B, A, 7)."""
    cls_prob = _f32(cls_prob)
    loc_pred = _f32(loc_pred)
    anchor = _f32(anchor)
    B, C, A = cls_prob.shape
    assert loc_pred.shape == (B, A * 5) and anchor.shape == (1, A, 4)
    var = _f32(variances)
    out = np.empty((B, A, 7), np.float32)
    valid = np.zeros((B,), np.int32)
    rc = lib().oracle_multibox_detection(_p(cls_prob), _p(loc_pred), _p(anchor), _p(out), B, A, C, threshold,
                                         int(bool(clip)), _p(var), nms_threshold, int(bool(force_suppress)),
                                         nms_topk, _p(valid, ctypes.c_int32), nthreads)
    if rc:
        raise OracleError(rc)
    return (out, valid) if return_valid else out


def head_layout(cls_heads, loc_heads, num_classes):
    """The layout shuffles of multibox_layer (symbol/common.py:399-432), numpy only: per scale the conv outputs
    (B, na*C, H, W) / (B, na*5, H, W) are transposed to NHWC and flattened, the scales concatenated, the class tensor
    reshaped to (B, A, C) and transposed to (B, C, A).  Returns (cls_preds (B, C, A), loc_preds (B, A*5))."""
    cls = [np.ascontiguousarray(np.transpose(_f32(h), (0, 2, 3, 1))).reshape(h.shape[0], -1) for h in cls_heads]
    loc = [np.ascontiguousarray(np.transpose(_f32(h), (0, 2, 3, 1))).reshape(h.shape[0], -1) for h in loc_heads]
    cls_preds = np.concatenate(cls, axis=1)
    cls_preds = cls_preds.reshape(cls_preds.shape[0], -1, num_classes)
    cls_preds = np.ascontiguousarray(np.transpose(cls_preds, (0, 2, 1)))
    return cls_preds, np.ascontiguousarray(np.concatenate(loc, axis=1))


def softmax_channel(cls_preds):
    """SoftmaxActivation(mode='channel') / SoftmaxOutput forward on (B, C, A): the softmax of multibox_target.cc:220-231
    per position (see oracle_softmax_channel; MXNet's own kernel is not in the reference tree)."""
    x = _f32(cls_preds)
    B, C, A = x.shape
    out = np.empty_like(x)
    lib().oracle_softmax_channel(_p(x), _p(out), B, C, A)
    return out


def multibox_detection_from_heads(cls_heads, loc_heads, anchor, num_classes, **kw):
    """symbol/symbol_builder.py:156-165: multibox_layer -> SoftmaxActivation(channel) -> MultiBoxDetection."""
    cls_preds, loc_preds = head_layout(cls_heads, loc_heads, num_classes)
    return multibox_detection(softmax_channel(cls_preds), loc_preds, anchor, **kw)


def smooth_l1(x, sigma=1.0):
    """mx.symbol.smooth_l1(scalar=sigma), element-wise in fp32: 0.5 (sigma x)^2 if |x| < 1/sigma^2 else |x| - 0.5/sigma^2
    (MXNet mshadow_op::smooth_l1_loss; documented formula, MXNet itself is not in the reference tree)."""
    x = _f32(x)
    s2 = np.float32(sigma) * np.float32(sigma)
    ax = np.abs(x)
    quad = np.float32(0.5) * (x * x) * s2
    lin = ax - np.float32(0.5) / s2
    return np.where(ax < np.float32(1.0) / s2, quad, lin).astype(np.float32)


def multibox_training_outputs(cls_preds, loc_preds, loc_target, loc_mask, cls_target, eps=1e-8):
    """Forward of the training graph after MultiBoxTarget (symbol/symbol_builder.py:82-88) and the statistics of
    MultiBoxMetric.update (train/metric.py:27-46):
        cls_prob  = SoftmaxOutput(cls_preds, cls_target, ignore_label=-1, multi_output)   -> channel softmax
        loc_loss  = MakeLoss(smooth_l1(loc_mask * (loc_preds - loc_target), scalar=1))    -> element-wise
        valid     = #(cls_target >= 0);  ce = sum -log(prob[label] + eps) over those;  sl1 = sum(loc_loss)
    Sums are accumulated in float64 here (numpy's float32 pairwise order is an implementation detail of numpy)."""
    prob = softmax_channel(cls_preds)
    diff = _f32(loc_mask) * (_f32(loc_preds) - _f32(loc_target))
    loc_loss = smooth_l1(diff, 1.0)
    label = _f32(cls_target)
    B, C, A = prob.shape
    stats = np.zeros((B, 3), np.float64)
    for b in range(B):
        m = np.nonzero(label[b] >= 0)[0]
        p = prob[b, label[b, m].astype(np.int64), m]
        stats[b, 0] = m.size
        stats[b, 1] = (-np.log(p + np.float32(eps))).astype(np.float64).sum()
        stats[b, 2] = loc_loss[b].astype(np.float64).sum()
    return prob, loc_loss, stats


def cpu_nms(dets, thresh, mode="cpu"):
    """cython/cpu_nms.pyx:17-68 (mode='cpu', suppress iff double(ovr) >= thresh) or the strict-greater rule
    of cython/nms_kernel.cu:71 and detect/nms.py:55 (mode='gpu').  Returns kept indices in score order."""
    dets = _f32(dets)
    n, dim = dets.shape
    order = np.ascontiguousarray(dets[:, 4].argsort()[::-1], dtype=np.int64)
    keep = np.empty((n,), np.int64)
    k = lib().oracle_cpu_nms(_p(dets), n, dim, _p(order, ctypes.c_int64), float(thresh),
                             0 if mode == "cpu" else 1, _p(keep, ctypes.c_int64))
    return [int(i) for i in keep[:k]]


def bbox_overlaps(boxes, query_boxes):
    """cython/bbox.pyx:15-55 (bbox_overlaps_cython): (N, 4) x (K, 4) float64 -> (N, K) float64."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float64)
    query_boxes = np.ascontiguousarray(query_boxes, dtype=np.float64)
    out = np.empty((boxes.shape[0], query_boxes.shape[0]), np.float64)
    lib().oracle_bbox_overlaps(_p(boxes, ctypes.c_double), boxes.shape[0], _p(query_boxes, ctypes.c_double),
                               query_boxes.shape[0], _p(out, ctypes.c_double))
    return out


def py_nms(dets, thresh):
    """detect/nms.py:24-58 restated with numpy (keeps ovr <= thresh)."""
    x1, y1, x2, y2, scores = (dets[:, i] for i in range(5))
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = scores.argsort()[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        rest = order[1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(0.0, xx2 - xx1 + 1)
        h = np.maximum(0.0, yy2 - yy1 + 1)
        inter = w * h
        ovr = inter / (areas[i] + areas[rest] - inter)
        order = rest[np.where(ovr <= thresh)[0]]
    return keep


def expf(x):
    x = _f32(x).ravel()
    y = np.empty_like(x)
    lib().oracle_expf_array(_p(x), _p(y), x.size)
    return y


def logf(x):
    x = _f32(x).ravel()
    y = np.empty_like(x)
    lib().oracle_logf_array(_p(x), _p(y), x.size)
    return y
