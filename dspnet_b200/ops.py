"""The three multibox operators with the reference's operator signatures.

Mirrors ``mx.contrib.{ndarray,symbol}.MultiBoxPrior / MultiBoxTarget / MultiBoxDetection`` as registered by
operator/multibox_prior.cc:96-99, multibox_target.cc:308-313 and multibox_detection.cc:194-199: same positional
inputs, same keyword names and defaults (the dmlc Param structs of operator/multibox_*-inl.h), same output shapes
(the InferShape methods) and the CPU operators' semantics.  Inputs are CUDA ``torch.Tensor`` (fp32, contiguous);
numpy arrays are accepted as HOST buffers: they are copied to the current CUDA device, the op runs there and the
results come back as numpy arrays (the host<->device round trip bench.py times as ``e2e``).

PyTorch is only the device-memory / stream plumbing; all compute is in libdspmb.so (hand-written sm_100a CUDA,
called through the C ABI of include/dspmb.h).  There is no CPU implementation in this package.
"""
import ast
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import DspmbError  # noqa: F401  (re-export)

__all__ = ["MultiBoxPrior", "MultiBoxTarget", "MultiBoxDetection", "MultiBoxDetectionFromHeads", "multibox_prior_concat",
           "DspmbError"]

_workspaces = {}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _workspace(kind, nbytes, device):
    """Scratch the reference would get from ResourceRequest::kTempSpace; cached per (op, device, stream)."""
    key = (kind, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _tuple(value, name):
    """Tuple parameters arrive as Python sequences or as the strings the reference builds
    (symbol/common.py:236-239,386-389: "(0.1,0.141)")."""
    if isinstance(value, str):
        value = ast.literal_eval(value)
    if isinstance(value, (int, float)):
        value = (value,)
    out = tuple(float(np.float32(v)) for v in value)
    if len(out) == 0:
        raise DspmbError(_lib.ERR_BAD_ARG, "%s must not be empty" % name)
    return out


def _require_cuda():
    if not torch.cuda.is_available():
        raise DspmbError(_lib.ERR_CUDA, "no CUDA device: dspnet_b200 has no CPU fallback")


def _as_device(x, name):
    """Returns (cuda fp32 contiguous tensor, came_from_host)."""
    if isinstance(x, np.ndarray):
        _require_cuda()
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        return t.cuda(non_blocking=True), True
    if not isinstance(x, torch.Tensor):
        raise TypeError("%s: expected torch.Tensor or numpy.ndarray, got %r" % (name, type(x)))
    if not x.is_cuda:
        _require_cuda()
        return x.to(dtype=torch.float32).contiguous().cuda(non_blocking=True), True
    if x.dtype != torch.float32:
        raise DspmbError(_lib.ERR_BAD_ARG, "%s: only float32 is supported (the only dtype the reference exercises)" % name)
    return x.contiguous(), False


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _back(outs, to_host):
    if not to_host:
        return outs
    return [o.cpu().numpy() for o in outs]


def MultiBoxPrior(data, sizes=(1.0,), ratios=(1.0,), clip=False, steps=(-1.0, -1.0), offsets=(0.5, 0.5), out=None):
    """Generate prior (anchor) boxes from the spatial shape of ``data`` -- operator/multibox_prior-inl.h:59-77
    (parameters), :171-193 (shape: ``(1, H*W*(len(sizes)+len(ratios)-1), 4)``), multibox_prior.cc:29-71.

    ``data`` may be any >= 4-D tensor (only ``shape[2:4]`` is read) or a plain ``(..., H, W)`` shape tuple.
    """
    shape = tuple(data) if isinstance(data, (tuple, list, torch.Size)) else tuple(data.shape)
    if len(shape) < 4:
        raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxPrior: input data should be 4D: batch-channel-y-x")
    sizes, ratios = _tuple(sizes, "sizes"), _tuple(ratios, "ratios")
    steps, offsets = _tuple(steps, "steps"), _tuple(offsets, "offsets")
    if len(steps) != 2 or len(offsets) != 2:
        raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxPrior: steps and offsets must be (y, x) pairs")
    _require_cuda()
    h, w = int(shape[2]), int(shape[3])
    device = data.device if isinstance(data, torch.Tensor) and data.is_cuda else torch.device("cuda", torch.cuda.current_device())
    n = h * w * (len(sizes) + len(ratios) - 1)
    if out is None:
        out = torch.empty((1, n, 4), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().dspmb_prior_f32(_ptr(out), h, w, _lib.float_array(sizes), len(sizes),
                                              _lib.float_array(ratios), len(ratios), steps[0], steps[1], offsets[0],
                                              offsets[1], int(bool(clip)), _stream()))
    return out


def multibox_prior_concat(feature_shapes, sizes, ratios, steps=None, offsets=(0.5, 0.5), clip=False, device=None):
    """The anchor branch of ``multibox_layer`` / ``multitask_layer`` (symbol/common.py:415-432): one MultiBoxPrior
    per feature map, flattened and concatenated to ``(1, A, 4)`` -- done here in a single launch.

    feature_shapes: [(H, W), ...]; sizes/ratios: per-map lists; steps: per-map scalar (used for y and x) or None/[]
    for the automatic 1/H, 1/W steps, exactly as the reference passes them.
    """
    _require_cuda()
    n = len(feature_shapes)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    sizes = [_tuple(s, "sizes") for s in sizes]
    ratios = [_tuple(r, "ratios") for r in ratios]
    if len(sizes) != n or len(ratios) != n:
        raise DspmbError(_lib.ERR_BAD_ARG, "multibox_prior_concat: need one sizes/ratios list per feature map")
    step_pairs = []
    for k in range(n):
        s = float(np.float32(steps[k])) if steps else -1.0
        step_pairs += [s, s]
    offs = list(_tuple(offsets, "offsets")) * n
    total = sum(h * w * (len(s) + len(r) - 1) for (h, w), s, r in zip(feature_shapes, sizes, ratios))
    out = torch.empty((1, total, 4), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().dspmb_prior_multi_f32(
            _ptr(out), n, _lib.int_array([h for h, _ in feature_shapes]), _lib.int_array([w for _, w in feature_shapes]),
            _lib.float_array([v for s in sizes for v in s]), _lib.int_array([len(s) for s in sizes]),
            _lib.float_array([v for r in ratios for v in r]), _lib.int_array([len(r) for r in ratios]),
            _lib.float_array(step_pairs), _lib.float_array(offs), int(bool(clip)), _stream()))
    return out


def MultiBoxTarget(anchor, label, cls_pred, overlap_threshold=0.5, ignore_label=-1.0, negative_mining_ratio=-1.0,
                   negative_mining_thresh=0.5, minimum_negative_samples=0, variances=(0.1, 0.1, 0.2, 0.2),
                   return_match=False, return_stats=False, check=True):
    """Compute multibox training targets -- operator/multibox_target-inl.h:59-80 (parameters), :213-238 (shapes),
    :89-171 + multibox_target.cc:72-284 (CPU semantics).

    anchor (1, A, 4), label (B, L, 6) ``[cls, xmin, ymin, xmax, ymax, dist]`` padded with -1 rows,
    cls_pred (B, C, A)  ->  [loc_target (B, A*5), loc_mask (B, A*5), cls_target (B, A)].
    ``return_match`` appends the (B, A) int32 matched-gt index (-1 for non-positives), ``return_stats`` the (B, 4)
    int32 [num_valid_gt, num_positive, num_negative, num_bipartite].  ``check`` synchronises and raises for the
    data-dependent CHECKs of the reference (label padding, too few mining candidates).
    """
    anchor, h0 = _as_device(anchor, "anchor")
    label, h1 = _as_device(label, "label")
    cls_pred, h2 = _as_device(cls_pred, "cls_pred")
    to_host = h0 and h1 and h2
    if anchor.dim() != 3 or anchor.shape[0] != 1 or anchor.shape[2] != 4 or anchor.shape[1] <= 0:
        raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxTarget: anchor should be a batch-shared (1, N, 4) tensor")
    if label.dim() != 3 or label.shape[1] <= 0 or label.shape[2] != 6:
        raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxTarget: label should be (batch, num_labels, 6) "
                                           "[cls-xmin-ymin-xmax-ymax-dist]")
    if cls_pred.dim() != 3 or cls_pred.shape[2] != anchor.shape[1] or cls_pred.shape[0] != label.shape[0]:
        raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxTarget: cls_pred should be (batch, num_classes, num_anchors)")
    var = _tuple(variances, "variances")
    if len(var) != 4:
        raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxTarget: variances must have 4 entries")
    B, L, W = label.shape
    A, C = anchor.shape[1], cls_pred.shape[1]
    dev = cls_pred.device
    loc_target = torch.empty((B, A * 5), dtype=torch.float32, device=dev)
    loc_mask = torch.empty((B, A * 5), dtype=torch.float32, device=dev)
    cls_target = torch.empty((B, A), dtype=torch.float32, device=dev)
    match = torch.empty((B, A), dtype=torch.int32, device=dev) if return_match else None
    stats = torch.empty((B, 4), dtype=torch.int32, device=dev) if return_stats else None
    L_ = _lib.lib()
    with torch.cuda.device(dev):
        nbytes = L_.dspmb_target_workspace_bytes(B, A, L, C)
        ws = _workspace("target", nbytes, dev)
        _lib.check(L_.dspmb_target_f32(_ptr(anchor), _ptr(label), _ptr(cls_pred), _ptr(loc_target), _ptr(loc_mask),
                                       _ptr(cls_target), B, A, L, W, C, overlap_threshold, ignore_label,
                                       negative_mining_ratio, negative_mining_thresh, int(minimum_negative_samples),
                                       _lib.float_array(var), _ptr(match), _ptr(stats), _ptr(ws), ws.numel(),
                                       _stream()))
        if check and B > 0:
            _lib.check(L_.dspmb_status(_ptr(ws), _stream()))
    outs = [loc_target, loc_mask, cls_target]
    if return_match:
        outs.append(match)
    if return_stats:
        outs.append(stats)
    return _back(outs, to_host)


def MultiBoxDetection(cls_prob, loc_pred, anchor, clip=True, threshold=0.01, background_id=0, nms_threshold=0.5,
                      force_suppress=False, variances=(0.1, 0.1, 0.2, 0.2), nms_topk=-1, return_valid_count=False):
    """Convert multibox predictions to detections -- operator/multibox_detection-inl.h:47-72 (parameters; like the
    reference, ``background_id`` is accepted and unused and there is no ``keep_topk``), :149-171 (shapes),
    :81-107 + multibox_detection.cc:53-169 (CPU semantics).

    cls_prob (B, C, A), loc_pred (B, A*5), anchor (1, A, 4)  ->  (B, A, 7) rows ``[id, score, xmin, ymin, xmax,
    ymax, dist]``; suppressed rows have id -1, unused rows are all -1.
    """
    cls_prob, h0 = _as_device(cls_prob, "cls_prob")
    loc_pred, h1 = _as_device(loc_pred, "loc_pred")
    anchor, h2 = _as_device(anchor, "anchor")
    to_host = h0 and h1 and h2
    if cls_prob.dim() != 3 or loc_pred.dim() != 2 or anchor.dim() != 3:
        raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxDetection: inputs are [cls_prob (B,C,A), loc_pred (B,A*5), anchor (1,A,4)]")
    B, C, A = cls_prob.shape
    if anchor.shape[1] != A or anchor.shape[2] != 4 or A <= 0:
        raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxDetection: number of anchors mismatch")
    if loc_pred.shape[0] != B or loc_pred.shape[1] != A * 5:
        raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxDetection: # anchors mismatch with # loc (5 per anchor)")
    var = _tuple(variances, "variances")
    if len(var) != 4:
        raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxDetection: variance size must be 4")
    dev = cls_prob.device
    out = torch.empty((B, A, 7), dtype=torch.float32, device=dev)
    valid = torch.empty((B,), dtype=torch.int32, device=dev) if return_valid_count else None
    L_ = _lib.lib()
    with torch.cuda.device(dev):
        nbytes = L_.dspmb_detection_workspace_bytes(B, A, C)
        ws = _workspace("detection", nbytes, dev)
        _lib.check(L_.dspmb_detection_f32(_ptr(cls_prob), _ptr(loc_pred), _ptr(anchor), _ptr(out), B, A, C, threshold,
                                          int(bool(clip)), _lib.float_array(var), nms_threshold,
                                          int(bool(force_suppress)), int(nms_topk), _ptr(valid), _ptr(ws), ws.numel(),
                                          _stream()))
    outs = [out] + ([valid] if return_valid_count else [])
    outs = _back(outs, to_host)
    return outs if return_valid_count else outs[0]


def MultiBoxDetectionFromHeads(cls_heads, loc_heads, anchor, num_classes, clip=True, threshold=0.01, background_id=0,
                               nms_threshold=0.5, force_suppress=False, variances=(0.1, 0.1, 0.2, 0.2), nms_topk=-1,
                               return_valid_count=False):
    """The inference tail of the SSD graph in one operator -- what
    symbol/symbol_builder.py:156-165 builds from multibox_layer's per-scale heads (symbol/common.py:399-432):
    transpose / Flatten / Concat / Reshape / transpose, ``SoftmaxActivation(mode='channel')`` and MultiBoxDetection.

    cls_heads[k] (B, na_k*C, H_k, W_k) and loc_heads[k] (B, na_k*5, H_k, W_k) are the conv outputs (CUDA tensors, NCHW),
    anchor (1, A, 4) the concatenated MultiBoxPrior boxes.  Returns (B, A, 7) like MultiBoxDetection; the (B, C, A)
    probability tensor is never materialised.  Built for C = 21 and C = 9."""
    if len(cls_heads) != len(loc_heads) or not cls_heads:
        raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxDetectionFromHeads: one class head and one loc head per scale")
    C = int(num_classes)
    cls_heads = [h.contiguous() for h in cls_heads]
    loc_heads = [h.contiguous() for h in loc_heads]
    dev = cls_heads[0].device
    B = cls_heads[0].shape[0]
    hw, na = [], []
    for ch, lh in zip(cls_heads, loc_heads):
        if not (ch.is_cuda and lh.is_cuda and ch.dtype == torch.float32 and lh.dtype == torch.float32 and ch.dim() == 4
                and lh.dim() == 4):
            raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxDetectionFromHeads: heads must be 4-D float32 CUDA tensors")
        n = ch.shape[1] // C
        if ch.shape[1] != n * C or lh.shape[1] != n * 5 or ch.shape[0] != B or lh.shape[0] != B or ch.shape[2:] != lh.shape[2:]:
            raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxDetectionFromHeads: head shapes do not match (B, na*C, H, W) / (B, na*5, H, W)")
        hw += [int(ch.shape[2]), int(ch.shape[3])]
        na.append(int(n))
    anchor, _ = _as_device(anchor, "anchor")
    A = anchor.shape[1]
    var = _tuple(variances, "variances")
    if len(var) != 4:
        raise DspmbError(_lib.ERR_BAD_ARG, "MultiBoxDetection: variance size must be 4")
    out = torch.empty((B, A, 7), dtype=torch.float32, device=dev)
    valid = torch.empty((B,), dtype=torch.int32, device=dev) if return_valid_count else None
    L_ = _lib.lib()
    k = len(cls_heads)
    cp = (ctypes.c_void_p * k)(*[h.data_ptr() for h in cls_heads])
    lp = (ctypes.c_void_p * k)(*[h.data_ptr() for h in loc_heads])
    hw_a, na_a = _lib.int_array(hw), _lib.int_array(na)
    with torch.cuda.device(dev):
        nbytes = L_.dspmb_detection_heads_workspace_bytes(B, A, C, hw_a, na_a, k)
        ws = _workspace("detection_heads", max(int(nbytes), 256), dev)
        _lib.check(L_.dspmb_detection_heads_f32(cp, lp, hw_a, na_a, k, _ptr(anchor), _ptr(out), B, A, C, threshold,
                                                int(bool(clip)), _lib.float_array(var), nms_threshold,
                                                int(bool(force_suppress)), int(nms_topk), _ptr(valid), _ptr(ws),
                                                ws.numel(), _stream()))
    return (out, valid) if return_valid_count else out
