"""The forward of the SSD training graph behind MultiBoxTarget and the reference's training metric, fused on the GPU
(SURVEY.md section 8f, row f2).

``multibox_training_outputs`` mirrors symbol/symbol_builder.py:82-88: ``cls_prob`` (forward of
``SoftmaxOutput(cls_preds, cls_target, ignore_label=-1, use_ignore=True, multi_output=True, normalization='valid')``,
i.e. the channel softmax) and ``loc_loss`` (``MakeLoss(smooth_l1(loc_target_mask * (loc_preds - loc_target), 1.0))``).
``MultiBoxMetric`` mirrors train/metric.py:7-75 (same ``update(labels, preds)`` / ``get()`` interface, same two
numbers), but ``update`` also accepts the per-image statistics the kernel produced, so that the (B, C, A) probability
tensor never has to leave the device.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import DspmbError

__all__ = ["multibox_training_outputs", "MultiBoxMetric"]

_ws = {}


def multibox_training_outputs(cls_preds, loc_preds, loc_target, loc_mask, cls_target, eps=1e-8, want_cls_prob=True,
                              want_loc_loss=True):
    """cls_preds (B, C, A), loc_preds / loc_target / loc_mask (B, A*5), cls_target (B, A): float32 CUDA tensors.
    Returns (cls_prob or None, loc_loss or None, stats) with stats (B, 4) float64 on the device:
    [valid_count, cross-entropy sum, smooth-L1 sum, number of loc_loss elements > 0]."""
    ts = (cls_preds, loc_preds, loc_target, loc_mask, cls_target)
    if not all(isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 for t in ts):
        raise DspmbError(_lib.ERR_BAD_ARG, "multibox_training_outputs: float32 CUDA tensors expected (no CPU fallback)")
    cls_preds, loc_preds, loc_target, loc_mask, cls_target = (t.contiguous() for t in ts)
    if cls_preds.dim() != 3:
        raise DspmbError(_lib.ERR_BAD_ARG, "multibox_training_outputs: cls_preds is (B, C, A)")
    B, C, A = cls_preds.shape
    if any(tuple(t.shape) != (B, A * 5) for t in (loc_preds, loc_target, loc_mask)) or tuple(cls_target.shape) != (B, A):
        raise DspmbError(_lib.ERR_BAD_ARG, "multibox_training_outputs: loc tensors are (B, A*5), cls_target is (B, A)")
    dev = cls_preds.device
    prob = torch.empty_like(cls_preds) if want_cls_prob else None
    loss = torch.empty_like(loc_preds) if want_loc_loss else None
    stats = torch.empty((B, 4), dtype=torch.float64, device=dev)
    L = _lib.lib()
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        nbytes = max(int(L.dspmb_multibox_loss_workspace_bytes(B, A)), 256)
        key = (dev.index, stream)
        ws = _ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _ws[key] = ws
        _lib.check(L.dspmb_multibox_loss_f32(p(cls_preds), p(loc_preds), p(loc_target), p(loc_mask), p(cls_target), p(prob),
                                             p(loss), p(stats), B, A, C, float(eps), p(ws), ws.numel(),
                                             ctypes.c_void_p(stream)))
    return prob, loss, stats


class MultiBoxMetric:
    """train/metric.py:7-75: 'CrossEntropy' and 'SmoothL1', each a running sum divided by the running valid count."""

    def __init__(self, eps=1e-8):
        self.eps = eps
        self.num = 2
        self.name = ['CrossEntropy', 'SmoothL1']
        self.reset()

    def reset(self):
        self.num_inst = [0] * self.num
        self.sum_metric = [0.0] * self.num

    def update_from_stats(self, stats):
        """stats: the (B, 4) tensor of multibox_training_outputs (32 bytes per image cross PCIe)."""
        s = stats.sum(dim=0).cpu().numpy() if isinstance(stats, torch.Tensor) else np.asarray(stats).sum(axis=0)
        self.sum_metric[0] += float(s[1])
        self.num_inst[0] += int(s[0])
        self.sum_metric[1] += float(s[2])
        self.num_inst[1] += int(s[0])

    def update(self, labels, preds):
        """The reference's signature: preds = [cls_prob (B,C,A), loc_loss (B,A*5), cls_label (B,A)] (labels unused,
        as in the reference).  CUDA tensors are reduced on the device; numpy arrays go through the reference's own
        arithmetic."""
        cls_prob, loc_loss, cls_label = preds[0], preds[1], preds[2]
        if isinstance(cls_prob, torch.Tensor):
            lab = cls_label.reshape(cls_label.shape[0], -1)
            valid = lab >= 0
            idx = lab.clamp(min=0).long().unsqueeze(1)
            prob = torch.gather(cls_prob, 1, idx).squeeze(1)
            ce = -torch.log(prob + np.float32(self.eps))
            self.sum_metric[0] += float(ce[valid].double().sum().item())
            n = int(valid.sum().item())
            self.num_inst[0] += n
            self.sum_metric[1] += float(loc_loss.double().sum().item())
            self.num_inst[1] += n
            return
        cls_prob, loc_loss, cls_label = (np.asarray(x) for x in (cls_prob, loc_loss, cls_label))
        valid_count = np.sum(cls_label >= 0)
        label = cls_label.flatten()
        mask = np.where(label >= 0)[0]
        indices = np.int64(label[mask])
        prob = cls_prob.transpose((0, 2, 1)).reshape((-1, cls_prob.shape[1]))
        prob = prob[mask, indices]
        self.sum_metric[0] += (-np.log(prob + self.eps)).sum()
        self.num_inst[0] += valid_count
        self.sum_metric[1] += np.sum(loc_loss)
        self.num_inst[1] += valid_count

    def get(self):
        names = ['%s' % self.name[i] for i in range(self.num)]
        values = [x / y if y != 0 else float('nan') for x, y in zip(self.sum_metric, self.num_inst)]
        return (names, values)
