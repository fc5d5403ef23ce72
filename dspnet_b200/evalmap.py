"""GPU side of the reference's detection evaluation.

* ``postfilter``  -- multi_solver.py:419-432: keep rows with ``id >= 0`` and ``score > 0.25``, pad to 200 rows.
* ``match_flags`` -- the per-image TP / FP matching inside ``MApMetric.update`` (evaluate/eval_metric.py:113-160).
* ``MApMetric`` / ``VOC07MApMetric`` -- same constructor arguments, ``update(labels, preds)``, ``get()`` and
  ``reset()`` as evaluate/eval_metric.py:4-277.  The IoU matching runs on the GPU; the record bookkeeping and the
  AP integration (a few hundred numbers) stay on the host, in the reference's order of operations.

CUDA tensors or numpy arrays are accepted; there is no CPU path for the two kernels.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops import DspmbError, _require_cuda


def _dev(x, dtype=torch.float32):
    if torch.is_tensor(x):
        if not x.is_cuda:
            x = x.cuda()
        return x.to(dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda().to(dtype)


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def postfilter(out, max_rows=200, score_thresh=0.25, valid_count=None):
    """(B, A, 7) detection output -> (rows (B, max_rows, 7) padded with -1, counts (B,) int32), CUDA tensors."""
    _require_cuda()
    out = _dev(out)
    if out.dim() != 3 or out.shape[2] != 7:
        raise DspmbError(_lib.ERR_BAD_ARG, "postfilter: (batch, rows, 7) detection output expected")
    B, A = out.shape[0], out.shape[1]
    rows = torch.empty((B, max_rows, 7), dtype=torch.float32, device=out.device)
    counts = torch.empty((B,), dtype=torch.int32, device=out.device)
    vc = None if valid_count is None else _dev(valid_count, torch.int32)
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().dspmb_detection_postfilter_f32(
            out.data_ptr(), vc.data_ptr() if vc is not None else None, B, A, int(max_rows), float(score_thresh),
            rows.data_ptr(), counts.data_ptr(), _stream(out)))
    return rows, counts


def match_flags(labels, preds, ovp_thresh=0.5, use_difficult=False):
    """labels (B, L, >=5), preds (B, M, >=6) -> flags (B, M) int32 CUDA tensor: 0 not recorded, 1 TP, 2 FP."""
    _require_cuda()
    labels, preds = _dev(labels), _dev(preds)
    if labels.dim() != 3 or preds.dim() != 3 or labels.shape[0] != preds.shape[0]:
        raise DspmbError(_lib.ERR_BAD_ARG, "match_flags: labels (B, L, W) and preds (B, M, P) expected")
    B, L, W = labels.shape
    M, P = preds.shape[1], preds.shape[2]
    flags = torch.zeros((B, M), dtype=torch.int32, device=preds.device)
    with torch.cuda.device(preds.device):
        _lib.check(_lib.lib().dspmb_map_match_f32(labels.data_ptr(), B, L, W, preds.data_ptr(), M, P, float(ovp_thresh),
                                                  int(bool(use_difficult)), flags.data_ptr(), _stream(preds)))
    return flags


def _pr_curve(record, gt_count):
    """Cumulative recall / precision of one class from its (score, flag) records (eval_metric.py:196-207): records
    with flag 0 are dropped, the rest ordered by descending score."""
    flags = record[:, 1].astype(int)
    kept = record[flags != 0]
    kept_flags = kept[kept[:, 0].argsort()[::-1], 1].astype(int)
    tp, fp = np.cumsum(kept_flags == 1), np.cumsum(kept_flags == 2)
    recall = tp / float(gt_count) if gt_count > 0 else tp * 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        precision = tp.astype(float) / (tp + fp)
    return recall, precision


def _ap_area(recall, precision):
    """Area under the monotone precision envelope (eval_metric.py:209-231)."""
    r = np.concatenate(([0.0], recall, [1.0]))
    p = np.concatenate(([0.0], precision, [0.0]))
    p = np.maximum.accumulate(p[::-1])[::-1]            # envelope from the right
    steps = np.flatnonzero(r[1:] != r[:-1])
    return np.sum((r[steps + 1] - r[steps]) * p[steps + 1])


def _ap_voc07(recall, precision):
    """PASCAL VOC 07 eleven-point AP (eval_metric.py:254-277)."""
    ap = 0.0
    for t in np.arange(0.0, 1.1, 0.1):
        reached = recall >= t
        ap += (np.max(precision[reached]) if np.sum(reached) != 0 else 0) / 11.0
    return ap


class MApMetric(object):
    """Mean AP for detection with the interface of evaluate/eval_metric.py:4-247 (constructor arguments, ``update``,
    ``get``, ``reset``, the ``records`` / ``counts`` dictionaries); the IoU matching of ``update`` runs on the GPU."""

    _ap = staticmethod(_ap_area)

    def __init__(self, ovp_thresh=0.5, use_difficult=False, class_names=None, pred_idx=0):
        if class_names is not None:
            if not isinstance(class_names, (list, tuple)) or not all(isinstance(n, str) for n in class_names):
                raise AssertionError("must provide names as str")
        self.class_names = class_names
        self.name = "mAP" if class_names is None else list(class_names) + ["mAP"]
        self.num = None if class_names is None else len(class_names) + 1
        self.ovp_thresh, self.use_difficult, self.pred_idx = ovp_thresh, use_difficult, int(pred_idx)
        self.reset()

    def reset(self):
        self.records, self.counts = dict(), dict()
        self.num_inst = 0 if self.num is None else [0] * self.num
        self.sum_metric = 0.0 if self.num is None else [0.0] * self.num

    def _push(self, cid, rows, gt_count):
        if cid in self.records:
            self.records[cid] = np.vstack((self.records[cid], rows))
            self.counts[cid] += gt_count
        else:
            self.records[cid], self.counts[cid] = rows, gt_count

    def update(self, labels, preds):
        """labels: [ (B, L, 5 or 6) ], preds: list whose entry ``pred_idx`` is (B, M, >=6) -- like the reference."""
        lab_t, pred_t = _dev(labels[0]), _dev(preds[self.pred_idx])
        flags = match_flags(lab_t, pred_t, self.ovp_thresh, self.use_difficult).cpu().numpy()
        lab, pred = lab_t.cpu().numpy(), pred_t.cpu().numpy()
        count_easy_only = (not self.use_difficult) and lab.shape[2] >= 6          # eval_metric.py:156-159
        for image_labels, image_preds, image_flags in zip(lab, pred, flags):
            pcls, lcls = image_preds[:, 0].astype(int), image_labels[:, 0].astype(int)
            # predicted classes in order of first appearance (:118-124), then the classes that only occur in the
            # labels (:168-176); ids < 0 are padding on both sides
            pred_order = [int(c) for c in pcls[np.sort(np.unique(pcls, return_index=True)[1])] if c >= 0]
            for cid in pred_order:
                mine = pcls == cid
                rows = np.column_stack((image_preds[mine, 1].astype(np.float64), image_flags[mine].astype(np.float64)))
                rows = rows[rows[:, 1] > 0]
                gts = image_labels[lcls == cid]
                if rows.size > 0:
                    self._push(cid, rows, int(np.sum(gts[:, 5] < 1)) if count_easy_only else gts.shape[0])
            label_order = [int(c) for c in lcls[np.sort(np.unique(lcls, return_index=True)[1])]]
            for cid in label_order:
                if cid >= 0 and cid not in pred_order:
                    self._push(cid, np.zeros((1, 2), dtype=np.float64), int(np.sum(lcls == cid)))

    def get(self):
        aps = {k: self._ap(*_pr_curve(v, self.counts[k])) for k, v in self.records.items()}
        mean_ap = np.mean(list(aps.values()))
        if self.num is None:
            self.num_inst, self.sum_metric = 1, mean_ap
            return (self.name, self.sum_metric / self.num_inst)
        for k, ap in aps.items():
            if k < self.num - 1:
                self.sum_metric[k], self.num_inst[k] = ap, 1
        self.sum_metric[-1], self.num_inst[-1] = mean_ap, 1
        values = [x / y if y != 0 else float("nan") for x, y in zip(self.sum_metric, self.num_inst)]
        return (["%s" % n for n in self.name], values)


class VOC07MApMetric(MApMetric):
    """11-point PASCAL VOC 07 AP -- evaluate/eval_metric.py:249-277."""

    _ap = staticmethod(_ap_voc07)
