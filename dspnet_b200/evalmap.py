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


class MApMetric(object):
    """Mean AP for detection -- evaluate/eval_metric.py:4-247 with the matching on the GPU."""

    def __init__(self, ovp_thresh=0.5, use_difficult=False, class_names=None, pred_idx=0):
        if class_names is None:
            self.num = None
            self.name = "mAP"
        else:
            assert isinstance(class_names, (list, tuple))
            for name in class_names:
                assert isinstance(name, str), "must provide names as str"
            self.name = list(class_names) + ["mAP"]
            self.num = len(class_names) + 1
        self.ovp_thresh = ovp_thresh
        self.use_difficult = use_difficult
        self.class_names = class_names
        self.pred_idx = int(pred_idx)
        self.reset()

    def reset(self):
        if self.num is None:
            self.num_inst = 0
            self.sum_metric = 0.0
        else:
            self.num_inst = [0] * self.num
            self.sum_metric = [0.0] * self.num
        self.records = dict()
        self.counts = dict()

    def update(self, labels, preds):
        """labels: [ (B, L, 5 or 6) ], preds: list whose entry ``pred_idx`` is (B, M, >=6) -- like the reference."""
        lab_t, pred_t = _dev(labels[0]), _dev(preds[self.pred_idx])
        flags = match_flags(lab_t, pred_t, self.ovp_thresh, self.use_difficult).cpu().numpy()
        lab, pred = lab_t.cpu().numpy(), pred_t.cpu().numpy()
        for b in range(pred.shape[0]):
            pcls = pred[b, :, 0].astype(int)
            lcls = lab[b, :, 0].astype(int)
            seen = []
            for c in pcls:  # the reference takes the classes in order of first appearance (:118-124)
                if c >= 0 and c not in seen:
                    seen.append(int(c))
            for cid in seen:
                rows = np.where(pcls == cid)[0]
                records = np.hstack((pred[b, rows, 1][:, np.newaxis].astype(np.float64),
                                     flags[b, rows][:, np.newaxis].astype(np.float64)))
                gts = lab[b][lcls == cid]
                if (not self.use_difficult) and gts.shape[1] >= 6:  # :156-159
                    gt_count = int(np.sum(gts[:, 5] < 1))
                else:
                    gt_count = gts.shape[0]
                records = records[np.where(records[:, -1] > 0)[0], :]
                if records.size > 0:
                    self._insert(cid, records, gt_count)
            rest = []
            for c in lcls:  # classes that occur only in the labels (:168-176)
                if c not in seen and c not in rest:
                    rest.append(int(c))
            for cid in rest:
                if cid >= 0:
                    self._insert(cid, np.array([[0, 0]], dtype=np.float64), int(np.sum(lcls == cid)))

    def get(self):
        self._update()
        if self.num is None:
            if self.num_inst == 0:
                return (self.name, float("nan"))
            return (self.name, self.sum_metric / self.num_inst)
        names = ["%s" % (self.name[i]) for i in range(self.num)]
        values = [x / y if y != 0 else float("nan") for x, y in zip(self.sum_metric, self.num_inst)]
        return (names, values)

    def _update(self):
        aps = []
        for k, v in self.records.items():
            recall, prec = self._recall_prec(v, self.counts[k])
            ap = self._average_precision(recall, prec)
            aps.append(ap)
            if self.num is not None and k < (self.num - 1):
                self.sum_metric[k] = ap
                self.num_inst[k] = 1
        if self.num is None:
            self.num_inst = 1
            self.sum_metric = np.mean(aps)
        else:
            self.num_inst[-1] = 1
            self.sum_metric[-1] = np.mean(aps)

    def _recall_prec(self, record, count):
        record = np.delete(record, np.where(record[:, 1].astype(int) == 0)[0], axis=0)
        sorted_records = record[record[:, 0].argsort()[::-1]]
        tp = np.cumsum(sorted_records[:, 1].astype(int) == 1)
        fp = np.cumsum(sorted_records[:, 1].astype(int) == 2)
        recall = tp * 0.0 if count <= 0 else tp / float(count)
        with np.errstate(divide="ignore", invalid="ignore"):
            prec = tp.astype(float) / (tp + fp)
        return recall, prec

    def _average_precision(self, rec, prec):
        mrec = np.concatenate(([0.0], rec, [1.0]))
        mpre = np.concatenate(([0.0], prec, [0.0]))
        for i in range(mpre.size - 1, 0, -1):
            mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
        i = np.where(mrec[1:] != mrec[:-1])[0]
        return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])

    def _insert(self, key, records, count):
        if key not in self.records:
            self.records[key] = records
            self.counts[key] = count
        else:
            self.records[key] = np.vstack((self.records[key], records))
            self.counts[key] += count


class VOC07MApMetric(MApMetric):
    """11-point PASCAL VOC 07 AP -- evaluate/eval_metric.py:249-277."""

    def _average_precision(self, rec, prec):
        ap = 0.0
        for t in np.arange(0.0, 1.1, 0.1):
            p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap += p / 11.0
        return ap
