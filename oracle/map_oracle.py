"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the reference's detection post-filter and mAP matching
(SURVEY.md 8f row f3): multi_solver.py:419-432 and evaluate/eval_metric.py:69-247 (+ VOC07 :249-277).

The matching is restated in the form the CUDA kernel uses -- one pass over the prediction rows in row order with a
`found` flag per label row -- which is equivalent to the reference's class-by-class loop because classes own
disjoint label rows; tests/test_evalmap.py checks it against the reference class itself (oracle/ref_map.py)."""
import numpy as np

F32 = np.float32


def postfilter(out, max_rows=200, score_thresh=0.25):
    """multi_solver.py:419-432: rows with id >= 0, then score > thresh, in row order, into (B, max_rows, 7) filled
    with -1.  (The reference raises if more than max_rows rows survive; here they are cut off.)"""
    out = np.asarray(out, dtype=F32)
    B = out.shape[0]
    pred = np.zeros((B, max_rows, out.shape[2]), F32) - F32(1.0)
    counts = np.zeros((B,), np.int32)
    for b in range(B):
        rows = out[b][out[b, :, 0] >= 0]
        rows = rows[rows[:, 1] > F32(score_thresh)][:max_rows]
        pred[b, : rows.shape[0]] = rows
        counts[b] = rows.shape[0]
    return pred, counts


def _iou(x, ys):
    """eval_metric.py:81-106, float32 elementwise like numpy evaluates it."""
    ixmin = np.maximum(ys[:, 0], x[0])
    iymin = np.maximum(ys[:, 1], x[1])
    ixmax = np.minimum(ys[:, 2], x[2])
    iymax = np.minimum(ys[:, 3], x[3])
    iw = np.maximum(ixmax - ixmin, F32(0.0))
    ih = np.maximum(iymax - iymin, F32(0.0))
    inters = iw * ih
    uni = (x[2] - x[0]) * (x[3] - x[1]) + (ys[:, 2] - ys[:, 0]) * (ys[:, 3] - ys[:, 1]) - inters
    with np.errstate(divide="ignore", invalid="ignore"):
        ious = inters / uni
    ious[uni < F32(1e-12)] = 0
    return ious


def match_flags(labels, preds, ovp_thresh=0.5, use_difficult=False):
    """flags (B, M) int32 per prediction row: 0 not recorded (id < 0, or matched to a difficult gt), 1 TP, 2 FP."""
    labels = np.asarray(labels, dtype=F32)
    preds = np.asarray(preds, dtype=F32)
    B, M = preds.shape[:2]
    flags = np.zeros((B, M), np.int32)
    for b in range(B):
        lab = labels[b]
        lcls = lab[:, 0].astype(int)
        found = np.zeros((lab.shape[0],), bool)
        for j in range(M):
            cid = int(preds[b, j, 0])
            if cid < 0:
                continue
            idx = np.where(lcls == cid)[0]
            if idx.size == 0:
                flags[b, j] = 2
                continue
            ious = _iou(preds[b, j, 2:], lab[idx, 1:5])
            a = int(np.argmax(ious))
            if ious[a] > F32(ovp_thresh):
                if (not use_difficult) and lab.shape[1] >= 6 and lab[idx[a], 5] > 0:
                    pass
                elif not found[idx[a]]:
                    flags[b, j] = 1
                    found[idx[a]] = True
                else:
                    flags[b, j] = 2
            else:
                flags[b, j] = 2
    return flags


class MApAccumulator(object):
    """Host side of MApMetric: turns (labels, preds, flags) into the reference's records / counts
    (eval_metric.py:113-176,233-246) and evaluates AP (:178-232, VOC07 :254-277)."""

    def __init__(self, use_difficult=False, voc07=False):
        self.use_difficult = use_difficult
        self.voc07 = voc07
        self.records = {}
        self.counts = {}

    def _push(self, cid, rows, gt_count):
        if cid in self.records:
            self.records[cid] = np.vstack((self.records[cid], rows))
            self.counts[cid] += gt_count
        else:
            self.records[cid], self.counts[cid] = rows, gt_count

    def update(self, labels, preds, flags):
        labels = np.asarray(labels, dtype=F32)
        preds = np.asarray(preds, dtype=F32)
        for b in range(preds.shape[0]):
            lab, pred, fl = labels[b], preds[b], flags[b]
            pcls = pred[:, 0].astype(int)
            lcls = lab[:, 0].astype(int)
            seen = []
            for c in pcls:  # classes in order of first appearance among the predictions
                if c >= 0 and c not in seen:
                    seen.append(int(c))
            for cid in seen:
                rows = np.where(pcls == cid)[0]
                rec = np.stack((pred[rows, 1].astype(np.float64), fl[rows].astype(np.float64)), axis=1)
                rec = rec[rec[:, 1] > 0]
                gts = lab[lcls == cid]
                easy_only = (not self.use_difficult) and gts.shape[1] >= 6
                if rec.size > 0:
                    self._push(cid, rec, int(np.sum(gts[:, 5] < 1)) if easy_only else gts.shape[0])
            rest = []
            for c in lcls:  # classes that only occur in the labels, in order of first appearance
                if c >= 0 and c not in seen and c not in rest:
                    rest.append(int(c))
            for cid in rest:
                self._push(cid, np.zeros((1, 2)), int(np.sum(lcls == cid)))

    @staticmethod
    def pr_curve(record, count):
        """:196-207 -- flag-0 rows dropped, descending score, cumulative TP / FP."""
        kept = record[record[:, 1].astype(int) != 0]
        fl = kept[np.argsort(kept[:, 0])[::-1], 1].astype(int)
        tp, fp = np.cumsum(fl == 1), np.cumsum(fl == 2)
        with np.errstate(divide="ignore", invalid="ignore"):
            return (tp / float(count) if count > 0 else tp * 0.0), tp.astype(float) / (tp + fp)

    def average_precision(self, rec, prec):
        if self.voc07:  # :254-277, eleven recall levels
            total = 0.0
            for t in np.arange(0.0, 1.1, 0.1):
                hit = rec >= t
                total += (prec[hit].max() if hit.any() else 0) / 11.0
            return total
        # :209-231, area under the right-to-left running maximum of the precision
        r = np.concatenate(([0.0], rec, [1.0]))
        p = np.concatenate(([0.0], prec, [0.0]))
        for k in range(p.size - 2, -1, -1):
            if p[k + 1] > p[k]:
                p[k] = p[k + 1]
        idx = np.nonzero(r[1:] != r[:-1])[0]
        return np.sum((r[idx + 1] - r[idx]) * p[idx + 1])

    def get(self):
        aps = [self.average_precision(*self.pr_curve(v, self.counts[k])) for k, v in self.records.items()]
        return "mAP", float(np.mean(aps)) if aps else float("nan")
