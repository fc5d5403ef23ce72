"""Seeded synthetic head tensors for the multibox path (numpy only).

All generators use ``numpy.random.Generator(PCG64(seed))`` with ``seed = 1000 * config_id + image_index`` so
that an image's tensors do not depend on the batch it is placed in -- which is what lets the multi-GPU tests
require bit-identical gathered outputs for any sharding of the batch.
"""
import numpy as np


def _rng(config_id, image_index):
    return np.random.Generator(np.random.PCG64(1000 * int(config_id) + int(image_index)))


def labels(config_id, batch, num_slots, num_classes, max_gt=8, first_image=0, edge_cases=True):
    """(B, L, 6) ``[cls, xmin, ymin, xmax, ymax, dist]`` rows, valid rows first, padding rows all -1
    (the producer contract of dataset/iterator.py:509-539).  With ``edge_cases`` global image 1 has no
    ground truth and global image 2 fills every slot."""
    out = np.full((batch, num_slots, 6), -1.0, np.float32)
    for b in range(batch):
        gi = first_image + b
        r = _rng(config_id, gi)
        g = int(r.integers(1, max_gt + 1))
        if edge_cases and gi == 1:
            g = 0
        if edge_cases and gi == 2:
            g = num_slots
        g = min(g, num_slots)
        cx, cy = r.uniform(.1, .9, g), r.uniform(.1, .9, g)
        w, h = r.uniform(.03, .6, g), r.uniform(.03, .6, g)
        out[b, :g, 0] = r.integers(0, num_classes - 1, g)
        out[b, :g, 1] = np.clip(cx - w / 2, 0, 1)
        out[b, :g, 2] = np.clip(cy - h / 2, 0, 1)
        out[b, :g, 3] = np.clip(cx + w / 2, 0, 1)
        out[b, :g, 4] = np.clip(cy + h / 2, 0, 1)
        out[b, :g, 5] = r.uniform(0, 1, g)
    return out


def cls_preds(config_id, batch, num_classes, num_anchors, first_image=0, bg_boost=4.0):
    """(B, C, A) raw logits for MultiBoxTarget: N(0,1) with the background channel raised by N(bg_boost,1)
    on 97 % of the anchors."""
    out = np.empty((batch, num_classes, num_anchors), np.float32)
    for b in range(batch):
        r = _rng(config_id + 1, first_image + b)
        x = r.standard_normal((num_classes, num_anchors), dtype=np.float32)
        bg = r.random(num_anchors) < 0.97
        x[0] += np.where(bg, r.normal(bg_boost, 1.0, num_anchors), 0).astype(np.float32)
        out[b] = x
    return out


def cls_prob(config_id, batch, num_classes, num_anchors, first_image=0, dense=False):
    """(B, C, A) channel-softmax probabilities for MultiBoxDetection.  Default: background boosted by N(8,1)
    on 97 % of the anchors and one foreground logit boosted by N(6,2) on the rest, so that valid_count is a
    few hundred to a few thousand and usually exceeds nms_topk=400.  ``dense``: plain N(0,1) logits, every
    anchor valid (stress row)."""
    out = np.empty((batch, num_classes, num_anchors), np.float32)
    for b in range(batch):
        r = _rng(config_id + 2, first_image + b)
        x = r.standard_normal((num_classes, num_anchors), dtype=np.float32)
        if not dense:
            bg = r.random(num_anchors) < 0.97
            x[0] += np.where(bg, r.normal(8.0, 1.0, num_anchors), 0).astype(np.float32)
            fg_cls = r.integers(1, num_classes, num_anchors)
            boost = np.where(bg, 0, r.normal(6.0, 2.0, num_anchors)).astype(np.float32)
            x[fg_cls, np.arange(num_anchors)] += boost
        x -= x.max(axis=0, keepdims=True)
        e = np.exp(x, dtype=np.float32)
        out[b] = e / e.sum(axis=0, keepdims=True, dtype=np.float32)
    return out


def loc_pred(config_id, batch, num_anchors, first_image=0):
    """(B, A*5): N(0,0.5) box channels, U(0,10) distance channel."""
    out = np.empty((batch, num_anchors, 5), np.float32)
    for b in range(batch):
        r = _rng(config_id + 3, first_image + b)
        out[b, :, :4] = r.normal(0, 0.5, (num_anchors, 4))
        out[b, :, 4] = r.uniform(0, 10, num_anchors)
    return out.reshape(batch, num_anchors * 5)


def nms_boxes(seed, n, with_class=False, num_classes=20):
    """(N,5) [x1,y1,x2,y2,score] pixel boxes with tie-free scores (a permuted linspace), for the standalone
    NMS sweep; with_class appends a class column -> (N,6)."""
    r = np.random.Generator(np.random.PCG64(seed))
    x1 = r.uniform(0, 1000, n)
    y1 = r.uniform(0, 1000, n)
    w = r.uniform(8, 200, n)
    h = r.uniform(8, 200, n)
    scores = r.permutation(np.linspace(0.01, 0.99, n))
    cols = [x1, y1, x1 + w, y1 + h, scores]
    if with_class:
        cols.append(r.integers(0, num_classes, n))
    return np.stack(cols, axis=1).astype(np.float32)


def heads_from_logits(preset, logits, loc):
    """Per-scale conv outputs that multibox_layer (symbol/common.py:399-432) would turn into the given
    ``logits`` (B, C, A) and ``loc`` (B, A*5): class heads (B, na*C, H, W) and loc heads (B, na*5, H, W), NCHW with
    channel = anchor_in_cell * C + class.  (The inverse of the transpose / Flatten / Concat / Reshape / transpose
    chain, so that head-fed and tensor-fed runs see the same numbers.)"""
    B, C, A = logits.shape
    loc = loc.reshape(B, A, 5)
    cls_heads, loc_heads, a0 = [], [], 0
    for fm in preset.maps:
        na = len(fm.sizes) + len(fm.ratios) - 1
        n = fm.height * fm.width * na
        blk = logits[:, :, a0:a0 + n].reshape(B, C, fm.height, fm.width, na)          # (B, C, H, W, na)
        cls_heads.append(np.ascontiguousarray(np.transpose(blk, (0, 4, 1, 2, 3)).reshape(B, na * C, fm.height, fm.width)))
        lb = loc[:, a0:a0 + n].reshape(B, fm.height, fm.width, na, 5)                    # (B, H, W, na, 5)
        loc_heads.append(np.ascontiguousarray(np.transpose(lb, (0, 3, 4, 1, 2)).reshape(B, na * 5, fm.height, fm.width)))
        a0 += n
    assert a0 == A
    return cls_heads, loc_heads


def det_logits(config_id, batch, num_classes, num_anchors, first_image=0):
    """(B, C, A) raw logits with the same mixture as cls_prob() (before its softmax): background boosted by N(8,1) on
    97 % of the anchors, one foreground logit boosted by N(6,2) on the rest."""
    out = np.empty((batch, num_classes, num_anchors), np.float32)
    for b in range(batch):
        r = _rng(config_id + 2, first_image + b)
        x = r.standard_normal((num_classes, num_anchors), dtype=np.float32)
        bg = r.random(num_anchors) < 0.97
        x[0] += np.where(bg, r.normal(8.0, 1.0, num_anchors), 0).astype(np.float32)
        fg_cls = r.integers(1, num_classes, num_anchors)
        boost = np.where(bg, 0, r.normal(6.0, 2.0, num_anchors)).astype(np.float32)
        x[fg_cls, np.arange(num_anchors)] += boost
        out[b] = x
    return out
