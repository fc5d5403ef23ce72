"""Host-side mirror of the anchor branch of the reference's graph builders.

``multibox_layer`` / ``multitask_layer`` (symbol/common.py:136-283, :286-433) emit, per feature map, a
``MultiBoxPrior`` node followed by Flatten, then Concat + Reshape to ``(1, A, 4)`` (:415-432).  The conv heads that
produce ``loc_preds`` / ``cls_preds`` stay the reference's (out of scope); only the anchor tensor and the layout
contract of the three tensors handed to MultiBoxTarget / MultiBoxDetection live here.
"""
from . import presets as _presets
from .ops import multibox_prior_concat


def multibox_anchors(preset, clip=False, device=None):
    """Anchors ``(1, A, 4)`` of a named preset (``ssd300``, ``ssd512``, ``ssd512_generic``, ``dspnet_cs``) or of a
    ``presets.Preset``, generated in one launch."""
    p = _presets.PRESETS[preset] if isinstance(preset, str) else preset
    steps = [fm.step for fm in p.maps]
    auto = all(s <= 0 for s in steps)
    return multibox_prior_concat([(fm.height, fm.width) for fm in p.maps], [fm.sizes for fm in p.maps],
                                 [fm.ratios for fm in p.maps], steps=None if auto else steps, clip=clip, device=device)


def head_shapes(preset, batch, loc_width=5):
    """Shapes of the tensors the reference graph hands to the ops: loc_preds (B, A*5), cls_preds (B, C, A),
    anchors (1, A, 4) (symbol/common.py:424-432) and label (B, L, 6) (dataset/iterator.py:553-603)."""
    p = _presets.PRESETS[preset] if isinstance(preset, str) else preset
    a = _presets.num_anchors(p)
    return {"loc_preds": (batch, a * loc_width), "cls_preds": (batch, p.num_classes, a), "anchors": (1, a, 4),
            "label": (batch, p.label_slots, 6)}
