"""Anchor presets of the reference's detection heads (shape/config facts only).

Each preset lists, per feature map, (H, W, sizes, ratios, step) exactly as the reference graphs hand them to
``MultiBoxPrior`` (symbol/common.py:415-420: ``steps=(s, s)`` or the auto-step sentinel ``(-1, -1)``):

* ``ssd300``        symbol/symbol_factory.py:31-39 (VGG16-reduced, 38..1 maps)            -> 8732 anchors
* ``ssd512``        symbol/legacy_vgg16_ssd_512.py:117-127 (64..1 maps)                   -> 24564 anchors
* ``ssd512_generic``symbol/symbol_factory.py:19-29 (last extra layer keeps a 2x2 map)     -> 24576 anchors
* ``dspnet_cs``     symbol/multitask_symbol_factory.py:68-81 with the first entry dropped
                    (symbol/multitask_symbol_builder.py:503-508), 512x1024 input, auto steps -> 12264 anchors
                    (matches the shape dump in utils.py:37)
"""
from collections import namedtuple

FeatureMap = namedtuple("FeatureMap", "height width sizes ratios step")
Preset = namedtuple("Preset", "name maps num_classes label_slots")

_R3 = (1.0, 2.0, 0.5)
_R5 = (1.0, 2.0, 0.5, 3.0, 1.0 / 3)


def _maps(hw, sizes, ratios, steps):
    return tuple(FeatureMap(h, w, tuple(s), tuple(r), st) for (h, w), s, r, st in zip(hw, sizes, ratios, steps))


_SSD300_SIZES = [[.1, .141], [.2, .272], [.37, .447], [.54, .619], [.71, .79], [.88, .961]]
_SSD512_SIZES = [[.07, .1025], [.15, .2121], [.3, .3674], [.45, .5196], [.6, .6708], [.75, .8216], [.9, .9721]]

PRESETS = {
    "ssd300": Preset(
        "ssd300",
        _maps([(38, 38), (19, 19), (10, 10), (5, 5), (3, 3), (1, 1)], _SSD300_SIZES,
              [_R3, _R5, _R5, _R5, _R3, _R3], [x / 300.0 for x in [8, 16, 32, 64, 100, 300]]),
        21, 58),
    "ssd512": Preset(
        "ssd512",
        _maps([(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2), (1, 1)], _SSD512_SIZES,
              [_R3, _R5, _R5, _R5, _R5, _R3, _R3], [x / 512.0 for x in [8, 16, 32, 64, 128, 256, 512]]),
        21, 58),
    "ssd512_generic": Preset(
        "ssd512_generic",
        _maps([(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2), (2, 2)], _SSD512_SIZES,
              [_R3, _R5, _R5, _R5, _R5, _R3, _R3], [x / 512.0 for x in [8, 16, 32, 64, 128, 256, 512]]),
        21, 58),
    "dspnet_cs": Preset(
        "dspnet_cs",
        _maps([(32, 64), (16, 32), (8, 16), (4, 8), (2, 4), (1, 2)], _SSD300_SIZES,
              [_R3, _R5, _R5, _R5, _R3, _R3], [-1.0] * 6),
        9, 200),
}


def anchors_per_location(fm):
    return len(fm.sizes) + len(fm.ratios) - 1


def num_anchors(preset):
    p = PRESETS[preset] if isinstance(preset, str) else preset
    return sum(fm.height * fm.width * anchors_per_location(fm) for fm in p.maps)
