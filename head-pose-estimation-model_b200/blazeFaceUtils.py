"""SSD anchor generation with the API of the reference's BlazePoser/blazeFaceUtils.py.

``SsdAnchorsCalculatorOptions`` (:3-51), ``Anchor`` (:53-60) and ``gen_anchors`` (:59-127) keep their
names and meaning (MediaPipe SsdAnchorsCalculator).  The implementation is array-based: anchors of
one stride group are produced with ``numpy`` broadcasting instead of nested Python loops; use
``anchor_table`` when an ``(A,4)`` array is more convenient than a list of ``Anchor`` objects.
"""
import math

import numpy as np


class SsdAnchorsCalculatorOptions:
    def __init__(self, input_size_width, input_size_height, min_scale, max_scale, num_layers, feature_map_width,
                 feature_map_height, strides, aspect_ratios, anchor_offset_x=0.5, anchor_offset_y=0.5,
                 reduce_boxes_in_lowest_layer=False, interpolated_scale_aspect_ratio=1.0, fixed_anchor_size=False):
        self.input_size_width, self.input_size_height = input_size_width, input_size_height
        self.min_scale, self.max_scale = min_scale, max_scale
        self.anchor_offset_x, self.anchor_offset_y = anchor_offset_x, anchor_offset_y
        self.num_layers = num_layers
        self.feature_map_width, self.feature_map_height = list(feature_map_width), list(feature_map_height)
        self.feature_map_width_size, self.feature_map_height_size = len(feature_map_width), len(feature_map_height)
        self.strides, self.strides_size = list(strides), len(strides)
        self.aspect_ratios, self.aspect_ratios_size = list(aspect_ratios), len(aspect_ratios)
        self.reduce_boxes_in_lowest_layer = reduce_boxes_in_lowest_layer
        self.interpolated_scale_aspect_ratio = interpolated_scale_aspect_ratio
        self.fixed_anchor_size = fixed_anchor_size

    def to_string(self):
        return "\n".join(f"{k}: {v}" for k, v in vars(self).items() if not k.endswith("_size") or k == "fixed_anchor_size")


class Anchor:
    __slots__ = ("x_center", "y_center", "h", "w")

    def __init__(self, x_center, y_center, h, w):
        self.x_center, self.y_center, self.h, self.w = x_center, y_center, h, w

    def to_string(self):
        return f"x_center: {self.x_center}, y_center: {self.y_center}, h: {self.h}, w: {self.w}"


def _layer_scale(o, layer):
    return o.min_scale + (o.max_scale - o.min_scale) * 1.0 * layer / (o.strides_size - 1.0)


def anchor_table(options) -> np.ndarray:
    """(A,4) float64 rows (x_center, y_center, h, w) in MediaPipe emission order; empty on bad options."""
    o = options
    if o.strides_size != o.num_layers:
        print("strides_size and num_layers must be equal.")
        return np.zeros((0, 4))
    chunks = []
    first = 0
    while first < o.strides_size:
        # consecutive layers sharing a stride contribute their boxes to the same grid cells
        heights, widths = [], []
        last = first
        while last < o.strides_size and o.strides[last] == o.strides[first]:
            scale = _layer_scale(o, last)
            if last == 0 and o.reduce_boxes_in_lowest_layer:
                pairs = [(1.0, 0.1), (2.0, scale), (0.5, scale)]
            else:
                pairs = [(ar, scale) for ar in o.aspect_ratios]
                if o.interpolated_scale_aspect_ratio > 0.0:
                    nxt = 1.0 if last == o.strides_size - 1 else _layer_scale(o, last + 1)
                    pairs.append((o.interpolated_scale_aspect_ratio, math.sqrt(scale * nxt)))
            for ar, sc in pairs:
                root = math.sqrt(ar)
                heights.append(sc / root)
                widths.append(sc * root)
            last += 1
        if o.feature_map_height_size > 0:
            fh, fw = o.feature_map_height[first], o.feature_map_width[first]
        else:
            fh = math.ceil(1.0 * o.input_size_height / o.strides[first])
            fw = math.ceil(1.0 * o.input_size_width / o.strides[first])
        k = len(heights)
        xc = (np.arange(fw, dtype=np.float64) + o.anchor_offset_x) * 1.0 / fw
        yc = (np.arange(fh, dtype=np.float64) + o.anchor_offset_y) * 1.0 / fh
        grid = np.empty((fh, fw, k, 4), dtype=np.float64)
        grid[..., 0] = xc[None, :, None]
        grid[..., 1] = yc[:, None, None]
        if o.fixed_anchor_size:
            grid[..., 2:] = 1.0
        else:
            grid[..., 2] = np.asarray(heights)[None, None, :]
            grid[..., 3] = np.asarray(widths)[None, None, :]
        chunks.append(grid.reshape(-1, 4))
        first = last
    return np.concatenate(chunks, axis=0) if chunks else np.zeros((0, 4))


def gen_anchors(options):
    return [Anchor(float(r[0]), float(r[1]), float(r[2]), float(r[3])) for r in anchor_table(options)]
