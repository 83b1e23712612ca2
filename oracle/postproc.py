"""ORACLE (test infrastructure, NOT product code) -- detector post-processing restated in numpy.

Follows, line by line in meaning (never in text):

* ``gen_anchors``                    ``BlazePoser/blazeFaceUtils.py:59-127`` with the options at
                                     ``BlazePoser/blazeFaceDetectorH5.py:236-241``
* ``filterDetections``               ``blazeFaceDetectorH5.py:319-327`` (threshold precomputed ``:85``)
* ``extractDetections``              ``blazeFaceDetectorH5.py:284-317`` (float64 arithmetic under NumPy 1.x
                                     scalar promotion)
* ``filterWithNonMaxSupression``     ``blazeFaceDetectorH5.py:329-357`` where
  ``tf.image.non_max_suppression`` is TensorFlow's NonMaxSuppressionV3 (un-vendored dependency
  ``tensorflow>=2.8.0``, files record Keras 2.13.1): float32 boxes, candidates by score descending
  with ties to the lower index, reject iff IoU with an already selected box is strictly greater
  than the threshold, IoU = 0 when either area <= 0, at most ``max_output_size`` results.

PARITY STATUS: unpinned by the reference (it has no tests and TensorFlow cannot run here);
pinned by hand-checked cases in ``tests/test_oracle_postproc.py``.

One deliberate definition: the reference's ``1/(1+np.exp(-x))`` on float32 uses NumPy's SIMD
``expf`` whose last-bit results depend on the host CPU.  To make *bit-exact kept indices* a
well-defined target, both this oracle and the CUDA kernel evaluate ``exp`` with the fully
specified float32 algorithm ``exp32`` below (IEEE round-to-nearest mul/add only, no FMA).
It agrees with a correctly rounded expf to <= 2 ulp, the same class as NumPy's.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

KEY_POINTS = 6
MAX_FACES = 100

# ------------------------------------------------------------------ anchors
def ssd_anchor_centres(input_w: int, input_h: int, strides: List[int], aspect_ratios=(1.0,),
                       interpolated_scale_aspect_ratio: float = 1.0, offset_x=0.5, offset_y=0.5,
                       reduce_boxes_in_lowest_layer=False):
    """MediaPipe SsdAnchorsCalculator, fixed_anchor_size=True => w=h=1. Returns (A,4) float64
    rows (x_center, y_center, h, w) in the reference's emission order."""
    n_layers = len(strides)
    rows = []
    layer = 0
    while layer < n_layers:
        per_cell = 0
        last = layer
        while last < n_layers and strides[last] == strides[layer]:
            if last == 0 and reduce_boxes_in_lowest_layer:
                per_cell += 3
            else:
                per_cell += len(aspect_ratios)
                if interpolated_scale_aspect_ratio > 0.0:
                    per_cell += 1
            last += 1
        fh = math.ceil(1.0 * input_h / strides[layer])
        fw = math.ceil(1.0 * input_w / strides[layer])
        for y in range(fh):
            for x in range(fw):
                for _ in range(per_cell):
                    rows.append(((x + offset_x) * 1.0 / fw, (y + offset_y) * 1.0 / fh, 1.0, 1.0))
        layer = last
    return np.asarray(rows, dtype=np.float64).reshape(-1, 4)


def blazeface_anchors(input_size: int = 128) -> np.ndarray:
    return ssd_anchor_centres(input_size, input_size, [8, 16, 16, 16])


# ------------------------------------------------------------------ deterministic float32 exp
_F = np.float32
_LOG2E = _F(1.4426950408889634)
_LN2_HI = _F(0.693359375)            # 9 significant bits: k*LN2_HI is exact for |k| < 2^15
_LN2_LO = _F(-2.12194440e-4)
_EXP_C = [_F(1.9875691500e-4), _F(1.3981999507e-3), _F(8.3334519073e-3),
          _F(4.1665795894e-2), _F(1.6666665459e-1), _F(5.0000001201e-1)]


def exp32(x: np.ndarray) -> np.ndarray:
    """Cephes-style expf using only float32 rn mul/add (mirrors hp_exp32 in csrc/postproc.cu)."""
    x = np.asarray(x, dtype=_F)
    x = np.minimum(np.maximum(x, _F(-87.0)), _F(88.0))
    k = np.rint(x * _LOG2E).astype(_F)
    r = x - k * _LN2_HI
    r = r - k * _LN2_LO
    p = _EXP_C[0]
    for c in _EXP_C[1:]:
        p = p * r + c            # float32 mul then float32 add (no fused op in numpy)
    r2 = r * r
    y = p * r2
    y = y + r
    y = y + _F(1.0)
    ki = k.astype(np.int32)
    scale = ((ki + 127).astype(np.uint32) << np.uint32(23)).view(_F)
    return (y * scale).astype(_F)


def sigmoid32(logits: np.ndarray) -> np.ndarray:
    e = exp32(-np.asarray(logits, dtype=_F))
    return (_F(1.0) / (_F(1.0) + e)).astype(_F)


# ------------------------------------------------------------------ detector post-processing
def logit_threshold(score_threshold: float) -> np.float32:
    """``np.log(t/(1-t))`` (blazeFaceDetectorH5.py:85); the comparison at :322 is float32-array vs
    float64-scalar which NumPy 1.x evaluates in float32."""
    return np.float32(np.log(score_threshold / (1 - score_threshold)))


def filter_detections(cls: np.ndarray, score_threshold: float):
    thr = logit_threshold(score_threshold)
    good = np.where(np.asarray(cls, dtype=np.float32) > thr)[0]
    return sigmoid32(cls[good]), good


def extract_detections(loc: np.ndarray, good: np.ndarray, anchors: np.ndarray, input_size: int = 128):
    """Float64 decode; elementwise numpy float64 ops round exactly like the reference's scalar ones."""
    n = good.shape[0]
    size = float(input_size)
    rows = loc[good].astype(np.float64).reshape(n, 16)
    ax = anchors[good, 0].astype(np.float64)
    ay = anchors[good, 1].astype(np.float64)
    cx = (rows[:, 0] + ax * size) / size
    cy = (rows[:, 1] + ay * size) / size
    w = rows[:, 2] / size
    h = rows[:, 3] / size
    kps = np.zeros((n, KEY_POINTS, 2), dtype=np.float64)
    for j in range(KEY_POINTS):
        kps[:, j, 0] = (rows[:, 4 + 2 * j] + ax * size) / size
        kps[:, j, 1] = (rows[:, 5 + 2 * j] + ay * size) / size
    boxes = np.stack([cx - w * 0.5, cy - h * 0.5, cx + w * 0.5, cy + h * 0.5], axis=1).reshape(n, 4)
    return boxes, kps


def _iou32(a: np.ndarray, b: np.ndarray) -> np.float32:
    f = np.float32
    ymin_i, ymax_i = min(a[0], a[2]), max(a[0], a[2])
    xmin_i, xmax_i = min(a[1], a[3]), max(a[1], a[3])
    ymin_j, ymax_j = min(b[0], b[2]), max(b[0], b[2])
    xmin_j, xmax_j = min(b[1], b[3]), max(b[1], b[3])
    area_i = f(f(ymax_i - ymin_i) * f(xmax_i - xmin_i))
    area_j = f(f(ymax_j - ymin_j) * f(xmax_j - xmin_j))
    if area_i <= 0 or area_j <= 0:
        return f(0.0)
    iy0, ix0 = max(ymin_i, ymin_j), max(xmin_i, xmin_j)
    iy1, ix1 = min(ymax_i, ymax_j), min(xmax_i, xmax_j)
    inter = f(max(f(iy1 - iy0), f(0.0)) * max(f(ix1 - ix0), f(0.0)))
    return f(inter / f(f(area_i + area_j) - inter))


def _iou32_many(a: np.ndarray, kept: np.ndarray) -> np.ndarray:
    """_iou32 of one box against an (m,4) array; elementwise float32 numpy ops are IEEE, hence bit-identical
    to the scalar version above (checked in tests/test_oracle_postproc.py)."""
    f = np.float32
    ymin_i, ymax_i = min(a[0], a[2]), max(a[0], a[2])
    xmin_i, xmax_i = min(a[1], a[3]), max(a[1], a[3])
    ymin_j, ymax_j = np.minimum(kept[:, 0], kept[:, 2]), np.maximum(kept[:, 0], kept[:, 2])
    xmin_j, xmax_j = np.minimum(kept[:, 1], kept[:, 3]), np.maximum(kept[:, 1], kept[:, 3])
    area_i = f(f(ymax_i - ymin_i) * f(xmax_i - xmin_i))
    area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j)
    iy = np.maximum(np.minimum(ymax_i, ymax_j) - np.maximum(ymin_i, ymin_j), f(0.0))
    ix = np.maximum(np.minimum(xmax_i, xmax_j) - np.maximum(xmin_i, xmin_j), f(0.0))
    inter = iy * ix
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / ((area_i + area_j) - inter)
    iou = np.where((area_j <= 0) | (area_i <= 0), f(0.0), iou)
    return iou.astype(f)


def tf_non_max_suppression(boxes, scores, max_output_size: int, iou_threshold: float, scalar: bool = False) -> np.ndarray:
    """Greedy hard NMS with TensorFlow NonMaxSuppressionV3 semantics (see module docstring)."""
    b = np.asarray(boxes).astype(np.float32).reshape(-1, 4)
    s = np.asarray(scores).astype(np.float32)
    thr = np.float32(iou_threshold)
    order = np.lexsort((np.arange(len(s)), -s.astype(np.float64)))   # score descending, ties -> lower index
    keep: List[int] = []
    kept_boxes = np.zeros((max(max_output_size, 1), 4), np.float32)
    for i in order:
        if len(keep) >= max_output_size:
            break
        if scalar:
            ok = not any(_iou32(b[i], b[j]) > thr for j in reversed(keep))
        else:
            ok = not (keep and bool(np.any(_iou32_many(b[i], kept_boxes[:len(keep)]) > thr)))
        if ok:
            kept_boxes[len(keep)] = b[i]
            keep.append(int(i))
    return np.asarray(keep, dtype=np.int32)


def pose_for_anchor(anchor_id: int, pose16: np.ndarray, pose8: np.ndarray) -> np.ndarray:
    """blazeFaceDetectorH5.py:342-353, generalised from 16/8 to the actual map widths."""
    n16 = pose16.shape[0] * pose16.shape[1] * 2
    if anchor_id < n16:
        cell = anchor_id // 2
        return pose16[cell // pose16.shape[1], cell % pose16.shape[1]]
    cell = (anchor_id - n16) // 6
    return pose8[cell // pose8.shape[1], cell % pose8.shape[1]]


def detect_postprocess(cls: np.ndarray, loc: np.ndarray, pose16: np.ndarray, pose8: np.ndarray,
                       anchors: np.ndarray, score_threshold=0.4, iou_threshold=0.3,
                       input_size: int = 128, max_faces: int = MAX_FACES):
    """cls (A,), loc (A,16), pose16 (H16,W16,3), pose8 (H8,W8,3) for ONE image ->
    dict(kept_anchor, boxes f64 (k,4), keypoints f64 (k,6,2), scores f32 (k,), poses f32 (k,3))."""
    scores, good = filter_detections(cls, score_threshold)
    boxes, kps = extract_detections(loc, good, anchors, input_size)
    sel = tf_non_max_suppression(boxes, scores, max_faces, iou_threshold)
    kept = good[sel].astype(np.int32)
    if sel.size == 0:
        poses = np.zeros((0, 3), dtype=np.float32)
    else:
        poses = np.stack([pose_for_anchor(int(a), pose16, pose8) for a in kept], axis=0).astype(np.float32)
    return {"kept_anchor": kept, "boxes": boxes[sel], "keypoints": kps[sel],
            "scores": scores[sel], "poses": poses}
