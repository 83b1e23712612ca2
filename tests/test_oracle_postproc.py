"""Hand-checked cases that pin the post-processing oracle (TF NonMaxSuppressionV3 semantics)."""
import numpy as np

from oracle import postproc as opp


def test_exp32_within_2ulp_of_true_exp():
    x = np.linspace(-87, 30, 400001).astype(np.float32)
    got = opp.exp32(x).astype(np.float64)
    true = np.exp(x.astype(np.float64))
    ulp = np.spacing(true.astype(np.float32)).astype(np.float64)
    assert (np.abs(got - true) / ulp).max() < 2.0
    assert opp.exp32(np.float32(0.0)) == np.float32(1.0)


def test_sigmoid_threshold_is_float32_compare():
    thr = opp.logit_threshold(0.4)
    assert thr.dtype == np.float32 and abs(float(thr) - np.log(0.4 / 0.6)) < 1e-7
    cls = np.array([thr, np.nextafter(thr, np.float32(1)), -5, 3], dtype=np.float32)
    scores, good = opp.filter_detections(cls, 0.4)
    assert good.tolist() == [1, 3]            # strict '>' on the float32 threshold
    assert scores.dtype == np.float32 and abs(scores[1] - 1 / (1 + np.exp(-3))) < 1e-6


def test_nms_basic_suppression_order_and_cap():
    boxes = np.array([[0, 0, 1, 1], [0, 0, 1, 0.9], [2, 2, 3, 3], [0, 0, 1, 1]], dtype=np.float64)
    scores = np.array([0.9, 0.8, 0.7, 0.9], dtype=np.float32)
    # ties (idx 0 and 3) -> lower index first; box 3 duplicates box 0 (IoU 1 > 0.3) and is dropped; box 1 IoU .9
    assert opp.tf_non_max_suppression(boxes, scores, 100, 0.3).tolist() == [0, 2]
    assert opp.tf_non_max_suppression(boxes, scores, 1, 0.3).tolist() == [0]
    assert opp.tf_non_max_suppression(boxes[:0], scores[:0], 100, 0.3).tolist() == []


def test_nms_strict_threshold_and_degenerate_boxes():
    # IoU exactly 1/3 with threshold 1/3 (float32) is NOT suppressed: strict '>'
    a = np.array([[0, 0, 1, 2], [0, 1, 1, 3]], dtype=np.float64)   # inter 1, union 3
    thr = float(np.float32(1.0) / np.float32(3.0))
    assert opp.tf_non_max_suppression(a, np.array([0.9, 0.8], np.float32), 10, thr).tolist() == [0, 1]
    assert opp.tf_non_max_suppression(a, np.array([0.9, 0.8], np.float32), 10, np.nextafter(np.float32(thr), np.float32(0))).tolist() == [0]
    # zero-area boxes have IoU 0 with everything, flipped corners are normalised
    z = np.array([[0, 0, 1, 1], [0.5, 0.5, 0.5, 0.9], [1, 1, 0, 0]], dtype=np.float64)
    assert opp.tf_non_max_suppression(z, np.array([0.9, 0.8, 0.7], np.float32), 10, 0.3).tolist() == [0, 1]


def test_decode_is_float64_then_boxes_cast_to_float32_in_nms():
    anchors = opp.blazeface_anchors(128)
    loc = np.zeros((896, 16), np.float32)
    loc[5] = np.arange(16, dtype=np.float32) * np.float32(1.1) + np.float32(0.3)
    boxes, kps = opp.extract_detections(loc, np.array([5]), anchors, 128)
    assert boxes.dtype == np.float64 and kps.shape == (1, 6, 2)
    ax, ay = anchors[5, 0], anchors[5, 1]
    cx = (float(loc[5, 0]) + ax * 128.0) / 128.0
    w = float(loc[5, 2]) / 128.0
    assert boxes[0, 0] == cx - w * 0.5 and boxes[0, 2] == cx + w * 0.5
    assert kps[0, 3, 1] == (float(loc[5, 11]) + ay * 128.0) / 128.0


def test_pose_lookup_by_anchor_cell():
    p16 = np.arange(16 * 16 * 3, dtype=np.float32).reshape(16, 16, 3)
    p8 = -np.arange(8 * 8 * 3, dtype=np.float32).reshape(8, 8, 3)
    assert np.array_equal(opp.pose_for_anchor(0, p16, p8), p16[0, 0])
    assert np.array_equal(opp.pose_for_anchor(35, p16, p8), p16[1, 1])      # cell 17
    assert np.array_equal(opp.pose_for_anchor(512, p16, p8), p8[0, 0])
    assert np.array_equal(opp.pose_for_anchor(512 + 6 * 9 + 5, p16, p8), p8[1, 1])
    assert np.array_equal(opp.pose_for_anchor(895, p16, p8), p8[7, 7])


def test_detect_postprocess_empty_and_full():
    anchors = opp.blazeface_anchors(128)
    p16, p8 = np.zeros((16, 16, 3), np.float32), np.zeros((8, 8, 3), np.float32)
    out = opp.detect_postprocess(np.full(896, -10, np.float32), np.zeros((896, 16), np.float32), p16, p8, anchors)
    assert out["kept_anchor"].size == 0 and out["poses"].shape == (0, 3) and out["boxes"].shape == (0, 4)
    rng = np.random.default_rng(7)
    cls = rng.normal(0, 2, 896).astype(np.float32)
    loc = np.zeros((896, 16), np.float32)
    loc[:, :2] = rng.uniform(-8, 8, (896, 2))
    loc[:, 2:4] = rng.uniform(16, 64, (896, 2))
    out = opp.detect_postprocess(cls, loc, p16, p8, anchors)
    k = out["kept_anchor"]
    assert 0 < len(k) <= 100 and len(set(k.tolist())) == len(k)
    assert np.all(np.diff(out["scores"]) <= 0)          # selection order = score descending


def test_vectorised_nms_equals_scalar_restatement():
    rng = np.random.default_rng(11)
    for trial in range(6):
        n = int(rng.integers(1, 400))
        c = rng.uniform(0, 1, (n, 2))
        wh = rng.uniform(0.0, 0.4, (n, 2)) * (rng.random((n, 2)) > 0.05)     # some zero-area boxes
        boxes = np.concatenate([c - wh / 2, c + wh / 2], axis=1)
        if trial % 2:
            boxes[:, [0, 2]] = boxes[:, [2, 0]]                              # flipped corners
        scores = np.round(rng.random(n), 2).astype(np.float32)               # many ties
        a = opp.tf_non_max_suppression(boxes, scores, 100, 0.3)
        b = opp.tf_non_max_suppression(boxes, scores, 100, 0.3, scalar=True)
        assert np.array_equal(a, b)
