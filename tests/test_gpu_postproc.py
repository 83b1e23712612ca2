"""GPU parity of threshold + decode + NMS + pose lookup: bit-exact against the numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import postproc as opp

pytestmark = pytest.mark.gpu


def _stress(B, size, seed, sigma=2.0):
    rng = np.random.default_rng(seed)
    g16, g8 = -(-size // 8), -(-size // 16)
    A = g16 * g16 * 2 + g8 * g8 * 6
    cls = rng.normal(0, sigma, (B, A)).astype(np.float32)
    loc = np.zeros((B, A, 16), np.float32)
    loc[..., :2] = rng.uniform(-8, 8, (B, A, 2))
    loc[..., 2:4] = rng.uniform(16, 64, (B, A, 2))
    loc[..., 4:] = rng.uniform(-20, 20, (B, A, 12))
    p16 = rng.normal(0, 20, (B, g16, g16, 3)).astype(np.float32)
    p8 = rng.normal(0, 20, (B, g8, g8, 3)).astype(np.float32)
    return cls, loc, p16, p8


def _run_fused(cls, loc, p16, p8, size, score_thr=0.4, iou_thr=0.3, max_out=100):
    from hpose_b200 import _lib
    from hpose_b200.device import default_context
    ctx = default_context()
    B = cls.shape[0]
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tc, tl, t16, t8 = d(cls), d(loc), d(p16), d(p8)
    cnt = torch.empty((B,), dtype=torch.int32, device="cuda")
    anc = torch.empty((B, max_out), dtype=torch.int32, device="cuda")
    boxes = torch.zeros((B, max_out, 4), dtype=torch.float64, device="cuda")
    kps = torch.zeros((B, max_out, 12), dtype=torch.float64, device="cuda")
    sc = torch.zeros((B, max_out), dtype=torch.float32, device="cuda")
    po = torch.zeros((B, max_out, 3), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().hp_decode_nms(ctx.handle, tc.data_ptr(), tl.data_ptr(), t16.data_ptr(), t8.data_ptr(), B, size, size,
                                        float(opp.logit_threshold(score_thr)), float(np.float32(iou_thr)), max_out,
                                        cnt.data_ptr(), anc.data_ptr(), boxes.data_ptr(), kps.data_ptr(), sc.data_ptr(),
                                        po.data_ptr(), ctx.stream_ptr()))
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in (cnt, anc, boxes, kps, sc, po)]


@pytest.mark.parametrize("size,B,seed,sigma", [(128, 24, 7, 2.0), (96, 16, 8, 2.0), (88, 9, 9, 4.0), (128, 6, 10, 8.0)])
def test_fused_decode_nms_bit_exact(size, B, seed, sigma):
    cls, loc, p16, p8 = _stress(B, size, seed, sigma)
    # force exact score ties and saturated scores (tie-break by lower anchor index)
    cls[0, 10] = cls[0, 200] = cls[0, 5] = np.float32(3.25)
    cls[1, :40] = np.float32(30.0)
    cnt, anc, boxes, kps, sc, po = _run_fused(cls, loc, p16, p8, size)
    anchors = opp.blazeface_anchors(size)
    for i in range(B):
        ref = opp.detect_postprocess(cls[i], loc[i], p16[i], p8[i], anchors, 0.4, 0.3, input_size=size)
        k = len(ref["kept_anchor"])
        assert cnt[i] == k, (i, cnt[i], k)
        assert np.array_equal(anc[i, :k], ref["kept_anchor"]), i            # NMS kept-anchor indices bit-exact
        assert np.all(anc[i, k:] == -1)
        assert np.array_equal(boxes[i, :k], ref["boxes"])                   # float64 decode bit-exact
        assert np.array_equal(kps[i, :k].reshape(k, 6, 2), ref["keypoints"])
        assert np.array_equal(sc[i, :k], ref["scores"])                     # float32 sigmoid bit-exact
        assert np.array_equal(po[i, :k], ref["poses"])


def test_edge_cases_empty_single_and_cap():
    size = 128
    cls, loc, p16, p8 = _stress(4, size, 3)
    cls[0, :] = -20                      # nothing passes
    cls[1, :] = -20; cls[1, 777] = 2.0   # exactly one
    cls[2, :] = 5.0                      # everything passes with identical scores -> first 100 non-overlapping by index
    loc[2, :, 2:4] = 1.0                 # tiny boxes: no overlaps
    cnt, anc, boxes, kps, sc, po = _run_fused(cls, loc, p16, p8, size)
    anchors = opp.blazeface_anchors(size)
    assert cnt[0] == 0 and np.all(anc[0] == -1)
    assert cnt[1] == 1 and anc[1, 0] == 777
    for i in (2, 3):
        ref = opp.detect_postprocess(cls[i], loc[i], p16[i], p8[i], anchors, 0.4, 0.3, input_size=size)
        assert cnt[i] == len(ref["kept_anchor"]) and np.array_equal(anc[i, :cnt[i]], ref["kept_anchor"])
    assert cnt[2] == 100
    c5 = _run_fused(cls, loc, p16, p8, size, max_out=5)
    assert c5[0][2] == 5 and np.array_equal(c5[1][2], anc[2, :5])


def test_other_thresholds():
    size = 96
    cls, loc, p16, p8 = _stress(8, size, 21, 3.0)
    anchors = opp.blazeface_anchors(size)
    for st, it in ((0.7, 0.5), (0.1, 0.05), (0.5, 0.9)):
        cnt, anc, *_ = _run_fused(cls, loc, p16, p8, size, st, it)
        for i in range(8):
            ref = opp.detect_postprocess(cls[i], loc[i], p16[i], p8[i], anchors, st, it, input_size=size)
            assert cnt[i] == len(ref["kept_anchor"]) and np.array_equal(anc[i, :cnt[i]], ref["kept_anchor"]), (st, it, i)


def test_facade_steps_match_oracle_one_image():
    """filterDetections / extractDetections / filterWithNonMaxSupression individually (reference method split)."""
    import os
    from conftest import GOLDEN
    from helpers import unified_fixture
    from hpose_b200 import keras_spec as K
    from hpose_b200.blazeFaceDetectorH5 import blazeFaceDetector
    from hpose_b200.unified import UnifiedModel
    _, w = unified_fixture()
    u = UnifiedModel(w, K.load_model(os.path.join(GOLDEN, "heads", "stoqa9pt.h5")),
                     K.load_model(os.path.join(GOLDEN, "heads", "hrchr82r.h5")))
    det = blazeFaceDetector(model=u)
    assert len(det.anchors) == 896 and det.inputHeight == 128 and abs(det.sigmoidScoreThreshold - np.log(0.4 / 0.6)) < 1e-12
    cls, loc, p16, p8 = _stress(1, 128, 33)
    anchors = opp.blazeface_anchors(128)
    scores, good = det.filterDetections(cls[0])
    rs, rg = opp.filter_detections(cls[0], 0.4)
    assert np.array_equal(good, rg) and np.array_equal(scores, rs)
    boxes, kps = det.extractDetections(loc[0], good)
    rb, rk = opp.extract_detections(loc[0], rg, anchors, 128)
    assert boxes.dtype == np.float64 and np.array_equal(boxes, rb) and np.array_equal(kps, rk)
    res = det.filterWithNonMaxSupression(boxes, kps, scores, good, p16[0], p8[0])
    ref = opp.detect_postprocess(cls[0], loc[0], p16[0], p8[0], anchors)
    assert np.array_equal(res.boxes, ref["boxes"]) and np.array_equal(res.keypoints, ref["keypoints"])
    assert np.array_equal(res.scores, ref["scores"]) and np.array_equal(res.poses, ref["poses"])
    empty = det.filterWithNonMaxSupression(boxes[:0], kps[:0], scores[:0], good[:0], p16[0], p8[0])
    assert empty.poses.shape == (0, 3) and empty.boxes.shape == (0, 4)


def test_detect_faces_end_to_end_vs_oracle():
    """uint8 BGR image -> Results through detectFaces (reference flow) and detectFacesBatch (fused flow)."""
    import os
    from conftest import GOLDEN
    from helpers import unified_fixture
    from hpose_b200 import keras_spec as K
    from hpose_b200.blazeFaceDetectorH5 import blazeFaceDetector
    from hpose_b200.unified import UnifiedModel
    from oracle.keras_graph import KerasGraph, to_torch
    graph, w = unified_fixture()
    u = UnifiedModel(w, K.load_model(os.path.join(GOLDEN, "heads", "stoqa9pt.h5")),
                     K.load_model(os.path.join(GOLDEN, "heads", "hrchr82r.h5")))
    det = blazeFaceDetector(scoreThreshold=0.4, iouThreshold=0.3, model=u)
    rng = np.random.default_rng(4)
    imgs = rng.integers(0, 256, size=(3, 128, 128, 3), dtype=np.uint8)
    imgs[1, 30:100, 30:100] = 200          # a bright blob so that some anchors fire with trained weights
    batch = det.detectFacesBatch(imgs)
    single = [det.detectFaces(im) for im in imgs]
    for a, b in zip(batch, single):
        assert np.array_equal(a.boxes, b.boxes) and np.array_equal(a.scores, b.scores) and np.array_equal(a.poses, b.poses)
    # oracle: same logits (from the CUDA path) through the numpy post-processing, and float64 graph for the poses
    x = ((imgs[..., ::-1].astype(np.float64) / 255.0).astype(np.float32) - np.float32(0.5)) / np.float32(0.5)
    out = u(x)
    anchors = opp.blazeface_anchors(128)
    with torch.no_grad():
        o64 = KerasGraph(graph, to_torch(w, torch.float64))(torch.tensor(x, dtype=torch.float64))
    for i in range(3):
        cls = np.concatenate([out[0][i, :, 0], out[1][i, :, 0]])
        loc = np.concatenate([out[2][i], out[3][i]])
        ref = opp.detect_postprocess(cls, loc, out[4][i], out[5][i], anchors)
        assert np.array_equal(batch[i].boxes, ref["boxes"]) and np.array_equal(batch[i].poses, ref["poses"])
        ref64 = opp.detect_postprocess(cls, loc, o64[4][i].numpy(), o64[5][i].numpy(), anchors)
        if len(ref64["poses"]):
            assert np.abs(batch[i].poses - ref64["poses"]).max() < 0.01      # angles within 0.01 degree
    assert det.detectFaces(np.zeros((64, 64, 3), np.uint8)).boxes.shape[1:] == (4,)   # any frame size (bicubic resize on the device)


@pytest.mark.parametrize("chunks", [1, 2, 4])
def test_detect_stream_matches_detect_device(chunks):
    """The pipelined serving loop (copies overlapped with compute on side streams, every batch optionally cut into
    slices) returns, batch by batch and in order, exactly what the synchronous path returns.  64 crops per batch keep every
    slice (16 crops = 576 rows of the 6 x 6 map) on the same Dense kernel as the whole batch: below 512 rows a layer takes the
    CUDA-core kernel, whose last-bit differences reorder near-tied scores in the NMS of these random-weight detections."""
    import torch
    from hpose_b200 import keras_spec as K, train_88
    from hpose_b200.attention_model import se_transformer_regr_head
    from hpose_b200.blazeFaceDetectorH5 import blazeFaceDetector
    from hpose_b200.unified import UnifiedModel, random_backbone
    K.reset_names(); K.set_seed(11)
    head16 = train_88.create_model()
    K.reset_names()
    head8 = se_transformer_regr_head(input_channels=96)
    det = blazeFaceDetector(model=UnifiedModel(random_backbone(seed=4, bias_scale=0.1), head16, head8), inputSize=96)
    rng = np.random.default_rng(3)
    host = [torch.from_numpy(rng.integers(0, 256, size=(64, 96, 96, 3), dtype=np.uint8)).pin_memory() for _ in range(4)]
    want = []
    for hb in host:
        out = det.detect_device(hb.cuda())
        want.append({k: out[k].cpu().numpy().copy() for k in ("count", "boxes", "scores", "poses", "keypoints")})
    got = []
    for res in det.detect_stream(iter(host), chunks=chunks):
        got.append({k: res[k].numpy().copy() for k in res})
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert np.array_equal(g["count"], w["count"])
        for i, c in enumerate(w["count"]):                 # entries beyond count[i] are uninitialised padding
            for k in ("boxes", "scores", "poses", "keypoints"):
                assert np.array_equal(g[k][i, :c], w[k][i, :c]), (k, i)


def test_detect_stream_reuses_its_pinned_slots():
    """A second detect_stream call on the same detector reuses the pinned host result tensors of the first one (pinning ~60 MB
    per slot costs milliseconds: a serving loop must not pay it per call) and returns the same results; a generator started
    while another one is still running gets private slots."""
    import torch
    from hpose_b200 import keras_spec as K, train_88
    from hpose_b200.attention_model import se_transformer_regr_head
    from hpose_b200.blazeFaceDetectorH5 import blazeFaceDetector
    from hpose_b200.unified import UnifiedModel, random_backbone
    K.reset_names(); K.set_seed(11)
    head16 = train_88.create_model()
    K.reset_names()
    head8 = se_transformer_regr_head(input_channels=96)
    det = blazeFaceDetector(model=UnifiedModel(random_backbone(seed=4, bias_scale=0.1), head16, head8), inputSize=96)
    rng = np.random.default_rng(5)
    host = [torch.from_numpy(rng.integers(0, 256, size=(64, 96, 96, 3), dtype=np.uint8)).pin_memory() for _ in range(5)]
    runs, ptrs = [], []
    for _ in range(2):
        got, pp = [], set()
        for res in det.detect_stream(iter(host)):
            got.append({k: res[k].numpy().copy() for k in res})
            pp.add(res["boxes"].data_ptr())
        runs.append(got)
        ptrs.append(pp)
    assert ptrs[0] == ptrs[1] and len(ptrs[0]) == 3
    for a, b in zip(*runs):
        assert np.array_equal(a["count"], b["count"])
        for i, c in enumerate(a["count"]):
            for k in ("boxes", "scores", "poses", "keypoints"):
                assert np.array_equal(a[k][i, :c], b[k][i, :c]), (k, i)
    # two generators at the same time: the second one must not share the first one's slots
    g1, g2 = det.detect_stream(iter(host)), det.detect_stream(iter(host))
    r1, r2 = next(g1), next(g2)
    assert r1["boxes"].data_ptr() != r2["boxes"].data_ptr()
    assert np.array_equal(r1["count"].numpy(), r2["count"].numpy())
    g1.close(); g2.close()


def test_kept_anchors_against_fp64_oracle_logits():
    """north_star asks for bit-exact kept-anchor indices against the reference path.  Here the whole CUDA path (fp32 logits from
    the tensor-core backbone -> CUDA decode + NMS) is compared with the float64 oracle graph -> numpy post-processing on the
    trained weights: the kept anchor ids must agree on every frame.  (A score within float32 rounding of the threshold or of
    another score could legitimately flip; the frames are checked for such near-ties and none of these has one.)"""
    import os
    from conftest import GOLDEN
    from helpers import unified_fixture
    from hpose_b200 import keras_spec as K
    from hpose_b200.blazeFaceDetectorH5 import blazeFaceDetector
    from hpose_b200.unified import UnifiedModel
    from oracle.keras_graph import KerasGraph, to_torch
    graph, w = unified_fixture()
    u = UnifiedModel(w, K.load_model(os.path.join(GOLDEN, "heads", "stoqa9pt.h5")),
                     K.load_model(os.path.join(GOLDEN, "heads", "hrchr82r.h5")))
    det = blazeFaceDetector(scoreThreshold=0.4, iouThreshold=0.3, model=u)
    rng = np.random.default_rng(21)
    yy, xx = np.mgrid[0:128, 0:128]
    imgs = rng.integers(0, 90, size=(24, 128, 128, 3)).astype(np.float64)
    for i in range(24):                                   # skin-toned ellipses with dark "eye" spots: enough for anchors to fire
        for _ in range(1 + i % 3):
            cy, cx, ry, rx = rng.uniform(30, 98), rng.uniform(30, 98), rng.uniform(14, 40), rng.uniform(10, 32)
            face = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1
            imgs[i][face] = np.array([120, 150, 210]) + rng.normal(0, 6, 3)
            for ex in (-0.4, 0.4):
                eye = ((yy - (cy - 0.25 * ry)) / (0.12 * ry)) ** 2 + ((xx - (cx + ex * rx)) / (0.18 * rx)) ** 2 < 1
                imgs[i][eye] = 30
    imgs = np.clip(imgs, 0, 255).astype(np.uint8)
    got = det.detect_device(imgs)
    cnt, anc = got["count"].cpu().numpy(), got["anchors"].cpu().numpy()
    x = ((imgs[..., ::-1].astype(np.float64) / 255.0).astype(np.float32) - np.float32(0.5)) / np.float32(0.5)
    with torch.no_grad():
        o64 = [t.numpy() for t in KerasGraph(graph, to_torch(w, torch.float64))(torch.tensor(x, dtype=torch.float64))]
    anchors = opp.blazeface_anchors(128)
    thr = np.log(0.4 / 0.6)
    faces = mismatched = near = 0
    for i in range(len(imgs)):
        cls = np.concatenate([o64[0][i, :, 0], o64[1][i, :, 0]])
        loc = np.concatenate([o64[2][i], o64[3][i]])
        ref = opp.detect_postprocess(cls.astype(np.float32), loc.astype(np.float32), o64[4][i].astype(np.float32),
                                     o64[5][i].astype(np.float32), anchors)
        faces += len(ref["kept_anchor"])
        s = np.sort(cls[cls > thr - 1e-3])
        near_tie = (np.abs(cls - thr) < 2e-4).any() or (len(s) > 1 and np.diff(s).min() < 2e-5)
        near += int(near_tie)
        same = cnt[i] == len(ref["kept_anchor"]) and np.array_equal(anc[i, :cnt[i]], ref["kept_anchor"])
        if not same:
            mismatched += 1
            assert near_tie, (i, anc[i, :cnt[i]], ref["kept_anchor"])
    assert faces >= 10, faces                              # the probe frames do trigger the trained detector
    assert mismatched == 0, (mismatched, near, faces)
