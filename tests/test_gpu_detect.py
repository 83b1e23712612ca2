"""GPU tests of the frames -> faces path: on-device bicubic resize, the single-call / CUDA-graph latency
path and the packed results (8f-4), the serving loop, and independence of the results from the batching."""
import os

import numpy as np
import pytest
import torch

from hpose_b200 import _lib

from oracle import postproc as opp
from oracle import preprocess as opre

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _trained_detector(**kw):
    from helpers import unified_fixture
    from hpose_b200 import keras_spec as K
    from hpose_b200.blazeFaceDetectorH5 import blazeFaceDetector
    from hpose_b200.unified import UnifiedModel
    _, w = unified_fixture()
    u = UnifiedModel(w, K.load_model(os.path.join(GOLDEN, "heads", "stoqa9pt.h5")),
                     K.load_model(os.path.join(GOLDEN, "heads", "hrchr82r.h5")))
    return blazeFaceDetector(model=u, **kw), u


def _random_detector(size=96, seed=4):
    from hpose_b200 import keras_spec as K, train_88
    from hpose_b200.attention_model import se_transformer_regr_head
    from hpose_b200.blazeFaceDetectorH5 import blazeFaceDetector
    from hpose_b200.unified import UnifiedModel, random_backbone
    K.reset_names(); K.set_seed(11)
    head16 = train_88.create_model()
    K.reset_names()
    head8 = se_transformer_regr_head(input_channels=96)
    return blazeFaceDetector(model=UnifiedModel(random_backbone(seed=seed, bias_scale=0.1), head16, head8), inputSize=size)


def _frames(rng, n, h, w):
    """Smooth blobs + noise: bicubic resampling of pure noise is all overshoot, real frames are not."""
    yy, xx = np.mgrid[0:h, 0:w]
    out = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        img = rng.integers(0, 80, size=(h, w, 3)).astype(np.float64)
        for _ in range(3):
            cy, cx, r = rng.uniform(0.2, 0.8) * h, rng.uniform(0.2, 0.8) * w, rng.uniform(0.1, 0.3) * min(h, w)
            img += 150.0 * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r))[..., None] * rng.uniform(0.3, 1.0, size=3)
        out[i] = np.clip(img, 0, 255).astype(np.uint8)
    return out


@pytest.mark.parametrize("hin,win,hout,wout", [(480, 640, 128, 128), (64, 64, 128, 128), (97, 131, 96, 96), (128, 128, 128, 128),
                                               (1080, 1920, 128, 128), (5, 3, 88, 88), (300, 200, 128, 96)])
def test_resize_preprocess_matches_oracle_bit_for_bit(hin, win, hout, wout):
    """hp_preprocess_resize_u8 against oracle/preprocess.py (blazeFaceDetectorH5.py:247-269 with TensorFlow's ResizeBicubic):
    the two evaluate the same float32 operations in the same order, so the comparison is exact (tolerance asked: 1e-5)."""
    from hpose_b200.device import default_context
    ctx = default_context()
    rng = np.random.default_rng(hin * 7 + win)
    frames = rng.integers(0, 256, size=(3, hin, win, 3), dtype=np.uint8)
    u8 = torch.from_numpy(frames).cuda()
    x = torch.full((3, hout, wout, 3), float("nan"), device="cuda")
    _lib.check(_lib.lib().hp_preprocess_resize_u8(ctx.handle, u8.data_ptr(), 3, hin, win, hout, wout, x.data_ptr(), ctx.stream_ptr()))
    got = x.cpu().numpy()
    for i in range(3):
        want = opre.prepare_input(frames[i], hout, wout)[0]
        assert np.abs(got[i] - want).max() <= 1e-5
        assert np.array_equal(got[i], want)


def test_detect_faces_on_frames_of_any_size():
    """detectFaces (one H2D, one graph launch, one D2H) == detectFacesStepwise (the reference's five calls) == the oracle's
    preprocessing + numpy post-processing of the same logits, on frames that need the bicubic resize; repeated calls replay
    the captured graph and return the same faces."""
    det, u = _trained_detector()
    rng = np.random.default_rng(5)
    frames = _frames(rng, 3, 240, 320)
    frames[1, 60:200, 90:230] = 210
    anchors = opp.blazeface_anchors(128)
    n_faces = 0
    for f in frames:
        first = det.detectFaces(f)
        step = det.detectFacesStepwise(f)
        x = opre.prepare_input(f, 128, 128)
        assert np.array_equal(det.prepareInputForInference(f), x)
        out = u(x)
        cls = np.concatenate([out[0][0, :, 0], out[1][0, :, 0]])
        loc = np.concatenate([out[2][0], out[3][0]])
        ref = opp.detect_postprocess(cls, loc, out[4][0], out[5][0], anchors)
        for r in (first, step):
            assert np.array_equal(r.boxes, ref["boxes"]) and np.array_equal(r.keypoints, ref["keypoints"])
            assert np.array_equal(r.scores, ref["scores"]) and np.array_equal(r.poses, ref["poses"])
        assert first.boxes.dtype == np.float64 and first.keypoints.shape[1:] == (6, 2) and first.poses.dtype == np.float32
        n_faces += len(first.scores)
    before = det.ctx.launch_count()
    again = [det.detectFaces(f) for f in frames]          # graph replays (captured on the second call above)
    assert det.ctx.launch_count() > before
    for f, r in zip(frames, again):
        s = det.detectFacesStepwise(f)
        assert np.array_equal(r.boxes, s.boxes) and np.array_equal(r.scores, s.scores) and np.array_equal(r.poses, s.poses)
    small = det.detectFaces(np.zeros((64, 64, 3), np.uint8))          # up-scaling path; the reference accepts any size too
    assert small.boxes.shape[1:] == (4,)
    with pytest.raises(ValueError):
        det.detectFaces(np.zeros((64, 64), np.uint8))


def test_packed_batch_equals_padded_batch_and_single_frames():
    det = _random_detector(96)
    rng = np.random.default_rng(9)
    frames = rng.integers(0, 256, size=(5, 120, 160, 3), dtype=np.uint8)
    packed = det.detectFacesBatch(frames)                               # hp_detect_frames, packed records
    padded = det.detectFacesBatch(torch.from_numpy(frames).cuda())      # hp_unified_forward, [B, 100, ...]
    assert len(packed) == 5
    for a, b in zip(packed, padded):
        assert len(a.scores) > 0
        assert np.array_equal(a.boxes, b.boxes) and np.array_equal(a.keypoints, b.keypoints)
        assert np.array_equal(a.scores, b.scores) and np.array_equal(a.poses, b.poses)
    one = det.detectFaces(frames[2])
    assert np.array_equal(one.boxes, packed[2].boxes) and np.array_equal(one.scores, packed[2].scores)
    assert np.array_equal(one.poses, packed[2].poses)


def test_result_capacity_truncates_without_overrun():
    """A result buffer smaller than sum(count): header.total reports what was found, header.written what fits."""
    det = _random_detector(96)
    L = _lib.lib()
    rng = np.random.default_rng(2)
    frames = torch.from_numpy(rng.integers(0, 256, size=(4, 96, 96, 3), dtype=np.uint8)).cuda()
    m = det.interpreter
    cap = 150
    nbytes = int(L.hp_detect_result_bytes(4, cap))
    res = torch.full(((nbytes + 7) // 8 + 64,), -1, dtype=torch.int64, device="cuda")
    _lib.check(L.hp_detect_frames(det.ctx.handle, m.head16.head_handle, m.head8.head_handle, frames.data_ptr(), 4, 96, 96, 96, 96,
                                  float(np.float32(det.sigmoidScoreThreshold)), 0.3, 100, res.data_ptr(), nbytes, 0, det.ctx.stream_ptr()))
    raw = res.cpu().numpy().view(np.uint8)
    hdr = raw[:32].view(np.int32)
    assert hdr[0] == hdr[4:8].sum() and hdr[0] > cap and hdr[1] == cap and hdr[2] == 4 and hdr[3] == cap
    assert np.all(res.cpu().numpy()[(nbytes + 7) // 8:] == -1)          # nothing written behind the buffer
    with pytest.raises(_lib.HposeError):
        L_small = 8
        _lib.check(L.hp_detect_frames(det.ctx.handle, m.head16.head_handle, m.head8.head_handle, frames.data_ptr(), 4, 96, 96, 96, 96,
                                      0.0, 0.3, 100, res.data_ptr(), L_small, 0, det.ctx.stream_ptr()))


def test_detect_stream_packed_and_buffer_lifetime():
    """The packed serving loop returns what the synchronous packed call returns, and a yielded result stays intact while the
    next TWO are produced (three rotating slots; ADVICE round 1: the two-slot version overwrote a result one step later)."""
    from hpose_b200.blazeFaceDetectorH5 import unpack_results
    det = _random_detector(96)
    rng = np.random.default_rng(3)
    host = [torch.from_numpy(rng.integers(0, 256, size=(16, 96, 96, 3), dtype=np.uint8)).pin_memory() for _ in range(6)]
    want = [det.detectFacesBatch(hb.numpy()) for hb in host]
    held, snapshots = [], []
    for i, res in enumerate(det.detect_stream(iter(host), packed=True)):
        assert res["total"] == int(res["count"].sum()) == len(res["faces"])
        held.append(res)
        snapshots.append({k: np.array(res[k], copy=True) for k in ("count", "faces")})
        for j in (i - 1, i - 2):                               # the two previous results are still untouched
            if j >= 0:
                assert np.array_equal(held[j]["count"], snapshots[j]["count"])
                assert np.array_equal(held[j]["faces"], snapshots[j]["faces"])
    assert len(snapshots) == 6
    for s, w in zip(snapshots, want):
        o = 0
        for c, r in zip(s["count"], w):
            f = s["faces"][o:o + c]
            assert np.array_equal(f["box"], r.boxes) and np.array_equal(f["score"], r.scores) and np.array_equal(f["pose"], r.poses)
            assert np.array_equal(f["keypoints"], r.keypoints)
            o += c
    # padded loop, three slots as well
    held, snaps = [], []
    for i, res in enumerate(det.detect_stream(iter(host))):
        held.append(res)
        snaps.append({k: res[k].numpy().copy() for k in res})
        for j in (i - 1, i - 2):
            if j >= 0:
                assert all(np.array_equal(held[j][k].numpy(), snaps[j][k]) for k in snaps[j])


def test_results_do_not_depend_on_the_batching():
    """Every layer runs the same kernel, row by row, whatever the batch size (the Dense layers no longer switch kernels below
    512 rows): a crop's faces are bit-identical whether it is processed alone, in a batch of 3 or in a batch of 70."""
    det = _random_detector(96, seed=8)
    rng = np.random.default_rng(12)
    frames = rng.integers(0, 256, size=(70, 96, 96, 3), dtype=np.uint8)
    big = det.detectFacesBatch(frames)
    small = det.detectFacesBatch(frames[:3])
    for i in range(3):
        one = det.detectFacesBatch(frames[i:i + 1])[0]
        for r in (small[i], one):
            assert np.array_equal(r.boxes, big[i].boxes) and np.array_equal(r.scores, big[i].scores)
            assert np.array_equal(r.poses, big[i].poses) and np.array_equal(r.keypoints, big[i].keypoints)
    # raw network outputs too
    x = torch.rand((70, 96, 96, 3), device="cuda") * 2 - 1
    a = det.interpreter.forward_device(x)
    b = det.interpreter.forward_device(x[:1].contiguous())
    for k in ("cls", "loc", "pose16", "pose8", "feat16", "feat8"):
        assert torch.equal(a[k][:1], b[k]), k


def test_bench_line_contract():
    """The default arm of bench.py prints one JSON line with the keys the driver and the judge read (small batch, few steps)."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "3", "--warmup", "3", "--batch", "256", "--no-extras"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["unit"] == "crops/s" and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["gpu_launches"] > 0 and "workload" in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 256 * 96 * 96 * 3 and e["d2h_bytes_per_step"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0
    assert d["clocks"]["sm_max_mhz"] > 0 and isinstance(d["clocks"]["reasons"], list)
