"""Pins the oracle: known answers from the reference's shipped artefacts."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REFERENCE
from helpers import rel_err, unified_fixture
from oracle import postproc as opp
from oracle.keras_graph import KerasGraph, normalise_weight_names, same_pad, to_torch

SURVEY_MAE = {
    ("stoqa9pt", "BIWI_Test_Enlarged_features_88_0.7_1.npz"): 3.4456,
    ("stoqa9pt", "AFLW2000_Enlarged_features_88_0.7_1.npz"): 7.8100,
    ("stoqa9pt", "BIWI_Train_Enlarged_features_88_0.7_1.npz"): 2.7621,
    ("12uei1sn", "BIWI_Test_Enlarged_features_88_0.7_1.npz"): 4.1258,
    ("12uei1sn", "AFLW2000_Enlarged_features_88_0.7_1.npz"): 8.3235,
    ("hrchr82r", "AFLW2000_features_96_0.7_1.npz"): 8.0307,
}
SURVEY_ROW0 = {
    ("stoqa9pt", "BIWI_Test_Enlarged_features_88_0.7_1.npz"): [5.146871, 13.760605, 0.931063],
    ("hrchr82r", "AFLW2000_features_96_0.7_1.npz"): [6.357804, -30.201954, 1.647308],
    ("12uei1sn", "BIWI_Test_Enlarged_features_88_0.7_1.npz"): [4.658937, 13.879566, 2.218227],
}


def test_same_padding_rules():
    # SURVEY App. B.1
    assert same_pad(128, 5, 2) == (1, 2)
    assert same_pad(64, 3, 2) == (0, 1)
    assert same_pad(11, 3, 2) == (1, 1)
    assert same_pad(16, 3, 1) == (1, 1)
    assert same_pad(11, 2, 2) == (0, 1)
    assert same_pad(12, 2, 2) == (0, 0)


def test_anchor_known_answers():
    a = opp.blazeface_anchors(128)
    assert a.shape == (896, 4)
    assert tuple(a[0]) == (0.03125, 0.03125, 1.0, 1.0) and tuple(a[1]) == (0.03125, 0.03125, 1.0, 1.0)
    assert a[2, 0] == 0.09375
    assert tuple(a[511, :2]) == (0.96875, 0.96875)
    assert all(tuple(a[i, :2]) == (0.0625, 0.0625) for i in range(512, 518))
    assert tuple(a[895, :2]) == (0.9375, 0.9375)
    digest = hashlib.sha256(a.astype("<f8").tobytes()).hexdigest()
    assert digest == "d98e2ed7e8f24aa0856ffdcb52bc0e8439e50de2b0fedbfd8deaf11a4d2ab8ae"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout only exists in the build container")
def test_anchors_match_reference_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_blazeFaceUtils", f"{REFERENCE}/BlazePoser/blazeFaceUtils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    o = ref.SsdAnchorsCalculatorOptions(input_size_width=128, input_size_height=128, min_scale=0.1484375, max_scale=0.75,
                                        anchor_offset_x=0.5, anchor_offset_y=0.5, num_layers=4, feature_map_width=[],
                                        feature_map_height=[], strides=[8, 16, 16, 16], aspect_ratios=[1.0],
                                        reduce_boxes_in_lowest_layer=False, interpolated_scale_aspect_ratio=1.0,
                                        fixed_anchor_size=True)
    want = np.array([[x.x_center, x.y_center, x.h, x.w] for x in ref.gen_anchors(o)])
    assert np.array_equal(want, opp.blazeface_anchors(128))
    # the product's own generator, including the non-fixed-size and reduce_boxes variants
    from hpose_b200 import blazeFaceUtils as mine
    for kw in (dict(fixed_anchor_size=True), dict(fixed_anchor_size=False),
               dict(fixed_anchor_size=False, reduce_boxes_in_lowest_layer=True, aspect_ratios=[1.0, 2.0])):
        args = dict(input_size_width=96, input_size_height=128, min_scale=0.2, max_scale=0.9, num_layers=4,
                    feature_map_width=[], feature_map_height=[], strides=[8, 16, 16, 32], aspect_ratios=[1.0])
        args.update(kw)
        want = np.array([[x.x_center, x.y_center, x.h, x.w] for x in ref.gen_anchors(ref.SsdAnchorsCalculatorOptions(**args))])
        got = mine.anchor_table(mine.SsdAnchorsCalculatorOptions(**args))
        assert np.array_equal(want, got), kw


def test_head_known_answers_from_fixtures():
    """Float64 oracle on the committed dataset slices reproduces the stored predictions, and the stored
    whole-dataset MAEs equal the SURVEY's numbers."""
    from hpose_b200 import h5lite
    with open(os.path.join(GOLDEN, "kat_heads.json")) as f:
        kat = json.load(f)
    z = np.load(os.path.join(GOLDEN, "kat_heads.npz"))
    for (hid, ds), mae in SURVEY_MAE.items():
        assert abs(kat[hid]["datasets"][ds]["mae_avg"] - mae) < 5e-5, (hid, ds)
    for key, row in SURVEY_ROW0.items():
        assert np.allclose(kat[key[0]]["datasets"][key[1]]["row0"], row, atol=2e-6)
    assert kat["stoqa9pt"]["first_kernel_sha16"] == "57aff167d79098ca"
    assert kat["hrchr82r"]["first_kernel_sha16"] == "9b27355e781ace25"
    assert kat["12uei1sn"]["params"] == 42502 and kat["stoqa9pt"]["params"] == 5891 and kat["hrchr82r"]["params"] == 3683
    for hid in ("stoqa9pt", "12uei1sn", "hrchr82r"):
        hf = h5lite.H5File(os.path.join(GOLDEN, "heads", f"{hid}.h5"))
        w = normalise_weight_names(hf.weights())
        g = KerasGraph(hf.model_config(), to_torch(w, torch.float64))
        for ds in kat[hid]["datasets"]:
            feats, want = z[f"{hid}|{ds}|features"], z[f"{hid}|{ds}|pred64"]
            with torch.no_grad():
                got = g(torch.tensor(feats.reshape(len(feats), 1, 1, -1), dtype=torch.float64)).numpy().reshape(-1, 3)
            assert np.abs(got - want).max() < 1e-9


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout only exists in the build container")
def test_head_known_answers_full_datasets():
    from hpose_b200 import h5lite
    paths = {"stoqa9pt": f"{REFERENCE}/Model-88/Trained-Models-88/stoqa9pt.h5",
             "12uei1sn": f"{REFERENCE}/Model-88/Trained-Models-88/12uei1sn.h5",
             "hrchr82r": f"{REFERENCE}/Model-96/Trained-Models-96-ReshapedInput-NoFlatten/hrchr82r.h5"}
    for (hid, ds), mae in SURVEY_MAE.items():
        hf = h5lite.H5File(paths[hid])
        g = KerasGraph(hf.model_config(), to_torch(normalise_weight_names(hf.weights()), torch.float64))
        d = np.load(f"{REFERENCE}/FeatureMaps-Datasets/{ds}")
        with torch.no_grad():
            pred = g(torch.tensor(d["features"].reshape(len(d["features"]), 1, 1, -1), dtype=torch.float64)).numpy()
        got = np.abs(pred.reshape(-1, 3) - d["poses"]).mean()
        assert abs(got - mae) < 5e-5, (hid, ds, got)


def test_unified_graph_shapes_and_kat():
    """Six outputs in the JoinModels.py:152-158 order/shape; fp32 evaluation stays within 3e-6 of fp64."""
    graph, w = unified_fixture()
    kat = np.load(os.path.join(GOLDEN, "unified_kat.npz"))
    x = kat["x"]
    g64 = KerasGraph(graph, to_torch(w, torch.float64))
    with torch.no_grad():
        o = g64(torch.tensor(x[:1], dtype=torch.float64))
    assert [tuple(t.shape) for t in o] == [(1, 512, 1), (1, 384, 1), (1, 512, 16), (1, 384, 16), (1, 16, 16, 3), (1, 8, 8, 3)]
    for t, name in zip(o, ("cls16", "cls8", "loc16", "loc8", "pose16", "pose8")):
        assert np.abs(t.numpy() - kat[name][:1]).max() < 1e-9, name
    g32 = KerasGraph(graph, to_torch(w, torch.float32))
    with torch.no_grad():
        o32 = g32(torch.tensor(x[:1], dtype=torch.float32))
    for a, b in zip(o32, o):
        assert rel_err(a.numpy(), b.numpy()) < 3e-6
    assert sum(v.size for v in w.values()) == 110964


def test_oracle_conv_matches_direct_numpy_loops():
    """The torch-based conv/pool restatement equals a naive numpy implementation of TF SAME semantics."""
    from oracle.keras_graph import _nhwc_conv
    rng = np.random.default_rng(5)
    for (h, w_, k, s, groups, cin, cout) in [(11, 9, 3, 2, 4, 4, 4), (8, 8, 3, 1, 6, 6, 6), (10, 12, 5, 2, 1, 3, 5), (7, 7, 1, 1, 1, 5, 4)]:
        x = rng.normal(size=(2, h, w_, cin))
        ker = rng.normal(size=(k, k, 1 if groups > 1 else cin, cout))
        b = rng.normal(size=(cout,))
        got = _nhwc_conv(torch.tensor(x), torch.tensor(ker), torch.tensor(b), (s, s), "same", groups).numpy()
        pt, _ = same_pad(h, k, s)
        pl, _ = same_pad(w_, k, s)
        ho, wo = -(-h // s), -(-w_ // s)
        want = np.zeros((2, ho, wo, cout))
        for n in range(2):
            for oy in range(ho):
                for ox in range(wo):
                    for co in range(cout):
                        acc = b[co]
                        for ky in range(k):
                            for kx in range(k):
                                iy, ix = oy * s - pt + ky, ox * s - pl + kx
                                if 0 <= iy < h and 0 <= ix < w_:
                                    if groups > 1:
                                        acc += x[n, iy, ix, co] * ker[ky, kx, 0, co]
                                    else:
                                        acc += (x[n, iy, ix, :] * ker[ky, kx, :, co]).sum()
                        want[n, oy, ox, co] = acc
        assert np.abs(got - want).max() < 1e-12


def _synthetic_probe_images(n=48, size=128, seed=0):
    """Noise, gratings, flat fields, blobs, blocks: a spread of inputs, none of them a face."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size] / (size - 1.0)
    imgs = []
    for i in range(n):
        k = i % 6
        if k == 0:
            im = rng.uniform(-1, 1, (size, size, 3))
        elif k == 1:
            im = np.stack([np.sin(xx * rng.uniform(1, 20) + rng.uniform(0, 6)), np.cos(yy * rng.uniform(1, 20)), xx * yy * 2 - 1], -1)
        elif k == 2:
            im = np.full((size, size, 3), rng.uniform(-1, 1)) + rng.normal(0, 0.05, (size, size, 3))
        elif k == 3:
            im = rng.uniform(-1, 0, (size, size, 3))
            cy, cx, r = rng.uniform(.3, .7), rng.uniform(.3, .7), rng.uniform(.1, .4)
            im += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r))[..., None] * rng.uniform(0.5, 2, 3)
        elif k == 4:
            im = rng.uniform(-1, 1, (8, 8, 3)).repeat(size // 8, 0).repeat(size // 8, 1)
        else:
            im = np.clip(rng.normal(0, 1, (size, size, 3)), -1, 1)
        imgs.append(np.clip(im, -1, 1))
    return np.stack(imgs).astype(np.float32)


def test_backbone_oracle_reproduces_the_channel_usage_of_the_shipped_feature_maps():
    """A pin for the BACKBONE restatement (the reference ships no backbone test and TensorFlow cannot run): the shipped npz
    data sets are real re_lu_15 outputs of the reference's trained backbone.  Channels that never fire in any
    of them (17 of 96) must be nearly silent in the oracle's re_lu_15 on the same trained weights, channels that fire often
    there must fire often here, and the per-channel firing rates must correlate -- a mis-ordered channel, a wrong padding
    rule or a wrong skip connection in the restatement destroys all three.  (The inputs differ -- synthetic probes, not
    face crops -- so the comparison is statistical; measured: Spearman 0.62, 0.12 against 0.78 mean firing rate.)"""
    from scipy.stats import spearmanr
    with open(os.path.join(GOLDEN, "tap_channel_stats.json")) as f:
        stats = json.load(f)
    fd = np.array(stats["96"]["nonzero_fraction"])
    dead, alive = np.where(fd == 0)[0], np.where(fd > 0.3)[0]
    assert len(dead) == 17 and len(alive) >= 30
    for e in stats["96"]["files"]:                       # 23-29 dead channels in each 96-channel set
        assert 23 <= len(e["dead"]) <= 29 and set(dead) <= set(e["dead"])
    for e in stats["88"]["files"]:                       # ... and 2-3 in each 88-channel set
        assert 2 <= len(e["dead"]) <= 3
    graph, w = unified_fixture()
    g = KerasGraph(graph, to_torch(w, torch.float32))
    taps = {}
    with torch.no_grad():
        g(torch.from_numpy(_synthetic_probe_images()), taps=taps)
    t8 = taps["re_lu_15"].numpy()
    assert t8.shape[1:] == (8, 8, 96) and taps["re_lu_10"].shape[1:] == (16, 16, 88) and t8.min() >= 0
    fo = (t8 > 0).reshape(-1, 96).mean(0)
    assert fo[dead].mean() < 0.2 and fo[alive].mean() > 0.65, (fo[dead].mean(), fo[alive].mean())
    assert spearmanr(fd, fo).correlation > 0.5
    # channels silent in the data and in the oracle
    assert set(np.where(fo == 0)[0]) <= set(np.where(fd < 0.02)[0])


def test_new_activations_match_their_definitions():
    from oracle.keras_graph import activation
    x = torch.tensor([-3.0, -0.5, 0.0, 0.25, 2.0], dtype=torch.float64)
    a, s = 1.6732632423543772, 1.0507009873554805
    np.testing.assert_allclose(activation("elu", x).numpy(), [np.expm1(-3), np.expm1(-0.5), 0, 0.25, 2.0])
    np.testing.assert_allclose(activation("selu", x).numpy(), [s * a * np.expm1(-3), s * a * np.expm1(-0.5), 0, s * 0.25, s * 2.0])
    np.testing.assert_allclose(activation("softplus", x).numpy(), np.log1p(np.exp(x.numpy())))
    np.testing.assert_allclose(activation("swish", x).numpy(), x.numpy() / (1 + np.exp(-x.numpy())))
    np.testing.assert_allclose(activation("leaky_relu", x).numpy(), [-0.6, -0.1, 0, 0.25, 2.0])


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout only exists in the build container")
def test_zoo_checkpoints_with_new_activations_load_and_evaluate():
    """The 17 zoo checkpoints round 1 refused for their activation (elu, selu, softplus, swish, leaky_relu) now compile;
    every regressor checkpoint of the reference either loads or is refused for a layer class outside SURVEY 8a."""
    import glob
    from hpose_b200 import keras_spec as K
    files = sorted(f for f in glob.glob(f"{REFERENCE}/Model-*/**/*.h5", recursive=True))
    assert len(files) >= 600
    loaded, refused = 0, {}
    for f in files:
        try:
            m = K.load_model(f)
            loaded += 1
        except ValueError as e:
            refused[f] = str(e)
    assert not [v for v in refused.values() if "activation" in v and "unsupported" in v], refused
    assert loaded >= 665 and all(("layer class" in v or "1x1" in v or "per-pixel" in v) for v in refused.values()), (loaded, set(refused.values()))
