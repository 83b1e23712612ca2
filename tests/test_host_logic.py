"""CPU tests of the host layer: spec compiler, Keras-format IO, builders, ABI surface, callbacks."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from helpers import head_oracle, rel_err, unified_fixture
from oracle.keras_graph import KerasGraph, to_torch


def test_library_builds_and_exports_every_declared_symbol(built_lib):
    from hpose_b200 import _lib
    header = open(os.path.join(ROOT, "include", "hpose.h")).read()
    declared = set(re.findall(r"\b(hp_[a-z0-9_]+)\s*\(", header))
    declared -= {"hp_status", "hp_impl"}
    lib = ctypes.CDLL(built_lib)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/hpose.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    # the dropout hash is host-callable without a GPU and must be a pure function
    lib.hp_dropout_hash.restype = ctypes.c_uint32
    lib.hp_dropout_hash.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
    a = lib.hp_dropout_hash(42, 0, 1, 2, 3)
    assert a == lib.hp_dropout_hash(42, 0, 1, 2, 3) and a != lib.hp_dropout_hash(42, 1, 1, 2, 3)


def test_no_gpu_means_loud_failure(built_lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hpose_b200 import _lib
    from hpose_b200.device import default_context
    with pytest.raises(_lib.HposeError):
        default_context()
    h = ctypes.c_void_p()
    rc = _lib.lib().hp_create(0, ctypes.byref(h))
    assert rc != 0 and b"no CPU fallback" in _lib.lib().hp_last_error()


def test_builders_match_reference_parameter_counts():
    from hpose_b200 import attention_model as am, keras_spec as K, train_88, train_96
    K.reset_names()
    assert am.se_transformer_regr_head().count_params() == 47328
    assert am.se_transformer_regr_head(88, 8, 4, 16, 64, 64).count_params() == 42502   # == 12uei1sn.h5
    assert train_88.create_model().count_params() == 5891                  # == stoqa9pt.h5
    assert am.create_modelC().count_params() == 979 + 1056 + 3738 + 129
    train_96.config.update(num_filters=64, dropout_rate=0.0, regularizer_rate=1e-5)
    m = train_96.create_model()
    assert m.count_params() == 96 * 64 + 64 + 64 * 3 + 3 and m.optimizer.kind == "adam"
    train_96.config.update(num_filters=-1)
    with pytest.raises(ValueError):
        train_96.create_model()
    train_96.config.update(num_filters=64)


def test_traced_config_is_keras_format_and_evaluates_in_oracle():
    from hpose_b200 import attention_model as am, keras_spec as K
    K.reset_names(); K.set_seed(3)
    m = am.se_transformer_regr_head(input_channels=24, reduction=4, num_heads=2, key_dim=4, ff_dim=8, hidden_channels=8)
    cfg = json.loads(m.to_json())
    names = [l["class_name"] for l in cfg["config"]["layers"]]
    assert names.count("MultiHeadAttention") == 1 and names.count("LayerNormalization") == 2 and names[0] == "InputLayer"
    g, _ = head_oracle(m)
    x = torch.randn(2, 3, 5, 24, dtype=torch.float64)
    with torch.no_grad():
        y = g(x)
    assert tuple(y.shape) == (2, 3, 5, 3)
    # flat parameter packing is a bijection
    flat = m.get_flat_weights()
    assert np.array_equal(m.program.pack(m.program.unpack(flat)), flat)


def test_h5_roundtrip_and_loader_errors(tmp_path):
    from hpose_b200 import keras_spec as K, train_88
    K.reset_names(); K.set_seed(11)
    m = train_88.create_model_skip_fc()
    m.compile(optimizer=K.SGD(learning_rate=2.8e-4), loss="mse", metrics=["mae"])
    p = str(tmp_path / "m.h5")
    m.save(p)
    m2 = K.load_model(p)
    assert m2.count_params() == m.count_params() and m2.optimizer.kind == "sgd"
    assert np.array_equal(m2.get_flat_weights(), m.get_flat_weights())
    assert abs(m2.optimizer.learning_rate - 2.8e-4) < 1e-9
    with pytest.raises(FileNotFoundError):
        K.load_model(str(tmp_path / "missing.h5"))
    # shipped checkpoints load and agree with the oracle's reading of the same file
    for hid, n in (("stoqa9pt", 5891), ("hrchr82r", 3683), ("12uei1sn", 42502)):
        mm = K.load_model(os.path.join(GOLDEN, "heads", f"{hid}.h5"))
        assert mm.count_params() == n


def test_generated_detector_graph_equals_shipped_graph():
    """unified.blazeface_graph_config reproduces the reference's unified model_config semantics."""
    from hpose_b200 import keras_spec as K
    from hpose_b200.unified import UnifiedModel, blazeface_graph_config, pack_backbone, unpack_backbone
    graph, w = unified_fixture()
    heads = {l["name"]: l for l in graph["config"]["layers"] if l["class_name"] == "Functional"}
    gen = blazeface_graph_config({"class_name": "Functional", "config": heads["model"]["config"]},
                                 {"class_name": "Functional", "config": heads["model_10"]["config"]})
    ref_names = [(l["class_name"], l["name"]) for l in graph["config"]["layers"]]
    gen_names = [(l["class_name"], l["name"]) for l in gen["config"]["layers"]]
    assert sorted(ref_names) == sorted(gen_names)
    x = torch.tensor(np.load(os.path.join(GOLDEN, "unified_kat.npz"))["x"][:1], dtype=torch.float64)
    wt = to_torch(w, torch.float64)
    with torch.no_grad():
        a = KerasGraph(graph, wt)(x)
        b = KerasGraph(gen, wt)(x)
    for s, t in zip(a, b):
        assert torch.equal(s, t)
    flat = pack_backbone(w)
    assert flat.size == 101390
    back = unpack_backbone(flat)
    assert all(np.array_equal(back[k], w[k]) for k in back)
    bad = dict(w); bad["conv2d_3/kernel"] = np.zeros((1, 1, 28, 31), np.float32)
    with pytest.raises(ValueError):
        pack_backbone(bad)


def test_unified_model_save_load(tmp_path):
    from hpose_b200 import keras_spec as K
    from hpose_b200.unified import UnifiedModel
    _, w = unified_fixture()
    h16 = K.load_model(os.path.join(GOLDEN, "heads", "stoqa9pt.h5"))
    h8 = K.load_model(os.path.join(GOLDEN, "heads", "hrchr82r.h5"))
    # nested head weights inside the unified file are bit-identical to the stand-alone checkpoints
    for k, v in h16.get_weights_dict().items():
        assert np.array_equal(v, w[f"model/{k}"])
    for k, v in h8.get_weights_dict().items():
        assert np.array_equal(v, w[f"model_10/{k}"])
    u = UnifiedModel(w, h16, h8)
    assert u.count_params() == 110964
    p = str(tmp_path / "unified.h5")
    u.save(p)
    u2 = UnifiedModel.load(p)
    assert np.array_equal(u2.backbone_flat, u.backbone_flat)
    assert np.array_equal(u2.head16.get_flat_weights(), h16.get_flat_weights())
    assert np.array_equal(u2.head8.get_flat_weights(), h8.get_flat_weights())
    with pytest.raises(FileNotFoundError):
        UnifiedModel.load(str(tmp_path / "nope.h5"))
    with pytest.raises(ValueError):
        UnifiedModel(w, h8, h16)


def test_join_models_errors(tmp_path):
    from hpose_b200.JoinModels import extract_id_from_path, join_models
    with pytest.raises(FileNotFoundError):
        join_models("a.h5", "b.h5", "c.h5", "re_lu_10", "re_lu_15", None)
    s = os.path.join(GOLDEN, "heads", "stoqa9pt.h5")
    with pytest.raises(ValueError):
        join_models(s, s, s, "re_lu_99", "re_lu_15", None)
    with pytest.raises(ValueError):     # a head checkpoint is not a BlazeFace detector
        join_models(s, s, s, "re_lu_10", "re_lu_15", None)
    assert extract_id_from_path("/x/y/stoqa9pt.h5") == "stoqa9pt" and extract_id_from_path("x.txt") is None


def test_spec_compiler_rejects_unsupported_graphs():
    from hpose_b200 import keras_spec as K
    K.reset_names()
    with pytest.raises(ValueError):
        K.Conv2D(8, kernel_size=3)
    x = K.Input((None, None, 8))
    with pytest.raises(ValueError):
        K.MultiHeadAttention(2, 4, value_dim=8)
    y = K.Conv2D(3, 1, activation="gelu")(x)                # outside the reference's zoo (relu, tanh, softsign, elu, selu, ...)
    with pytest.raises(ValueError):
        K.Model(x, y)
    assert K.Model(x, K.Conv2D(3, 1, activation="selu")(x)).count_params() == 27


def test_callbacks_follow_keras_rules():
    from hpose_b200 import keras_spec as K

    class Fake:
        stop_training = False
        w = 0
        saved = []

        def get_flat_weights(self): return self.w
        def set_flat_weights(self, w): self.w = w
        def save(self, p): self.saved.append(p)

    m = Fake()
    es = K.EarlyStopping(monitor="val_loss", patience=2, min_delta=1e-3, restore_best_weights=True)
    ck = K.ModelCheckpoint("best.h5", monitor="val_loss", save_best_only=True)
    for cb in (es, ck):
        cb.set_model(m); cb.on_train_begin()
    vals = [1.0, 0.9, 0.8995, 0.8999, 0.95]     # improvements < min_delta do not reset patience
    for e, v in enumerate(vals):
        m.w = e
        for cb in (es, ck):
            cb.on_epoch_end(e, {"val_loss": v})
        if m.stop_training:
            break
    assert m.stop_training and e == 3 and m.w == 1          # restored to the epoch-1 weights
    assert len(m.saved) == 3                                 # 1.0, 0.9, 0.8995 were strict improvements


def test_npz_contract(tmp_path):
    from hpose_b200.utilities import load_dataset, load_dataset_with_weights, save_dataset
    f = np.abs(np.random.default_rng(0).normal(size=(9, 96))).astype(np.float32)
    p = np.array([[0, 0, 0], [70, 0, 0], [0, 65, 3]] * 3, dtype=np.float64)
    path = str(tmp_path / "d.npz")
    save_dataset(path, f, p)
    ff, pp = load_dataset(path)
    assert ff.dtype == np.float32 and pp.dtype == np.float64 and ff.shape == (9, 96) and pp.shape == (9, 3)
    d = load_dataset_with_weights(path)
    assert d["weights"][0] == 1.0 and abs(d["weights"][1] - 0.5 ** 2) < 1e-12 and abs(d["weights"][2] - 0.5) < 1e-9
    with pytest.raises(FileNotFoundError):
        load_dataset(str(tmp_path / "none.npz"))


def test_shard_bounds_cover_everything():
    from hpose_b200.parallel import shard_bounds
    for n in (0, 1, 7, 128, 4096, 4099):
        for world in (1, 2, 4, 8):
            cuts = [shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_join_models_happy_path(tmp_path):
    """join_models (reference JoinModels.py:5-90): detector .h5 + two regressor .h5 -> unified model with the six outputs in
    the reference order, saved as a Keras-format .h5 that loads back bit-identically and carries the shipped unified graph."""
    from hpose_b200 import h5lite
    from hpose_b200.JoinModels import join_models
    from hpose_b200.unified import UnifiedModel, backbone_weight_specs
    graph, w = unified_fixture()
    det_path = str(tmp_path / "face_detector.h5")
    det_w = {}
    for key, _ in backbone_weight_specs():
        layer, var = key.split("/", 1)
        det_w[f"{layer}/{layer}/{var}:0"] = w[key]
    det_cfg = {"class_name": "Functional", "config": {"name": "blazeface_front", "layers": [], "input_layers": [], "output_layers": []}}
    h5lite.write_h5(det_path, det_w, det_cfg)
    r1 = os.path.join(GOLDEN, "heads", "stoqa9pt.h5")
    r2 = os.path.join(GOLDEN, "heads", "hrchr82r.h5")
    out = str(tmp_path / "reg1-stoqa9pt-reg2-hrchr82r.h5")
    u = join_models(det_path, r1, r2, "re_lu_10", "re_lu_15", out, metadata={"note": "test"})
    assert u._metadata == {"note": "test"} and os.path.exists(out) and not os.path.exists(out + ".tmp")
    assert u.count_params() == 110964                                 # the shipped unified model
    back = UnifiedModel.load(out)
    assert np.array_equal(back.backbone_flat, u.backbone_flat)
    assert np.array_equal(back.head16.get_flat_weights(), u.head16.get_flat_weights())
    assert np.array_equal(back.head8.get_flat_weights(), u.head8.get_flat_weights())
    cfg = h5lite.H5File(out).model_config()
    assert [o[0] for o in cfg["config"]["output_layers"]] == [
        "tf_op_layer_classificators_1", "tf_op_layer_classificators_2", "tf_op_layer_regressors_1", "tf_op_layer_regressors_2",
        "model", "model_10"]                                          # JoinModels.py:152-158 / blazeFaceDetectorH5.py:273-278
    # the joined graph evaluates, in the oracle, to what the reference's own unified graph gives on the same weights
    x = torch.tensor(np.load(os.path.join(GOLDEN, "unified_kat.npz"))["x"][:1], dtype=torch.float64)
    wj = {}
    for k, v in h5lite.H5File(out).weights().items():
        parts = k.split("/")
        parts[-1] = parts[-1].split(":")[0]
        if len(parts) >= 3 and parts[0] == parts[1]:
            parts = parts[1:]
        wj["/".join(parts)] = v
    with torch.no_grad():
        a = KerasGraph(graph, to_torch(w, torch.float64))(x)
        b = KerasGraph(cfg, to_torch(wj, torch.float64))(x)
    for s, t in zip(a, b):
        assert torch.equal(s, t)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout only exists in the build container")
def test_real_unified_h5_files_load():
    """Every unified .h5 the reference ships loads through h5lite + UnifiedModel.load; the selected one equals the fixture."""
    import glob
    from hpose_b200.unified import UnifiedModel, pack_backbone
    _, w = unified_fixture()
    files = sorted(glob.glob("/root/reference/BlazePoser/UnifiedModels/*.h5"))
    assert len(files) == 4
    for f in files:
        u = UnifiedModel.load(f)
        assert u.backbone_flat.size == 101390 and u.head16.program.in_channels == 88 and u.head8.program.in_channels == 96
        assert np.array_equal(u.backbone_flat, pack_backbone(w))      # one detector, four head pairs
        if f.endswith("reg1-stoqa9pt-reg2-hrchr82r-selected.h5"):
            for k, v in u.head16.get_weights_dict().items():
                assert np.array_equal(v, w[f"model/{k}"])
            for k, v in u.head8.get_weights_dict().items():
                assert np.array_equal(v, w[f"model_10/{k}"])


def _chain_desc(built_lib, first, nblk, H, W, tail):
    lib = ctypes.CDLL(built_lib)
    out = np.zeros(264, np.uint32)
    rc = lib.hp_debug_chain_describe(first, nblk, H, W, tail, out.ctypes.data_as(ctypes.c_void_p))
    return rc, out


@pytest.mark.parametrize("size", [64, 88, 96, 100, 104, 128])
def test_chain_lane_tables_cover_every_pixel_column_once(built_lib, size):
    """Host logic of the chain kernels (csrc/blocks_chain.cu: chain_lane_table): every TMEM lane in use owns exactly one
    (image, strip, column) of the tile / one output pixel of the stride-2 tail block, nothing twice, nothing missing."""
    H16, H8 = -(-size // 8), -(-size // 16)
    for first, nblk, H, tail in ((6, 5, H16, 1), (12, 4, H8, 0)):
        rc, d = _chain_desc(built_lib, first, nblk, H, H, tail)
        assert rc == 0, (size, first)
        TR, NI, PS, lanes, tail_lanes, smem = [int(v) for v in d[:6]]
        assert smem <= 227 * 1024 and 1 <= lanes <= 128 and PS % 8 == 4
        strips = -(-H // TR)
        got = {(int(v) & 255, (int(v) >> 8) & 255, int(v) >> 16) for v in d[8:8 + lanes]}
        assert got == {(im, s, x) for im in range(NI) for s in range(strips) for x in range(H)}
        assert not d[8 + lanes:136].any()
        if tail:
            Ho = -(-H // 2)
            assert tail_lanes == NI * Ho * Ho <= 128
            got = {(int(v) & 255, (int(v) >> 8) & 255, (int(v) >> 16) & 255) for v in d[136:136 + tail_lanes]}
            assert got == {(im, y, x) for im in range(NI) for y in range(Ho) for x in range(Ho)}


def _wavefronts(residues):
    """shared-memory wavefronts of one LDS.128 of a warp: the 8 lanes of a quarter warp share a wavefront unless two of them hit the
    same bank group (16-byte chunk index mod 8)"""
    tot = 0
    for q in range(0, len(residues), 8):
        grp = residues[q:q + 8]
        tot += max(grp.count(r) for r in set(grp))
    return tot


def test_chain_lane_tables_are_bank_conflict_free_at_96(built_lib):
    """At the headline size (12 x 12 and 6 x 6 maps) the lane tables make every 128-bit shared-memory access of the chain kernels
    conflict-free: the bank group of a pixel is (linear pixel index x odd chunk count) mod 8, the tail block's swapped lanes start
    one 16-byte chunk further."""
    for first, nblk, H, tail in ((6, 5, 12, 1), (12, 4, 6, 0)):
        rc, d = _chain_desc(built_lib, first, nblk, H, H, tail)
        assert rc == 0
        TR, NI, PS, lanes, tail_lanes = [int(v) for v in d[:5]]
        chunks = PS // 4
        assert chunks % 2 == 1
        res = []
        for v in d[8:8 + lanes]:
            im, s, x = int(v) & 255, (int(v) >> 8) & 255, int(v) >> 16
            res.append((((im * (H + 1) + s * TR) * H + x) * chunks) % 8)
        # 12 x 12: 4 wavefronts per warp; 6 x 6 (7 images of 18 lanes): the residue classes are not equally full, one group has two lanes
        # in a class (17 wavefronts for 126 lanes; the natural lane order needs 31)
        assert _wavefronts(res) <= -(-lanes // 8) + (1 if H == 6 else 0), (H, res)
        if tail:
            res = []
            for v in d[136:136 + tail_lanes]:
                im, y, x, swp = int(v) & 255, (int(v) >> 8) & 255, (int(v) >> 16) & 255, int(v) >> 31
                res.append((((im * (H + 1) + 2 * y) * H + 2 * x) * chunks + swp) % 8)
            assert _wavefronts(res) == -(-tail_lanes // 8), res


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs without a GPU (the CPU restatement on a bounded sample) and prints ONE JSON line with the
    keys the driver reads; under torchrun only rank 0 prints."""
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--cpu-sample", "4"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "crops/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]
