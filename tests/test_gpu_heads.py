"""GPU parity of the regressor heads: forward (angles within 0.01 degree), one fit step
(gradients / updated weights vs float64 autograd + Keras optimizer formulas) and fit/evaluate plumbing."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from helpers import DropoutSource, head_oracle, rel_err, synthetic_features, synthetic_poses
from oracle.keras_graph import KerasGraph, keras_train_step, to_torch

pytestmark = pytest.mark.gpu
ANGLE_TOL = 0.01


def test_shipped_checkpoints_on_dataset_slices():
    """model.predict on the committed npz slices: within 0.01 degree of the float64 oracle and the same MAE."""
    from hpose_b200 import keras_spec as K
    z = np.load(os.path.join(GOLDEN, "kat_heads.npz"))
    with open(os.path.join(GOLDEN, "kat_heads.json")) as f:
        kat = json.load(f)
    for hid in ("stoqa9pt", "12uei1sn", "hrchr82r"):
        m = K.load_model(os.path.join(GOLDEN, "heads", f"{hid}.h5"))
        for ds in kat[hid]["datasets"]:
            feats, want, poses = z[f"{hid}|{ds}|features"], z[f"{hid}|{ds}|pred64"], z[f"{hid}|{ds}|poses"]
            got = m.predict(feats.reshape(len(feats), 1, 1, -1), verbose=0).reshape(-1, 3)
            assert np.abs(got - want).max() < ANGLE_TOL, (hid, ds, np.abs(got - want).max())
            assert abs(np.abs(got - poses).mean() - np.abs(want - poses).mean()) < 1e-3


def test_evaluate_head_pose_model_contract(tmp_path):
    from hpose_b200.evaluate import evaluate_head_pose_model
    from hpose_b200.utilities import save_dataset
    z = np.load(os.path.join(GOLDEN, "kat_heads.npz"))
    key = "hrchr82r|AFLW2000_features_96_0.7_1.npz"
    p = str(tmp_path / "slice.npz")
    save_dataset(p, z[f"{key}|features"], z[f"{key}|poses"])
    m = evaluate_head_pose_model(os.path.join(GOLDEN, "heads", "hrchr82r.h5"), p, verbose=False)
    err = z[f"{key}|pred64"] - z[f"{key}|poses"]
    assert set(m) == {"MAE", "MSE"} and set(m["MAE"]) == {"yaw", "pitch", "roll", "average"}
    assert abs(m["MAE"]["average"] - np.abs(err).mean()) < 1e-3
    assert abs(m["MSE"]["average"] - (err ** 2).mean()) / (err ** 2).mean() < 1e-4
    with pytest.raises(ValueError):
        evaluate_head_pose_model(os.path.join(GOLDEN, "heads", "stoqa9pt.h5"), p, verbose=False)


@pytest.mark.parametrize("hw", [(1, 1), (6, 6), (12, 12), (16, 16), (5, 3)])
def test_attention_head_forward_spatial(hw):
    """SE + transformer head on real spatial maps (tokens = H*W up to 256), trained 12uei1sn weights and a
    random 96-channel default head."""
    from hpose_b200 import keras_spec as K
    from hpose_b200.attention_model import se_transformer_regr_head
    H, W = hw
    m = K.load_model(os.path.join(GOLDEN, "heads", "12uei1sn.h5"))
    x = synthetic_features(3 * H * W, 88, seed=H * 31 + W).reshape(3, H, W, 88)
    g, _ = head_oracle(m)
    with torch.no_grad():
        want = g(torch.tensor(x, dtype=torch.float64)).numpy()
    got = m.predict(x)
    assert got.shape == (3, H, W, 3) and np.abs(got - want).max() < ANGLE_TOL, np.abs(got - want).max()
    K.reset_names(); K.set_seed(5)
    m2 = se_transformer_regr_head(input_channels=96)
    x2 = synthetic_features(2 * H * W, 96, seed=7, sigma=0.55, p=0.31).reshape(2, H, W, 96)
    g2, _ = head_oracle(m2)
    with torch.no_grad():
        want2 = g2(torch.tensor(x2, dtype=torch.float64)).numpy()
    assert rel_err(m2.predict(x2), want2) < 1e-4


def test_all_builders_forward():
    from hpose_b200 import attention_model as am, keras_spec as K, train_88, train_96
    train_96.config.update(num_filters=48, dropout_rate=0.1, regularizer_rate=1e-5)
    K.reset_names(); K.set_seed(9)
    models = [am.create_modelC(), am.create_model_complex(1e-6, 1e-4), train_88.create_model(),
              train_88.create_model_skip_fc(), train_88.bestmodelV1(), train_96.create_model()]
    for m in models:
        c = m.program.in_channels
        x = synthetic_features(2 * 4 * 4, c, seed=c).reshape(2, 4, 4, c)
        g, _ = head_oracle(m)
        with torch.no_grad():
            want = g(torch.tensor(x, dtype=torch.float64)).numpy()
        assert rel_err(m.predict(x), want) < 1e-4, m.name


def _dropout_source(model, seed, step, n):
    from hpose_b200 import _lib
    lib = _lib.lib()

    def fn(layer_name, n_img, c, rate):
        op_id = model.program.dropout_ops[layer_name]
        u = np.array([[lib.hp_dropout_hash(seed, step, op_id, i, j) for j in range(c)] for i in range(n_img)], dtype=np.uint32)
        keep = (u >> 8).astype(np.float32) * np.float32(1.0 / 16777216.0) >= np.float32(rate)
        return keep.astype(np.float64)
    return DropoutSource(fn)


def _train_parity(model, opt, x, y, steps, seed=123, gtol=2e-4, wtol=2e-5):
    """Run `steps` optimizer steps on the GPU and in the float64 oracle from the same start."""
    from hpose_b200 import _lib
    model.compile(optimizer=opt, loss="mse", metrics=["mae"])
    g, params = head_oracle(model, torch.float64, requires_grad=True)
    ocfg = {"name": opt.kind, "learning_rate": opt.learning_rate, "beta_1": opt.beta_1, "beta_2": opt.beta_2,
            "epsilon": opt.epsilon}
    state = {}
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    keys = [k for k, _, _ in model.program.layout]
    for step in range(steps):
        loss_ref, mae_ref, grads_ref = keras_train_step(g, params, x, y, ocfg, state,
                                                        dropout=_dropout_source(model, seed, step, len(x)))
        # oracle gradients include the L2 term; the library reports data gradients, so add 2*l2*w back
        w_before = model.get_weights_dict()
        loss, mae = model.train_on_device(xt, yt, seed=seed)
        ctx, head = model._device
        gflat = np.empty(model.program.n_params, np.float32)
        _lib.check(_lib.lib().hp_head_get_grads(ctx.handle, head, gflat.ctypes.data, gflat.size))
        got_g = model.program.unpack(gflat)
        l2 = {}
        for o in model.program.ops:
            if o["op"] == _lib.HP_OP_DENSE:
                l2[f"{o['layer']}/kernel"], l2[f"{o['layer']}/bias"] = o["l2_w"], o["l2_b"]
        gscale = max(float(grads_ref[k].abs().max()) for k in keys) + 1e-30
        for k in keys:
            want = grads_ref[k].numpy()
            have = got_g[k].astype(np.float64) + 2.0 * l2.get(k, 0.0) * w_before[k].astype(np.float64)
            assert np.abs(have - want).max() <= gtol * gscale, (step, k, np.abs(have - want).max(), gscale)
        assert abs(loss - loss_ref) <= 1e-4 * max(1.0, abs(loss_ref)), (step, loss, loss_ref)
        assert abs(mae - mae_ref) <= 1e-4 * max(1.0, abs(mae_ref)), (step, mae, mae_ref)
        w_after = model.get_weights_dict()
        for k in keys:
            assert np.abs(w_after[k] - params[k].detach().numpy()).max() <= wtol * max(1.0, float(params[k].abs().max())), (step, k)


@pytest.mark.parametrize("optname", ["sgd", "adam", "adamax"])
def test_train_step_train96_model(optname):
    """train_96.create_model: L2 on kernels and biases, dropout on hidden and output (config 4 shapes)."""
    from hpose_b200 import keras_spec as K, train_96
    train_96.config.update(num_filters=64, dropout_rate=0.25, regularizer_rate=1e-3, optimizer=optname)
    K.reset_names(); K.set_seed(1)
    m = train_96.create_model()
    x = synthetic_features(128, 96, seed=2, sigma=0.55, p=0.31).reshape(128, 1, 1, 96)
    y = synthetic_poses(128, seed=3).reshape(128, 1, 1, 3)
    opt = {"sgd": K.SGD(2.8e-4), "adam": K.Adam(2.8e-4), "adamax": K.Adamax(2.8e-4)}[optname]
    _train_parity(m, opt, x, y, steps=3)


def test_train_step_residual_and_skip_models():
    from hpose_b200 import attention_model as am, keras_spec as K, train_88
    K.reset_names(); K.set_seed(2)
    for m in (am.create_model_complex(1e-4, 0.2), train_88.create_model_skip_fc(), am.create_modelC()):
        x = synthetic_features(96, 88, seed=5).reshape(96, 1, 1, 88)
        y = synthetic_poses(96, seed=6).reshape(96, 1, 1, 3)
        _train_parity(m, K.SGD(2.8e-4), x, y, steps=2)


@pytest.mark.parametrize("hw", [(1, 1), (3, 4)])
def test_train_step_attention_head(hw):
    """se_transformer_regr_head forward+backward: T=1 (the reference's training shape) and T=12 tokens."""
    from hpose_b200 import keras_spec as K
    from hpose_b200.attention_model import se_transformer_regr_head
    H, W = hw
    K.reset_names(); K.set_seed(4)
    m = se_transformer_regr_head(input_channels=88, reduction=8, num_heads=4, key_dim=16, ff_dim=64, hidden_channels=64)
    n = 40
    x = synthetic_features(n * H * W, 88, seed=8).reshape(n, H, W, 88)
    y = np.repeat(synthetic_poses(n, seed=9).reshape(n, 1, 1, 3), H, axis=1).repeat(W, axis=2).copy()
    _train_parity(m, K.Adam(1e-3), x, y, steps=2, gtol=5e-4)


def test_fit_reduces_loss_and_checkpoints(tmp_path):
    """model.fit loop services: history keys, early stopping, best-checkpoint, evaluate; loss goes down."""
    from hpose_b200 import keras_spec as K, train_96
    train_96.config.update(num_filters=32, dropout_rate=0.0, regularizer_rate=1e-6, optimizer="adam")
    K.reset_names(); K.set_seed(12)
    m = train_96.create_model()
    m.optimizer.learning_rate = 5e-3
    rng = np.random.default_rng(0)
    x = synthetic_features(600, 96, seed=1, sigma=0.55, p=0.31)
    wtrue = rng.normal(0, 3, (96, 3))
    y = (x @ wtrue).astype(np.float32)
    ck = str(tmp_path / "best.h5")
    hist = m.fit(x[:500].reshape(-1, 1, 1, 96), y[:500].reshape(-1, 1, 1, 3), epochs=12, batch_size=128,
                 validation_data=(x[500:].reshape(-1, 1, 1, 96), y[500:].reshape(-1, 1, 1, 3)),
                 callbacks=[K.ModelCheckpoint(ck, monitor="val_loss", save_best_only=True),
                            K.EarlyStopping(monitor="val_loss", patience=40, min_delta=1e-3, restore_best_weights=True)],
                 verbose=0)
    assert set(hist.history) == {"loss", "mae", "val_loss", "val_mae"} and len(hist.history["loss"]) == 12
    assert hist.history["loss"][-1] < 0.5 * hist.history["loss"][0]
    loss, mae = m.evaluate(x[500:].reshape(-1, 1, 1, 96), y[500:].reshape(-1, 1, 1, 3), verbose=0)
    assert abs(loss - hist.history["val_loss"][-1]) < 1e-3 * max(1.0, loss)
    best = K.load_model(ck)
    l2, _ = best.evaluate(x[500:].reshape(-1, 1, 1, 96), y[500:].reshape(-1, 1, 1, 3), verbose=0)
    assert abs(l2 - min(hist.history["val_loss"])) < 1e-3 * max(1.0, l2)


def test_head_create_rejects_bad_programs():
    from hpose_b200 import _lib
    from hpose_b200.device import default_context
    ctx = default_context()
    ops = (_lib.hp_head_op * 1)()
    ops[0].op, ops[0].in0, ops[0].out, ops[0].cin, ops[0].cout = _lib.HP_OP_DENSE, 0, 1, 8, 4
    ops[0].w_off, ops[0].b_off = 0, 1000
    regs = (_lib.hp_head_reg * 2)()
    regs[0].channels, regs[1].channels = 8, 4
    head = C.c_void_p()
    rc = _lib.lib().hp_head_create(ctx.handle, ops, 1, regs, 2, 1, 36, C.byref(head))
    assert rc == -1 and b"out of range" in _lib.lib().hp_last_error()


DENSE_SHAPES = [(4096, 88, 64, 0), (4097, 88, 34, 0), (1000, 96, 102, 0), (777, 20, 3, 0), (5000, 128, 128, 0), (640, 64, 3, 0),
                (4100, 88, 64, 3), (900, 96, 128, 3), (513, 24, 16, 1), (3000, 64, 64, 4), (300, 88, 64, 0)]


@pytest.mark.parametrize("M,K,N,n2", DENSE_SHAPES)
@pytest.mark.parametrize("act", ["linear", "softsign", "relu"])
def test_dense_layer_matches_fp64(M, K, N, n2, act):
    """Dense / 1x1-conv layer (tensor-core 3xTF32 kernel where the shape allows it, CUDA cores otherwise) and the fused
    Dense -> narrow Dense pair against an fp64 torch reference: partial last tile, K and N padding, every activation."""
    from hpose_b200 import _lib
    from hpose_b200.device import default_context
    ctx = default_context()
    lib = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = torch.randn((M, K), generator=g, device="cuda")
    W = torch.randn((K, N), generator=g, device="cuda") / K ** 0.5
    b = torch.randn((N,), generator=g, device="cuda")
    W2 = torch.randn((N, max(n2, 1)), generator=g, device="cuda") / N ** 0.5
    b2 = torch.randn((max(n2, 1),), generator=g, device="cuda")
    fn = {"linear": lambda v: v, "softsign": torch.nn.functional.softsign, "relu": torch.relu}[act]
    want = fn(x.double() @ W.double() + b.double())
    if n2:
        want = want @ W2.double() + b2.double()
    y = torch.full((M, n2 if n2 else N), float("nan"), device="cuda")
    _lib.check(lib.hp_debug_dense(ctx.handle, x.data_ptr(), M, K, W.data_ptr(), b.data_ptr(), N, _lib.HP_ACT[act], y.data_ptr(),
                                  W2.data_ptr() if n2 else None, b2.data_ptr() if n2 else None, n2, _lib.HP_ACT["linear"], None))
    torch.cuda.synchronize()
    # 3xTF32 drops the a_lo * w_lo term (2^-22 per product): ~4e-6 absolute on pre-activations of magnitude 1-5 with K = 88;
    # north_star asks for 1e-4 on feature maps
    err = float((y.double() - want).abs().max() / want.abs().max())
    assert err < 1e-5, (M, K, N, n2, act, err)


# ------------------------------------------------------------------------------ hp_head_train_run (many steps, CUDA graph)
def _make_train96(seed, rate=0.25, opt="adam"):
    from hpose_b200 import keras_spec as K, train_96
    train_96.config.update(num_filters=64, dropout_rate=rate, regularizer_rate=1e-5, optimizer=opt)
    K.reset_names(); K.set_seed(seed)
    m = train_96.create_model()
    m.optimizer.learning_rate = 2.8e-4
    return m


@pytest.mark.parametrize("optname,graph", [("adam", True), ("adam", False), ("adamax", True), ("sgd", True)])
def test_train_run_equals_single_steps(optname, graph):
    """hp_head_train_run (device-resident data set, on-device gather of idx[...], step counter / Adam step sizes / loss
    accumulators in device memory, one CUDA-graph launch per step) takes the same steps as hp_head_train_step called once
    per batch from the host -- dropout masks included (they are keyed by the step index, read from device memory here)."""
    n, bs, steps = 128 * 7 + 40, 128, 7
    x = synthetic_features(n, 96, seed=2, sigma=0.55, p=0.31).reshape(n, 1, 1, 96)
    y = synthetic_poses(n, seed=3).reshape(n, 1, 1, 3)
    perm = np.random.default_rng(5).permutation(n).astype(np.int32)
    xt, yt, pt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(perm).cuda()
    a, b = _make_train96(1, opt=optname), _make_train96(1, opt=optname)
    want_loss = want_mae = 0.0
    for s in range(steps):
        idx = torch.from_numpy(perm[s * bs:(s + 1) * bs].astype(np.int64)).cuda()
        loss, mae = a.train_on_device(xt.index_select(0, idx), yt.index_select(0, idx), seed=77)
        want_loss, want_mae = want_loss + loss * bs, want_mae + mae * bs
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        got_loss, got_mae = b.train_run_device(xt, yt, pt, 0, bs, steps, seed=77, graph=graph)
        tail = b.train_run_device(xt, yt, pt, steps * bs, 40, 1, seed=77, graph=graph)      # the partial batch of an epoch
    side.synchronize()
    idx = torch.from_numpy(perm[steps * bs:].astype(np.int64)).cuda()
    tl, tm = a.train_on_device(xt.index_select(0, idx), yt.index_select(0, idx), seed=77)
    wa, wb = a.get_weights_dict(), b.get_weights_dict()
    for k in wa:
        assert np.abs(wa[k] - wb[k]).max() <= 1e-7 * max(1.0, float(np.abs(wa[k]).max())), k
    assert abs(got_loss - want_loss) <= 1e-5 * abs(want_loss) and abs(got_mae - want_mae) <= 1e-5 * abs(want_mae)
    assert abs(tail[0] - tl * 40) <= 1e-5 * abs(tl * 40) and abs(tail[1] - tm * 40) <= 1e-5 * abs(tm * 40)


def test_train_run_attention_head_graph_and_errors():
    """The attention head (LayerNorm / MHA backward, T = 4 tokens) through the captured step; argument checks."""
    from hpose_b200 import _lib, keras_spec as K
    from hpose_b200.attention_model import se_transformer_regr_head

    def make():
        K.reset_names(); K.set_seed(4)
        m = se_transformer_regr_head(input_channels=88, reduction=8, num_heads=2, key_dim=8, ff_dim=16, hidden_channels=16)
        m.compile(optimizer=K.Adam(1e-3), loss="mse", metrics=["mae"])
        return m
    n, bs = 96, 32
    x = synthetic_features(n * 4, 88, seed=8).reshape(n, 2, 2, 88)
    y = np.repeat(synthetic_poses(n, seed=9).reshape(n, 1, 1, 3), 2, axis=1).repeat(2, axis=2).copy()
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    a, b = make(), make()
    for s in range(3):
        a.train_on_device(xt[s * bs:(s + 1) * bs].contiguous(), yt[s * bs:(s + 1) * bs].contiguous(), seed=5)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        b.train_run_device(xt, yt, None, 0, bs, 3, seed=5, graph=True)          # idx = None: identity order
        with pytest.raises(_lib.HposeError):
            b.train_run_device(xt, yt, None, 64, bs, 2, seed=5)                 # runs past the end of the data set
        with pytest.raises(_lib.HposeError):
            b.train_run_device(xt, yt, None, 0, bs, 1, rank=0, world=2, seed=5)  # two ranks without a communicator
    side.synchronize()
    wa, wb = a.get_weights_dict(), b.get_weights_dict()
    for k in wa:
        # the attention backward accumulates with atomics (order varies run to run) and Adam normalises the ~1e-9 gradient of
        # the key bias (zero in exact arithmetic: softmax is shift-invariant) to +-lr: compare on the scale of one Adam step
        assert np.abs(wa[k] - wb[k]).max() <= 3e-3 * 1e-3 * 3, k


@pytest.mark.parametrize("act", ["elu", "selu", "softplus", "leaky_relu", "swish"])
def test_new_activations_forward_and_training(act):
    """The activations of the 17 zoo checkpoints round 1 refused: inference (tensor-core Dense kernel with the activation in
    its epilogue and in the fused narrow layer) against the float64 oracle, and a training step (swish: inference only)."""
    from hpose_b200 import _lib, keras_spec as K
    K.reset_names(); K.set_seed(21)
    x_in = K.Input((None, None, 88))
    h = K.Conv2D(32, 1, activation=act)(x_in)
    h = K.Conv2D(16, 1, activation=act)(h)
    m = K.Model(x_in, K.Conv2D(3, 1)(h))
    x = synthetic_features(600 * 4, 88, seed=5).reshape(600, 2, 2, 88) - 0.3          # both signs reach the activation
    g, _ = head_oracle(m)
    with torch.no_grad():
        want = g(torch.tensor(x, dtype=torch.float64)).numpy()
    got = m.predict(x)
    assert rel_err(got, want) < 2e-5
    small = m.predict(x[:3])                                                          # same kernel, same bits, whatever the batch
    assert np.array_equal(small, got[:3])
    y = synthetic_poses(600, seed=6).reshape(600, 1, 1, 3).repeat(2, 1).repeat(2, 2).copy()
    if act == "swish":
        m.compile(optimizer=K.SGD(1e-4), loss="mse", metrics=["mae"])
        with pytest.raises(_lib.HposeError):
            m.train_on_device(torch.from_numpy(x[:64]).cuda(), torch.from_numpy(y[:64]).cuda())
    else:
        _train_parity(m, K.Adam(1e-3), x[:64], y[:64], steps=2)
