"""GPU parity: CUDA backbone (fast and naive kernel families) vs the float64 oracle.
Tolerance (north_star): feature maps within 1e-4 relative (max-abs error / max-abs value per map)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from helpers import rel_err, unified_fixture
from oracle.keras_graph import KerasGraph, to_torch

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _ctx():
    from hpose_b200.device import default_context
    return default_context()


def _oracle_taps(graph, w, x):
    taps = {}
    g = KerasGraph(graph, to_torch(w, torch.float64))
    with torch.no_grad():
        outs = g(torch.tensor(x, dtype=torch.float64), taps=taps)
    return outs, taps


def _read_act(ctx, xt, blk, shape):
    from hpose_b200 import _lib
    dst = torch.empty(shape, dtype=torch.float32, device=xt.device)
    B, H, W, _ = xt.shape
    _lib.check(_lib.lib().hp_backbone_read_activation(ctx.handle, xt.data_ptr(), B, H, W, blk, dst.data_ptr(),
                                                      dst.numel(), ctx.stream_ptr()))
    torch.cuda.synchronize()
    return dst.cpu().numpy()


@pytest.mark.parametrize("impl", ["naive", "fast", "cpasync", "tma"])
def test_trained_weights_128_all_layers(impl):
    """Shipped detector weights, reference input size; every block output is compared."""
    from hpose_b200 import _lib
    from hpose_b200.unified import pack_backbone
    graph, w = unified_fixture()
    kat = np.load(os.path.join(GOLDEN, "unified_kat.npz"))
    x = kat["x"]
    ctx = _ctx()
    ctx.set_impl({"naive": _lib.HP_IMPL_NAIVE, "fast": _lib.HP_IMPL_FAST, "cpasync": _lib.HP_IMPL_CPASYNC, "tma": _lib.HP_IMPL_TMA}[impl])
    try:
        flat = pack_backbone(w)
        _lib.check(_lib.lib().hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
        xt = torch.from_numpy(x).to(ctx.torch_device)
        _, taps = _oracle_taps(graph, w, x)
        errs = {}
        want = taps["conv2d"].numpy()
        errs["stem"] = rel_err(_read_act(ctx, xt, -1, want.shape), want)
        for blk in range(16):
            want = taps["re_lu" if blk == 0 else f"re_lu_{blk}"].numpy()
            errs[f"blk{blk}"] = rel_err(_read_act(ctx, xt, blk, want.shape), want)
        bad = {k: v for k, v in errs.items() if not v < TOL}
        assert not bad, f"{impl}: per-layer relative errors {errs}"
    finally:
        ctx.set_impl(_lib.HP_IMPL_FAST)


@pytest.mark.parametrize("size,batch", [(128, 2), (96, 5), (88, 3), (64, 1), (120, 2), (100, 2), (104, 1)])
def test_unified_outputs_match_oracle(size, batch):
    """All six outputs of the unified graph (trained weights) at several input sizes incl. odd maps (88 -> 11 -> 6)."""
    from hpose_b200 import keras_spec as K
    from hpose_b200.unified import UnifiedModel
    graph, w = unified_fixture()
    u = UnifiedModel(w, K.load_model(os.path.join(GOLDEN, "heads", "stoqa9pt.h5")),
                     K.load_model(os.path.join(GOLDEN, "heads", "hrchr82r.h5")))
    rng = np.random.default_rng(size)
    x = rng.uniform(-1, 1, size=(batch, size, size, 3)).astype(np.float32)
    got = u(x)
    want, taps = _oracle_taps(graph, w, x)
    names = ("cls16", "cls8", "loc16", "loc8", "pose16", "pose8")
    errs = {n: rel_err(a, b.numpy()) for n, a, b in zip(names, got, want)}
    assert all(v < TOL for v in errs.values()), errs
    # angles within 0.01 degree (north_star)
    assert np.abs(got[4] - want[4].numpy()).max() < 0.01 and np.abs(got[5] - want[5].numpy()).max() < 0.01
    dev = u.forward_device(torch.from_numpy(x).cuda())
    assert rel_err(dev["feat16"].cpu().numpy(), taps["re_lu_10"].numpy()) < TOL
    assert rel_err(dev["feat8"].cpu().numpy(), taps["re_lu_15"].numpy()) < TOL


def test_golden_unified_kat():
    """Committed float64 known answers (tests/golden/unified_kat.npz, generated from the shipped .h5)."""
    from hpose_b200 import keras_spec as K
    from hpose_b200.unified import UnifiedModel
    _, w = unified_fixture()
    kat = np.load(os.path.join(GOLDEN, "unified_kat.npz"))
    u = UnifiedModel(w, K.load_model(os.path.join(GOLDEN, "heads", "stoqa9pt.h5")),
                     K.load_model(os.path.join(GOLDEN, "heads", "hrchr82r.h5")))
    got = u(kat["x"])
    for name, a in zip(("cls16", "cls8", "loc16", "loc8", "pose16", "pose8"), got):
        assert rel_err(a, kat[name]) < TOL, name
    assert np.abs(got[4] - kat["pose16"]).max() < 0.01 and np.abs(got[5] - kat["pose8"]).max() < 0.01


def test_random_weights_fast_equals_naive_large_batch():
    """Random-init weights (bench recipe) at B=67, 96x96: the fast kernels agree with the naive CUDA kernels
    (size-independent cross-check at a batch the oracle would take long for) and with the oracle on 2 images."""
    from hpose_b200 import _lib, keras_spec as K
    from hpose_b200.unified import blazeface_graph_config, pack_backbone, random_backbone
    ctx = _ctx()
    w = random_backbone(seed=1234, bias_scale=0.05)
    flat = pack_backbone(w)
    _lib.check(_lib.lib().hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
    B, S = 67, 96
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand((B, S, S, 3), generator=g, device="cuda") * 2 - 1
    A = _lib.lib().hp_num_anchors(S, S)
    res = {}
    for impl in (_lib.HP_IMPL_NAIVE, _lib.HP_IMPL_FAST, _lib.HP_IMPL_CPASYNC):
        ctx.set_impl(impl)
        f16 = torch.empty((B, 12, 12, 88), device="cuda"); f8 = torch.empty((B, 6, 6, 96), device="cuda")
        cls = torch.empty((B, A), device="cuda"); loc = torch.empty((B, A, 16), device="cuda")
        _lib.check(_lib.lib().hp_backbone_forward(ctx.handle, x.data_ptr(), B, S, S, f16.data_ptr(), f8.data_ptr(),
                                                  cls.data_ptr(), loc.data_ptr(), ctx.stream_ptr()))
        torch.cuda.synchronize()
        res[impl] = [t.cpu().numpy() for t in (f16, f8, cls, loc)]
    ctx.set_impl(_lib.HP_IMPL_FAST)
    for impl in (_lib.HP_IMPL_FAST, _lib.HP_IMPL_CPASYNC):
        for a, b, n in zip(res[impl], res[_lib.HP_IMPL_NAIVE], ("feat16", "feat8", "cls", "loc")):
            assert rel_err(a, b) < 2e-5, (impl, n)
    # oracle on the first two images
    K.reset_names()
    dummy16 = K.Model(*(lambda i: (i, K.Conv2D(3, 1)(i)))(K.Input((None, None, 88))))
    K.reset_names()
    dummy8 = K.Model(*(lambda i: (i, K.Conv2D(3, 1)(i)))(K.Input((None, None, 96))))
    cfg = blazeface_graph_config(dummy16._config, dummy8._config)
    ww = dict(w)
    for nm, d in (("model", dummy16), ("model_10", dummy8)):
        for k, v in d.get_weights_dict().items():
            ww[f"{nm}/{k}"] = v
    taps = {}
    with torch.no_grad():
        KerasGraph(cfg, to_torch(ww, torch.float64))(x[:2].double().cpu(), taps=taps)
    assert rel_err(res[_lib.HP_IMPL_FAST][0][:2], taps["re_lu_10"].numpy()) < TOL
    assert rel_err(res[_lib.HP_IMPL_FAST][1][:2], taps["re_lu_15"].numpy()) < TOL


def test_argument_errors():
    from hpose_b200 import _lib
    ctx = _ctx()
    x = torch.zeros((1, 8, 8, 3), device="cuda")
    rc = _lib.lib().hp_backbone_forward(ctx.handle, x.data_ptr(), 1, 8, 8, None, None, None, None, None)
    assert rc == -1 and b"H, W >= 16" in _lib.lib().hp_last_error()
    bad = np.zeros(10, np.float32)
    assert _lib.lib().hp_backbone_load_weights(ctx.handle, bad.ctypes.data, 10, 0) == -1


def test_preprocess_matches_reference_arithmetic():
    from hpose_b200 import _lib
    ctx = _ctx()
    img = np.random.default_rng(1).integers(0, 256, size=(3, 40, 24, 3), dtype=np.uint8)
    u8 = torch.from_numpy(img).cuda()
    x = torch.empty((3, 40, 24, 3), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().hp_preprocess_u8(ctx.handle, u8.data_ptr(), 3, 40, 24, x.data_ptr(), ctx.stream_ptr()))
    want = ((img[..., ::-1].astype(np.float64) / 255.0).astype(np.float32) - np.float32(0.5)) / np.float32(0.5)
    assert np.array_equal(x.cpu().numpy(), want)


# (TR, NSTG, npipe, nsets, nbuf): nbuf == 0 -> pipelined kernel (early blocks only); nbuf > 0 -> warp-specialised kernel, where the
# npipe slot carries the number of epilogue warp sets, NSTG 0 = automatic and nbuf % 16 caps the ring of halo buffers
TC_GEOMETRIES = [(2, 2, 4, 1, 0), (2, 2, 3, 2, 0), (4, 2, 2, 2, 0), (4, 1, 2, 2, 0), (4, 1, 1, 1, 0),
                 (4, 0, 2, 2, 4), (4, 0, 2, 3, 3), (4, 2, 1, 2, 2), (2, 0, 2, 2, 4), (2, 0, 2, 3, 2), (2, 2, 1, 3, 4),
                 (2, 0, 2, 2, 4 + 32), (2, 0, 2, 3, 4 + 32), (2, 0, 1, 4, 3 + 32), (4, 0, 2, 2, 4 + 32), (4, 0, 2, 3, 4 + 32),   # + 32: k-step work units
                 (4, 0, 2, 2, 4 + 32 + 128), (4, 0, 2, 2, 3 + 32 + 256), (2, 0, 2, 3, 4 + 32 + 128), (2, 0, 2, 3, 4 + 16 + 128),   # + 64 n: n issuer warps
                 (1, 0, 2, 3, 4), (1, 0, 1, 3, 2), (1, 0, 1, 2, 3), (1, 0, 2, 4, 4)]   # TR 1: pixel-per-lane kernel (maps of <= 128 pixels)
TC_BLOCK_CHANNELS = {0: 24, 1: 28, 2: 32, 3: 36, 4: 42, 5: 48, 6: 56, 7: 64, 8: 72, 9: 80, 10: 88, 11: 96, 12: 96, 15: 96}


@pytest.mark.parametrize("geom", TC_GEOMETRIES)
@pytest.mark.parametrize("size,batch", [(96, 37), (88, 5), (128, 3)])
def test_tensor_core_block_geometries(geom, size, batch):
    """Every instantiated geometry of the tensor-core BlazeBlock kernels reproduces the naive CUDA kernels on every
    block (stride-2 blocks: any TR selects the stride-2 kernel with the given warp sets / buffers) (random-init weights; batch 37 gives the persistent CTAs different tile counts and a partial last
    multi-image tile, 88 gives partial bands and odd widths).  Geometries that do not fit a block (TMEM / shared
    memory / not instantiated) must be refused with an error, never run wrong."""
    from hpose_b200 import _lib
    from hpose_b200.unified import pack_backbone, random_backbone
    ctx = _ctx()
    lib = _lib.lib()
    flat = pack_backbone(random_backbone(seed=99, bias_scale=0.1))
    _lib.check(lib.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
    g = torch.Generator(device="cuda").manual_seed(size)
    x = torch.rand((batch, size, size, 3), generator=g, device="cuda") * 2 - 1
    TR, NSTG, npipe, nsets, nbuf = geom
    if nbuf == 0 and not (lib.hp_build_features() & 1):
        pytest.skip("pipelined first-generation kernel: only in builds with -DHP_LEGACY_KERNELS")
    ran = 0
    try:
        for blk, c in TC_BLOCK_CHANNELS.items():
            H = -(-size // 2)
            for b in (2, 5, 11):
                if blk >= b:                     # output map of the block (2, 5, 11 are the stride-2 blocks)
                    H = -(-H // 2)
            shape = (batch, H, H, c)
            ctx.set_impl(_lib.HP_IMPL_NAIVE)
            want = _read_act(ctx, x, blk, shape)
            ctx.set_impl(_lib.HP_IMPL_FAST)
            strips = -(-H // TR)
            bands = -(-strips // max(1, 128 // H))
            BH = -(-strips // bands) * TR
            _lib.check(lib.hp_debug_set_tc(ctx.handle, blk, TR, NSTG, BH, npipe, nsets, nbuf))
            try:
                got = _read_act(ctx, x, blk, shape)
            except _lib.HposeError as e:
                assert e.code in (-1, -5), e          # refused: not instantiated / does not fit
                continue
            finally:
                _lib.check(lib.hp_debug_set_tc(ctx.handle, blk, 0, 0, 0, 0, 0, 0))
            ran += 1
            assert rel_err(got, want) < 2e-5, (geom, blk, rel_err(got, want))
    finally:
        ctx.set_impl(_lib.HP_IMPL_FAST)
    assert ran >= 1 or size == 128, f"geometry {geom} ran on no block"


# BH, nbuf, nout, gather sets + 16 x issuers (+ 128 / 512: two / three epilogue sets; + 256: the first-generation kernel, where + 128
# places the issuers on SM sub-partition 3)
@pytest.mark.parametrize("cfg", [(0, 0, 0, 0), (0, 2, 2, 2 + 32), (4, 3, 2, 3 + 32 + 128), (0, 4, 3, 4 + 32), (0, 0, 0, 3 + 48 + 512),
                                 (0, 0, 0, 256), (0, 2, 2, 2 + 32 + 128 + 256), (4, 3, 2, 3 + 32 + 256)])
@pytest.mark.parametrize("size,batch", [(96, 37), (88, 5), (128, 3), (64, 2), (120, 2), (100, 2), (104, 2), (160, 1)])
def test_tensor_core_stem(cfg, size, batch):
    """The implicit-GEMM (tcgen05, split fp16) stem reproduces the naive CUDA stem for every pipeline geometry of both
    generations (band height, input buffers, output stages, gather / epilogue warp sets, issuers); 88 and 120 give partial
    last bands and partial M-tiles."""
    from hpose_b200 import _lib
    from hpose_b200.unified import pack_backbone, random_backbone
    ctx = _ctx()
    lib = _lib.lib()
    flat = pack_backbone(random_backbone(seed=5, bias_scale=0.2))
    _lib.check(lib.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
    g = torch.Generator(device="cuda").manual_seed(size + 1)
    x = torch.rand((batch, size, size, 3), generator=g, device="cuda") * 2 - 1
    shape = (batch, size // 2, size // 2, 24)
    try:
        ctx.set_impl(_lib.HP_IMPL_NAIVE)
        want = _read_act(ctx, x, -1, shape)
        ctx.set_impl(_lib.HP_IMPL_FAST)
        _lib.check(lib.hp_debug_set_stem_tc(ctx.handle, *cfg))
        got = _read_act(ctx, x, -1, shape)
        assert rel_err(got, want) < 2e-5, rel_err(got, want)
        _lib.check(lib.hp_debug_set_stem_tc(ctx.handle, -1, 0, 0, 0))       # CUDA-core stem still agrees
        assert rel_err(_read_act(ctx, x, -1, shape), want) < 2e-5
    finally:
        _lib.check(lib.hp_debug_set_stem_tc(ctx.handle, 0, 0, 0, 0))
        ctx.set_impl(_lib.HP_IMPL_FAST)


def test_full_size_batch_properties():
    """BASELINE size (4096 crops of 96x96, random-init weights): the tensor-core path agrees with the naive CUDA kernels,
    is deterministic run to run, and every image is independent of its batch neighbours (a 64-image slice pushed through
    on its own is bit-identical to the same slice of the full batch): persistent tiles never mix images."""
    from hpose_b200 import _lib
    from hpose_b200.unified import pack_backbone, random_backbone
    ctx = _ctx()
    lib = _lib.lib()
    flat = pack_backbone(random_backbone(seed=1234))
    _lib.check(lib.hp_backbone_load_weights(ctx.handle, flat.ctypes.data, flat.size, 0))
    B, S = 4096, 96
    g = torch.Generator(device="cuda").manual_seed(2026)
    x = torch.rand((B, S, S, 3), generator=g, device="cuda") * 2 - 1
    A = lib.hp_num_anchors(S, S)

    def run(xx, impl):
        n = xx.shape[0]
        ctx.set_impl(impl)
        f16 = torch.empty((n, 12, 12, 88), device="cuda"); f8 = torch.empty((n, 6, 6, 96), device="cuda")
        cls = torch.empty((n, A), device="cuda"); loc = torch.empty((n, A, 16), device="cuda")
        _lib.check(lib.hp_backbone_forward(ctx.handle, xx.data_ptr(), n, S, S, f16.data_ptr(), f8.data_ptr(), cls.data_ptr(),
                                           loc.data_ptr(), ctx.stream_ptr()))
        torch.cuda.synchronize()
        return f16, f8, cls, loc

    try:
        fast = run(x, _lib.HP_IMPL_FAST)
        again = run(x, _lib.HP_IMPL_FAST)
        for a, b in zip(fast, again):
            assert torch.equal(a, b), "not deterministic"
        part = run(x[1000:1064].contiguous(), _lib.HP_IMPL_FAST)
        for a, b in zip(fast, part):
            assert torch.equal(a[1000:1064], b), "an image depends on its batch neighbours"
        naive = run(x, _lib.HP_IMPL_NAIVE)
        for a, b, n in zip(fast, naive, ("feat16", "feat8", "cls", "loc")):
            err = float((a - b).abs().max() / b.abs().max())
            assert err < 2e-5, (n, err)
    finally:
        ctx.set_impl(_lib.HP_IMPL_FAST)


def test_two_models_share_a_context_without_mixing_weights():
    """The backbone weights live in the per-GPU context: a second model on the same context used to overwrite the
    first one's weights silently.  Each model now reloads its own when the context's weight generation is not the one it saw."""
    from hpose_b200 import keras_spec as K, train_88
    from hpose_b200.attention_model import se_transformer_regr_head
    from hpose_b200.unified import UnifiedModel, random_backbone

    def make(seed):
        K.reset_names(); K.set_seed(seed)
        h16 = train_88.create_model()
        K.reset_names()
        return UnifiedModel(random_backbone(seed=seed, bias_scale=0.1), h16, se_transformer_regr_head(input_channels=96))
    a, b = make(1), make(2)
    x = np.random.default_rng(0).uniform(-1, 1, (2, 96, 96, 3)).astype(np.float32)
    ra = a(x)
    rb = b(x)                                   # loads b's backbone into the shared context
    ra2 = a(x)                                  # a must notice and reload its own
    assert not np.array_equal(ra[0], rb[0])
    for u, v in zip(ra, ra2):
        assert np.array_equal(u, v)
    for u, v in zip(rb, b(x)):
        assert np.array_equal(u, v)


def test_stem_input_range_is_guarded():
    """The tensor-core stem splits its input into fp16 parts (|x| <= 65504).  An input outside that range sets a sticky status
    flag instead of producing a silent inf, and the host model re-runs the batch with the fp32 stem."""
    from hpose_b200 import _lib, keras_spec as K, train_88
    from hpose_b200.attention_model import se_transformer_regr_head
    from hpose_b200.unified import UnifiedModel, random_backbone
    ctx = _ctx()
    K.reset_names(); K.set_seed(3)
    h16 = train_88.create_model()
    K.reset_names()
    bbw = random_backbone(seed=3, bias_scale=0.1)
    m = UnifiedModel(bbw, h16, se_transformer_regr_head(input_channels=96))
    x = np.random.default_rng(1).uniform(-1, 1, (2, 96, 96, 3)).astype(np.float32)
    flags = C.c_uint(7)
    m.forward_device(torch.from_numpy(x).cuda())
    _lib_check_status = lambda: (_lib.check(_lib.lib().hp_backbone_status(ctx.handle, C.byref(flags), ctx.stream_ptr())), flags.value)[1]
    assert _lib_check_status() == 0
    bad = x.copy()
    bad[1, 40, 41, 2] = 1.0e6                   # finite in fp32, infinite in fp16
    m.forward_device(torch.from_numpy(bad).cuda())
    assert _lib_check_status() == 1 and _lib_check_status() == 0          # reported once, then cleared
    nan = x.copy()
    nan[0, 0, 0, 0] = np.nan
    m.forward_device(torch.from_numpy(nan).cuda())
    assert _lib_check_status() == 1
    # the host API re-runs a flagged batch with the fp32 stem: finite outputs that match the float64 oracle
    out = m(bad)
    w = dict(bbw)
    for name, head in (("model", m.head16), ("model_10", m.head8)):
        for k, v in head.get_weights_dict().items():
            w[f"{name}/{k}"] = v
    with torch.no_grad():
        ref = [t.numpy() for t in KerasGraph(m.config(), to_torch(w, torch.float64))(torch.tensor(bad, dtype=torch.float64))]
    for got, want in zip(out, ref):
        assert np.isfinite(got).all() and rel_err(got, want) < 1e-4



def test_chain_fp16_range_is_guarded():
    """The fused chain kernels (blocks 6-15) multiply split-fp16 operands.  A depthwise output beyond the fp16 range (here: the
    depthwise kernel of block 8 scaled by 3e6, everything finite in fp32) sets HP_STATUS_CHAIN_RANGE instead of clamping an inf
    to zero in the ReLU, and the host model re-runs the batch with the 3xTF32 chain kernels: outputs match the float64 oracle."""
    from hpose_b200 import _lib, keras_spec as K, train_88
    from hpose_b200.attention_model import se_transformer_regr_head
    from hpose_b200.unified import UnifiedModel, random_backbone
    ctx = _ctx()
    K.reset_names(); K.set_seed(4)
    h16 = train_88.create_model()
    K.reset_names()
    bbw = random_backbone(seed=4, bias_scale=0.1)
    x = np.random.default_rng(2).uniform(-1, 1, (2, 96, 96, 3)).astype(np.float32)
    flags = C.c_uint(7)
    status = lambda: (_lib.check(_lib.lib().hp_backbone_status(ctx.handle, C.byref(flags), ctx.stream_ptr())), flags.value)[1]
    m = UnifiedModel(bbw, h16, se_transformer_regr_head(input_channels=96))
    m.forward_device(torch.from_numpy(x).cuda())
    assert status() == 0
    big = dict(bbw)
    big["depthwise_conv2d_8/depthwise_kernel"] = bbw["depthwise_conv2d_8/depthwise_kernel"] * np.float32(3e6)
    big["conv2d_9/kernel"] = bbw["conv2d_9/kernel"] * np.float32(1e-6)          # the pointwise conv behind it brings the values back
    mb = UnifiedModel(big, h16, se_transformer_regr_head(input_channels=96))
    mb.forward_device(torch.from_numpy(x).cuda())
    assert status() == 4 and status() == 0                                      # reported once, then cleared
    out = mb(x)
    assert status() == 0
    w = dict(big)
    for name, head in (("model", mb.head16), ("model_10", mb.head8)):
        for k, v in head.get_weights_dict().items():
            w[f"{name}/{k}"] = v
    with torch.no_grad():
        ref = [t.numpy() for t in KerasGraph(mb.config(), to_torch(w, torch.float64))(torch.tensor(x, dtype=torch.float64))]
    for got, want in zip(out, ref):
        assert np.isfinite(got).all() and rel_err(got, want) < 1e-4
