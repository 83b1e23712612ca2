"""The unified detector + pose graph (the artefact JoinModels.py:5-90 produces and
blazeFaceDetectorH5.py:102 loads): BlazeFace-front backbone weights packed for libhpose plus two
regressor heads attached at ``re_lu_10`` (16x16x88) and ``re_lu_15`` (8x8x96).

Also generates the Keras ``model_config`` of that graph (layer names as in SURVEY Appendix A) so
that unified ``.h5`` files written here stay readable by Keras-style tooling.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import numpy as np

from . import _lib, h5lite
from .keras_spec import Model

BLOCKS = [(24, 24, 1), (24, 28, 1), (28, 32, 2), (32, 36, 1), (36, 42, 1), (42, 48, 2), (48, 56, 1), (56, 64, 1),
          (64, 72, 1), (72, 80, 1), (80, 88, 1), (88, 96, 2), (96, 96, 1), (96, 96, 1), (96, 96, 1), (96, 96, 1)]
DET_HEADS = [("conv2d_17", 88, 2), ("conv2d_18", 96, 6), ("conv2d_19", 88, 32), ("conv2d_20", 96, 96)]


def _sfx(base: str, i: int) -> str:
    return base if i == 0 else f"{base}_{i}"


def backbone_weight_specs():
    """[(keras key, shape)] in the packed order of hp_backbone_load_weights (layout_id 0)."""
    specs = [("conv2d/kernel", (5, 5, 3, 24)), ("conv2d/bias", (24,))]
    for i, (cin, cout, _) in enumerate(BLOCKS):
        dw, pw = _sfx("depthwise_conv2d", i), f"conv2d_{i + 1}"
        specs += [(f"{dw}/depthwise_kernel", (3, 3, cin, 1)), (f"{dw}/bias", (cin,)),
                  (f"{pw}/kernel", (1, 1, cin, cout)), (f"{pw}/bias", (cout,))]
    for name, cin, cout in (DET_HEADS[0], DET_HEADS[1], DET_HEADS[2], DET_HEADS[3]):
        specs += [(f"{name}/kernel", (1, 1, cin, cout)), (f"{name}/bias", (cout,))]
    return specs


def pack_backbone(weights: Dict[str, np.ndarray]) -> np.ndarray:
    parts = []
    for key, shape in backbone_weight_specs():
        if key not in weights:
            raise ValueError(f"Layer weight '{key}' not found in face detector model")
        w = np.asarray(weights[key], np.float32)
        if tuple(w.shape) != shape:
            raise ValueError(f"'{key}' has shape {w.shape}, expected {shape}: not a BlazeFace-front detector")
        parts.append(w.reshape(-1))
    flat = np.concatenate(parts)
    assert flat.size == _lib.HP_BACKBONE_PARAMS
    return flat


def unpack_backbone(flat: np.ndarray) -> Dict[str, np.ndarray]:
    out, off = {}, 0
    for key, shape in backbone_weight_specs():
        n = int(np.prod(shape))
        out[key] = np.asarray(flat[off:off + n], np.float32).reshape(shape).copy()
        off += n
    return out


def random_backbone(seed: int = 1234, bias_scale: float = 0.0) -> Dict[str, np.ndarray]:
    """Glorot-uniform kernels, zero biases (SURVEY 8d synthetic-weights recipe)."""
    rng = np.random.default_rng(seed)
    out = {}
    for key, shape in backbone_weight_specs():
        if key.endswith("bias"):
            out[key] = (rng.standard_normal(shape) * bias_scale).astype(np.float32)
        else:
            if key.endswith("depthwise_kernel"):
                fan_in, fan_out = 9, 9
            else:
                fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            out[key] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
    return out


# ---------------------------------------------------------------------- Keras graph of the detector
def _node(*srcs):
    return [[[s, 0, 0, {}] for s in srcs]]


def blazeface_graph_config(head16_cfg: dict, head8_cfg: dict, head16_name="model", head8_name="model_10",
                           input_size: int = 128) -> dict:
    layers: List[dict] = []

    def add(cls, name, cfg, inbound, cfg_name=None):
        c = {"name": cfg_name or name, "trainable": True, "dtype": "float32"}
        c.update(cfg)
        layers.append({"class_name": cls, "config": c, "name": name, "inbound_nodes": inbound})

    def conv(name, filters, k, s, pad, act, src):
        add("Conv2D", name, {"filters": filters, "kernel_size": [k, k], "strides": [s, s], "padding": pad,
                             "data_format": "channels_last", "dilation_rate": [1, 1], "groups": 1, "activation": act,
                             "use_bias": True}, _node(src))

    add("InputLayer", "input", {"batch_input_shape": [None, input_size, input_size, 3], "sparse": False,
                                "ragged": False}, [])
    conv("conv2d", 24, 5, 2, "same", "relu", "input")
    cur = "conv2d"
    n_pad = n_pool = 0
    for i, (cin, cout, s) in enumerate(BLOCKS):
        dw, pw = _sfx("depthwise_conv2d", i), f"conv2d_{i + 1}"
        add("DepthwiseConv2D", dw, {"kernel_size": [3, 3], "strides": [s, s], "padding": "same",
                                    "data_format": "channels_last", "dilation_rate": [1, 1], "groups": 1,
                                    "activation": "linear", "use_bias": True, "depth_multiplier": 1}, _node(cur))
        skip = cur
        if s == 2:
            mp = _sfx("max_pooling2d", n_pool)
            n_pool += 1
            add("MaxPooling2D", mp, {"pool_size": [2, 2], "padding": "same", "strides": [2, 2],
                                     "data_format": "channels_last"}, _node(cur))
            skip = mp
        conv(pw, cout, 1, 1, "valid", "linear", dw)
        if cout > cin:
            pname = _sfx("Pad", n_pad)
            n_pad += 1
            add("TensorFlowOpLayer", f"tf_op_layer_{pname}",
                {"node_def": {"name": pname, "op": "Pad", "input": [f"{skip}/out", f"{pname}/paddings"],
                              "attr": {"T": {"type": "DT_FLOAT"}, "Tpaddings": {"type": "DT_INT32"}}},
                 "constants": {"1": [[0, 0], [0, 0], [0, 0], [0, cout - cin]]}}, _node(skip), cfg_name=pname)
            skip = f"tf_op_layer_{pname}"
        add("Add", _sfx("add", i), {}, _node(pw, skip))
        add("ReLU", _sfx("re_lu", i), {"max_value": None, "negative_slope": 0.0, "threshold": 0.0}, _node(_sfx("add", i)))
        cur = _sfx("re_lu", i)
    g16, g8 = input_size // 8, input_size // 16
    for (name, cin, cout), src in zip(DET_HEADS, ("re_lu_10", "re_lu_15", "re_lu_10", "re_lu_15")):
        conv(name, cout, 1, 1, "same", "linear", src)
    add("Reshape", "reshape", {"target_shape": [g16, g16, 88]}, _node("re_lu_10"))
    add("Reshape", "reshape_1", {"target_shape": [g8, g8, 96]}, _node("re_lu_15"))
    for op_name, src, shape in (("classificators_1", "conv2d_17", [1, g16 * g16 * 2, 1]),
                                ("classificators_2", "conv2d_18", [1, g8 * g8 * 6, 1]),
                                ("regressors_1", "conv2d_19", [1, g16 * g16 * 2, 16]),
                                ("regressors_2", "conv2d_20", [1, g8 * g8 * 6, 16])):
        add("TensorFlowOpLayer", f"tf_op_layer_{op_name}",
            {"node_def": {"name": op_name, "op": "Reshape", "input": [f"{src}/BiasAdd", f"{op_name}/shape"],
                          "attr": {"T": {"type": "DT_FLOAT"}, "Tshape": {"type": "DT_INT32"}}},
             "constants": {"1": shape}}, _node(src), cfg_name=op_name)
    for hname, hcfg, src in ((head16_name, head16_cfg, "reshape"), (head8_name, head8_cfg, "reshape_1")):
        inner = dict(hcfg["config"] if hcfg.get("class_name") in ("Functional", "Model") else hcfg)
        inner["name"] = hname
        layers.append({"class_name": "Functional", "config": inner, "name": hname, "inbound_nodes": _node(src)})
    return {"class_name": "Functional",
            "config": {"name": "model_unified", "trainable": True, "layers": layers,
                       "input_layers": [["input", 0, 0]],
                       "output_layers": [["tf_op_layer_classificators_1", 0, 0], ["tf_op_layer_classificators_2", 0, 0],
                                         ["tf_op_layer_regressors_1", 0, 0], ["tf_op_layer_regressors_2", 0, 0],
                                         [head16_name, 1, 0], [head8_name, 1, 0]]},
            "keras_version": "2.13.1", "backend": "tensorflow"}


def _normalise(h5_weights: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    out = {}
    for k, v in h5_weights.items():
        parts = k.split("/")
        parts[-1] = parts[-1].split(":")[0]
        if len(parts) >= 3 and parts[0] == parts[1]:
            parts = parts[1:]
        out["/".join(parts)] = v
    return out


class UnifiedModel:
    """Callable with the contract of the Keras unified model at blazeFaceDetectorH5.py:272-278:
    ``model(x)`` -> [cls16 (B,A16,1), cls8 (B,A8,1), loc16 (B,A16,16), loc8 (B,A8,16),
    pose16 (B,H16,W16,3), pose8 (B,H8,W8,3)], batch-generalised."""

    def __init__(self, backbone_weights: Dict[str, np.ndarray], head16: Model, head8: Model,
                 head16_name="model", head8_name="model_10"):
        if head16.program.in_channels != 88 or head8.program.in_channels != 96:
            raise ValueError("regressor1 must take 88 channels (re_lu_10) and regressor2 96 channels (re_lu_15)")
        if head16.program.out_channels != 3 or head8.program.out_channels != 3:
            raise ValueError("regressors must output 3 channels (yaw, pitch, roll)")
        self.backbone_flat = pack_backbone(backbone_weights)
        self.head16, self.head8 = head16, head8
        self.head16_name, self.head8_name = head16_name, head8_name
        self._ctx = None
        self._generation = -1             # hp_backbone_generation right after this model's own load

    # ---- persistence
    @classmethod
    def load(cls, path: str) -> "UnifiedModel":
        if not os.path.exists(path):
            raise FileNotFoundError(f"Model file not found: {path}")
        f = h5lite.H5File(path)
        cfg = f.model_config()
        w = _normalise(f.weights())
        nested = [l for l in cfg["config"]["layers"] if l["class_name"] in ("Functional", "Model")]
        if len(nested) != 2:
            raise ValueError(f"{path}: expected two nested regressor models, found {len(nested)}")
        by_out = {o[0]: i for i, o in enumerate(cfg["config"]["output_layers"])}
        nested.sort(key=lambda l: by_out.get(l["name"], 99))
        heads = []
        for l in nested:
            pre = l["name"] + "/"
            hw = {k[len(pre):]: v for k, v in w.items() if k.startswith(pre)}
            heads.append(Model(_config={"class_name": "Functional", "config": l["config"]}, _weights=hw))
        return cls(w, heads[0], heads[1], nested[0]["name"], nested[1]["name"])

    def config(self) -> dict:
        return blazeface_graph_config(self.head16._config, self.head8._config, self.head16_name, self.head8_name)

    def save(self, path: str):
        h5w = {}
        for key, arr in unpack_backbone(self.backbone_flat).items():
            layer, var = key.split("/", 1)
            h5w[f"{layer}/{layer}/{var}:0"] = arr
        for hname, head in ((self.head16_name, self.head16), (self.head8_name, self.head8)):
            for key, arr in head.get_weights_dict().items():
                h5w[f"{hname}/{key}:0"] = arr
        h5lite.write_h5(path, h5w, self.config())

    def count_params(self) -> int:
        return int(self.backbone_flat.size + self.head16.count_params() + self.head8.count_params())

    # ---- device
    def to_device(self, ctx=None):
        """Make this model's weights the ones on the device.  The backbone weights live in the per-GPU context, which several
        models may share: the context remembers whose weights it holds and a model that finds another owner reloads its own
        (a second model on the same context used to leave the first one running on the wrong backbone)."""
        from .device import default_context
        ctx = ctx or default_context()
        if self._ctx is not ctx:
            self.head16.to_device(ctx)
            self.head8.to_device(ctx)
            self._ctx = ctx
        L = _lib.lib()
        if self._generation != L.hp_backbone_generation(ctx.handle):
            _lib.check(L.hp_backbone_load_weights(ctx.handle, self.backbone_flat.ctypes.data, self.backbone_flat.size, 0))
            self._generation = L.hp_backbone_generation(ctx.handle)
        return self

    def forward_device(self, x):
        """x: CUDA float32 (B,H,W,3) in [-1,1] -> dict of CUDA tensors (cls, loc, feat16, feat8, pose16, pose8)."""
        import torch
        self.to_device()
        ctx = self._ctx
        x = x.contiguous().float()
        B, H, W, _ = x.shape
        A = _lib.lib().hp_num_anchors(H, W)
        H16, W16, H8, W8 = -(-H // 8), -(-W // 8), -(-H // 16), -(-W // 16)
        dev = x.device
        out = {"cls": torch.empty((B, A), dtype=torch.float32, device=dev),
               "loc": torch.empty((B, A, 16), dtype=torch.float32, device=dev),
               "feat16": torch.empty((B, H16, W16, 88), dtype=torch.float32, device=dev),
               "feat8": torch.empty((B, H8, W8, 96), dtype=torch.float32, device=dev)}
        _lib.check(_lib.lib().hp_backbone_forward(ctx.handle, x.data_ptr(), B, H, W, out["feat16"].data_ptr(),
                                                  out["feat8"].data_ptr(), out["cls"].data_ptr(), out["loc"].data_ptr(),
                                                  ctx.stream_ptr()))
        out["pose16"] = self.head16.predict_device(out["feat16"])
        out["pose8"] = self.head8.predict_device(out["feat8"])
        return out

    def __call__(self, x):
        import torch
        from .device import default_context
        ctx = default_context()
        xt = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(ctx.torch_device)
        if xt.dim() != 4 or xt.shape[-1] != 3:
            raise ValueError(f"expected input (B,H,W,3), got {tuple(xt.shape)}")
        # The tensor-core stem and the fused chain kernels multiply split-fp16 operands: an input (|x| > 65504, inf, NaN) or a
        # depthwise output of blocks 6-15 outside the fp16 range raises a status flag instead of a silent inf; the batch is then
        # redone with the fp32 stem / the 3xTF32 chain kernels, which have no such limit.
        L = _lib.lib()
        flags = C.c_uint(0)
        stem_fp32 = chain_tf32 = False
        try:
            for _ in range(3):
                o = self.forward_device(xt)
                _lib.check(L.hp_backbone_status(ctx.handle, C.byref(flags), ctx.stream_ptr()))
                redo = False
                if (flags.value & 1) and not stem_fp32:
                    stem_fp32 = redo = True
                    _lib.check(L.hp_debug_set_stem_tc(ctx.handle, -1, 0, 0, 0))
                if (flags.value & 4) and not chain_tf32:
                    chain_tf32 = redo = True
                    _lib.check(L.hp_debug_set_chain(ctx.handle, 2 + 4, 0, 0))
                if not redo:
                    break
        finally:
            if stem_fp32:
                _lib.check(L.hp_debug_set_stem_tc(ctx.handle, 0, 0, 0, 0))
            if chain_tf32:
                _lib.check(L.hp_debug_set_chain(ctx.handle, 2, 0, 0))
        B = xt.shape[0]
        A16 = o["feat16"].shape[1] * o["feat16"].shape[2] * 2
        cls, loc = o["cls"].cpu().numpy(), o["loc"].cpu().numpy()
        return [cls[:, :A16, None], cls[:, A16:, None], loc[:, :A16], loc[:, A16:],
                o["pose16"].cpu().numpy(), o["pose8"].cpu().numpy()]
