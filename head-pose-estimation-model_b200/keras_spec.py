"""Keras-shaped host objects for the regressor heads, backed by libhpose (no TensorFlow).

The reference builds its heads with the Keras functional API (``Model-88/attention_model.py:16-169``,
``Model-88/train_88.py:66-253``, ``Model-96/train_96.py:65-110``) and drives them through the Keras
``Model`` duck type: ``compile`` / ``fit`` / ``evaluate`` / ``predict`` / ``save`` / ``count_params`` /
``to_json`` (``train_96.py:99-109,175-196``, ``test.py:21-34``).  This module offers the same
vocabulary -- ``Input``, ``Conv2D``, ``Dense``, ``SpatialDropout2D``, ``Add``, ``Multiply``,
``GlobalAveragePooling2D``, ``Reshape``, ``Lambda``, ``MultiHeadAttention``, ``LayerNormalization``,
``Activation``, ``Model``, ``load_model`` -- but a ``Model`` here is only a *spec* (the same JSON
``model_config`` Keras writes into its ``.h5`` files) plus a flat float32 parameter vector; all
arithmetic happens in CUDA through the C ABI (``hp_head_*``).
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib, h5lite
from ._lib import (HP_ACT, HP_OP_ACT, HP_OP_ADD, HP_OP_DENSE, HP_OP_DROPOUT, HP_OP_GAP, HP_OP_LAYERNORM,
                   HP_OP_MHA, HP_OP_MULCH, HP_OPT, hp_head_op, hp_head_reg, hp_opt_config)

_NAME_COUNTS: Dict[str, int] = {}
_INIT_RNG = np.random.default_rng(1234)


def set_seed(seed: int):
    """Seeds weight initialisation (the reference seeds TF/NumPy at train_96.py:20-23)."""
    global _INIT_RNG
    _INIT_RNG = np.random.default_rng(seed)


def reset_names():
    _NAME_COUNTS.clear()


def _auto_name(base: str) -> str:
    n = _NAME_COUNTS.get(base, 0)
    _NAME_COUNTS[base] = n + 1
    return base if n == 0 else f"{base}_{n}"


def _l2_value(reg) -> float:
    if not reg:
        return 0.0
    if isinstance(reg, dict):
        cfg = reg.get("config", reg)
        return float(cfg.get("l2", 0.0) or 0.0)
    return float(getattr(reg, "l2", 0.0))


class _L2:
    def __init__(self, l2=0.01):
        self.l2 = float(l2)

    def get_config(self):
        return {"module": "keras.regularizers", "class_name": "L2", "config": {"l2": self.l2},
                "registered_name": None}


class regularizers:  # keras.regularizers.l2(...)
    L2 = _L2

    @staticmethod
    def l2(l2=0.01):
        return _L2(l2)


class initializers:  # keras.initializers.GlorotUniform()
    class GlorotUniform:
        def __init__(self, seed=None):
            self.seed = seed


class KTensor:
    """Symbolic tensor: (producing layer, node index) + static shape (None = dynamic)."""

    def __init__(self, layer: "Layer", node: int, shape: Tuple[Optional[int], ...]):
        self.layer, self.node, self.shape = layer, node, tuple(shape)


class Layer:
    class_name = "Layer"
    base_name = "layer"

    def __init__(self, name: Optional[str] = None, **cfg):
        self.name = name or _auto_name(self.base_name)
        self.cfg = cfg
        self.inbound: List[List[KTensor]] = []
        self.kwargs_in: List[dict] = []

    def __call__(self, inputs, *extra, **kw):
        ins = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        ins += [e for e in extra if isinstance(e, KTensor)]
        self.inbound.append(ins)
        return KTensor(self, len(self.inbound) - 1, self.out_shape([t.shape for t in ins]))

    def out_shape(self, shapes):
        return shapes[0]

    def get_config(self) -> dict:
        d = {"name": self.name, "trainable": True, "dtype": "float32"}
        d.update(self.cfg)
        return d

    def param_shapes(self, in_shapes) -> List[Tuple[str, Tuple[int, ...], str]]:
        """[(variable name, shape, initialiser kind)]"""
        return []


class InputLayer(Layer):
    class_name, base_name = "InputLayer", "input"


def Input(shape, name=None) -> KTensor:
    lay = InputLayer(name=name or _auto_name("input"), batch_input_shape=[None] + list(shape), sparse=False,
                     ragged=False)
    lay.inbound.append([])
    return KTensor(lay, 0, (None,) + tuple(shape))


def _reg_cfg(reg):
    if reg is None:
        return None
    return reg.get_config() if hasattr(reg, "get_config") else reg


class Conv2D(Layer):
    class_name, base_name = "Conv2D", "conv2d"

    def __init__(self, filters, kernel_size=1, strides=1, padding="valid", activation=None, use_bias=True,
                 kernel_initializer=None, kernel_regularizer=None, bias_regularizer=None, name=None, **_):
        ks = [kernel_size, kernel_size] if isinstance(kernel_size, int) else list(kernel_size)
        st = [strides, strides] if isinstance(strides, int) else list(strides)
        if ks != [1, 1] or st != [1, 1]:
            raise ValueError("regressor heads only use 1x1 convolutions (SURVEY 8a-a5)")
        super().__init__(name, filters=int(filters), kernel_size=ks, strides=st, padding=padding,
                         data_format="channels_last", dilation_rate=[1, 1], groups=1,
                         activation=activation or "linear", use_bias=bool(use_bias),
                         kernel_regularizer=_reg_cfg(kernel_regularizer), bias_regularizer=_reg_cfg(bias_regularizer))

    def out_shape(self, shapes):
        return shapes[0][:-1] + (self.cfg["filters"],)

    def param_shapes(self, in_shapes):
        cin = in_shapes[0][-1]
        return [("kernel", (1, 1, cin, self.cfg["filters"]), "glorot"), ("bias", (self.cfg["filters"],), "zeros")]


class Dense(Layer):
    class_name, base_name = "Dense", "dense"

    def __init__(self, units, activation=None, use_bias=True, kernel_regularizer=None, bias_regularizer=None,
                 name=None, **_):
        super().__init__(name, units=int(units), activation=activation or "linear", use_bias=bool(use_bias),
                         kernel_regularizer=_reg_cfg(kernel_regularizer), bias_regularizer=_reg_cfg(bias_regularizer))

    def out_shape(self, shapes):
        return shapes[0][:-1] + (self.cfg["units"],)

    def param_shapes(self, in_shapes):
        cin = in_shapes[0][-1]
        return [("kernel", (cin, self.cfg["units"]), "glorot"), ("bias", (self.cfg["units"],), "zeros")]


class SpatialDropout2D(Layer):
    class_name, base_name = "SpatialDropout2D", "spatial_dropout2d"

    def __init__(self, rate, name=None, **_):
        super().__init__(name, rate=float(rate), noise_shape=None, seed=None)


class Dropout(SpatialDropout2D):
    class_name, base_name = "Dropout", "dropout"


class Add(Layer):
    class_name, base_name = "Add", "add"


class Multiply(Layer):
    class_name, base_name = "Multiply", "multiply"

    def out_shape(self, shapes):
        return max(shapes, key=len)


class Activation(Layer):
    class_name, base_name = "Activation", "activation"

    def __init__(self, activation, name=None, **_):
        super().__init__(name, activation=activation)


class ReLU(Layer):
    class_name, base_name = "ReLU", "re_lu"


class GlobalAveragePooling2D(Layer):
    class_name, base_name = "GlobalAveragePooling2D", "global_average_pooling2d"

    def __init__(self, name=None, **_):
        super().__init__(name, data_format="channels_last", keepdims=False)

    def out_shape(self, shapes):
        return (shapes[0][0], shapes[0][-1])


class Reshape(Layer):
    class_name, base_name = "Reshape", "reshape"

    def __init__(self, target_shape, name=None, **_):
        super().__init__(name, target_shape=list(target_shape))

    def out_shape(self, shapes):
        return (shapes[0][0],) + tuple(self.cfg["target_shape"])


class Flatten(Layer):
    class_name, base_name = "Flatten", "flatten"

    def out_shape(self, shapes):
        return (shapes[0][0], shapes[0][-1])


class Lambda(Layer):
    """Only the two token reshapes of attention_model.py:42-49,66-74 exist in the reference:
    one input -> (B,H,W,C)->(B,HW,C); two inputs -> back to the spatial shape of the second."""
    class_name, base_name = "Lambda", "lambda"

    def __init__(self, function=None, name=None, **_):
        super().__init__(name, function="token_reshape", function_type="raw", arguments={})

    def out_shape(self, shapes):
        if len(shapes) == 1:
            return (shapes[0][0], None, shapes[0][-1])
        return shapes[1][:-1] + (shapes[0][-1],)


class LayerNormalization(Layer):
    class_name, base_name = "LayerNormalization", "layer_normalization"

    def __init__(self, axis=-1, epsilon=1e-3, name=None, **_):
        super().__init__(name, axis=[axis] if isinstance(axis, int) else list(axis), epsilon=float(epsilon),
                         center=True, scale=True)

    def param_shapes(self, in_shapes):
        c = in_shapes[0][-1]
        return [("gamma", (c,), "ones"), ("beta", (c,), "zeros")]


class MultiHeadAttention(Layer):
    class_name, base_name = "MultiHeadAttention", "multi_head_attention"

    def __init__(self, num_heads, key_dim, value_dim=None, dropout=0.0, use_bias=True, name=None, **_):
        if value_dim not in (None, key_dim):
            raise ValueError("value_dim != key_dim is not used by the reference and unsupported")
        if dropout:
            raise ValueError("attention dropout is not used by the reference and unsupported")
        super().__init__(name, num_heads=int(num_heads), key_dim=int(key_dim), value_dim=int(key_dim), dropout=0.0,
                         use_bias=bool(use_bias), output_shape=None, attention_axes=[1])

    def __call__(self, query, value=None, key=None, **kw):
        value = value if value is not None else query
        if value.layer is not query.layer or value.node != query.node or key is not None:
            raise ValueError("only self-attention (query is value) is supported (attention_model.py:52-55)")
        self.inbound.append([query])
        self.kwargs_in.append({"value": [query.layer.name, query.node, 0]})
        return KTensor(self, len(self.inbound) - 1, query.shape)

    def param_shapes(self, in_shapes):
        c = in_shapes[0][-1]
        h, d = self.cfg["num_heads"], self.cfg["key_dim"]
        out = []
        for p in ("query", "key", "value"):
            out += [(f"{p}/kernel", (c, h, d), "glorot_einsum"), (f"{p}/bias", (h, d), "zeros")]
        out += [("attention_output/kernel", (h, d, c), "glorot_einsum_out"), ("attention_output/bias", (c,), "zeros")]
        return out


LAYER_CLASSES = {c.class_name: c for c in (InputLayer, Conv2D, Dense, SpatialDropout2D, Dropout, Add, Multiply,
                                           Activation, ReLU, GlobalAveragePooling2D, Reshape, Flatten, Lambda,
                                           LayerNormalization, MultiHeadAttention)}


def _glorot(shape, kind, rng) -> np.ndarray:
    """Keras GlorotUniform: limit = sqrt(6/(fan_in+fan_out)); fans per SURVEY App. B.5."""
    if kind == "zeros":
        return np.zeros(shape, np.float32)
    if kind == "ones":
        return np.ones(shape, np.float32)
    if kind == "glorot":
        recept = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
        fan_in, fan_out = shape[-2] * recept, shape[-1] * recept
    elif kind == "glorot_einsum":      # (C,h,d): Keras _compute_fans on a 3-D shape: receptive = C
        fan_in, fan_out = shape[1] * shape[0], shape[2] * shape[0]
    else:                              # (h,d,C)
        fan_in, fan_out = shape[1] * shape[0], shape[2] * shape[0]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


# ====================================================================== graph config <-> program
def _layer_name(l: dict) -> str:
    return l.get("name", l["config"].get("name"))


class HeadProgram:
    """hp_head_op program + parameter layout compiled from a Keras Functional config."""

    def __init__(self, config: dict):
        if config.get("class_name") in ("Functional", "Model"):
            config = config["config"]
        self.config = config
        self.ops: List[dict] = []
        self.regs: List[Tuple[int, int]] = []          # (channels, per_image)
        self.layout: List[Tuple[str, int, Tuple[int, ...]]] = []   # (weight key, offset, shape)
        self.n_params = 0
        self.dropout_ops: Dict[str, int] = {}          # layer name -> op_id
        self._compile()

    def _new_reg(self, ch, per_image):
        self.regs.append((int(ch), int(per_image)))
        return len(self.regs) - 1

    def _alloc(self, key, shape):
        off = self.n_params
        self.layout.append((key, off, tuple(int(s) for s in shape)))
        self.n_params += int(np.prod(shape))
        return off

    def _compile(self):
        cfg = self.config
        layers = cfg["layers"]
        if len(cfg["input_layers"]) != 1 or len(cfg["output_layers"]) != 1:
            raise ValueError("a regressor head has exactly one input and one output")
        val: Dict[Tuple[str, int], int] = {}
        for l in layers:
            cn, c, name = l["class_name"], l["config"], _layer_name(l)
            if cn == "InputLayer":
                shape = c["batch_input_shape"]
                if len(shape) != 4:
                    raise ValueError(f"head input must be (None,H,W,C), got {shape}")
                val[(name, 0)] = self._new_reg(shape[-1], 0)
                self.in_channels = int(shape[-1])
                continue
            for node_idx, node in enumerate(l["inbound_nodes"]):
                ins = [val[(r[0], r[1])] for r in node]
                out = self._emit(cn, c, name, ins)
                val[(name, node_idx)] = out
        o = cfg["output_layers"][0]
        self.out_reg = val[(o[0], o[1])]
        if self.regs[self.out_reg][1]:
            raise ValueError("head output must be a per-pixel map")
        # outputs may alias register 0 only for degenerate graphs
        if self.out_reg == 0:
            raise ValueError("head has no computation")
        self.out_channels = self.regs[self.out_reg][0]

    def _emit(self, cn, c, name, ins) -> int:
        regs = self.regs
        r0 = ins[0]
        ch0, pi0 = regs[r0]

        def op(**kw):
            d = dict(op=0, in0=0, in1=-1, out=0, cin=0, cout=0, act=0, w_off=0, b_off=0, heads=0, key_dim=0,
                     op_id=len(self.ops), fparam=0.0, l2_w=0.0, l2_b=0.0, layer=name)
            d.update(kw)
            self.ops.append(d)
            return d["out"]

        if cn in ("Conv2D", "Dense"):
            if cn == "Conv2D" and (list(c["kernel_size"]) != [1, 1] or list(c["strides"]) != [1, 1]):
                raise ValueError(f"{name}: only 1x1 stride-1 convolutions are supported in heads")
            cout = int(c["filters"] if cn == "Conv2D" else c["units"])
            act = c.get("activation") or "linear"
            if act not in HP_ACT:
                raise ValueError(f"{name}: activation '{act}' unsupported")
            kshape = (1, 1, ch0, cout) if cn == "Conv2D" else (ch0, cout)
            w = self._alloc(f"{name}/kernel", kshape)
            b = self._alloc(f"{name}/bias", (cout,))
            if not c.get("use_bias", True):
                raise ValueError(f"{name}: use_bias=False unsupported")
            return op(op=HP_OP_DENSE, in0=r0, out=self._new_reg(cout, pi0), cin=ch0, cout=cout, act=HP_ACT[act],
                      w_off=w, b_off=b, l2_w=_l2_value(c.get("kernel_regularizer")),
                      l2_b=_l2_value(c.get("bias_regularizer")))
        if cn in ("SpatialDropout2D", "Dropout"):
            o = op(op=HP_OP_DROPOUT, in0=r0, out=self._new_reg(ch0, pi0), cin=ch0, cout=ch0, fparam=float(c["rate"]))
            self.dropout_ops[name] = self.ops[-1]["op_id"]
            return o
        if cn == "Add":
            if len(ins) != 2:
                raise ValueError(f"{name}: Add takes two inputs")
            pi = int(regs[ins[0]][1] and regs[ins[1]][1])
            return op(op=HP_OP_ADD, in0=ins[0], in1=ins[1], out=self._new_reg(ch0, pi), cin=ch0, cout=ch0)
        if cn == "Multiply":
            if len(ins) != 2:
                raise ValueError(f"{name}: Multiply takes two inputs")
            a, b = ins
            if regs[a][1] and not regs[b][1]:
                a, b = b, a
            if regs[a][1] or not regs[b][1]:
                raise ValueError(f"{name}: only feature-map x per-image-gate products are supported")
            return op(op=HP_OP_MULCH, in0=a, in1=b, out=self._new_reg(regs[a][0], 0), cin=regs[a][0], cout=regs[a][0])
        if cn in ("Activation", "ReLU"):
            act = "relu" if cn == "ReLU" else c["activation"]
            if act not in HP_ACT:
                raise ValueError(f"{name}: activation '{act}' unsupported")
            return op(op=HP_OP_ACT, in0=r0, out=self._new_reg(ch0, pi0), cin=ch0, cout=ch0, act=HP_ACT[act])
        if cn == "GlobalAveragePooling2D":
            return op(op=HP_OP_GAP, in0=r0, out=self._new_reg(ch0, 1), cin=ch0, cout=ch0)
        if cn in ("Reshape", "Flatten", "Lambda"):
            return r0     # pure re-views of [rows][channels] data (token flatten / (1,1,C) / identity)
        if cn == "LayerNormalization":
            g = self._alloc(f"{name}/gamma", (ch0,))
            b = self._alloc(f"{name}/beta", (ch0,))
            return op(op=HP_OP_LAYERNORM, in0=r0, out=self._new_reg(ch0, pi0), cin=ch0, cout=ch0, w_off=g, b_off=b,
                      fparam=float(c["epsilon"]))
        if cn == "MultiHeadAttention":
            h, d = int(c["num_heads"]), int(c["key_dim"])
            if int(c.get("value_dim") or d) != d or float(c.get("dropout", 0.0)) != 0.0:
                raise ValueError(f"{name}: value_dim/dropout variants unsupported")
            w0 = None
            for p in ("query", "key", "value"):
                w = self._alloc(f"{name}/{p}/kernel", (ch0, h, d))
                self._alloc(f"{name}/{p}/bias", (h, d))
                w0 = w if w0 is None else w0
            self._alloc(f"{name}/attention_output/kernel", (h, d, ch0))
            self._alloc(f"{name}/attention_output/bias", (ch0,))
            return op(op=HP_OP_MHA, in0=r0, out=self._new_reg(ch0, 0), cin=ch0, cout=ch0, w_off=w0, heads=h, key_dim=d)
        raise ValueError(f"layer class {cn} is not part of the regressor-head vocabulary")

    # ---- parameter (un)packing
    def pack(self, weights: Dict[str, np.ndarray]) -> np.ndarray:
        flat = np.zeros(self.n_params, np.float32)
        for key, off, shape in self.layout:
            if key not in weights:
                raise KeyError(f"missing weight '{key}'")
            w = np.asarray(weights[key], np.float32)
            if tuple(w.shape) != shape:
                raise ValueError(f"weight '{key}' has shape {w.shape}, expected {shape}")
            flat[off:off + w.size] = w.reshape(-1)
        return flat

    def unpack(self, flat: np.ndarray) -> Dict[str, np.ndarray]:
        return {key: np.asarray(flat[off:off + int(np.prod(shape))], np.float32).reshape(shape).copy()
                for key, off, shape in self.layout}

    def c_arrays(self, inference: bool):
        ops = self.ops
        arr_ops = (hp_head_op * len(ops))()
        for i, o in enumerate(ops):
            a = arr_ops[i]
            for f, _ in hp_head_op._fields_:
                setattr(a, f, o[f])
            if inference and o["op"] == HP_OP_DROPOUT:
                a.fparam = 0.0
        arr_regs = (hp_head_reg * len(self.regs))()
        for i, (ch, pi) in enumerate(self.regs):
            arr_regs[i].channels, arr_regs[i].per_image = ch, pi
        return arr_ops, arr_regs


# ====================================================================== optimizers / callbacks
class _Optimizer:
    kind = "sgd"

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, **_):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon

    def c_config(self) -> hp_opt_config:
        return hp_opt_config(HP_OPT[self.kind], float(self.learning_rate), float(self.beta_1), float(self.beta_2),
                             float(self.epsilon))

    def get_config(self):
        name = {"sgd": "SGD", "adam": "Adam", "adamax": "Adamax"}[self.kind]
        cfg = {"name": name, "learning_rate": float(np.float32(self.learning_rate))}
        if self.kind == "sgd":
            cfg.update(decay=0.0, momentum=0.0, nesterov=False)
        else:
            cfg.update(beta_1=self.beta_1, beta_2=self.beta_2, epsilon=self.epsilon)
        return {"class_name": name, "config": cfg}


class SGD(_Optimizer):
    kind = "sgd"

    def __init__(self, learning_rate=0.01, momentum=0.0, nesterov=False, **kw):
        if momentum or nesterov:
            raise ValueError("SGD momentum is not used by the reference and unsupported")
        super().__init__(learning_rate, **kw)


class Adam(_Optimizer):
    kind = "adam"


class Adamax(_Optimizer):
    kind = "adamax"


class optimizers:
    SGD, Adam, Adamax = SGD, Adam, Adamax


def _make_optimizer(opt) -> _Optimizer:
    if isinstance(opt, _Optimizer):
        return opt
    if isinstance(opt, str):
        return {"sgd": SGD, "adam": Adam, "adamax": Adamax}[opt.lower()]()
    if isinstance(opt, dict):
        cls = {"sgd": SGD, "adam": Adam, "adamax": Adamax}[opt["class_name"].lower()]
        c = opt.get("config", {})
        return cls(learning_rate=c.get("learning_rate", 0.001), beta_1=c.get("beta_1", 0.9),
                   beta_2=c.get("beta_2", 0.999), epsilon=c.get("epsilon", 1e-7))
    raise ValueError(f"unknown optimizer {opt!r}")


class Callback:
    model: "Model" = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass


class History(Callback):
    def __init__(self):
        self.history: Dict[str, List[float]] = {}
        self.epoch: List[int] = []

    def on_epoch_end(self, epoch, logs=None):
        self.epoch.append(epoch)
        for k, v in (logs or {}).items():
            self.history.setdefault(k, []).append(v)


class EarlyStopping(Callback):
    """keras.callbacks.EarlyStopping(monitor, patience, min_delta, restore_best_weights) for 'min' monitors
    (train_96.py:159-164): improvement iff value < best - min_delta."""

    def __init__(self, monitor="val_loss", patience=0, min_delta=0.0, restore_best_weights=False, **_):
        self.monitor, self.patience, self.min_delta = monitor, int(patience), abs(float(min_delta))
        self.restore_best_weights = restore_best_weights

    def on_train_begin(self, logs=None):
        self.best, self.wait, self.best_weights, self.stopped_epoch, self.best_epoch = math.inf, 0, None, 0, 0

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        if self.restore_best_weights and self.best_weights is None:
            self.best_weights = self.model.get_flat_weights()
        self.wait += 1
        if cur < self.best - self.min_delta:
            self.best, self.best_epoch, self.wait = cur, epoch, 0
            if self.restore_best_weights:
                self.best_weights = self.model.get_flat_weights()
        if self.wait >= self.patience and epoch > 0:
            self.stopped_epoch = epoch
            self.model.stop_training = True
            if self.restore_best_weights and self.best_weights is not None:
                self.model.set_flat_weights(self.best_weights)


class ModelCheckpoint(Callback):
    """keras.callbacks.ModelCheckpoint(filepath, monitor, save_best_only) (train_96.py:154-158)."""

    def __init__(self, filepath, monitor="val_loss", save_best_only=False, **_):
        self.filepath, self.monitor, self.save_best_only = filepath, monitor, save_best_only
        self.best = math.inf

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if self.save_best_only:
            if cur is None or not cur < self.best:
                return
            self.best = cur
        if getattr(self.model, "_fit_rank", 0) != 0:
            return                                     # data-parallel fit: the weights are identical on every rank, rank 0 writes
        self.model.save(self.filepath.format(epoch=epoch + 1, **(logs or {})))


class callbacks:
    Callback, History, EarlyStopping, ModelCheckpoint = Callback, History, EarlyStopping, ModelCheckpoint


# ====================================================================== Model
class Model:
    """Keras-``Model``-shaped handle on a regressor head living on one B200."""

    def __init__(self, inputs=None, outputs=None, name=None, *, _config: Optional[dict] = None,
                 _weights: Optional[Dict[str, np.ndarray]] = None):
        if _config is None:
            _config, _weights = self._trace(inputs, outputs, name or _auto_name("model"))
        if _config.get("class_name") not in ("Functional", "Model"):
            raise ValueError("only Functional model configs are supported")
        self._config = _config
        self.name = _config["config"].get("name", "model")
        self.program = HeadProgram(_config)
        missing = [k for k, _, _ in self.program.layout if k not in (_weights or {})]
        if missing:
            raise KeyError(f"weights missing for {missing[:3]}...")
        self._flat = self.program.pack(_weights)
        self._device = None           # (ctx, head handle for training, head handle for inference)
        self.optimizer: Optional[_Optimizer] = None
        self.loss = None
        self.metrics_names = ["loss"]
        self.stop_training = False
        self.history = None
        self._train_steps = 0

    # ---- tracing the functional graph into a Keras-format config
    @staticmethod
    def _trace(inputs, outputs, name):
        if inputs is None or outputs is None:
            raise ValueError("Model(inputs, outputs) required")
        ins = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        outs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
        order: List[Layer] = []
        seen = set()

        def visit(t: KTensor):
            if id(t.layer) in seen:
                return
            for node in t.layer.inbound:
                for src in node:
                    visit(src)
            seen.add(id(t.layer))
            order.append(t.layer)

        for o in outs:
            visit(o)
        layers, weights = [], {}
        for lay in order:
            nodes = []
            for ni, node in enumerate(lay.inbound):
                kw = lay.kwargs_in[ni] if ni < len(lay.kwargs_in) else {}
                nodes.append([[s.layer.name, s.node, 0, dict(kw) if j == 0 else {}] for j, s in enumerate(node)])
            if isinstance(lay, InputLayer):
                nodes = []
            layers.append({"class_name": lay.class_name, "config": lay.get_config(), "name": lay.name,
                           "inbound_nodes": nodes})
            if lay.inbound and lay.inbound[0]:
                shapes = [s.shape for s in lay.inbound[0]]
                for var, shape, kind in lay.param_shapes(shapes):
                    weights[f"{lay.name}/{var}"] = _glorot(shape, kind, _INIT_RNG)
        cfg = {"class_name": "Functional",
               "config": {"name": name, "trainable": True, "layers": layers,
                          "input_layers": [[t.layer.name, 0, 0] for t in ins],
                          "output_layers": [[t.layer.name, t.node, 0] for t in outs]},
               "keras_version": "2.13.1", "backend": "tensorflow"}
        return cfg, weights

    # ---- introspection
    def count_params(self) -> int:
        return int(self.program.n_params)

    def to_json(self) -> str:
        return json.dumps(self._config)

    def get_config(self) -> dict:
        return self._config["config"]

    @property
    def input_shape(self):
        return (None, None, None, self.program.in_channels)

    def summary(self, print_fn=print):
        print_fn(f'Model: "{self.name}"')
        for key, off, shape in self.program.layout:
            print_fn(f"  {key:50s} {str(shape):18s} {int(np.prod(shape))}")
        print_fn(f"Total params: {self.count_params()}")

    # ---- weights
    def get_weights_dict(self) -> Dict[str, np.ndarray]:
        return self.program.unpack(self.get_flat_weights())

    def get_weights(self) -> List[np.ndarray]:
        d = self.get_weights_dict()
        return [d[k] for k, _, _ in self.program.layout]

    def set_weights(self, weights):
        if isinstance(weights, dict):
            self.set_flat_weights(self.program.pack(weights))
        else:
            keys = [k for k, _, _ in self.program.layout]
            self.set_flat_weights(self.program.pack(dict(zip(keys, weights))))

    def get_flat_weights(self) -> np.ndarray:
        if self._device is not None:
            ctx, head = self._device
            out = np.empty(self.program.n_params, np.float32)
            _lib.check(_lib.lib().hp_head_get_weights(ctx.handle, head, out.ctypes.data, out.size))
            self._flat = out
        return self._flat.copy()

    def set_flat_weights(self, flat: np.ndarray):
        flat = np.ascontiguousarray(flat, np.float32)
        if flat.size != self.program.n_params:
            raise ValueError(f"expected {self.program.n_params} parameters, got {flat.size}")
        self._flat = flat.copy()
        if self._device is not None:
            ctx, head = self._device
            _lib.check(_lib.lib().hp_head_set_weights(ctx.handle, head, self._flat.ctypes.data, flat.size))

    # ---- device residency
    def to_device(self, ctx=None):
        """Create the hp_head on ``ctx`` (default: the process-wide context of the current CUDA device)."""
        from .device import default_context
        ctx = ctx or default_context()
        if self._device is not None and self._device[0] is ctx:
            return self
        if self._device is not None:
            self.get_flat_weights()
            self.release()
        ops, regs = self.program.c_arrays(inference=False)
        head = C.c_void_p()
        _lib.check(_lib.lib().hp_head_create(ctx.handle, ops, len(self.program.ops), regs, len(self.program.regs),
                                             self.program.out_reg, self.program.n_params, C.byref(head)))
        self._device = (ctx, head)
        _lib.check(_lib.lib().hp_head_set_weights(ctx.handle, head, self._flat.ctypes.data, self._flat.size))
        return self

    def release(self):
        if self._device is not None:
            ctx, head = self._device
            _lib.lib().hp_head_destroy(ctx.handle, head)
            self._device = None

    @property
    def head_handle(self):
        self.to_device()
        return self._device[1]

    # ---- inference (test.py:34, JoinModels regressor call)
    def predict_device(self, feat):
        """feat: CUDA float32 torch tensor (B,H,W,C) -> CUDA tensor (B,H,W,out)."""
        import torch
        self.to_device()
        ctx, head = self._device
        if feat.dim() != 4 or feat.shape[-1] != self.program.in_channels:
            raise ValueError(f"expected input (B,H,W,{self.program.in_channels}), got {tuple(feat.shape)}")
        feat = feat.contiguous().float()
        B, H, W, _ = feat.shape
        out = torch.empty((B, H, W, self.program.out_channels), dtype=torch.float32, device=feat.device)
        _lib.check(_lib.lib().hp_head_forward(ctx.handle, head, feat.data_ptr(), B, H, W, out.data_ptr(),
                                              ctx.stream_ptr()))
        return out

    def predict(self, x, batch_size=None, verbose=0, **_) -> np.ndarray:
        import torch
        from .device import default_context
        ctx = default_context()
        x = np.ascontiguousarray(x, np.float32)
        if x.ndim != 4:
            raise ValueError(f"expected (N,H,W,C) input, got shape {x.shape}")
        outs = []
        chunk = int(batch_size) if batch_size else max(1, min(len(x), (1 << 22) // max(1, x.shape[1] * x.shape[2])))
        for i in range(0, len(x), chunk):
            xt = torch.from_numpy(x[i:i + chunk]).to(ctx.torch_device)
            outs.append(self.predict_device(xt).cpu().numpy())
        return np.concatenate(outs, axis=0) if outs else np.zeros((0,) + x.shape[1:3] + (self.program.out_channels,),
                                                                   np.float32)

    __call__ = predict

    # ---- training (train_96.py:99-109,175-187)
    def compile(self, optimizer="sgd", loss="mse", metrics=None, **_):
        if loss not in ("mse", "mean_squared_error"):
            raise ValueError("only the 'mse' loss of the reference configs is implemented")
        self.optimizer = _make_optimizer(optimizer)
        self.loss = "mse"
        self.metrics_names = ["loss"] + [m for m in (metrics or [])]
        if any(m not in ("mae", "mean_absolute_error") for m in self.metrics_names[1:]):
            raise ValueError("only the 'mae' metric of the reference configs is implemented")

    def train_on_device(self, x, y, n_global=None, seed=0, want_loss=True):
        """One optimizer step on CUDA tensors x (n,H,W,C), y (n,H,W,3). Returns (loss, mae) or None."""
        if self.optimizer is None:
            raise RuntimeError("call compile() before training")
        self.to_device()
        ctx, head = self._device
        n, H, W, _ = x.shape
        res = (C.c_float * 2)()
        opt = self.optimizer.c_config()
        _lib.check(_lib.lib().hp_head_train_step(ctx.handle, head, x.data_ptr(), y.data_ptr(), n, H, W,
                                                 int(n_global or n), C.byref(opt), C.c_uint64(seed),
                                                 C.cast(res, C.c_void_p) if want_loss else None, ctx.stream_ptr()))
        self._train_steps += 1
        return (float(res[0]), float(res[1])) if want_loss else None

    def evaluate_device(self, x, y):
        self.to_device()
        ctx, head = self._device
        n, H, W, _ = x.shape
        res = (C.c_float * 3)()
        _lib.check(_lib.lib().hp_head_evaluate(ctx.handle, head, x.data_ptr(), y.data_ptr(), n, H, W,
                                               C.cast(res, C.c_void_p), ctx.stream_ptr()))
        return float(res[0]) + float(res[2]), float(res[1])

    def evaluate(self, x, y, batch_size=None, verbose=0, **_):
        import torch
        from .device import default_context
        ctx = default_context()
        xt = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(ctx.torch_device)
        yt = torch.from_numpy(np.ascontiguousarray(y, np.float32)).to(ctx.torch_device)
        loss, mae = self.evaluate_device(xt, yt)
        if verbose:
            print(f"loss: {loss:.4f} - mae: {mae:.4f}")
        return [loss, mae]

    def train_run_device(self, x_all, y_all, idx, first_item, batch_global, n_steps, rank=0, world=1, seed=0, graph=True,
                         want_sums=True):
        """``n_steps`` optimizer steps on a device-resident data set (``hp_head_train_run``): step s uses the global batch
        ``idx[first_item + s * batch_global ...]`` (``idx``: CUDA int32 tensor or None), of which this rank takes rows
        rank, rank + world, ...  Returns (sum of loss * batch_global, sum of mae * batch_global) or None."""
        if self.optimizer is None:
            raise RuntimeError("call compile() before training")
        self.to_device()
        ctx, head = self._device
        n_items, H, W, _ = x_all.shape
        sums = (C.c_double * 2)()
        opt = self.optimizer.c_config()
        _lib.check(_lib.lib().hp_head_train_run(
            ctx.handle, head, x_all.data_ptr(), y_all.data_ptr(), idx.data_ptr() if idx is not None else None, int(n_items),
            int(first_item), int(batch_global), int(n_steps), int(rank), int(world), H, W, C.byref(opt), C.c_uint64(seed),
            _lib.HP_TRAIN_GRAPH if graph else 0, C.cast(sums, C.c_void_p) if want_sums else None, ctx.stream_ptr()))
        self._train_steps += int(n_steps)
        return (float(sums[0]), float(sums[1])) if want_sums else None

    def fit(self, x, y, epochs=1, batch_size=32, validation_data=None, callbacks=None, verbose=1, shuffle=True,
            seed=42, distributed=None, graph=True, **_):
        """model.fit (train_96.py:175-183): shuffled mini-batches (final partial batch included), per-epoch validation,
        callbacks.  The data set stays on the device; an epoch is one ``hp_head_train_run`` call for the full batches (each
        step one CUDA-graph launch, no host round trip: loss and mae are accumulated on the device and read once per epoch)
        plus one for the partial batch.

        With ``distributed`` (a ``parallel.DataParallel`` object) every global batch is split across the ranks (rank r takes
        rows r, r + world, ... of it; a rank left without rows in a short final batch contributes zeros) and the gradients
        are all-reduced inside the step; the weights stay identical on all ranks."""
        import torch
        from .device import default_context
        if self.optimizer is None:
            raise RuntimeError("call compile() before fit()")
        ctx = default_context()
        dev = ctx.torch_device
        rank, world = (distributed.rank, distributed.world_size) if distributed else (0, 1)
        if world > 1 and not distributed._comm_ready:
            distributed.init_gradient_comm()               # collective: every rank reaches fit() (never train un-reduced)
        self._fit_rank = rank
        xt = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(dev)
        yt = torch.from_numpy(np.ascontiguousarray(y, np.float32)).to(dev)
        val = None
        if validation_data is not None:
            val = (torch.from_numpy(np.ascontiguousarray(validation_data[0], np.float32)).to(dev),
                   torch.from_numpy(np.ascontiguousarray(validation_data[1], np.float32)).to(dev))
        hist = History()
        cbs = [hist] + list(callbacks or [])
        for cb in cbs:
            cb.set_model(self)
            cb.on_train_begin()
        self.stop_training = False
        n = xt.shape[0]
        batch_size = int(batch_size)
        n_full, tail = divmod(n, batch_size)
        rng = np.random.default_rng(seed)
        perm_t = torch.empty(n, dtype=torch.int32, device=dev)            # one buffer for all epochs: the captured step reads it
        side = torch.cuda.Stream(dev)                                      # graph capture needs a non-default stream
        side.wait_stream(torch.cuda.current_stream(dev))
        for epoch in range(int(epochs)):
            perm = rng.permutation(n) if shuffle else np.arange(n)
            with torch.cuda.stream(side):
                perm_t.copy_(torch.from_numpy(perm.astype(np.int32)))
                tot_loss = tot_mae = 0.0
                if n_full:
                    a, b = self.train_run_device(xt, yt, perm_t, 0, batch_size, n_full, rank, world, seed, graph)
                    tot_loss, tot_mae = tot_loss + a, tot_mae + b
                if tail:
                    a, b = self.train_run_device(xt, yt, perm_t, n_full * batch_size, tail, 1, rank, world, seed, graph)
                    tot_loss, tot_mae = tot_loss + a, tot_mae + b
                logs = {"loss": tot_loss / n, "mae": tot_mae / n}
                if val is not None:
                    vl, vm = self.evaluate_device(val[0], val[1])
                    logs.update(val_loss=vl, val_mae=vm)
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} - " + " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()))
            for cb in cbs:
                cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        torch.cuda.current_stream(dev).wait_stream(side)
        for cb in cbs:
            cb.on_train_end()
        self.history = hist
        return hist

    # ---- persistence (Keras HDF5 layout)
    def training_config(self) -> Optional[dict]:
        if self.optimizer is None:
            return None
        return {"loss": "mse",
                "metrics": [[{"class_name": "MeanMetricWrapper",
                              "config": {"name": "mae", "dtype": "float32", "fn": "mean_absolute_error"}}]],
                "weighted_metrics": None, "loss_weights": None, "optimizer_config": self.optimizer.get_config()}

    def save(self, path: str):
        wd = self.get_weights_dict()
        h5w = {}
        for key, arr in wd.items():
            layer, var = key.split("/", 1)
            h5w[f"{layer}/{layer}/{var}:0"] = arr
        h5lite.write_h5(path, h5w, self._config, self.training_config())


def load_model(path: str) -> Model:
    """tf.keras.models.load_model for regressor-head checkpoints (test.py:21, JoinModels.py:30-31)."""
    if not os.path.exists(path):
        raise FileNotFoundError(f"Model file not found: {path}")
    f = h5lite.H5File(path)
    cfg = f.model_config()
    weights = {}
    for k, v in f.weights().items():
        parts = k.split("/")
        parts[-1] = parts[-1].split(":")[0]
        if len(parts) >= 3 and parts[0] == parts[1]:
            parts = parts[1:]
        weights["/".join(parts)] = v
    m = Model(_config=cfg, _weights=weights)
    tc = f.training_config()
    if tc and tc.get("optimizer_config"):
        try:
            m.compile(optimizer=tc["optimizer_config"], loss="mse", metrics=["mae"])
        except (ValueError, KeyError):
            pass
    return m


def model_from_json(text: str) -> Model:
    """keras.models.model_from_json (utilities.py:37-41): fresh Glorot weights for the given graph."""
    cfg = json.loads(text)
    prog = HeadProgram(cfg)
    weights = {}
    for key, _, shape in prog.layout:
        if key.endswith("gamma"):
            kind = "ones"
        elif key.endswith("kernel"):
            kind = "glorot" if len(shape) != 3 else "glorot_einsum"
        else:
            kind = "zeros"
        weights[key] = _glorot(shape, kind, _INIT_RNG)
    return Model(_config=cfg, _weights=weights)
