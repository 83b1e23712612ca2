"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the Keras graphs.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package never does.

PARITY STATUS: **unpinned for the backbone** -- the reference (TensorFlow/Keras 2.13,
un-vendored: ``requirements.txt:2``) cannot be installed or run in this environment and
ships no tests or golden vectors, so this file restates the Keras layer semantics from
SURVEY.md Appendix B and is anchored on (a) the reference's own serialized graph
(``model_config`` JSON inside ``BlazePoser/UnifiedModels/*.h5`` -- this interpreter
evaluates that JSON directly, layer by layer), (b) the trained weights in the shipped
``.h5`` files and (c) label-MAE known answers on the shipped ``.npz`` datasets
(``tests/test_oracle_kat.py``) which pin the regressor heads.

What is restated, with the reference call sites:

* the unified detector+pose graph evaluated at ``BlazePoser/blazeFaceDetectorH5.py:272``
  (``self.interpreter(x)``), built by ``JoinModels.py:5-90``;
* the regressor heads built by ``Model-88/attention_model.py:16-169``,
  ``Model-88/train_88.py:66-253`` and ``Model-96/train_96.py:65-110`` and run through
  ``model.predict`` (``Model-96/test.py:34``) / ``model.fit`` (``train_96.py:175``);
* Keras semantics: SAME padding (asymmetric, extra on bottom/right), ReLU/softsign/tanh/
  sigmoid, MultiHeadAttention einsum layout, LayerNormalization(eps) over the last axis,
  SpatialDropout2D masks of shape (B,1,1,C), L2 regularisers ``l * sum(w**2)``.

Everything runs on torch CPU tensors (float64 for parity truth, float32 for the timed CPU
baseline) so that gradients for the train-step parity come from autograd.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- helpers
def same_pad(n: int, k: int, s: int):
    """TF 'SAME' padding: out = ceil(n/s); total = max((out-1)*s + k - n, 0); before = total//2."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def activation(name: str, x: torch.Tensor) -> torch.Tensor:
    if name in (None, "linear"):
        return x
    if name == "relu":
        return torch.relu(x)
    if name == "tanh":
        return torch.tanh(x)
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "softsign":
        return x / (1.0 + x.abs())
    if name == "elu":                                   # keras.activations.elu, alpha = 1
        return torch.where(x > 0, x, torch.expm1(x))
    if name == "selu":                                  # keras.activations.selu: scale * elu(x, alpha)
        return 1.0507009873554805 * torch.where(x > 0, x, 1.6732632423543772 * torch.expm1(x))
    if name == "softplus":
        return torch.nn.functional.softplus(x)
    if name == "swish":
        return x * torch.sigmoid(x)
    if name == "leaky_relu":                            # tf.nn.leaky_relu, alpha = 0.2 (how the name gets into a Keras 2.13 config)
        return torch.where(x > 0, x, 0.2 * x)
    raise NotImplementedError(f"activation {name}")


def _nhwc_conv(x, kernel, bias, strides, padding, groups=1):
    """x NHWC, kernel HWIO (Keras layout)."""
    kh, kw = kernel.shape[0], kernel.shape[1]
    sh, sw = strides
    xin = x.permute(0, 3, 1, 2)
    if padding == "same":
        pt, pb = same_pad(x.shape[1], kh, sh)
        pl, pr = same_pad(x.shape[2], kw, sw)
        xin = F.pad(xin, (pl, pr, pt, pb))
    w = kernel.permute(3, 2, 0, 1)
    y = F.conv2d(xin, w, bias, stride=(sh, sw), groups=groups)
    return y.permute(0, 2, 3, 1)


class DropoutSource:
    """Supplies SpatialDropout2D keep-masks; tests pass the same generator the CUDA path uses."""

    def __init__(self, fn: Optional[Callable[[str, int, int, float], np.ndarray]] = None):
        self.fn = fn

    def mask(self, layer_name: str, n: int, c: int, rate: float) -> Optional[np.ndarray]:
        if self.fn is None:
            return None
        return self.fn(layer_name, n, c, rate)


class KerasGraph:
    """Evaluates a Keras-2.x Functional ``model_config`` dict on torch CPU tensors.

    ``weights`` maps ``"<layer>/<var>"`` (e.g. ``"conv2d_3/kernel"``, nested
    ``"model/conv2d/bias"``, MHA ``"multi_head_attention/query/kernel"``) to torch tensors.
    """

    def __init__(self, config: dict, weights: Dict[str, torch.Tensor], prefix: str = ""):
        if config.get("class_name") in ("Functional", "Model"):
            config = config["config"]
        self.cfg = config
        self.w = weights
        self.prefix = prefix
        self.layers = config["layers"]
        self.inputs = [n[0] for n in config["input_layers"]]
        self.outputs = [(n[0], n[1]) for n in config["output_layers"]]

    # ---- weights
    def _get(self, layer: str, var: str) -> torch.Tensor:
        key = f"{self.prefix}{layer}/{var}"
        if key not in self.w:
            raise KeyError(f"missing weight {key}")
        return self.w[key]

    def weight_keys(self) -> List[str]:
        keys = []
        for l in self.layers:
            cn, name = l["class_name"], l.get("name", l["config"].get("name"))
            if cn in ("Conv2D", "Dense"):
                keys += [f"{self.prefix}{name}/kernel", f"{self.prefix}{name}/bias"]
            elif cn == "DepthwiseConv2D":
                keys += [f"{self.prefix}{name}/depthwise_kernel", f"{self.prefix}{name}/bias"]
            elif cn == "LayerNormalization":
                keys += [f"{self.prefix}{name}/gamma", f"{self.prefix}{name}/beta"]
            elif cn == "MultiHeadAttention":
                for p in ("query", "key", "value", "attention_output"):
                    keys += [f"{self.prefix}{name}/{p}/kernel", f"{self.prefix}{name}/{p}/bias"]
            elif cn in ("Functional", "Model"):
                keys += KerasGraph(l["config"], self.w, f"{self.prefix}{name}/").weight_keys()
        return keys

    # ---- regularisation (Keras adds l*sum(w^2) for every regularised tensor)
    def l2_penalty(self) -> torch.Tensor:
        total = torch.zeros((), dtype=torch.float64)
        for l in self.layers:
            cfg = l["config"]
            lname = l.get("name", cfg.get("name"))
            for regname, var in (("kernel_regularizer", "kernel"), ("bias_regularizer", "bias")):
                reg = cfg.get(regname)
                if not reg:
                    continue
                rc = reg.get("config", reg)
                lam = float(rc.get("l2", 0.0) or 0.0)
                if lam:
                    w = self._get(lname, var)
                    total = total + lam * (w.double() ** 2).sum()
        return total

    # ---- evaluation
    def __call__(self, *args: torch.Tensor, training: bool = False,
                 dropout: Optional[DropoutSource] = None, taps: Optional[dict] = None):
        vals: Dict[tuple, torch.Tensor] = {}
        call_count: Dict[str, int] = {}
        for name, x in zip(self.inputs, args):
            vals[(name, 0)] = x
        for l in self.layers:
            cn, cfg, name = l["class_name"], l["config"], l.get("name", l["config"].get("name"))
            if cn == "InputLayer":
                continue
            for node_idx, node in enumerate(l["inbound_nodes"]):
                ins = []
                for ref in node:
                    ins.append(vals[(ref[0], ref[1])])
                    if len(ref) > 3 and isinstance(ref[3], dict):
                        for kw in ref[3].values():  # MHA passes value=... as kwarg
                            if isinstance(kw, list) and len(kw) == 3:
                                ins.append(vals[(kw[0], kw[1])])
                out = self._apply(cn, cfg, name, ins, training, dropout)
                # nested Functional models are called once at build (node 0) and once when
                # joined (node 1); the outer graph references node index as stored
                vals[(name, node_idx)] = out
                if cn in ("Functional", "Model"):
                    vals[(name, 1)] = out
                    vals[(name, 0)] = out
                if taps is not None:
                    taps[self.prefix + name] = out
        outs = [vals[o] for o in self.outputs]
        return outs if len(outs) > 1 else outs[0]

    def _apply(self, cn, cfg, name, ins, training, dropout):
        x = ins[0]
        if cn == "Conv2D":
            y = _nhwc_conv(x, self._get(name, "kernel"), self._get(name, "bias") if cfg.get("use_bias", True) else None,
                           cfg["strides"], cfg["padding"])
            return activation(cfg.get("activation"), y)
        if cn == "DepthwiseConv2D":
            k = self._get(name, "depthwise_kernel")          # (kh,kw,C,1)
            c = k.shape[2]
            kk = k.permute(0, 1, 3, 2)                        # -> HWIO with I=1,O=C for groups=C
            y = _nhwc_conv(x, kk, self._get(name, "bias") if cfg.get("use_bias", True) else None,
                           cfg["strides"], cfg["padding"], groups=c)
            return activation(cfg.get("activation"), y)
        if cn == "MaxPooling2D":
            ph, pw = cfg["pool_size"]
            sh, sw = cfg["strides"]
            xin = x.permute(0, 3, 1, 2)
            if cfg["padding"] == "same":
                pt, pb = same_pad(x.shape[1], ph, sh)
                pl, pr = same_pad(x.shape[2], pw, sw)
                xin = F.pad(xin, (pl, pr, pt, pb), value=float("-inf"))
            return F.max_pool2d(xin, (ph, pw), (sh, sw)).permute(0, 2, 3, 1)
        if cn == "TensorFlowOpLayer":
            op = cfg["node_def"]["op"]
            const = cfg.get("constants", {})
            c1 = const.get("1", const.get(1))
            if op == "Pad":
                pads = c1
                flat = []
                for lo, hi in reversed(pads):
                    flat += [int(lo), int(hi)]
                return F.pad(x, flat)
            if op == "Reshape":
                shape = [int(v) for v in c1]
                # batch-generalise [1, A, K] -> [B, A*?, K]: keep the last dim
                return x.reshape(x.shape[0], -1, shape[-1])
            raise NotImplementedError(op)
        if cn == "Add":
            y = ins[0]
            for t in ins[1:]:
                y = y + t
            return y
        if cn == "Multiply":
            y = ins[0]
            for t in ins[1:]:
                y = y * t
            return y
        if cn == "ReLU":
            return torch.relu(x)
        if cn == "Activation":
            return activation(cfg["activation"], x)
        if cn == "Reshape":
            ts = cfg["target_shape"]
            if len(ts) == 3 and x.dim() == 4 and ts[-1] == x.shape[-1]:
                return x   # identity re-shape of a feature map (JoinModels.py:58,62); keeps H,W generic
            return x.reshape(x.shape[0], *ts)
        if cn == "Flatten":
            return x.reshape(x.shape[0], -1)
        if cn in ("SpatialDropout2D", "Dropout"):
            rate = float(cfg["rate"])
            if not training or rate <= 0.0:
                return x
            m = dropout.mask(self.prefix + name, x.shape[0], x.shape[-1], rate) if dropout else None
            if m is None:
                raise RuntimeError("training with dropout needs a DropoutSource")
            mt = torch.as_tensor(m, dtype=x.dtype).reshape(x.shape[0], 1, 1, x.shape[-1])
            return x * mt / (1.0 - rate)
        if cn == "GlobalAveragePooling2D":
            return x.mean(dim=(1, 2))
        if cn == "Dense":
            y = x @ self._get(name, "kernel")
            if cfg.get("use_bias", True):
                y = y + self._get(name, "bias")
            return activation(cfg.get("activation"), y)
        if cn == "Lambda":
            if len(ins) == 1:      # reshape_flat (attention_model.py:42-49)
                return x.reshape(x.shape[0], x.shape[1] * x.shape[2], x.shape[3])
            t, orig = ins          # reshape_back (attention_model.py:66-74)
            return t.reshape(orig.shape[0], orig.shape[1], orig.shape[2], t.shape[2])
        if cn == "LayerNormalization":
            eps = float(cfg["epsilon"])
            mu = x.mean(dim=-1, keepdim=True)
            var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
            y = (x - mu) / torch.sqrt(var + eps)
            return y * self._get(name, "gamma") + self._get(name, "beta")
        if cn == "MultiHeadAttention":
            q_in = ins[0]
            v_in = ins[1] if len(ins) > 1 else ins[0]
            k_in = ins[2] if len(ins) > 2 else v_in
            d = int(cfg["key_dim"])
            wq, bq = self._get(name, "query/kernel"), self._get(name, "query/bias")
            wk, bk = self._get(name, "key/kernel"), self._get(name, "key/bias")
            wv, bv = self._get(name, "value/kernel"), self._get(name, "value/bias")
            wo, bo = self._get(name, "attention_output/kernel"), self._get(name, "attention_output/bias")
            q = torch.einsum("btc,chd->bthd", q_in, wq) + bq
            k = torch.einsum("bsc,chd->bshd", k_in, wk) + bk
            v = torch.einsum("bsc,chd->bshd", v_in, wv) + bv
            q = q * (1.0 / math.sqrt(d))
            s = torch.einsum("bshd,bthd->bhts", k, q)
            p = torch.softmax(s, dim=-1)
            o = torch.einsum("bhts,bshd->bthd", p, v)
            return torch.einsum("bthd,hdc->btc", o, wo) + bo
        if cn in ("Functional", "Model"):
            sub = KerasGraph(cfg, self.w, f"{self.prefix}{name}/")
            return sub(*ins, training=training, dropout=dropout)
        raise NotImplementedError(f"layer class {cn}")


def to_torch(weights: Dict[str, np.ndarray], dtype=torch.float64, requires_grad=False) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in weights.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        if requires_grad:
            t.requires_grad_(True)
        out[k] = t
    return out


def normalise_weight_names(h5_weights: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """``conv2d/conv2d/kernel:0`` -> ``conv2d/kernel``; nested ``model/conv2d/kernel:0`` stays nested.

    Keras writes top-level layer weights as ``<layer>/<layer>/<var>:0`` and nested-model
    weights as ``<model>/<layer>/<var>:0``.
    """
    out = {}
    for k, v in h5_weights.items():
        parts = k.split("/")
        parts[-1] = parts[-1].split(":")[0]
        if len(parts) >= 3 and parts[0] == parts[1]:
            parts = parts[1:]
        out["/".join(parts)] = v
    return out


# --------------------------------------------------------------------------- training step restatement
def keras_train_step(graph: KerasGraph, params: Dict[str, torch.Tensor], x: np.ndarray, y: np.ndarray,
                     opt: dict, state: dict, dropout: Optional[DropoutSource] = None,
                     sample_weight_sum: Optional[float] = None):
    """One ``model.fit`` step (``train_96.py:175``, ``train_88.py:355``) in float64.

    loss = mean((pred-y)^2 over all elements) + sum_l l2*sum(w^2); metric mae.
    Optimizer updates follow Keras 2.13.  ``params`` must be leaf
    tensors with requires_grad; they are updated in place.  Returns (loss, mae, grads).
    """
    for p in params.values():
        if p.grad is not None:
            p.grad = None
    xt = torch.as_tensor(x, dtype=torch.float64)
    yt = torch.as_tensor(y, dtype=torch.float64)
    pred = graph(xt, training=True, dropout=dropout)
    mse = ((pred - yt) ** 2).mean()
    loss = mse + graph.l2_penalty()
    mae = (pred - yt).abs().mean()
    loss.backward()
    grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}
    apply_optimizer(params, grads, opt, state)
    return float(loss.detach()), float(mae.detach()), grads


def apply_optimizer(params, grads, opt: dict, state: dict):
    kind = opt["name"].lower()
    lr = float(opt["learning_rate"])
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    with torch.no_grad():
        for k, p in params.items():
            g = grads[k]
            if kind == "sgd":
                p -= lr * g
            elif kind == "adam":
                b1, b2, eps = opt.get("beta_1", 0.9), opt.get("beta_2", 0.999), opt.get("epsilon", 1e-7)
                m = state.setdefault("m", {}).setdefault(k, torch.zeros_like(p))
                v = state.setdefault("v", {}).setdefault(k, torch.zeros_like(p))
                m += (g - m) * (1 - b1)
                v += (g * g - v) * (1 - b2)
                alpha = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
                p -= alpha * m / (torch.sqrt(v) + eps)
            elif kind == "adamax":
                b1, b2, eps = opt.get("beta_1", 0.9), opt.get("beta_2", 0.999), opt.get("epsilon", 1e-7)
                m = state.setdefault("m", {}).setdefault(k, torch.zeros_like(p))
                u = state.setdefault("v", {}).setdefault(k, torch.zeros_like(p))
                m += (g - m) * (1 - b1)
                torch.maximum(b2 * u, g.abs(), out=u)
                p -= (lr / (1 - b1 ** t)) * m / (u + eps)
            else:
                raise NotImplementedError(kind)
