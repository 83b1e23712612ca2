"""Shared test helpers: fixtures loading and oracle construction (tests may import oracle/)."""
import json
import os

import numpy as np
import torch

from oracle.keras_graph import DropoutSource, KerasGraph, normalise_weight_names, to_torch  # noqa: F401

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unified_fixture():
    with open(os.path.join(GOLDEN, "unified_graph.json")) as f:
        graph = json.load(f)
    with np.load(os.path.join(GOLDEN, "unified_weights.npz")) as z:
        weights = {k: z[k].astype(np.float32) for k in z.files}
    return graph, weights


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def head_oracle(model, dtype=torch.float64, requires_grad=False):
    """KerasGraph over the model's own Keras-format config and current weights."""
    w = to_torch(model.get_weights_dict(), dtype, requires_grad)
    return KerasGraph(model._config, w), w


def synthetic_features(n, c, seed, sigma=1.1, p=0.29):
    """relu(N(0,sigma)) * Bernoulli(p): matches the shipped feature statistics."""
    rng = np.random.default_rng(seed)
    x = np.maximum(rng.normal(0, sigma, size=(n, c)), 0) * (rng.random((n, c)) < p)
    return x.astype(np.float32)


def synthetic_poses(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.normal([15, -6, -1], [27.5, 27, 12.5], size=(n, 3))).astype(np.float32)
