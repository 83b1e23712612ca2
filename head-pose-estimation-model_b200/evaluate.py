"""``evaluate_head_pose_model`` with the contract of the reference's Model-96/test.py:9-69."""
import numpy as np

from .keras_spec import load_model

ANGLES = ("yaw", "pitch", "roll")


def evaluate_head_pose_model(model_path, dataset_path, verbose=True):
    """Load a head checkpoint and an npz dataset, predict on the GPU and report per-angle MAE / MSE.

    Returns ``{'MAE': {yaw, pitch, roll, average}, 'MSE': {...}}`` exactly like the reference."""
    model = load_model(model_path)
    with np.load(dataset_path) as data:
        features, truth = data["features"], data["poses"]
    n, c = features.shape
    if c != model.program.in_channels:
        raise ValueError(f"dataset has {c} channels, model expects {model.program.in_channels}")
    pred = model.predict(features.reshape(n, 1, 1, c), verbose=0).reshape(n, 3)
    err = pred.astype(np.float64) - truth
    mae, mse = np.abs(err).mean(axis=0), np.square(err).mean(axis=0)
    metrics = {"MAE": {a: float(mae[i]) for i, a in enumerate(ANGLES)},
               "MSE": {a: float(mse[i]) for i, a in enumerate(ANGLES)}}
    metrics["MAE"]["average"] = float(mae.mean())
    metrics["MSE"]["average"] = float(mse.mean())
    if verbose:
        for kind in ("MAE", "MSE"):
            print(kind + ": " + ", ".join(f"{k} {v:.4f}" for k, v in metrics[kind].items()))
    return metrics
