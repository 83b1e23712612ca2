"""Dataset helpers with the reference's npz contract (Model-96/utilities.py:31-34,43-77;
Model-88/utilities.py:35-38): ``features`` (N,C) float32 post-ReLU tap vectors, ``poses`` (N,3)
float64 degrees [yaw, pitch, roll]."""
import os

import numpy as np


def load_dataset(dataset_path):
    if not os.path.exists(dataset_path):
        raise FileNotFoundError(dataset_path)
    with np.load(dataset_path) as z:
        return z["features"], z["poses"]


def save_dataset(dataset_path, features, poses):
    """Writes the same uncompressed layout the reference's extractor produces."""
    np.savez(dataset_path, features=np.asarray(features, np.float32), poses=np.asarray(poses, np.float64))


def load_dataset_with_weights(npz_path):
    """Adds per-sample weights w = 1 if delta <= 60 deg else 0.5**((delta-60)/5),
    delta = arccos(cos(pitch) cos(yaw)) (utilities.py:43-77)."""
    features, poses = load_dataset(npz_path)
    yaw, pitch = np.deg2rad(poses[:, 0]), np.deg2rad(poses[:, 1])
    delta = np.rad2deg(np.arccos(np.clip(np.cos(pitch) * np.cos(yaw), -1.0, 1.0)))
    weights = np.where(delta > 60.0, 0.5 ** ((delta - 60.0) / 5.0), 1.0)
    return {"features": features, "poses": poses, "weights": weights}


def train_val_split(features, poses, test_size=0.2, random_state=42):
    """sklearn.model_selection.train_test_split(test_size, random_state) as used at train_96.py:142-146."""
    from sklearn.model_selection import train_test_split
    return train_test_split(features, poses, test_size=test_size, random_state=random_state)
