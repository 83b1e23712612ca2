"""Model-96 training entry points (reference: Model-96/train_96.py).

``config`` (:42-59), ``create_model`` (:65-110) and ``train`` (:113-209) keep their names; the three
CLI overrides of ``__main__`` (:215-238) are kept in ``main``.  wandb is out of scope.
"""
import argparse
import os

import numpy as np

from . import keras_spec as K
from .utilities import load_dataset, train_val_split

RANDOM_SEED = 42

config = {
    "learning_rate": 0.00028,
    "batch_size": 128,
    "total_epochs": 10000,
    "early_stopping_patience": 40,
    "early_stopping_min_delta": 0.001,
    "optimizer": "adam",
    "loss_function": "mse",
    "performance_metrics": ["mae"],
    "save_best_only": True,
    "monitor_metric": "val_loss",
    "dropout_rate": -1,
    "regularizer_rate": -1,
    "num_filters": -1,
}


def create_model():
    """96 -> num_filters (tanh) -> dropout -> 3 -> dropout; L2 on kernels AND biases; compiled."""
    for key in ("dropout_rate", "regularizer_rate", "num_filters"):
        if config[key] < 0:
            raise ValueError(f"config['{key}'] must be set (the reference's default -1 is a sentinel)")
    l2 = K.regularizers.l2(config["regularizer_rate"])
    feat = K.Input(shape=(None, None, 96))
    hidden = K.Conv2D(filters=config["num_filters"], kernel_size=1, padding="same", activation="tanh",
                      kernel_initializer=K.initializers.GlorotUniform(), bias_regularizer=l2,
                      kernel_regularizer=l2)(feat)
    hidden = K.SpatialDropout2D(config["dropout_rate"])(hidden)
    angles = K.Conv2D(filters=3, kernel_size=1, padding="same", activation=None,
                      kernel_initializer=K.initializers.GlorotUniform(), bias_regularizer=l2,
                      kernel_regularizer=l2)(hidden)
    angles = K.SpatialDropout2D(config["dropout_rate"])(angles)
    model = K.Model(inputs=feat, outputs=angles)
    opt = {"adamax": K.Adamax, "sgd": K.SGD}.get(config["optimizer"], K.Adam)(learning_rate=config["learning_rate"])
    model.compile(optimizer=opt, loss=config["loss_function"], metrics=config["performance_metrics"])
    return model


def train(features_dir=None, out_dir=None, run_id="run", max_epochs=None, verbose=1, distributed=None,
          train_file="BIWI_train_features_96.npz", test_file="BIWI_test_features_96.npz",
          aflw_file="AFLW2000_features_96_0.7_1.npz"):
    K.set_seed(RANDOM_SEED)
    d = features_dir or os.getenv("FEATUREMAPS_DIR_PATH", "")
    feats, poses = load_dataset(os.path.join(d, train_file))
    feats, poses = feats.reshape(-1, 1, 1, 96), poses.reshape(-1, 1, 1, 3)
    tr_x, va_x, tr_y, va_y = train_val_split(feats, poses, 0.2, 42)
    out_dir = out_dir or os.getenv("TRAINED_MODELS_96_RESHAPEDINPUT_NOFLATTEN_PATH", ".")
    os.makedirs(out_dir, exist_ok=True)
    cbs = [K.ModelCheckpoint(os.path.join(out_dir, f"{run_id}.h5"), monitor=config["monitor_metric"],
                             save_best_only=config["save_best_only"]),
           K.EarlyStopping(monitor=config["monitor_metric"], patience=config["early_stopping_patience"],
                           min_delta=config["early_stopping_min_delta"], restore_best_weights=True)]
    model = create_model()
    hist = model.fit(tr_x, tr_y, epochs=max_epochs or config["total_epochs"], batch_size=config["batch_size"],
                     validation_data=(va_x, va_y), callbacks=cbs, verbose=verbose, distributed=distributed)
    summary = {"total_parameters": model.count_params(), "model_architecture": model.to_json()}
    for tag, fname in (("test", test_file), ("test_AFLW2000", aflw_file)):
        path = os.path.join(d, fname)
        if os.path.exists(path):
            fx, fy = load_dataset(path)
            loss, mae = model.evaluate(fx.reshape(-1, 1, 1, 96), fy.reshape(-1, 1, 1, 3), verbose=0)
            summary[f"{tag}_loss"], summary[f"{tag}_mae"] = loss, mae
    best = int(np.argmin(hist.history["val_loss"]))
    summary.update(best_epoch=best + 1, best_epoch_train_loss=hist.history["loss"][best],
                   best_epoch_train_mae=hist.history["mae"][best], best_epoch_val_loss=hist.history["val_loss"][best],
                   best_epoch_val_mae=hist.history["val_mae"][best])
    return model, hist, summary


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--dropout_rate", type=float, default=config["dropout_rate"])
    p.add_argument("--regularizer_rate", type=float, default=config["regularizer_rate"])
    p.add_argument("--num_filters", type=int, default=config["num_filters"])
    config.update(vars(p.parse_args(argv)))
    return train()


if __name__ == "__main__":
    main()
