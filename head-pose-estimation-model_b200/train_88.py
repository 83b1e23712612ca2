"""Model-88 training entry points (reference: Model-88/train_88.py).

``config`` (:45-64), the head builders ``create_model`` (:66-158), ``create_model_skip_fc`` (:163-223),
``bestmodelV1`` (:226-253) and ``train`` (:256-397) keep their names and meaning; wandb logging and the
matplotlib histogram are experiment bookkeeping and out of scope.
"""
import os

import numpy as np

from . import keras_spec as K
from .attention_model import create_model_complex, create_modelC, se_transformer_regr_head  # noqa: F401
from .utilities import load_dataset, train_val_split

config = {
    "learning_rate": 0.00028,
    "batch_size": 128,
    "total_epochs": 1000000,
    "early_stopping_patience": 40,
    "early_stopping_min_delta": 0.001,
    "optimizer": "sgd",
    "loss_function": "mse",
    "performance_metrics": ["mae"],
    "save_best_only": True,
    "monitor_metric": "val_loss",
    "dropout_rate": 0.0001,
    "filtersnum": 64,
    "regularizer_rate": 1e-6,
}


def _conv(x, width, activation, l2):
    return K.Conv2D(filters=width, kernel_size=1, padding="same", activation=activation, kernel_regularizer=l2,
                    kernel_initializer=K.initializers.GlorotUniform())(x)


def create_model():
    """88 -> filtersnum (softsign) -> 3 (linear), SpatialDropout2D after both convs."""
    l2 = K.regularizers.l2(config["regularizer_rate"])
    feat = K.Input(shape=(None, None, 88))
    h = K.SpatialDropout2D(config["dropout_rate"])(_conv(feat, config["filtersnum"], "softsign", l2))
    out = K.SpatialDropout2D(config["dropout_rate"])(_conv(h, 3, "linear", l2))
    return K.Model(inputs=feat, outputs=out)


bestmodelV1 = create_model  # identical architecture in the reference (train_88.py:226-253)


def create_model_skip_fc():
    """88 -> 32 -> 64 -> 32 (+skip from the first 32) -> 3, softsign, dropout after each stage."""
    l2 = K.regularizers.l2(config["regularizer_rate"])
    dr = config["dropout_rate"]
    feat = K.Input(shape=(None, None, 88))
    a = K.SpatialDropout2D(dr)(_conv(feat, 32, "softsign", l2))
    b = K.SpatialDropout2D(dr)(_conv(a, 64, "softsign", l2))
    c = K.SpatialDropout2D(dr)(K.Add()([_conv(b, 32, "softsign", l2), a]))
    return K.Model(inputs=feat, outputs=_conv(c, 3, "linear", l2), name="FC_Skip_Regressor")


def _optimizer():
    if config["optimizer"] == "sgd":
        return K.SGD(learning_rate=config["learning_rate"])
    return K.Adam(learning_rate=config["learning_rate"])


def train(model=None, features_dir=None, out_dir="Trained-Models-88", run_id="run", max_epochs=None, verbose=1,
          distributed=None):
    """Load the BIWI 88-channel feature sets, split 80/20 (seed 42), fit with early stopping and
    best-checkpointing, then evaluate on BIWI test and AFLW2000.  Returns (model, history, summary)."""
    d = features_dir or os.getenv("FEATUREMAPS_DIR_PATH", "")
    parts = [load_dataset(os.path.join(d, "BIWI_Train_Enlarged_features_88_0.7_1.npz"))]
    extra = os.path.join(d, "BIWI_NoTrack_Enlarged_features_88_0.7_1.npz")
    if os.path.exists(extra):   # listed in the reference's .MISSING_LARGE_BLOBS; used when present
        parts.append(load_dataset(extra))
    feats = np.concatenate([p[0] for p in parts], axis=0).reshape(-1, 1, 1, 88)
    poses = np.concatenate([p[1] for p in parts], axis=0).reshape(-1, 1, 1, 3)
    tr_x, va_x, tr_y, va_y = train_val_split(feats, poses, 0.2, 42)
    model = model or create_model_complex(config["regularizer_rate"], config["dropout_rate"])
    model.compile(optimizer=_optimizer(), loss=config["loss_function"], metrics=config["performance_metrics"])
    os.makedirs(out_dir, exist_ok=True)
    cbs = [K.ModelCheckpoint(os.path.join(out_dir, f"{run_id}.h5"), monitor=config["monitor_metric"],
                             save_best_only=config["save_best_only"]),
           K.EarlyStopping(monitor=config["monitor_metric"], patience=config["early_stopping_patience"],
                           min_delta=config["early_stopping_min_delta"], restore_best_weights=True)]
    hist = model.fit(tr_x, tr_y, epochs=max_epochs or config["total_epochs"], batch_size=config["batch_size"],
                     validation_data=(va_x, va_y), callbacks=cbs, verbose=verbose, distributed=distributed)
    summary = {"total_parameters": model.count_params()}
    for tag, fname in (("test", "BIWI_Test_Enlarged_features_88_0.7_1.npz"),
                       ("test_AFLW2000", "AFLW2000_Enlarged_features_88_0.7_1.npz")):
        path = os.path.join(d, fname)
        if os.path.exists(path):
            fx, fy = load_dataset(path)
            loss, mae = model.evaluate(fx.reshape(-1, 1, 1, 88), fy.reshape(-1, 1, 1, 3), verbose=0)
            summary[f"{tag}_loss"], summary[f"{tag}_mae"] = loss, mae
    best = int(np.argmin(hist.history["val_loss"]))
    summary.update(best_epoch=best + 1, best_epoch_val_loss=hist.history["val_loss"][best],
                   best_epoch_val_mae=hist.history["val_mae"][best])
    return model, hist, summary
