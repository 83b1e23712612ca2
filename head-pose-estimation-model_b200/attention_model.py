"""Regressor-head builders with the signatures of the reference's Model-88/attention_model.py.

``se_transformer_regr_head`` (attention_model.py:16-80), ``create_modelC`` (:82-95) and
``create_model_complex`` (:97-169) return ``keras_spec.Model`` objects: graph specs whose forward,
backward and optimizer steps run in CUDA through libhpose.
"""
from . import keras_spec as K


def _se_gate(x, channels, squeezed):
    """Squeeze-and-excitation channel gate: GAP -> Dense(relu) -> Dense(sigmoid) -> broadcast multiply."""
    pooled = K.GlobalAveragePooling2D()(x)
    gate = K.Dense(channels, activation="sigmoid")(K.Dense(squeezed, activation="relu")(pooled))
    return K.Multiply()([x, K.Reshape((1, 1, channels))(gate)])


def se_transformer_regr_head(input_channels=88, reduction=16, num_heads=4, key_dim=16, ff_dim=64,
                             hidden_channels=128):
    """(B,H,W,C) -> (B,H,W,3): SE gate, one post-norm transformer encoder block over the H*W tokens,
    then Conv1x1(hidden, relu) -> Conv1x1(3)."""
    feat = K.Input(shape=(None, None, input_channels))
    gated = _se_gate(feat, input_channels, input_channels // reduction)
    tokens = K.Lambda()(gated)
    attended = K.MultiHeadAttention(num_heads=num_heads, key_dim=key_dim)(tokens, tokens)
    y = K.LayerNormalization()(K.Add()([tokens, attended]))
    ff = K.Dense(input_channels)(K.Dense(ff_dim, activation="relu")(y))
    y = K.LayerNormalization()(K.Add()([y, ff]))
    spatial = K.Lambda()([y, feat])
    hidden = K.Conv2D(hidden_channels, kernel_size=1, activation="relu")(spatial)
    angles = K.Conv2D(3, kernel_size=1, activation=None)(hidden)
    return K.Model(inputs=feat, outputs=angles, name="SE_Transformer_Regr")


def create_modelC():
    """SE gate (88 -> 11 -> 88) followed by Conv1x1(42, relu) -> Conv1x1(3)."""
    feat = K.Input((None, None, 88))
    gated = _se_gate(feat, 88, 11)
    hidden = K.Conv2D(42, 1, activation="relu")(gated)
    return K.Model(feat, K.Conv2D(3, 1, activation=None)(hidden))


def create_model_complex(reg, dr):
    """Residual 1x1-conv regressor: 88 -> 16, three (conv, conv, add, relu) blocks of width 16,
    16 -> 8 -> 3; softsign activations, SpatialDropout2D(dr) after every hidden conv, L2(reg) on kernels."""
    l2 = K.regularizers.l2(reg)

    def conv(x, width, activation="softsign", drop=True):
        y = K.Conv2D(width, kernel_size=1, padding="same", activation=activation, kernel_regularizer=l2)(x)
        return K.SpatialDropout2D(dr)(y) if drop else y

    feat = K.Input(shape=(None, None, 88))
    x = conv(feat, 16)
    for _ in range(3):
        x = K.Activation("relu")(K.Add()([x, conv(conv(x, 16), 16)]))
    x = conv(x, 8)
    return K.Model(inputs=feat, outputs=conv(x, 3, activation=None, drop=False), name="Complex_Conv_Skip_Model")
