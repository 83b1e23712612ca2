"""ORACLE (test infrastructure, NOT product code) -- ``prepareInputForInference`` restated in numpy.

Follows ``BlazePoser/blazeFaceDetectorH5.py:247-269``:

    img = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)            # :249   channel reversal
    img = img / 255.0                                       # :254   float64
    img_resized = tf.image.resize(img, [H, W], method='bicubic', preserve_aspect_ratio=False)   # :255
    img_input = (img_resized.numpy() - 0.5) / 0.5           # :262   float32 (weak Python scalars)
    reshape_img = img_input.reshape(1, H, W, C)             # :265

``tf.image.resize(method='bicubic')`` is the un-vendored dependency ``tensorflow>=2.8.0`` (requirements.txt:2; the
``.h5`` files record Keras 2.13.1).  With ``antialias=False`` (the default) it is the ``ResizeBicubic`` CPU kernel with
``half_pixel_centers=True``, whose published algorithm (tensorflow/core/kernels/image/resize_bicubic_op.cc) is restated
here:

* ``scale = float(in) / float(out)``; source coordinate of output index o: ``(float(o) + 0.5f) * scale - 0.5f``;
  ``in_loc = floor``, ``delta = coordinate - in_loc`` (all float32);
* the four tap weights come from a TABLE of 1025 entries (``kTableSize = 1 << 10``) of the Keys cubic with
  ``A = -0.5``, indexed by ``offset = lrintf(delta * 1024)`` (round half to even) -- not from the exact polynomial
  at ``delta``; table entries are ``((A + 2) x - (A + 3)) x x + 1`` for the inner taps and ``((A x - 5 A) x + 8 A) x - 4 A``
  at ``x + 1`` for the outer ones, evaluated in double and stored as float;
* taps outside the image get weight 0 (their index is clamped) and the remaining weights are renormalised to sum 1;
* every output pixel: the 4 rows are combined first (``v0 w0 + v1 w1 + v2 w2 + v3 w3``, float32, left to right, input
  cast to float32), then the 4 resulting columns the same way; the output is float32 whatever the input type.

PARITY STATUS: unpinned by the reference (it has no tests and TensorFlow cannot run here).  Pinned by properties that
hold for the TensorFlow kernel by construction and are hand-checkable (``tests/test_oracle_preprocess.py``): identity
when the sizes agree, exact reproduction of constant and (in the interior) linear images, the weight table against
the closed-form Keys kernel at the table nodes, the exact 2:1 downscale weights (-1/16, 9/16, 9/16, -1/16), partition of
unity at the borders.
"""
from __future__ import annotations

import numpy as np

TABLE_SIZE = 1 << 10


def _coeff_table() -> np.ndarray:
    """float32 [(TABLE_SIZE + 1) * 2]: entry 2 i = inner-tap weight at x = i / 1024, 2 i + 1 = outer-tap weight at x + 1.
    The kernel's ``InitCoeffsTable(const double a)`` evaluates the polynomials in double (x is a float that holds i / 1024
    exactly) and rounds on the store into the float table."""
    a = -0.5
    t = np.empty((TABLE_SIZE + 1) * 2, np.float32)
    for i in range(TABLE_SIZE + 1):
        x = float(np.float32(i * 1.0 / TABLE_SIZE))
        t[2 * i] = np.float32(((a + 2) * x - (a + 3)) * x * x + 1)
        x += 1.0
        t[2 * i + 1] = np.float32(((a * x - 5 * a) * x + 8 * a) * x - 4 * a)
    return t


COEFFS = _coeff_table()


def taps(in_size: int, out_size: int):
    """Per output index: 4 clamped source indices (int64 [out, 4]) and 4 float32 weights ([out, 4])."""
    f = np.float32
    scale = f(f(in_size) / f(out_size))
    idx = np.empty((out_size, 4), np.int64)
    wgt = np.empty((out_size, 4), np.float32)
    for o in range(out_size):
        loc_f = f(f(f(o) + f(0.5)) * scale) - f(0.5)
        loc_f = f(loc_f)
        loc = int(np.floor(loc_f))
        delta = f(loc_f - f(loc))
        off = int(np.rint(f(delta * f(TABLE_SIZE))))     # lrintf: round half to even, like np.rint
        cand = (loc - 1, loc, loc + 1, loc + 2)
        raw = (COEFFS[off * 2 + 1], COEFFS[off * 2], COEFFS[(TABLE_SIZE - off) * 2], COEFFS[(TABLE_SIZE - off) * 2 + 1])
        w = []
        for k in range(4):
            b = min(max(cand[k], 0), in_size - 1)
            idx[o, k] = b
            w.append(raw[k] if b == cand[k] else f(0.0))
        s = f(f(f(w[0] + w[1]) + w[2]) + w[3])
        if abs(s) >= f(1000.0) * np.finfo(np.float32).tiny:
            inv = f(f(1.0) / s)
            w = [f(v * inv) for v in w]
        wgt[o] = w
    return idx, wgt


def resize_bicubic(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img: (H, W, C) any real dtype -> (out_h, out_w, C) float32, TensorFlow ResizeBicubic(half_pixel_centers=True)."""
    v = np.asarray(img).astype(np.float32)               # static_cast<float>(input)
    yi, yw = taps(v.shape[0], out_h)
    xi, xw = taps(v.shape[1], out_w)
    # rows first: [out_h, W, C]
    rows = v[yi[:, 0]] * yw[:, 0, None, None]
    for k in range(1, 4):
        rows = rows + v[yi[:, k]] * yw[:, k, None, None]  # float32 products and sums, left to right
    out = rows[:, xi[:, 0]] * xw[None, :, 0, None]
    for k in range(1, 4):
        out = out + rows[:, xi[:, k]] * xw[None, :, k, None]
    return out.astype(np.float32)


def prepare_input(image_bgr_u8: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """blazeFaceDetectorH5.py:247-269 for one HxWx3 uint8 BGR frame -> (1, out_h, out_w, 3) float32 in [-1, 1]."""
    img = np.asarray(image_bgr_u8)[..., ::-1]
    img = img / 255.0                                     # float64
    res = resize_bicubic(img, out_h, out_w)               # float32
    x = (res - np.float32(0.5)) / np.float32(0.5)
    return x.reshape(1, out_h, out_w, 3).astype(np.float32)
