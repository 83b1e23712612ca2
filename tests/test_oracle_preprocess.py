"""CPU tests of the pre-processing oracle (oracle/preprocess.py): properties TensorFlow's ResizeBicubic(half_pixel_centers)
has by construction, hand-checkable cases, and the EMA filter of the facade (reference blazeFaceDetectorH5.py:16-35)."""
import numpy as np
import pytest

from oracle import preprocess as P


def keys(x, a=-0.5):
    x = abs(x)
    if x <= 1:
        return ((a + 2) * x - (a + 3)) * x * x + 1
    if x < 2:
        return ((a * x - 5 * a) * x + 8 * a) * x - 4 * a
    return 0.0


def test_table_is_the_keys_cubic_at_its_nodes():
    assert P.COEFFS.dtype == np.float32 and P.COEFFS.size == 2050
    for i in (0, 1, 17, 256, 512, 768, 1023, 1024):
        x = i / 1024
        assert P.COEFFS[2 * i] == np.float32(keys(x))
        assert P.COEFFS[2 * i + 1] == np.float32(keys(x + 1))
    assert P.COEFFS[0] == 1 and P.COEFFS[1] == 0 and P.COEFFS[2048] == 0 and P.COEFFS[2049] == 0


def test_two_to_one_weights_and_indices():
    idx, w = P.taps(256, 128)
    # interior: source coordinate 2 o + 0.5 -> taps 2o-1 .. 2o+2 with the delta = 0.5 weights (-1/16, 9/16, 9/16, -1/16)
    assert np.array_equal(idx[5], [9, 10, 11, 12])
    assert np.array_equal(w[5], np.float32([-0.0625, 0.5625, 0.5625, -0.0625]))
    # first output: the tap at -1 is outside -> weight 0, the others renormalised by 1 / (1 - (-1/16)) = 16/17
    assert np.array_equal(idx[0], [0, 0, 1, 2]) and w[0][0] == 0
    np.testing.assert_allclose(w[0][1:], np.float32([0.5625, 0.5625, -0.0625]) * np.float32(16 / 17), rtol=2e-7)
    assert np.array_equal(idx[127], [253, 254, 255, 255]) and w[127][3] == 0


@pytest.mark.parametrize("n_in,n_out", [(128, 128), (480, 128), (64, 128), (37, 96), (1, 8), (2, 5), (1000, 7)])
def test_weights_are_a_partition_of_unity(n_in, n_out):
    idx, w = P.taps(n_in, n_out)
    assert idx.min() >= 0 and idx.max() <= n_in - 1
    np.testing.assert_allclose(w.astype(np.float64).sum(1), 1.0, atol=3e-7)


def test_identity_when_sizes_agree():
    img = np.random.default_rng(0).integers(0, 256, (37, 53, 3), dtype=np.uint8)
    assert np.array_equal(P.resize_bicubic(img.astype(np.float64), 37, 53), img.astype(np.float32))
    x = P.prepare_input(img, 37, 53)
    want = ((img[..., ::-1].astype(np.float64) / 255.0).astype(np.float32) - np.float32(0.5)) / np.float32(0.5)
    assert x.shape == (1, 37, 53, 3) and x.dtype == np.float32 and np.array_equal(x[0], want)


def test_constant_and_linear_images_are_reproduced():
    c = np.full((20, 30, 3), 0.3)
    assert np.abs(P.resize_bicubic(c, 128, 96) - np.float32(0.3)).max() < 1e-7
    yy, xx = np.mgrid[0:64, 0:48]
    lin = (0.5 * xx + 0.25 * yy)[..., None].astype(np.float64)
    r = P.resize_bicubic(lin, 128, 96)
    oy, ox = (np.arange(128) + 0.5) * 0.5 - 0.5, (np.arange(96) + 0.5) * 0.5 - 0.5
    exact = 0.5 * ox[None, :] + 0.25 * oy[:, None]
    assert np.abs(r[..., 0] - exact)[4:-4, 4:-4].max() < 1e-5          # the Keys cubic reproduces linear functions


def test_downscale_has_no_antialias_and_overshoots():
    # a one-pixel-wide bright column between the sampled taps vanishes at 4:1 (no antialias), a step edge overshoots
    img = np.zeros((16, 64, 1))
    img[:, 4] = 1.0                                  # output 1 samples around x = 5.5: taps 4..7 -> weight -1/16 on column 4
    r = P.resize_bicubic(img, 16, 16)
    assert r[0, 1, 0] == np.float32(-0.0625)
    assert r[0, 3:, 0].max() == 0


def test_channel_order_and_range():
    img = np.zeros((128, 128, 3), np.uint8)
    img[..., 0] = 255                                # blue in BGR
    x = P.prepare_input(img, 128, 128)[0]
    assert np.all(x[..., 2] == 1.0) and np.all(x[..., 0] == -1.0) and np.all(x[..., 1] == -1.0)


def test_ema_filter_follows_the_reference_recurrence():
    from hpose_b200.blazeFaceDetectorH5 import EMAFilter
    f = EMAFilter(0.25, initial_value=7.0)
    assert f.state == 7.0 and not f.initialized
    assert f.update(10.0) == 10.0 and f.initialized            # the first measurement replaces the initial value
    assert f.update(20.0) == 0.25 * 20.0 + 0.75 * 10.0
    assert f.update(-4.0) == 0.25 * -4.0 + 0.75 * 12.5
    assert EMAFilter(1.0).update(3.0) == 3.0
    with pytest.raises(AssertionError):
        EMAFilter(0.0)
    with pytest.raises(AssertionError):
        EMAFilter(1.5)
