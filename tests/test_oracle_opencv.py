"""The oracle's restatement of the OpenCV primitives the reference calls (GaussianBlur, meanStdDev,
mean, Mat::mul ...) against golden outputs of cv2 4.13 (tests/golden/blur_cv2.npz,
made by tests/golden/make_golden.py)."""
import numpy as np
import pytest


def test_gaussian_kernel_bit_exact(oracle, golden):
    g = golden("blur_cv2.npz")
    for sigma in (0.5, 1.0, 1.5, 2.0, 3.0):
        assert np.array_equal(oracle.gaussian_kernel(sigma), g[f"kernel_{sigma}"]), sigma
    assert len(oracle.gaussian_kernel(1.0)) == 9  # ksize = cvRound(8 sigma + 1) | 1


@pytest.mark.parametrize("name", ["a", "b"])
@pytest.mark.parametrize("sigma", [1.0, 2.0])
def test_gaussian_blur_bit_exact(oracle, golden, name, sigma):
    """ksize >= 7 and a row length that is a multiple of the SIMD width: the oracle reproduces the op
    order of OpenCV's AVX2/FMA row and column filters BIT-EXACTLY (cv2 4.13, x86-64)."""
    g = golden("blur_cv2.npz")
    out = oracle.gaussian_blur(g[f"img_{name}"], sigma)
    ref = g[f"blur_{name}_{sigma}"]
    assert np.array_equal(out, ref), float(np.abs(out - ref).max())


@pytest.mark.parametrize("name,sigma", [("a", 0.5), ("b", 0.5), ("c", 0.5), ("c", 1.0), ("c", 2.0), ("d", 1.0), ("d", 2.0)])
def test_gaussian_blur_ragged_or_small_kernel(oracle, golden, name, sigma):
    """Ragged widths (OpenCV's scalar tail columns) and the 5-tap small-kernel path use another
    association order inside OpenCV: agreement to a few f32 ulps of the image maximum."""
    g = golden("blur_cv2.npz")
    out = oracle.gaussian_blur(g[f"img_{name}"], sigma)
    ref = g[f"blur_{name}_{sigma}"]
    assert np.abs(out - ref).max() <= 4 * np.finfo(np.float32).eps * np.abs(ref).max()


def test_gaussian_blur_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    img = rng.random((50, 72)).astype(np.float32)
    assert np.array_equal(oracle.gaussian_blur(img, 1.0), cv2.GaussianBlur(img, (0, 0), 1.0))


def test_reflect101_border(oracle):
    # impulse at a corner: REFLECT_101 does not duplicate the border pixel
    img = np.zeros((12, 12), np.float32)
    img[1, 0] = 1.0
    k = oracle.gaussian_kernel(1.0).astype(np.float64)
    out = oracle.gaussian_blur(img, 1.0)
    # output row 0 reads input row 1 directly (tap +1) and through the reflection -1 -> 1 (tap -1)
    assert abs(out[0, 0] - (k[3] + k[5]) * k[4]) < 1e-7
    assert abs(out[1, 0] - (k[4] + k[2]) * k[4]) < 1e-7   # tap 0 and tap -2 (1-2 = -1 -> 1)
    assert abs(out[2, 0] - (k[3] + k[1]) * k[4]) < 1e-7   # tap -1 and tap -3 (2-3 = -1 -> 1)
    assert abs(out[6, 0]) < 1e-12 or abs(out[6, 0] - 0.0) < 1e-7


def test_mean_stddev(oracle, golden):
    g = golden("blur_cv2.npz")
    m, s = oracle.mean_stddev(g["blur_a_1.0"])
    assert abs(m - g["meanstd_a"][0]) <= 1e-15 * abs(m) + 1e-18
    assert abs(s - g["meanstd_a"][1]) <= 1e-13 * abs(s)


def test_contrast_variance_chain(oracle, golden):
    """contrast_Variance / contrast_MeanSquare (local_focus_funcs.cpp:9-44) through the oracle's FE entry point
    is exercised in test_oracle_fe.py; here the raw op chain on a cv2-produced image pair."""
    g = golden("blur_cv2.npz")
    img, deriv = g["blur_a_1.0"], g["var_deriv"]
    N = img.size
    mean = img.astype(np.float64).sum() / N
    zm = (img * np.float32(2.0) + np.float32(-2.0 * mean)).astype(np.float32)
    for c in range(3):
        ch = deriv[..., c]
        mc = ch.astype(np.float64).sum() / N
        d = (ch + np.float32(-mc)).astype(np.float32)
        gc = (zm * d).astype(np.float32).astype(np.float64).sum() / N
        assert abs(gc - g["var_grad"][c]) <= 1e-6 * abs(g["var_grad"][c]) + 1e-9
