"""Bearing-vector LUT (CMaxSLAM::precomputeBearingVectors, src/cmax_slam.cpp:106-120): numpy restatement against
cv2 4.13 golden vectors (CPU), device kernel against both (GPU)."""
import numpy as np
import pytest

from oracle import lut_py


def _cams(g):
    for name in ("a", "b"):
        W, H = (int(v) for v in g[f"{name}_WH"])
        yield name, W, H, g[f"{name}_K"].reshape(3, 3), g[f"{name}_D"], g[f"{name}_R"].reshape(3, 3), g[f"{name}_P"].reshape(3, 4), g[f"{name}_rect"]


def test_restatement_matches_cv2_undistort_points(golden):
    g = golden("lut_cv2.npz")
    for name, W, H, K, D, R, P, rect in _cams(g):
        ys, xs = np.mgrid[0:H, 0:W]
        uv = np.stack([xs.ravel(), ys.ravel()], 1)
        mine = lut_py.undistort_points(uv, K, D, R, P)
        ulp = np.spacing(np.abs(rect).astype(np.float32))
        diff = np.abs(mine.astype(np.float64) - rect.astype(np.float64))
        assert (diff <= ulp).all(), (name, diff.max())                    # float32 outputs: at most one ulp apart
        assert (diff == 0).mean() > 0.99, (name, (diff == 0).mean())     # and almost all identical


def test_undistorted_camera_is_the_pinhole_lut():
    from cmax_slam_b200 import synth
    K4 = (200.0, 201.0, 120.0, 90.0)
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]])
    lut = lut_py.bearing_vectors(240, 180, K, np.zeros(5), np.eye(3), np.hstack([K, np.zeros((3, 1))]))
    assert np.array_equal(lut, synth.bearing_lut(240, 180, K4).reshape(-1, 3))


@pytest.mark.gpu
def test_device_lut_matches_cv2_and_restatement(golden):
    from cmax_slam_b200.stream import precompute_bearing_vectors
    from cmax_slam_b200 import synth
    g = golden("lut_cv2.npz")
    for name, W, H, K, D, R, P, rect in _cams(g):
        dev = precompute_bearing_vectors(W, H, K, D, R, P)
        ref = lut_py.bearing_vectors(W, H, K, D, R, P)
        # same doubles as the restatement wherever the float32 rectified pixel agrees (>= 99 % bit-identical), never
        # more than one float32 ulp of the pixel coordinate away
        same = (dev == ref).all(axis=1).mean()
        assert same > 0.99, (name, same)
        px = np.stack([dev[:, 0] * P[0, 0] + P[0, 2] + P[0, 3], dev[:, 1] * P[1, 1] + P[1, 2] + P[1, 3]], 1)
        assert (np.abs(px - rect.astype(np.float64)) <= 2 * np.spacing(np.abs(rect).astype(np.float32)) + 1e-9).all(), name
        assert (dev[:, 2] == 1.0).all()
    K4 = (200.0, 201.0, 120.0, 90.0)
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]])
    dev = precompute_bearing_vectors(240, 180, K, np.zeros(5))
    assert np.array_equal(dev, synth.bearing_lut(240, 180, K4).reshape(-1, 3))     # no distortion: exact pinhole rays
