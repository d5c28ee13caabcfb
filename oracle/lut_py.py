"""TEST INFRASTRUCTURE.  numpy restatement of CMaxSLAM::precomputeBearingVectors (src/cmax_slam.cpp:106-120) =
image_geometry::PinholeCameraModel::rectifyPoint + projectPixelTo3dRay over cv::undistortPoints (both un-vendored:
ROS image_geometry, OpenCV).  Pinned against cv2 4.13 in tests/golden/lut_cv2.npz (tests/test_lut.py)."""
import numpy as np


def undistort_points(uv, K, D, R, P, iters=5):
    """cvUndistortPointsInternal for CV_32FC2 input (float in, double arithmetic, float out)."""
    uv = np.asarray(uv, dtype=np.float32).astype(np.float64)
    k = np.zeros(12)
    k[:len(D)] = D
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    ifx, ify = 1.0 / fx, 1.0 / fy
    x = (uv[:, 0] - cx) * ifx
    y = (uv[:, 1] - cy) * ify
    x0, y0 = x.copy(), y.copy()
    dead = np.zeros(len(x), bool)
    for _ in range(iters):
        r2 = x * x + y * y
        icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
        neg = (icdist < 0) & ~dead
        dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2
        dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2
        xn = (x0 - dx) * icdist
        yn = (y0 - dy) * icdist
        upd = ~dead & ~neg
        x = np.where(upd, xn, x); y = np.where(upd, yn, y)
        x = np.where(neg, x0, x); y = np.where(neg, y0, y)
        dead |= neg
    RR = P[:3, :3] @ R
    xx = RR[0, 0] * x + RR[0, 1] * y + RR[0, 2]
    yy = RR[1, 0] * x + RR[1, 1] * y + RR[1, 2]
    ww = 1.0 / (RR[2, 0] * x + RR[2, 1] * y + RR[2, 2])
    return np.stack([(xx * ww).astype(np.float32), (yy * ww).astype(np.float32)], 1)


def bearing_vectors(W, H, K, D, R, P):
    K, R, P = np.asarray(K, float).reshape(3, 3), np.asarray(R, float).reshape(3, 3), np.asarray(P, float).reshape(3, 4)
    D = np.asarray(D, float).ravel()
    ys, xs = np.mgrid[0:H, 0:W]
    uv = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float64)
    if np.any(D != 0):
        rect = undistort_points(uv, K, D, R, P).astype(np.float64)
    else:
        rect = uv
    out = np.ones((W * H, 3))
    out[:, 0] = (rect[:, 0] - P[0, 2] - P[0, 3]) / P[0, 0]
    out[:, 1] = (rect[:, 1] - P[1, 2] - P[1, 3]) / P[1, 1]
    return out
