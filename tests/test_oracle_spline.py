"""The oracle's restated Sophus / basalt spline math against the REAL reference code
(thirdparty/basalt-headers, compiled into oracle/_ref; golden copy in tests/golden/spline_ref.npz)."""
import ctypes as C

import numpy as np
import pytest


@pytest.mark.parametrize("order", [2, 4])
def test_spline_eval_vs_golden(oracle, golden, order):
    g = golden("spline_ref.npz")
    knots, t0, dt = g[f"knots_{order}"], int(g["t0"]), int(g["dt"])
    for i, t in enumerate(g[f"t_{order}"]):
        q, R, idx, J = oracle.spline_eval(order, knots, t0, dt, int(t))
        assert idx == g[f"idx_{order}"][i]
        assert np.abs(q - g[f"q_{order}"][i]).max() < 5e-15
        assert np.abs(R - g[f"R_{order}"][i]).max() < 5e-15
        assert np.abs(J - g[f"J_{order}"][i]).max() < 1e-14


def test_exp_log_jacobians_vs_golden(oracle, golden):
    g = golden("spline_ref.npz")
    L = oracle.lib()
    for i, w in enumerate(g["w"]):
        w = np.ascontiguousarray(w)
        q = np.zeros(4)
        L.orc_so3_exp(oracle._d(w), oracle._d(q))
        assert np.abs(q - g["exp_w"][i]).max() < 1e-15
        lg = np.zeros(3)
        L.orc_so3_log(oracle._d(np.ascontiguousarray(g["exp_w"][i])), oracle._d(lg))
        assert np.abs(lg - g["log_exp_w"][i]).max() < 1e-15


def test_spline_time_range(oracle):
    knots = np.tile(np.array([0.0, 0, 0, 1]), (6, 1))
    t0, dt = 1000, 100
    for order in (2, 4):
        last = t0 + (6 - order + 1) * dt - 1   # maxTimeNs(), so3_spline.h:120-122
        assert oracle.spline_eval(order, knots, t0, dt, last) is not None
        assert oracle.spline_eval(order, knots, t0, dt, last + 1) is None
        assert oracle.spline_eval(order, knots, t0, dt, t0 - 1) is None


@pytest.mark.parametrize("order", [2, 4])
def test_spline_live_reference(oracle, order):
    """Against the real basalt code when oracle/_ref was built (container with /root/reference, or
    the prebuilt .so that travels to the GPU box)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(order)
    from cmax_slam_b200 import synth
    K = 9
    knots = np.zeros((K, 4)); knots[0] = [0, 0, 0, 1]
    for k in range(1, K):
        knots[k] = synth._qmul(knots[k - 1], synth._qexp(rng.normal(0, 0.4, 3)))
        knots[k] /= np.linalg.norm(knots[k])
    t0, dt = 1_600_000_000_000_000_000, 100_000_000
    for t in rng.integers(t0, t0 + (K - order + 1) * dt - 1, 300):
        a = oracle.spline_eval(order, knots, t0, dt, int(t))
        b = oracle.ref_spline_eval(order, knots, t0, dt, int(t))
        assert a[2] == b[2]
        assert np.abs(a[0] - b[0]).max() < 5e-15 and np.abs(a[1] - b[1]).max() < 5e-15
        assert np.abs(a[3] - b[3]).max() < 1e-14


def test_spline_jacobian_numeric(oracle):
    """Same check as basalt's own SplineTest.SO3CUBSplineEvaluateKnots (test_spline.cpp:95-132): analytic
    d_val_d_knot vs central differences of log(R(x) R^-1) under LEFT knot perturbations exp(x)*knot."""
    from cmax_slam_b200 import synth
    rng = np.random.default_rng(3)
    L = oracle.lib()
    for order in (2, 4):
        K = 7
        knots = np.zeros((K, 4)); knots[0] = [0, 0, 0, 1]
        for k in range(1, K):
            knots[k] = synth._qmul(knots[k - 1], synth._qexp(rng.normal(0, 0.3, 3)))
            knots[k] /= np.linalg.norm(knots[k])
        t0, dt = 0, 1_000_000
        t = int(2.37 * dt)
        q0, R0, idx, J = oracle.spline_eval(order, knots, t0, dt, t)
        eps = 1e-6
        for k in range(order):
            num = np.zeros((3, 3))
            for a in range(3):
                res = []
                for sgn in (+1, -1):
                    kn = knots.copy()
                    e = np.zeros(3); e[a] = sgn * eps
                    kn[idx + k] = synth._qmul(synth._qexp(e), kn[idx + k])
                    q1 = oracle.spline_eval(order, kn, t0, dt, t)[0]
                    d = synth._qmul(q1, synth._qconj(q0))
                    res.append(synth._qlog(d))
                num[:, a] = (res[0] - res[1]) / (2 * eps)
            assert np.abs(num - J[k]).max() < 1e-6


def test_ros_batch_mid_time(oracle):
    """time_first + (time_last - time_first) * 0.5 with ros::Duration rounding (fromSec: floor + round)."""
    L = oracle.lib()
    cases = [((1600000000, 999999000), (1600000001, 1000), (1600000001, 0)),
             ((10, 0), (10, 1), (10, 1)),            # 0.5 ns rounds half away from zero -> 1
             ((10, 0), (10, 3), (10, 2)),            # 1.5 ns -> 2
             ((10, 5), (10, 5), (10, 5)),
             ((10, 999999999), (12, 1), (11, 500000000))]
    for (s0, n0), (s1, n1), (es, en) in cases:
        s, n = C.c_uint32(), C.c_uint32()
        L.orc_batch_mid_time(s0, n0, s1, n1, C.byref(s), C.byref(n))
        assert (s.value, n.value) == (es, en), ((s0, n0), (s1, n1), (s.value, n.value))
