"""Global-map upkeep: setUpdateTimesIG / updateIG / updateIGp.
CPU: properties of the oracle restatement.  GPU: device-resident map against the oracle (bytes and floats exact)."""
import numpy as np
import pytest

from cmax_slam_b200 import synth

K_T = (60.0, 61.0, 31.5, 23.5)


def test_fov_mask_properties(oracle):
    lut = synth.bearing_lut(64, 48, K_T)
    times = np.zeros((64, 128), np.uint8)
    oracle.set_update_times(lut, 64, 48, 128, 64, [0, 0, 0, 1.0], 3, times)
    assert set(np.unique(times)) <= {0, 1} and times.sum() > 0
    ys, xs = np.nonzero(times)
    # identity pose looks along +z: the footprint is centred on the panorama centre
    assert abs(xs.mean() - 64) < 3 and abs(ys.mean() - 32) < 3
    t2 = times.copy()
    oracle.set_update_times(lut, 64, 48, 128, 64, [0, 0, 0, 1.0], 3, t2)
    assert np.array_equal(t2, 2 * times)                      # visit counts accumulate
    sat = np.full((64, 128), 255, np.uint8)
    oracle.set_update_times(lut, 64, 48, 128, 64, [0, 0, 0, 1.0], 3, sat)
    assert sat.max() == 255                                     # CV_8U saturation
    r0 = np.zeros((64, 128), np.uint8)
    oracle.set_update_times(lut, 64, 48, 128, 64, [0, 0, 0, 1.0], 0, r0)
    assert 0 < r0.sum() < times.sum()                           # radius dilates


def test_update_ig_is_gated_by_visit_count(oracle):
    IG = np.ones((4, 4), np.float32)
    il = np.full((4, 4), 2.0, np.float32)
    times = np.array([[0, 10, 11, 255]] * 4, np.uint8)
    oracle.update_ig(IG, il, times, 10)
    assert np.array_equal(IG[:, 0], [3] * 4) and np.array_equal(IG[:, 1], [3] * 4)
    assert np.array_equal(IG[:, 2], [1] * 4) and np.array_equal(IG[:, 3], [1] * 4)


@pytest.mark.gpu
def test_device_map_matches_oracle(oracle):
    from cmax_slam_b200.backend import EventWarperCMax
    w = synth.make_be_window(20000, 8, 128, 64, 17, order=2, sensor=(64, 48), K4=K_T, n_landmarks=300, n_fixed=1)
    be = EventWarperCMax(64, 48, w.lut, 128, 64, spline_order=2)
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, None, 0.0)
    be.resetIG()
    rng = np.random.default_rng(1)
    rots = np.concatenate([w.knots_xyzw[:3], [[0, 0, 0, 1.0]]])
    IG_ref = np.zeros((64, 128), np.float32)
    times_ref = np.zeros((64, 128), np.uint8)
    for rnd in range(3):                                   # three "windows": mark FOV, then update IG at some x
        be.setUpdateTimesIG(rots, radius=3)
        for q in rots:
            oracle.set_update_times(w.lut, 64, 48, 128, 64, q, 3, times_ref)
        x = rng.normal(0, 0.01, 21)
        be.updateIG(x, max_update_times=8)
        ilo, _ = be.local_iwe(x)
        oracle.update_ig(IG_ref, ilo, times_ref, 8)
        IG, times = be.getIG()
        assert np.array_equal(times, times_ref)             # byte-exact visit counts
        assert np.abs(IG - IG_ref).max() <= 4e-6 * max(1.0, float(IG_ref.max()))   # IL_old re-scattered: f32 atomic order
    assert times_ref.max() > 8                              # the gate was exercised
    # IGp <- IG on the device, alpha from updateAlpha; evaluation must equal one with the host copy uploaded
    be.updateIGp(alpha=float("nan"))
    c_dev, g_dev = be.eval(None, True)
    a_dev = be.alpha
    IG, _ = be.getIG()
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IG, float("nan"))
    c_up, g_up = be.eval(None, True)
    assert abs(a_dev - be.alpha) <= 1e-6 * abs(be.alpha)
    assert abs(c_dev - c_up) <= 1e-6 * abs(c_up) and np.abs(g_dev - g_up).max() <= 1e-5 * np.abs(g_up).max()
    # setMap round trip
    be.setMap(IG_ref, times_ref)
    IG2, t2 = be.getIG()
    assert np.array_equal(IG2, IG_ref) and np.array_equal(t2, times_ref)
    be.close()
