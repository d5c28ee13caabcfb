"""GPU parity of the back-end CUDA path (device So3Spline, equirectangular warp, IL_old/IL_new scatter,
blur, variance, adjoint / dense gradient) against the CPU oracle -- through the C ABI."""
import numpy as np
import pytest

from cmax_slam_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-5
K_T = (60.0, 61.0, 31.5, 23.5)


def _window(n, order, seed=9, K=8, pano=(128, 64), sensor=(64, 48), K4=K_T, nl=300):
    return synth.make_be_window(n, K, pano[0], pano[1], seed, order=order, sensor=sensor, K4=K4, n_landmarks=nl,
                                n_fixed=1 if order == 2 else 3)


def _pair(oracle, w, IGp=None, alpha=0.5, grad_mode=1, sample_rate=1, batch_size=100, sigma=1.0, measure=0):
    from cmax_slam_b200.backend import EventWarperCMax
    be = EventWarperCMax(w.sensor_width, w.sensor_height, w.lut, w.pano_width, w.pano_height, blur_sigma=sigma,
                         event_batch_size=batch_size, event_sample_rate=sample_rate, spline_order=w.spline_order,
                         contrast_measure=measure, grad_mode=grad_mode)
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, alpha)
    a = oracle.be_args(w.events, w.lut, w.sensor_width, w.sensor_height, w.pano_width, w.pano_height, w.knots_xyzw,
                       w.t0_ns, w.dt_ns, w.spline_order, w.n_fixed, w.tnext, IGp, alpha, batch_size=batch_size,
                       sample_rate=sample_rate, blur_sigma=sigma, measure=measure)
    return be, a


def _check(be, oracle, a, x, images=True):
    ro = oracle.be_eval(a, x, True, images=images, cells=True)
    assert np.array_equal(be.warped_cells(x), ro["cells"])                        # bit-exact integer work
    c, g = be.eval(x, True)
    c0, _ = be.eval(x, False)
    assert abs(c - ro["contrast"]) <= RTOL * abs(ro["contrast"]) + 1e-12
    assert abs(c0 - ro["contrast"]) <= RTOL * abs(ro["contrast"]) + 1e-12
    assert np.abs(g - ro["grad"]).max() <= RTOL * np.abs(ro["grad"]).max() + 1e-12
    if images:
        ilo, iln = be.local_iwe(x)
        tol = 4e-6 * max(1.0, float(ro["il_old"].max()), float(ro["il_new"].max()))
        assert np.abs(ilo - ro["il_old"]).max() <= tol and np.abs(iln - ro["il_new"]).max() <= tol
        assert np.abs(be.computeImageOfWarpedEvents(x) - ro["iwe"]).max() <= 4e-6 * max(1.0, float(ro["iwe"].max()))
        bands = be.derivative_bands(x, True)
        assert np.abs(bands - ro["bands"]).max() <= 4e-6 * max(1.0, float(np.abs(ro["bands"]).max()))
    return ro


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("grad_mode", [0, 1])
def test_be_parity(oracle, order, grad_mode):
    w = _window(20000, order)
    rng = np.random.default_rng(order)
    IGp = np.abs(rng.normal(0, 0.3, (64, 128))).astype(np.float32)
    be, a = _pair(oracle, w, IGp, 0.5, grad_mode)
    P = 3 * (8 - w.n_fixed)
    _check(be, oracle, a, None)
    _check(be, oracle, a, rng.normal(0, 0.02, P))
    be.close()


@pytest.mark.parametrize("measure", [0, 1])
def test_be_parity_variants(oracle, measure):
    """sample rate 3, batch 64, trailing single-event batch (n % bs == 1), no IGp, sigma 0.7, mean-square."""
    w = _window(64 * 90 + 1, 4, seed=12)
    assert len(w.events) % 64 == 1
    be, a = _pair(oracle, w, None, 0.0, 1, sample_rate=3, batch_size=64, sigma=0.7, measure=measure)
    x = np.random.default_rng(1).normal(0, 0.02, 3 * (8 - w.n_fixed))
    ro = _check(be, oracle, a, x)
    assert ro["cells"][-1] == -2            # the reference loop never visits that event
    be.close()


@pytest.mark.parametrize("order", [2, 4])
def test_be_golden_fixture(oracle, golden, order):
    from cmax_slam_b200.backend import EventWarperCMax
    g = golden("be_oracle.npz")
    lut = synth.bearing_lut(64, 48, K_T)
    for mode in (0, 1):
        be = EventWarperCMax(64, 48, lut, 128, 64, spline_order=order, grad_mode=mode, event_sample_rate=2 if order == 4 else 1)
        be.set_window(g[f"events_{order}"], g[f"knots_{order}"], int(g[f"t0_{order}"]), int(g[f"dt_{order}"]),
                      1 if order == 2 else 3, tuple(g[f"tnext_{order}"]), g[f"IGp_{order}"], 0.5)
        x = g[f"x_{order}"]
        c, gr = be.eval(x, True)
        assert abs(c - float(g[f"contrast_{order}"])) <= RTOL * float(g[f"contrast_{order}"])
        assert np.abs(gr - g[f"grad_{order}"]).max() <= RTOL * np.abs(g[f"grad_{order}"]).max()
        assert np.array_equal(be.warped_cells(x), g[f"cells_{order}"])
        be.close()


@pytest.mark.parametrize("order", [2, 4])
def test_device_spline_pose_table(oracle, order):
    """Device So3Spline<N>::evaluate (+ f32 knot Jacobians) vs the oracle's spline, itself pinned to
    the real basalt code (tests/test_oracle_spline.py)."""
    w = _window(5000, order, seed=21)
    be, _ = _pair(oracle, w)
    P = 3 * (8 - w.n_fixed)
    x = np.random.default_rng(5).normal(0, 0.05, P)
    R, Jk, idx = be.batch_poses(x)
    kn = w.knots_xyzw.copy()
    for i in range(w.n_fixed, 8):
        q = np.zeros(4)
        oracle.lib().orc_so3_exp(oracle._d(np.ascontiguousarray(x[3 * (i - w.n_fixed):3 * (i - w.n_fixed) + 3])), oracle._d(q))
        kn[i] = synth._qmul(q, kn[i]); kn[i] /= np.linalg.norm(kn[i])
    import ctypes as C
    nb = len(idx)
    assert nb == (len(w.events) + 99) // 100
    for b in range(0, nb, 3):
        e0, e1 = w.events[b * 100], w.events[min((b + 1) * 100, len(w.events)) - 1]
        s, ns = C.c_uint32(), C.c_uint32()
        oracle.lib().orc_batch_mid_time(int(e0["sec"]), int(e0["nsec"]), int(e1["sec"]), int(e1["nsec"]), C.byref(s), C.byref(ns))
        t_ns = s.value * 10**9 + ns.value
        q, Ro, io, Jo = oracle.spline_eval(order, kn, w.t0_ns, w.dt_ns, t_ns)
        assert idx[b] == io
        assert np.abs(R[b] - Ro).max() < 1e-13
        Jref = np.concatenate([Jo[k] for k in range(order)], axis=1)
        assert np.abs(Jk[b] - Jref.astype(np.float32)).max() <= 2e-6 * max(1.0, np.abs(Jref).max())
    be.close()


def test_be_alpha_from_first_eval(oracle):
    """alpha = NaN -> updateAlpha on the first evaluation of the window, then frozen (:201-210)."""
    w = _window(15000, 2, seed=33)
    rng = np.random.default_rng(2)
    IGp = np.abs(rng.normal(0, 0.5, (64, 128))).astype(np.float32)
    be, a = _pair(oracle, w, IGp, float("nan"))
    assert np.isnan(be.alpha)
    x0 = np.zeros(3 * (8 - w.n_fixed))
    c, _ = be.eval(x0, False)
    ro = oracle.be_eval(a, x0, False, images=True)     # a.alpha is NaN here; only il_old/il_new are used
    alpha = oracle.update_alpha(IGp, ro["il_old"] + ro["il_new"])
    assert abs(be.alpha - alpha) <= 1e-6 * alpha
    a2 = oracle.be_args(w.events, w.lut, 64, 48, 128, 64, w.knots_xyzw, w.t0_ns, w.dt_ns, 2, w.n_fixed, w.tnext, IGp, alpha)
    x = rng.normal(0, 0.02, len(x0))
    r2 = oracle.be_eval(a2, x, True)
    c, g = be.eval(x, True)
    assert abs(be.alpha - alpha) <= 1e-6 * alpha       # frozen
    assert abs(c - r2["contrast"]) <= RTOL * r2["contrast"]
    assert np.abs(g - r2["grad"]).max() <= RTOL * np.abs(r2["grad"]).max()
    # no IGp content -> alpha = 0
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, np.zeros((64, 128), np.float32), float("nan"))
    be.eval(x0, False)
    assert be.alpha == 0.0
    be.close()


def test_be_errors_and_edges(oracle):
    from cmax_slam_b200._capi import CmaxbError
    from cmax_slam_b200.backend import EventWarperCMax
    w = _window(1000, 2, seed=4)
    be = EventWarperCMax(64, 48, w.lut, 128, 64, spline_order=2)
    with pytest.raises(CmaxbError) as e:
        be.eval(None)
    assert e.value.code == -6
    with pytest.raises(CmaxbError) as e:                       # too few knots for the event times
        be.set_window(w.events, w.knots_xyzw[:4], w.t0_ns, w.dt_ns, 1, w.tnext, None, 0.0)
    assert e.value.code == -5
    bad = w.events.copy(); bad["y"][3] = 48
    with pytest.raises(CmaxbError) as e:
        be.set_window(bad, w.knots_xyzw, w.t0_ns, w.dt_ns, 1, w.tnext, None, 0.0)
    assert e.value.code == -3
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, 1, w.tnext, None, 0.0)
    with pytest.raises(CmaxbError) as e:
        be.eval(np.zeros(5))                                   # wrong parameter count
    assert e.value.code == -1
    # empty window and single event (never visited by the reference loop)
    for ev in (w.events[:0], w.events[:1]):
        be.set_window(ev, w.knots_xyzw, w.t0_ns, w.dt_ns, 1, w.tnext, None, 0.0)
        c, g = be.eval(None, True)
        assert c == 0.0 and np.all(g == 0)
    # everything fixed: no parameters, value still defined
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, 8, w.tnext, None, 0.0)
    a = oracle.be_args(w.events, w.lut, 64, 48, 128, 64, w.knots_xyzw, w.t0_ns, w.dt_ns, 2, 8, w.tnext)
    c, g = be.eval(None, True)
    assert abs(c - oracle.be_eval(a, None, False)["contrast"]) <= RTOL * c and len(g) == 0
    be.close()


def test_be_scale_1m_events_64_knots(oracle):
    """C4-shaped window at 1e6 events (pano 1280x720, 64 knots): direct parity of value + gradient
    (the oracle needs a few seconds: 189 dense bands), and size-independent properties."""
    w = synth.be_config("C4", scale=0.1)
    rng = np.random.default_rng(4)
    IGp = np.abs(rng.normal(0, 0.3, (720, 1280))).astype(np.float32)
    be, a = _pair(oracle, w, IGp, 0.5)
    x = rng.normal(0, 0.01, 3 * 63)
    ro = _check(be, oracle, a, x, images=False)
    ilo, iln = be.local_iwe(x)
    assert abs(ilo.astype(np.float64).sum() + iln.astype(np.float64).sum() - ro["n_inbounds"]) <= 1e-6 * ro["n_inbounds"]
    be.close()


def test_be_gsl_callbacks_on_device(oracle):
    from cmax_slam_b200 import backend
    w = _window(5000, 2, seed=6)
    be, a = _pair(oracle, w)
    ro = oracle.be_eval(a, None, True)
    f, df = backend.global_contrast_fdf(None, be)
    assert abs(f + ro["contrast"]) <= RTOL * ro["contrast"] and np.abs(df + ro["grad"]).max() <= RTOL * np.abs(ro["grad"]).max()
    be.close()


def test_be_full_size_c4_properties():
    """BASELINE config C4 at FULL size (1e7 events, 64 knots, 1280x720): the oracle would need minutes (189 dense
    bands), so the CUDA path is checked through size-independent properties: every in-bounds event casts votes that
    sum to 1 (checksum = integer count of in-bounds cells); value-only == value of f+g; the analytic gradient equals
    the central difference of the value along a random direction; begin + end_launch + end_fetch == eval."""
    w = synth.be_config("C4", scale=1.0)
    rng = np.random.default_rng(4)
    IGp = np.abs(rng.normal(0, 0.3, (720, 1280))).astype(np.float32)
    from cmax_slam_b200.backend import EventWarperCMax
    be = EventWarperCMax(w.sensor_width, w.sensor_height, w.lut, 1280, 720, spline_order=2)
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
    x = rng.normal(0, 0.01, 3 * 63)
    c, g = be.eval(x, True)
    c0, _ = be.eval(x, False)
    assert abs(c - c0) <= 1e-7 * abs(c) and np.isfinite(g).all()
    cells = be.warped_cells(x)
    n_in = int((cells >= 0).sum())
    ilo, iln = be.local_iwe(x)
    assert abs(ilo.astype(np.float64).sum() + iln.astype(np.float64).sum() - n_in) <= 1e-6 * n_in
    # directional derivative at x = 0 (the gradient is taken w.r.t. a fresh left perturbation of the updated knots,
    # so only there d/dx coincides with it exactly; the value is piecewise smooth: bilinear votes, bounds test)
    d = rng.normal(0, 1, len(x)); d /= np.linalg.norm(d)
    h = 2e-4
    _, g0 = be.eval(np.zeros(len(x)), True)
    cp, _ = be.eval(h * d, False)
    cm, _ = be.eval(-h * d, False)
    fd = (cp - cm) / (2 * h)
    assert abs(fd - g0 @ d) <= 2e-2 * max(abs(fd), np.abs(g0).max()), (fd, g0 @ d)
    be.eval_begin(x, True)
    be.eval_end_launch()
    c2, g2 = be.eval_end_fetch()
    assert abs(c2 - c) <= 1e-7 * abs(c) and np.abs(g2 - g).max() <= 1e-6 * np.abs(g).max()
    be.close()
