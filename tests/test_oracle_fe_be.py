"""Oracle self-consistency: properties the domain offers + the reference's quirks (SURVEY.md section 7
"Reproducing reference quirks") + regression pins (tests/golden/{fe,be}_oracle.npz)."""
import numpy as np
import pytest

from cmax_slam_b200 import synth

K_T = (60.0, 61.0, 31.5, 23.5)


def _fe(oracle, n=4000, seed=3, **kw):
    pk = synth.make_fe_packet(n, 64, 48, K_T, seed, 120)
    return pk, oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, 64, 48, K_T, **kw)


def test_fe_golden_regression(oracle, golden):
    g = golden("fe_oracle.npz")
    ev = g["events"]
    lut = synth.bearing_lut(64, 48, tuple(g["K"]))
    for m in (0, 1):
        a = oracle.fe_args(ev, float(g["t_ref_sec"]), lut, 64, 48, tuple(g["K"]), measure=m)
        for i, om in enumerate(g["omegas"]):
            r = oracle.fe_eval(a, om, True, images=True, cells=True)
            assert r["contrast"] == float(g[f"contrast_{i}_{m}"])
            assert np.array_equal(r["grad"], g[f"grad_{i}_{m}"])
            if m == 0:
                assert np.array_equal(r["cells"], g[f"cells_{i}"])
                assert np.array_equal(r["iwe"], g[f"iwe_{i}"])
                assert np.array_equal(r["deriv"], g[f"deriv_{i}"])


def test_fe_contrast_peaks_at_true_motion(oracle):
    pk, a = _fe(oracle, 20000)
    c_true = oracle.fe_eval(a, pk.omega_true, False)["contrast"]
    assert c_true > 1.2 * oracle.fe_eval(a, [0, 0, 0], False)["contrast"]
    for d in np.eye(3) * 0.4:
        assert c_true > oracle.fe_eval(a, pk.omega_true + d, False)["contrast"]


@pytest.mark.parametrize("measure", [0, 1])
def test_fe_gradient_matches_finite_differences(oracle, measure):
    pk, a = _fe(oracle, 20000, measure=measure)
    w = pk.omega_true + np.array([0.15, -0.1, 0.2])
    g = oracle.fe_eval(a, w, True)["grad"]
    fd = np.array([(oracle.fe_eval(a, w + e, False)["contrast"] - oracle.fe_eval(a, w - e, False)["contrast"]) / 2e-3
                   for e in np.eye(3) * 1e-3])
    assert np.abs(g - fd).max() < 0.05 * np.abs(fd).max()


def test_fe_votes_sum_to_inbounds_count(oracle):
    pk, a = _fe(oracle, 8000, blur_sigma=0.0)
    r = oracle.fe_eval(a, [2.0, 3.0, -4.0], True, images=True, cells=True)
    assert r["n_inbounds"] == int((r["cells"] >= 0).sum())
    assert abs(r["iwe_raw"].astype(np.float64).sum() - r["n_inbounds"]) < 1e-2
    # each derivative corner quadruple sums to zero -> derivative images sum to ~0
    assert abs(r["deriv_raw"].astype(np.float64).sum()) < 1e-1
    assert np.array_equal(r["iwe"], r["iwe_raw"])   # sigma <= 0: no blur (:32)


def test_fe_bounds_and_truncation_quirk(oracle):
    """xx = (int)x truncates toward zero and the accepted range is 1 <= xx < W-2, 1 <= yy < H-2
    (local_image_warped_events.cpp:139-142)."""
    W, H = 16, 12
    Kc = (10.0, 10.0, 0.0, 0.0)   # pixel == 10 * bearing
    lut = np.zeros((W * H, 3)); lut[:, 2] = 1.0
    pts = [(0.95, 1.0), (1.0, 1.0), (13.999, 5.5), (14.0, 5.5), (5.5, 9.999), (5.5, 10.0), (-0.5, 3.0), (1.5, 0.999)]
    ev = np.zeros(len(pts), synth.EVENT_DTYPE)
    for i, (px, py) in enumerate(pts):
        ev[i]["x"], ev[i]["y"] = i, 0
        lut[i] = [px / 10.0, py / 10.0, 1.0]
        ev[i]["sec"], ev[i]["nsec"] = 100, 0
    a = oracle.fe_args(ev, 100.0, lut, W, H, Kc, batch_size=100, blur_sigma=0.0)
    cells = oracle.fe_eval(a, [0, 0, 0], False, cells=True)["cells"]
    expect = [-1, 1 * W + 1, 5 * W + 13, -1, 9 * W + 5, -1, -1, -1]
    assert list(cells) == expect


def test_fe_batch_shares_one_dt(oracle):
    """All events of a batch are warped with the batch mid-time (:67-76): shuffling timestamps inside
    a batch while keeping its first and last stamp leaves the result unchanged."""
    pk, a = _fe(oracle, 1000, batch_size=100)
    r0 = oracle.fe_eval(a, [1.0, -2.0, 3.0], True, cells=True)
    ev = pk.events.copy()
    for b in range(0, 1000, 100):
        ev["nsec"][b + 1:b + 99] = ev["nsec"][b]          # interior stamps irrelevant
        ev["sec"][b + 1:b + 99] = ev["sec"][b]
    a2 = oracle.fe_args(ev, pk.t_ref_sec, pk.lut, 64, 48, K_T, batch_size=100)
    r1 = oracle.fe_eval(a2, [1.0, -2.0, 3.0], True, cells=True)
    assert np.array_equal(r0["cells"], r1["cells"]) and r0["contrast"] == r1["contrast"]
    # while another batch size changes it
    a3 = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, 64, 48, K_T, batch_size=10)
    assert oracle.fe_eval(a3, [1.0, -2.0, 3.0], False)["contrast"] != r0["contrast"]


def test_fe_errors(oracle):
    pk, a = _fe(oracle, 300)
    ev = pk.events.copy()
    ev[[0, 99]] = ev[[99, 0]]       # first batch now spans a negative time (CHECK_GE, :72)
    a2 = oracle.fe_args(ev, pk.t_ref_sec, pk.lut, 64, 48, K_T)
    with pytest.raises(RuntimeError, match="non-negative"):
        oracle.fe_eval(a2, [0, 0, 1], False)
    ev = pk.events.copy(); ev["x"][5] = 64
    a3 = oracle.fe_args(ev, pk.t_ref_sec, pk.lut, 64, 48, K_T)
    with pytest.raises(RuntimeError, match="outside the sensor"):
        oracle.fe_eval(a3, [0, 0, 1], False)


def test_fe_empty_packet(oracle):
    pk, _ = _fe(oracle, 10)
    a = oracle.fe_args(pk.events[:0], pk.t_ref_sec, pk.lut, 64, 48, K_T)
    r = oracle.fe_eval(a, [1, 2, 3], True)
    assert r["contrast"] == 0.0 and np.all(r["grad"] == 0)


def test_fe_batch_api_matches_single(oracle):
    pk, a = _fe(oracle, 3000)
    oms = synth.fe_hypotheses(pk, 5)
    c, g = oracle.fe_eval_batch(a, oms, True, n_threads=3)
    for i, om in enumerate(oms):
        r = oracle.fe_eval(a, om, True)
        assert r["contrast"] == c[i] and np.array_equal(r["grad"], g[i])


# ---------------------------------------------------------------------------------------------- BE
def _be(oracle, n=4000, order=2, seed=9, K=8, **kw):
    w = synth.make_be_window(n, K, 128, 64, seed, order=order, sensor=(64, 48), K4=K_T, n_landmarks=300,
                             n_fixed=1 if order == 2 else 3)
    a = oracle.be_args(w.events, w.lut, 64, 48, 128, 64, w.knots_xyzw, w.t0_ns, w.dt_ns, order, w.n_fixed, w.tnext, **kw)
    return w, a


@pytest.mark.parametrize("order", [2, 4])
def test_be_golden_regression(oracle, golden, order):
    g = golden("be_oracle.npz")
    lut = synth.bearing_lut(64, 48, K_T)
    a = oracle.be_args(g[f"events_{order}"], lut, 64, 48, 128, 64, g[f"knots_{order}"], int(g[f"t0_{order}"]),
                       int(g[f"dt_{order}"]), order, 1 if order == 2 else 3, tuple(g[f"tnext_{order}"]), g[f"IGp_{order}"], 0.5,
                       sample_rate=2 if order == 4 else 1)
    r = oracle.be_eval(a, g[f"x_{order}"], True, images=True, cells=True)
    assert r["contrast"] == float(g[f"contrast_{order}"])
    assert np.array_equal(r["grad"], g[f"grad_{order}"])
    assert np.array_equal(r["cells"], g[f"cells_{order}"])
    assert np.array_equal(r["iwe"], g[f"iwe_{order}"])


@pytest.mark.parametrize("order", [2, 4])
def test_be_gradient_matches_finite_differences(oracle, order):
    w, a = _be(oracle, 20000, order)
    P = 3 * (8 - w.n_fixed)
    g = oracle.be_eval(a, None, True)["grad"]
    fd = np.zeros(P)
    for i in range(P):
        e = np.zeros(P); e[i] = 1e-4
        fd[i] = (oracle.be_eval(a, e, False)["contrast"] - oracle.be_eval(a, -e, False)["contrast"]) / 2e-4
    assert np.abs(g - fd).max() < 0.05 * np.abs(fd).max()


def test_be_gradient_ignores_Jl_of_x(oracle):
    """The BE gradient is taken w.r.t. a fresh left perturbation of the ALREADY-UPDATED knots
    (...analytical.cpp:33, trajectory.cpp:236): eval(x) on knots K == eval(0) on knots exp(x)K."""
    w, a = _be(oracle, 5000, 2)
    P = 3 * (8 - w.n_fixed)
    x = np.random.default_rng(0).normal(0, 0.05, P)
    r1 = oracle.be_eval(a, x, True)
    kn = w.knots_xyzw.copy()
    for i in range(w.n_fixed, 8):
        q = synth._qmul(synth._qexp(x[3 * (i - w.n_fixed):3 * (i - w.n_fixed) + 3]), kn[i])
        kn[i] = q / np.linalg.norm(q)
    a2 = oracle.be_args(w.events, w.lut, 64, 48, 128, 64, kn, w.t0_ns, w.dt_ns, 2, w.n_fixed, w.tnext)
    r2 = oracle.be_eval(a2, None, True)
    assert abs(r1["contrast"] - r2["contrast"]) < 1e-9 * r1["contrast"]
    assert np.abs(r1["grad"] - r2["grad"]).max() < 1e-6 * np.abs(r1["grad"]).max()


def test_be_trailing_single_event_batch_is_skipped(oracle):
    """`for (beg = begin; beg < end-1; beg += bs)` (event_pano_warper.cpp:188-196)."""
    w, _ = _be(oracle, 401)
    assert len(w.events) == 401
    a = oracle.be_args(w.events, w.lut, 64, 48, 128, 64, w.knots_xyzw, w.t0_ns, w.dt_ns, 2, 1, w.tnext, batch_size=100)
    cells = oracle.be_eval(a, None, False, cells=True)["cells"]
    assert cells[400] == -2 and np.all(cells[:400] != -2)
    a = oracle.be_args(w.events[:400], w.lut, 64, 48, 128, 64, w.knots_xyzw, w.t0_ns, w.dt_ns, 2, 1, w.tnext, batch_size=100)
    assert np.all(oracle.be_eval(a, None, False, cells=True)["cells"] != -2)
    # 402 events: the trailing batch has two events and IS visited
    w2, _ = _be(oracle, 402)
    a = oracle.be_args(w2.events, w2.lut, 64, 48, 128, 64, w2.knots_xyzw, w2.t0_ns, w2.dt_ns, 2, 1, w2.tnext, batch_size=100)
    assert np.all(oracle.be_eval(a, None, False, cells=True)["cells"] != -2)


def test_be_sample_rate_strides_inside_each_batch(oracle):
    w, _ = _be(oracle, 1000)
    a = oracle.be_args(w.events, w.lut, 64, 48, 128, 64, w.knots_xyzw, w.t0_ns, w.dt_ns, 2, 1, w.tnext, batch_size=100, sample_rate=3)
    cells = oracle.be_eval(a, None, False, cells=True)["cells"]
    visited = np.nonzero(cells != -2)[0]
    expect = np.array([b + j for b in range(0, 1000, 100) for j in range(0, 100, 3)])
    assert np.array_equal(visited, expect)


def test_be_old_new_split_and_alpha(oracle):
    w, a = _be(oracle, 6000)
    r = oracle.be_eval(a, None, False, images=True, cells=True)
    t = w.events["sec"].astype(np.int64) * 10**9 + w.events["nsec"]
    tn = w.tnext[0] * 10**9 + w.tnext[1]
    inb = r["cells"] >= 0
    assert abs(r["il_old"].astype(np.float64).sum() - (inb & (t < tn)).sum()) < 1e-2
    assert abs(r["il_new"].astype(np.float64).sum() - (inb & (t >= tn)).sum()) < 1e-2
    IGp = np.zeros((64, 128), np.float32)
    assert oracle.update_alpha(IGp, r["il_old"] + r["il_new"]) == 0.0          # countNonZero(IGp) < 1
    IL = r["il_old"] + r["il_new"]
    assert abs(oracle.update_alpha(IL, IL) - 1.0) < 1e-12                       # equal densities


def test_be_time_outside_spline_is_an_error(oracle):
    w, _ = _be(oracle, 500)
    a = oracle.be_args(w.events, w.lut, 64, 48, 128, 64, w.knots_xyzw[:4], w.t0_ns, w.dt_ns, 2, 1, w.tnext)
    with pytest.raises(RuntimeError, match="outside the spline"):
        oracle.be_eval(a, None, False)
