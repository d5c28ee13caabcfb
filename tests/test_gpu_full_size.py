"""Parity at the FULL sizes of BASELINE.json (round-1 verdict: direct oracle parity stopped at 10 % scale; C3 and C5 had no
test at their shape).  The oracle's value-only entry (the reference's df == nullptr path: no derivative bands) with
per-event cells is affordable at 1e7 / 5e7 events, so the integer work is checked bit for bit on every event and the
contrast to 1e-5 at full size; gradients at full size through a directional finite difference of the oracle-checked
contrast.  C3: all 256 hypotheses on the full 1M-event packet against the oracle's batch entry."""
import numpy as np
import pytest

from cmax_slam_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def test_c3_256_hypotheses_on_the_full_packet(oracle):
    from cmax_slam_b200.frontend import AngVelEstimatorCMax
    pk = synth.fe_config("C3")
    assert len(pk.events) == 1_000_000
    oms = synth.fe_hypotheses(pk, 256, seed=3, sigma=0.5)
    a = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K)
    import os
    co, go = oracle.fe_eval_batch(a, oms, True, n_threads=min(32, os.cpu_count() or 1))
    fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut, max_hypotheses=32)
    fe.set_packet(pk.events, pk.t_ref_sec)
    c = np.zeros(256); g = np.zeros((256, 3))
    for c0 in range(0, 256, 32):                     # one GPU's share of C3 = 32 hypotheses per launch
        c[c0:c0 + 32], g[c0:c0 + 32] = fe.eval_batch(oms[c0:c0 + 32], True)
    assert np.abs(c - co).max() <= RTOL * np.abs(co).max()
    assert (np.abs(c - co) <= RTOL * np.abs(co)).all()
    gmax = np.abs(go).max(axis=1, keepdims=True)
    assert (np.abs(g - go) <= RTOL * gmax + 1e-7 * np.abs(go).max()).all()
    # the same hypotheses one per launch over the throughput lanes (cmaxb_fe_eval_launch), value only
    cv = []
    out = 0
    for i in range(64):
        fe.eval_launch(oms[i:i + 1], False); out += 1
        if out >= 6:
            cv.append(fe.eval_fetch()[0][0]); out -= 1
    while out:
        cv.append(fe.eval_fetch()[0][0]); out -= 1
    assert np.abs(np.array(cv) - co[:64]).max() <= RTOL * np.abs(co).max()
    # cells of three hypotheses, every event
    for i in (0, 100, 255):
        ro = oracle.fe_eval(a, oms[i], False, cells=True)
        assert np.array_equal(fe.warped_cells(oms[i]), ro["cells"])
    fe.close()


def _be_full(oracle, name, seed):
    from cmax_slam_b200.backend import EventWarperCMax
    w = synth.be_config(name, device="cuda")
    rng = np.random.default_rng(seed)
    IGp = np.abs(rng.normal(0, 0.3, (w.pano_height, w.pano_width))).astype(np.float32)
    K = len(w.knots_xyzw)
    x = rng.normal(0, 0.01, 3 * (K - w.n_fixed))
    be = EventWarperCMax(w.sensor_width, w.sensor_height, w.lut, w.pano_width, w.pano_height, spline_order=2)
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
    a = oracle.be_args(w.events, w.lut, w.sensor_width, w.sensor_height, w.pano_width, w.pano_height, w.knots_xyzw, w.t0_ns,
                       w.dt_ns, 2, w.n_fixed, w.tnext, IGp, 0.5)
    ro = oracle.be_eval(a, x, False, cells=True)                       # value-only oracle pass: no bands
    cells = be.warped_cells(x)
    assert np.array_equal(cells, ro["cells"]), int((cells != ro["cells"]).sum())     # bit-exact on EVERY event (atan2 / asin included)
    c, g = be.eval(x, True)
    cv, _ = be.eval(x, False)
    assert abs(c - ro["contrast"]) <= RTOL * ro["contrast"] and abs(cv - ro["contrast"]) <= RTOL * ro["contrast"]
    ilo, iln = be.local_iwe(x)
    n_in = int((cells >= 0).sum())
    assert n_in == ro["n_inbounds"]
    assert abs(ilo.astype(np.float64).sum() + iln.astype(np.float64).sum() - n_in) <= 1e-6 * n_in      # sum of votes = in-bounds events
    # gradient at full size: directional derivative against a central difference of the (oracle-checked) contrast
    d = rng.normal(0, 1, len(x)); d /= np.linalg.norm(d)
    h = 2e-4
    fd = (be.eval(x + h * d, False)[0] - be.eval(x - h * d, False)[0]) / (2 * h)
    assert abs(g @ d - fd) <= 2e-3 * max(abs(fd), np.abs(g).max())
    be.close()


def test_c4_full_size_cells_and_contrast(oracle):
    _be_full(oracle, "C4", 4)


def test_c5_full_size_cells_and_contrast(oracle):
    _be_full(oracle, "C5", 5)
