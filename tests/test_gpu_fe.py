"""GPU parity of the front-end CUDA path against the CPU oracle -- through the C ABI.

Bars (north_star): integer work (per-event cell index, in-bounds count) bit-exact; float variance /
gradient within 1e-5 relative (the gradient with an absolute floor of 1e-5 * max|g|: components cross
zero at the optimum).  Float images: f32 atomics reorder the sums, so pixels agree to a few ulps of
the largest pixel."""
import numpy as np
import pytest

from cmax_slam_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-5
K_T = (60.0, 61.0, 31.5, 23.5)


def _mk(pk, **kw):
    from cmax_slam_b200.frontend import AngVelEstimatorCMax
    fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut, **kw)
    fe.set_packet(pk.events, pk.t_ref_sec)
    return fe


def _check(fe, oracle, a, om, images=True):
    ro = oracle.fe_eval(a, om, True, images=images, cells=True)
    cells = fe.warped_cells(om)
    assert np.array_equal(cells, ro["cells"])                                    # bit-exact integer work
    c, g = fe.eval(om, True)
    c0, g0 = fe.eval(om, False)
    assert g0 is None
    assert abs(c - ro["contrast"]) <= RTOL * abs(ro["contrast"]) + 1e-12
    assert abs(c0 - ro["contrast"]) <= RTOL * abs(ro["contrast"]) + 1e-12
    assert np.abs(g - ro["grad"]).max() <= RTOL * np.abs(ro["grad"]).max() + 1e-12
    if images:
        raw = fe.computeImageOfWarpedEvents(om, blurred=False)
        tol = 4e-6 * max(1.0, float(ro["iwe_raw"].max()))
        assert np.abs(raw - ro["iwe_raw"]).max() <= tol
        assert abs(raw.astype(np.float64).sum() - ro["n_inbounds"]) <= 1e-6 * max(1, ro["n_inbounds"])
        iwe, d = fe.computeImageOfWarpedEvents(om, with_deriv=True, blurred=True)
        assert np.abs(iwe - ro["iwe"]).max() <= tol
        assert np.abs(d - ro["deriv"]).max() <= 4e-6 * max(1.0, float(np.abs(ro["deriv"]).max()))
    return ro


@pytest.mark.parametrize("grad_mode", [0, 1])
@pytest.mark.parametrize("measure", [0, 1])
def test_fe_parity_c1(oracle, grad_mode, measure):
    pk = synth.fe_config("C1", scale=0.3)
    fe = _mk(pk, grad_mode=grad_mode, contrast_measure=measure)
    a = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K, measure=measure)
    for om in (pk.omega_true, np.zeros(3), pk.omega_true + np.array([0.3, -0.2, 0.25])):
        _check(fe, oracle, a, om)
    fe.close()


@pytest.mark.parametrize("grad_mode", [0, 1])
@pytest.mark.parametrize("sigma", [0.0, 0.7, 2.0])
def test_fe_parity_ragged_image_and_blur_sizes(oracle, grad_mode, sigma):
    """346x260 (live DAVIS, not a multiple of the 32x32 tile) and other blur radii."""
    pk = synth.make_fe_packet(30011, 346, 260, (250.0, 251.0, 170.3, 128.9), 17, 900)
    fe = _mk(pk, grad_mode=grad_mode, blur_sigma=sigma, event_batch_size=64)
    a = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, 346, 260, pk.K, batch_size=64, blur_sigma=sigma)
    _check(fe, oracle, a, pk.omega_true + np.array([0.1, 0.2, -0.3]))
    fe.close()


def test_fe_golden_fixture(oracle, golden):
    from cmax_slam_b200.frontend import AngVelEstimatorCMax
    g = golden("fe_oracle.npz")
    lut = synth.bearing_lut(64, 48, tuple(g["K"]))
    for m in (0, 1):
        for mode in (0, 1):
            fe = AngVelEstimatorCMax(64, 48, tuple(g["K"]), lut, contrast_measure=m, grad_mode=mode)
            fe.set_packet(g["events"], float(g["t_ref_sec"]))
            for i, om in enumerate(g["omegas"]):
                c, gr = fe.eval(om, True)
                assert abs(c - float(g[f"contrast_{i}_{m}"])) <= RTOL * abs(float(g[f"contrast_{i}_{m}"]))
                assert np.abs(gr - g[f"grad_{i}_{m}"]).max() <= RTOL * np.abs(g[f"grad_{i}_{m}"]).max()
                if m == 0:
                    assert np.array_equal(fe.warped_cells(om), g[f"cells_{i}"])
            fe.close()


def test_fe_full_size_c2_parity_and_properties(oracle):
    """BASELINE config C2 (1M events, 640x480) at full size: direct parity (the oracle needs ~0.1 s)
    plus size-independent properties."""
    pk = synth.fe_config("C2")
    a = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K)
    om = pk.omega_true + np.array([0.2, -0.1, 0.15])
    for mode in (0, 1):
        fe = _mk(pk, grad_mode=mode)
        _check(fe, oracle, a, om, images=(mode == 0))
        # votes sum to the in-bounds count (a checksum of the whole scatter)
        cells = fe.warped_cells(om)
        raw = fe.computeImageOfWarpedEvents(om, blurred=False)
        assert abs(raw.astype(np.float64).sum() - (cells >= 0).sum()) <= 1e-6 * len(cells)
        # contrast is maximal near the true motion; analytic gradient vs finite differences of the GPU value
        c_true, _ = fe.eval(pk.omega_true, False)
        assert c_true > fe.eval(np.zeros(3), False)[0] and c_true > fe.eval(pk.omega_true + 0.5, False)[0]
        c, g = fe.eval(om, True)
        fd = np.array([(fe.eval(om + e, False)[0] - fe.eval(om - e, False)[0]) / 2e-3 for e in np.eye(3) * 1e-3])
        assert np.abs(g - fd).max() < 0.05 * np.abs(fd).max()
        # repeatability (atomics reorder f32 sums: not bit-exact, but far inside the tolerance)
        c2, g2 = fe.eval(om, True)
        assert abs(c - c2) <= 1e-7 * abs(c) and np.abs(g - g2).max() <= 1e-6 * np.abs(g).max()
        fe.close()


def test_fe_scatter_is_additive_over_packets(oracle):
    """Linearity: the raw IWE of a packet is the sum of the raw IWEs of its batch-aligned halves
    (same t_ref), for every omega."""
    pk = synth.fe_config("C1", scale=0.4)
    n = (len(pk.events) // 200) * 100
    om = np.array([0.5, -0.7, 1.9])
    fe = _mk(pk)
    whole = fe.computeImageOfWarpedEvents(om, blurred=False).astype(np.float64)
    fe.set_packet(pk.events[:n], pk.t_ref_sec)
    a = fe.computeImageOfWarpedEvents(om, blurred=False).astype(np.float64)
    fe.set_packet(pk.events[n:], pk.t_ref_sec)
    b = fe.computeImageOfWarpedEvents(om, blurred=False).astype(np.float64)
    assert np.abs(whole - (a + b)).max() <= 4e-6 * whole.max()
    fe.close()


@pytest.mark.parametrize("bs", [7, 100, 1024, 3000])
def test_fe_batch_sizes_of_the_packet_preparation(oracle, bs):
    """The binning's counting pass forms the batch time offsets when its chunk is a multiple of the batch size (7: more batches
    per chunk than threads; 100: the reference's; 1024: two per chunk); 3000 exceeds the chunk: separate kernel."""
    pk = synth.fe_config("C1", scale=0.2)
    fe = _mk(pk, event_batch_size=bs)
    a = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K, batch_size=bs)
    om = pk.omega_true + np.array([0.2, -0.1, 0.3])
    _check(fe, oracle, a, om)
    fe.close()


def test_fe_edge_cases(oracle):
    from cmax_slam_b200._capi import CmaxbError
    from cmax_slam_b200.frontend import AngVelEstimatorCMax
    pk = synth.make_fe_packet(1234, 64, 48, K_T, 3, 100)
    fe = AngVelEstimatorCMax(64, 48, K_T, pk.lut, max_hypotheses=3)
    with pytest.raises(CmaxbError) as e:                       # eval before set_packet
        fe.eval([0, 0, 0])
    assert e.value.code == -6
    # empty packet: zero image, zero contrast, zero gradient (like the oracle)
    fe.set_packet(pk.events[:0], pk.t_ref_sec)
    c, g = fe.eval([1, 2, 3], True)
    assert c == 0.0 and np.all(g == 0)
    # ragged last batch (1234 % 100 != 0), single event, and a packet warped fully out of the image
    for ev, om in ((pk.events, [0.3, 0.2, 0.1]), (pk.events[:1], [0.3, 0.2, 0.1]), (pk.events, [900.0, 0, 0])):
        fe.set_packet(ev, pk.t_ref_sec)
        a = oracle.fe_args(ev, pk.t_ref_sec, pk.lut, 64, 48, K_T)
        ro = oracle.fe_eval(a, om, True, cells=True)
        assert np.array_equal(fe.warped_cells(om), ro["cells"])
        c, g = fe.eval(om, True)
        assert abs(c - ro["contrast"]) <= RTOL * abs(ro["contrast"]) + 1e-12
        assert np.abs(g - ro["grad"]).max() <= RTOL * np.abs(ro["grad"]).max() + 1e-12
    # reference aborts / throws -> error codes
    bad = pk.events.copy(); bad["x"][7] = 64
    with pytest.raises(CmaxbError) as e:
        fe.set_packet(bad, pk.t_ref_sec)
    assert e.value.code == -3
    bad = pk.events.copy(); bad[[0, 99]] = bad[[99, 0]]
    with pytest.raises(CmaxbError) as e:
        fe.set_packet(bad, pk.t_ref_sec)
    assert e.value.code == -4
    with pytest.raises(CmaxbError) as e:                       # packet was rejected: no evaluation possible
        fe.eval([0, 0, 0])
    fe.set_packet(pk.events, pk.t_ref_sec)
    with pytest.raises(CmaxbError) as e:
        fe.eval_batch(np.zeros((4, 3)))                        # k > max_hypotheses
    assert e.value.code == -1
    fe.close()


@pytest.mark.parametrize("grad_mode", [0, 1])
def test_fe_hypothesis_batch_matches_singles(oracle, grad_mode):
    pk = synth.fe_config("C1", scale=0.3)
    fe = _mk(pk, grad_mode=grad_mode, max_hypotheses=8)
    oms = synth.fe_hypotheses(pk, 8)
    c, g = fe.eval_batch(oms, True)
    a = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K)
    co, go = oracle.fe_eval_batch(a, oms, True, n_threads=4)
    assert np.abs(c - co).max() <= RTOL * np.abs(co).max()
    assert np.abs(g - go).max() <= RTOL * np.abs(go).max()
    fe.eval_launch(oms[:5], False)
    c5, _ = fe.eval_fetch()
    assert np.abs(c5 - co[:5]).max() <= RTOL * np.abs(co).max()
    fe.close()


def test_gsl_callbacks_on_device(oracle):
    from cmax_slam_b200 import frontend
    pk = synth.fe_config("C1", scale=0.1)
    fe = _mk(pk)
    a = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K)
    ro = oracle.fe_eval(a, [0.1, 0.2, 0.3], True)
    f, df = frontend.local_contrast_fdf([0.1, 0.2, 0.3], fe)
    assert abs(f + ro["contrast"]) <= RTOL * ro["contrast"] and np.abs(df + ro["grad"]).max() <= RTOL * np.abs(ro["grad"]).max()
    # value-only evaluations sum the squared blurred image; gradient evaluations get S2 = <I, B^T B I> / 2 from the adjoint
    # image (csrc/fe_fused.cuh): the two agree to the f32 rounding of the filter coefficients
    assert abs(frontend.local_contrast_f([0.1, 0.2, 0.3], fe) - f) <= 2e-6 * abs(f)
    fe.close()


def test_native_library_is_what_ran():
    from cmax_slam_b200 import _capi
    n0 = _capi.launch_count()
    pk = synth.fe_config("C1", scale=0.05)
    fe = _mk(pk)
    fe.eval(pk.omega_true, True)
    fe.close()
    assert _capi.launch_count() - n0 >= 3      # validate + batch-dt kernels + the fused evaluation kernel
    assert "libcmax_b200.so" in open("/proc/self/maps").read()


def test_fe_more_hypotheses_than_one_launch_holds(oracle):
    """k = 40 > 32 hypotheses: the fused kernel is launched in chunks; rows must line up with the single evaluations."""
    pk = synth.fe_config("C1", scale=0.1)
    fe = _mk(pk, max_hypotheses=40)
    oms = synth.fe_hypotheses(pk, 40, sigma=0.3)
    c, g = fe.eval_batch(oms, True)
    a = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K)
    co, go = oracle.fe_eval_batch(a, oms, True, n_threads=8)
    assert np.abs(c - co).max() <= RTOL * np.abs(co).max()
    assert np.abs(g - go).max() <= RTOL * np.abs(go).max()
    cv, _ = fe.eval_batch(oms[:33], False)
    assert np.abs(cv - co[:33]).max() <= RTOL * np.abs(co).max()
    fe.close()


def test_fe_async_packet_upload_and_deferred_verdict(oracle):
    """cmaxb_fe_set_packet_async: evaluation right after the queued upload; a bad packet is reported by the
    next evaluation instead of by set_packet."""
    from cmax_slam_b200._capi import CmaxbError
    pk = synth.fe_config("C1", scale=0.1)
    fe = _mk(pk)
    om = pk.omega_true + 0.1
    a = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K)
    ro = oracle.fe_eval(a, om, True)
    fe.set_packet(pk.events, pk.t_ref_sec, wait=False)
    c, g = fe.eval(om, True)
    assert abs(c - ro["contrast"]) <= RTOL * ro["contrast"] and np.abs(g - ro["grad"]).max() <= RTOL * np.abs(ro["grad"]).max()
    bad = pk.events.copy(); bad["x"][11] = pk.width
    fe.set_packet(bad, pk.t_ref_sec, wait=False)            # accepted: nothing has been checked yet
    with pytest.raises(CmaxbError) as e:
        fe.eval(om, True)
    assert e.value.code == -3
    with pytest.raises(CmaxbError) as e:                     # the packet is now marked unusable
        fe.eval(om, True)
    assert e.value.code == -6
    fe.set_packet(pk.events, pk.t_ref_sec)
    c2, _ = fe.eval(om, False)
    assert abs(c2 - ro["contrast"]) <= RTOL * ro["contrast"]
    fe.close()
