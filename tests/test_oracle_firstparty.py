"""The oracle's restatement of the WHOLE hot path against the reference's OWN translation units --
src/frontend/local_image_warped_events.cpp, src/frontend/local_focus_funcs.cpp, src/backend/event_pano_warper.cpp,
src/backend/global_focus_funcs.cpp (+ trajectory.cpp, image_geom_util.cpp, equirectangular_camera.h, real basalt / Sophus /
Eigen) -- compiled UNMODIFIED with the stand-in headers of oracle/stubs/ (ROS time, dvs_msgs::Event, a cv::Mat whose image
primitives are the oracle's cv2-pinned restatements, glog CHECKs) into oracle/_ref/libref_{fe,focus,warper}.so.
Same compiler, same flags, sequential event order on both sides: images, contrast and gradient are BIT-EXACT.
Golden copy (tests/golden/hotpath_firstparty.npz) for machines without the reference tree; live tests when oracle/_ref is built."""
import numpy as np
import pytest

from cmax_slam_b200 import synth


def _fe_packet():
    pk = synth.make_fe_packet(4000, 48, 36, (40.0, 41.0, 23.5, 17.5), 71, 120)
    return pk


def test_front_end_images_contrast_gradient_bit_exact(oracle, golden):
    g = golden("hotpath_firstparty.npz")
    pk = _fe_packet()
    sec, nsec = (int(v) for v in g["fe_tref"])
    assert float(sec) + 1e-9 * float(nsec) == pk.t_ref_sec
    for k in range(3):
        om = g[f"fe{k}_omega"]
        for m in (0, 1):
            a = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, 48, 36, pk.K, measure=m)
            o = oracle.fe_eval(a, om, True, images=True)
            assert np.array_equal(o["iwe"], g[f"fe{k}_iwe"]) and np.array_equal(o["deriv"], g[f"fe{k}_deriv"]), (k, m)
            assert o["contrast"] == float(g[f"fe{k}_contrast{m}"]) and np.array_equal(o["grad"], g[f"fe{k}_grad{m}"]), (k, m)


@pytest.mark.parametrize("order", [2, 4])
def test_back_end_images_alpha_contrast_gradient_bit_exact(oracle, golden, order):
    g = golden("hotpath_firstparty.npz")
    w = synth.make_be_window(5000, 6, 64, 32, 72 + order, order=order, sensor=(32, 24), K4=(30.0, 30.5, 15.5, 11.5), n_landmarks=150,
                             n_fixed=1 if order == 2 else 3)
    t_beg, dtk = g[f"be{order}_tbeg"]
    t0_ns, dt_ns = int(1e9 * t_beg), int(1e9 * dtk)            # the temporary trajectory's origin (trajectory.cpp:60-63)
    IG = g[f"be{order}_IG"]
    a0 = oracle.be_args(w.events, w.lut, 32, 24, 64, 32, w.knots_xyzw, t0_ns, dt_ns, order, w.n_fixed, w.tnext, IG, 0.0)
    o0 = oracle.be_eval(a0, None, False, images=True)
    alpha = oracle.update_alpha(IG, o0["il_old"] + o0["il_new"])
    assert alpha == float(g[f"be{order}_alpha"])
    a = oracle.be_args(w.events, w.lut, 32, 24, 64, 32, w.knots_xyzw, t0_ns, dt_ns, order, w.n_fixed, w.tnext, IG, alpha)
    o = oracle.be_eval(a, None, True, images=True)
    for k in ("il_old", "il_new", "iwe", "bands"):
        assert np.array_equal(o[k], g[f"be{order}_{k}"]), k
    assert o["contrast"] == float(g[f"be{order}_contrast"]) and np.array_equal(o["grad"], g[f"be{order}_grad"])


def test_live_front_end_against_reference_translation_units(oracle):
    if not oracle.have_ref_firstparty():
        pytest.skip("oracle/_ref first-party libraries not built (reference tree absent)")
    rng = np.random.default_rng(3)
    pk = synth.fe_config("C1", scale=0.1)
    sec = int(np.floor(pk.t_ref_sec)); nsec = int(round((pk.t_ref_sec - sec) * 1e9))
    # ragged last batch (not a multiple of 100), a non-default batch size and no blur
    for n, bs, sigma in ((len(pk.events), 100, 1.0), (7777, 100, 1.0), (5001, 37, 0.0), (2500, 100, 2.0)):
        ev = pk.events[:n]
        om = pk.omega_true + rng.normal(0, 0.3, 3)
        iwe, der = oracle.ref1p_fe_images(ev, (sec, nsec), pk.lut, pk.width, pk.height, pk.K, om, True, blur_sigma=sigma, batch_size=bs)
        for m in (0, 1):
            a = oracle.fe_args(ev, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K, bs, sigma, m)
            o = oracle.fe_eval(a, om, True, images=True)
            assert np.array_equal(o["iwe"], iwe) and np.array_equal(o["deriv"], der), (n, bs, sigma)
            c, g = oracle.ref1p_fe_contrast(iwe, der, m)
            assert o["contrast"] == c and np.array_equal(o["grad"], g), (n, bs, sigma, m)
            cv, _ = oracle.ref1p_fe_contrast(iwe, None, m)
            assert cv == oracle.fe_eval(a, om, False)["contrast"]


@pytest.mark.parametrize("order,sample_rate,n_events", [(2, 1, 20000), (4, 1, 20000), (2, 3, 15001), (4, 2, 10101)])
def test_live_back_end_against_reference_translation_units(oracle, order, sample_rate, n_events):
    """Two consecutive 'windows' on one EventWarper: first iteration with an empty map (alpha = 0), updateIG + FOV marks, then a
    first iteration with the accumulated map; strides inside batches, the skipped trailing one-event batch (n = 10101)."""
    if not oracle.have_ref_firstparty():
        pytest.skip("oracle/_ref first-party libraries not built (reference tree absent)")
    K_T = (60.0, 61.0, 31.5, 23.5)
    w = synth.make_be_window(n_events, 8, 128, 64, 13 + order, order=order, sensor=(64, 48), K4=K_T, n_landmarks=300,
                             n_fixed=1 if order == 2 else 3)
    dtk, t_beg = w.dt_ns / 1e9, w.t0_ns / 1e9
    t0_ns, dt_ns = int(1e9 * t_beg), int(1e9 * dtk)
    rw = oracle.RefEventWarper(w.lut, 64, 48, 128, 64, order=order, sample_rate=sample_rate, max_update_times=1)
    IG = np.zeros((64, 128), np.float32)
    times = np.zeros((64, 128), np.uint8)
    rng = np.random.default_rng(9)
    for it in range(2):
        x = rng.normal(0, 0.02, 3 * (8 - w.n_fixed))
        from cmax_slam_b200 import trajectory as T
        knots = T.incremental_update(w.knots_xyzw, w.n_fixed, x)             # exp(x_i) * K_i (pinned in test_traj_firstparty.py)
        r = rw.eval(w.events, t_beg, dtk, knots, w.n_fixed, w.tnext, True, True)
        a0 = oracle.be_args(w.events, w.lut, 64, 48, 128, 64, w.knots_xyzw, t0_ns, dt_ns, order, w.n_fixed, w.tnext, IG, 0.0, 100, sample_rate)
        o0 = oracle.be_eval(a0, x, False, images=True)
        alpha = oracle.update_alpha(IG, o0["il_old"] + o0["il_new"])
        assert alpha == r["alpha"], it
        a = oracle.be_args(w.events, w.lut, 64, 48, 128, 64, w.knots_xyzw, t0_ns, dt_ns, order, w.n_fixed, w.tnext, IG, alpha, 100, sample_rate)
        o = oracle.be_eval(a, x, True, images=True)
        for k in ("il_old", "il_new", "iwe", "bands"):
            assert np.array_equal(o[k], r[k]), (it, k)
        c, g = oracle.ref1p_be_contrast(r["iwe"], r["bands"], 0)
        assert o["contrast"] == c and np.array_equal(o["grad"], g)
        # map upkeep: updateIG (gated by the visit counts) and setUpdateTimesIG
        rw.update_ig()
        oracle.update_ig(IG, o["il_old"], times, 1)
        for q in (knots[0], knots[3]):
            rw.mark_fov(q, 3)
            oracle.set_update_times(w.lut, 64, 48, 128, 64, q, 3, times)
        IG_ref, times_ref = rw.get_map()
        assert np.array_equal(IG, IG_ref) and np.array_equal(times, times_ref), it
    assert times.max() >= 2                                                    # the update gate (max_update_times = 1) was exercised
    rw.close()
