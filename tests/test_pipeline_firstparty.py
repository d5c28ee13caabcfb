"""The reference run AS A WHOLE on the CPU -- its node classes, warp, focus functions and its own optimiser glue
(src/frontend/local_optim_contrast_gsl.cpp, src/backend/global_optim_contrast_gsl*.cpp) compiled unmodified with the stand-in
headers of oracle/stubs/ (ROS, OpenCV, glog, and a GSL interface whose Fletcher-Reeves minimiser is restated) into
oracle/_ref/libref_full.so -- against
  (a) the Python restatement pipeline (oracle/pipeline_py.py = gsl_fr.py + pgo_py.py over the oracle cost): angular velocity of
      every packet IDENTICAL, control poses to 1e-15, the panoramic map and its visit counts identical;
  (b) the library's own Fletcher-Reeves loop (csrc/optim.cu through cmaxb_optimize_callback) driven with the oracle cost: the same
      angular velocities, bit for bit.
Golden copy of the reference run: tests/golden/pipeline_firstparty.npz; live run when oracle/_ref is built.  Host code only."""
import numpy as np
import pytest

from cmax_slam_b200 import synth

K_T = (120.0, 122.0, 63.0, 47.0)


def _seq(seed):
    return synth.make_be_window(120000, 9, 256, 128, seed, order=2, sensor=(128, 96), K4=K_T, n_landmarks=800, knot_sigma=0.1)


def _qdist(a, b):
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    c = np.minimum(np.linalg.norm(a - b, axis=1), np.linalg.norm(a + b, axis=1))
    return float((4 * np.arcsin(np.clip(c / 2, 0, 1))).max())


@pytest.mark.parametrize("tag,order,seed", [("lin", 2, 31), ("cub", 4, 32)])
def test_python_restatement_pipeline_equals_the_reference_run(oracle, golden, tag, order, seed):
    from oracle import pipeline_py
    g = golden("pipeline_firstparty.npz")
    w = _seq(seed)
    r = pipeline_py.run(w.events, w.lut, 128, 96, K_T, 256, 128, order)
    assert len(r["packets"]) == len(g[f"{tag}_omegas"]) and len(r["windows"]) == int(g[f"{tag}_n_win"]) >= 2
    for i, (tp, om) in enumerate(r["packets"]):
        assert tuple(tp) == tuple(int(v) for v in g[f"{tag}_stamps"][i])
        assert np.array_equal(om, g[f"{tag}_omegas"][i]), (i, om, g[f"{tag}_omegas"][i])       # bit for bit
    for i, (rep, kn) in enumerate(r["windows"]):
        v = g[f"{tag}_win{i}"]
        assert (int(v[9]), int(v[10]), int(v[11]), int(v[12])) == (rep["n_ctrl_poses"], rep["idx_cp_traj_beg"], rep["idx_cp_opt_beg"], rep["num_cp_opt"])
        assert rep["optimized"] == 1
        assert _qdist(kn, g[f"{tag}_knots{i}"]) <= 1e-14, (i, _qdist(kn, g[f"{tag}_knots{i}"]))
        assert _qdist(rep["pose_latest"][1], g[f"{tag}_latest{i}"]) <= 1e-14
    assert np.array_equal(r["IG"], g[f"{tag}_IG"]) and np.array_equal(r["times"], g[f"{tag}_times"])


def test_library_solver_loop_reproduces_the_reference_front_end_solves(oracle, golden):
    """csrc/optim.cu's loop (cmaxb_optimize_callback) over the oracle's front-end cost, packet after packet with the
    reference's warm start: the angular velocities of the reference run, bit for bit."""
    from cmax_slam_b200 import _capi
    from oracle import pipeline_py
    g = golden("pipeline_firstparty.npz")
    w = _seq(31)
    stats = []

    def solver(f, fdf, x0):
        x, st = _capi.optimize_callback(f, fdf, x0, (0.1, 0.05, 50, 1e-3, 1e-4))     # local_optim_contrast_gsl.cpp:106-122
        stats.append(st)
        return x

    # the back-end does not feed back into the front-end: stop after the packets of the first 60 % of the stream
    n = int(0.6 * len(w.events))
    r = pipeline_py.run(w.events[:n], w.lut, 128, 96, K_T, 256, 128, 2, fe_solver=solver, win_size=1e6)
    assert len(r["packets"]) >= 15
    for i, (tp, om) in enumerate(r["packets"]):
        assert np.array_equal(om, g["lin_omegas"][i]), (i, om, g["lin_omegas"][i])
    assert all(s["f_evals"] > s["g_evals"] > 0 for s in stats)


@pytest.mark.parametrize("degree,seed", [(1, 41), (3, 42)])
def test_live_reference_run_equals_python_restatement_pipeline(oracle, degree, seed):
    if not oracle.have_ref_full():
        pytest.skip("oracle/_ref/libref_full.so not built (reference tree absent)")
    from oracle import pipeline_py
    w = synth.make_be_window(90000, 8, 256, 128, seed, order=2, sensor=(128, 96), K4=K_T, n_landmarks=600, knot_sigma=0.08)
    ref = oracle.RefNode(128, 96, K_T, np.zeros((1, 3)), dt_ang_vel=0.01, num_events_per_packet=6000, dt_knots=0.05, spline_degree=degree,
                         pano_height=128, min_ev_rate=10, max_update_times=30, full=True)
    ref.events(w.events)
    n_pk, n_win, _ = ref.counts()
    r = pipeline_py.run(w.events, w.lut, 128, 96, K_T, 256, 128, 4 if degree == 3 else 2)
    assert len(r["packets"]) == n_pk and len(r["windows"]) == n_win >= 1
    for i, (tp, om) in enumerate(r["packets"]):
        assert np.array_equal(om, ref.packet_omega(i)), i
    for i, (rep, kn) in enumerate(r["windows"]):
        v, _, lq, knots = ref.window(i)
        assert _qdist(kn, knots) <= 1e-14
    IG, times = ref.get_map(256, 128)
    assert np.array_equal(r["IG"], IG) and np.array_equal(r["times"], times)
    ref.close()


def test_live_reference_run_with_a_silent_gap_and_skipped_windows(oracle):
    """A 0.3 s hole in the event stream: the front-end meets a packet spanning more than 10 dt (angular velocity forced to zero,
    no solve) and the back-end meets windows with fewer events than min_num_ev_per_win_ (no solve, no map update) -- the reference
    run and the Python restatement pipeline still agree bit for bit."""
    if not oracle.have_ref_full():
        pytest.skip("oracle/_ref/libref_full.so not built (reference tree absent)")
    from oracle import pipeline_py
    w = synth.make_be_window(100000, 16, 256, 128, 51, order=2, sensor=(128, 96), K4=K_T, n_landmarks=600, knot_sigma=0.08)
    ev = w.events
    t = ev["sec"].astype(np.int64) * 1_000_000_000 + ev["nsec"].astype(np.int64)
    t0 = int(t[0])
    keep = (t < t0 + 250_000_000) | (t > t0 + 550_000_000)          # drop 0.3 s of events
    ev = ev[keep]
    rate = 4000                                                        # events / s needed for a window to be solved (0.2 s -> 800 events)
    ref = oracle.RefNode(128, 96, K_T, np.zeros((1, 3)), dt_ang_vel=0.01, num_events_per_packet=6000, dt_knots=0.05, spline_degree=1,
                         pano_height=128, min_ev_rate=rate, max_update_times=30, full=True)
    ref.events(ev)
    n_pk, n_win, _ = ref.counts()
    r = pipeline_py.run(ev, w.lut, 128, 96, K_T, 256, 128, 2, min_ev_rate=rate)
    assert len(r["packets"]) == n_pk and len(r["windows"]) == n_win >= 3
    zero = 0
    for i, (tp, om) in enumerate(r["packets"]):
        assert np.array_equal(om, ref.packet_omega(i)), i
        zero += int(not om.any())
    assert zero >= 1                                                   # the packet across the hole was not solved
    skipped = 0
    for i, (rep, kn) in enumerate(r["windows"]):
        v, _, lq, knots = ref.window(i)
        assert _qdist(kn, knots) <= 1e-14, i
        skipped += int(rep["optimized"] == 0)
    assert 1 <= skipped < n_win                                        # some windows skipped, some solved
    IG, times = ref.get_map(256, 128)
    assert np.array_equal(r["IG"], IG) and np.array_equal(r["times"], times)
    ref.close()
