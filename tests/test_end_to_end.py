"""The whole data path on the device (cmax_slam_b200.pipeline.CMaxSLAM = the reference node's wiring over the C ABI):
a synthetic rotating camera, events delivered in messages, front-end solve per packet, back-end solve per window.
Checked against the ground truth of the generator (the per-piece parity tests pin each stage to the oracle)."""
import numpy as np
import pytest

from cmax_slam_b200 import synth

pytestmark = pytest.mark.gpu
K_T = (120.0, 122.0, 63.0, 47.0)


def test_synthetic_sequence_front_to_back():
    from cmax_slam_b200.pipeline import CMaxSLAM
    # 0.9 s of rotation at ~2-4 rad/s seen by a 128x96 sensor (f = 120 px): 18 linear-spline knots every 50 ms.  The
    # motion inside one packet must be several pixels: with sub-pixel motion the integer event coordinates make
    # omega = 0 the sharpest image (every vote lands on one pixel), for the reference as for us.
    w = synth.make_be_window(300000, 19, 256, 128, 31, order=2, sensor=(128, 96), K4=K_T, n_landmarks=800, knot_sigma=0.1)
    slam = CMaxSLAM(128, 96, K_T, w.lut, num_events_per_packet=6000, dt_ang_vel=0.01, backend_time_window_size=0.2,
                    backend_sliding_window_stride=0.1, dt_knots=0.05, spline_degree=1, pano_height=128, max_update_times=30)
    slam.ang_vel = np.array([0.05, -0.05, 0.05])                # not exactly 0: the cost is discontinuous there
    for i in range(0, len(w.events), 5000):                     # messages of 5000 events
        slam.eventsCallback(w.events[i:i + 5000])
    assert len(slam.ang_vels) >= 60 and len(slam.windows) >= 5
    # front-end: angular velocity against the generator's (piecewise constant between knots, body frame)
    errs = []
    for (ts, om, stats) in slam.ang_vels[2:-2]:
        t_ns = ts[0] * 1_000_000_000 + ts[1]
        seg = int((t_ns - w.t0_ns) // w.dt_ns)
        if seg < 0 or seg >= len(w.knots_xyzw) - 1:
            continue
        d = synth._qlog(synth._qmul(synth._qconj(w.knots_xyzw[seg][None, :]), w.knots_xyzw[seg + 1][None, :]))[0] / (w.dt_ns * 1e-9)
        # a packet (3000 events ~ 18 ms) may straddle a knot, where the true velocity jumps: compare away from knots
        frac = ((t_ns - w.t0_ns) % w.dt_ns) / w.dt_ns
        if 0.3 < frac < 0.7:
            errs.append(np.abs(om - d).max())
            assert stats is not None and stats["cost_final"] <= stats["cost_initial"]
    assert len(errs) > 10 and np.median(errs) < 0.25, (np.median(errs), max(errs))       # CPU oracle dry run: 0.12 rad/s (of 2-4)
    # back-end: every window solved, contrast improved, indices slide by cp_stride = 2
    for k, rep in enumerate(slam.windows):
        assert rep["optimized"] == 1 and rep["idx_cp_traj_beg"] == 2 * k
        assert rep["opt"]["cost_final"] <= rep["opt"]["cost_initial"]
    # trajectory against the generator, up to the unknown global rotation at the first stamp
    q, t0_ns, dt_ns = slam.trajectory()
    assert len(q) >= 12
    def rel(a, b):
        return synth._qmul(synth._qconj(a[None, :]), b[None, :])[0]
    worst = 0.0
    for i in range(2, len(q) - 2):
        t_ns = t0_ns + i * dt_ns                                   # control pose i of a linear spline = pose at knot time
        seg = (t_ns - w.t0_ns) // w.dt_ns
        u = ((t_ns - w.t0_ns) % w.dt_ns) / w.dt_ns
        if seg + 1 >= len(w.knots_xyzw):
            break
        dd = synth._qlog(rel(w.knots_xyzw[seg], w.knots_xyzw[seg + 1])[None, :])[0]
        truth = synth._qmul(w.knots_xyzw[seg][None, :], synth._qexp((dd * u)[None, :]))[0]
        t_ns0 = t0_ns + 2 * dt_ns
        seg0 = (t_ns0 - w.t0_ns) // w.dt_ns
        u0 = ((t_ns0 - w.t0_ns) % w.dt_ns) / w.dt_ns
        d0 = synth._qlog(rel(w.knots_xyzw[seg0], w.knots_xyzw[seg0 + 1])[None, :])[0]
        truth0 = synth._qmul(w.knots_xyzw[seg0][None, :], synth._qexp((d0 * u0)[None, :]))[0]
        est_rel = rel(q[2], q[i])                                   # rotation from control pose 2 to i
        tru_rel = rel(truth0, truth)
        ang = np.linalg.norm(synth._qlog(rel(tru_rel, est_rel)[None, :])[0])
        worst = max(worst, ang)
    assert worst < 0.08, worst                                      # CPU oracle dry run (scratch/e2e_cpu_dryrun.py): 0.03 rad
    IG = slam.getIG()
    assert IG.sum() > 1e4 and np.isfinite(IG).all()
    slam.close()
