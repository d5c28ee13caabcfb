"""The reference's OWN trajectory code (src/backend/trajectory.cpp: Linear/CubicTrajectory, compiled unmodified with ROS /
OpenCV / glog stand-ins into oracle/_ref/libref_traj.so; golden copy tests/golden/traj_firstparty.npz) against
  * the library's trajectory initialisation (csrc/traj_init.cu: cmaxb_traj_*),
  * the window time origin and the knot update the back-end pipeline / cmaxb_be_set_window use (csrc/pgo.cu),
  * the oracle's spline wrapper incl. the f32 repacking of the knot Jacobians (oracle/cmax_oracle.cpp).
Host code only: no device needed."""
import numpy as np
import pytest

from cmax_slam_b200 import trajectory as T


def _qdist(a, b):
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    c = np.minimum(np.linalg.norm(a - b, axis=1), np.linalg.norm(a + b, axis=1))
    return float((4 * np.arcsin(np.clip(c / 2, 0, 1))).max())


def test_generate_ctrl_poses_matches_reference_source(golden):
    g = golden("traj_firstparty.npz")
    for i in range(int(g["n_gen"])):
        order, dtk = g[f"gen{i}_in"]
        tb, te = tuple(int(v) for v in g[f"gen{i}_tb"]), tuple(int(v) for v in g[f"gen{i}_te"])
        ours = T.generate_ctrl_poses(int(order), float(dtk), g[f"gen{i}_stamps"], g[f"gen{i}_poses"], tb, te)
        ref = g[f"gen{i}_ctrl"]
        assert ours.shape == ref.shape, (i, ours.shape, ref.shape)             # same number of control poses
        assert _qdist(ours, ref) <= 2e-9, (i, _qdist(ours, ref))


def test_window_evaluation_matches_reference_source(oracle, golden):
    """CopyAndIncrementalUpdate + evaluate: time origin int64(1e9 * (t_beg + idx * dt)), left-multiplicative knot update,
    value, first knot index and the f32 Jacobian blocks handed to the event warper."""
    g = golden("traj_firstparty.npz")
    for i in range(int(g["n_win"])):
        order, dtk, idx_traj, idx_opt = g[f"win{i}_in"]
        order, idx_traj, idx_opt = int(order), int(idx_traj), int(idx_opt)
        knots, drotv = g[f"win{i}_knots"], g[f"win{i}_drotv"]
        t_traj = tuple(int(v) for v in g[f"win{i}_ttraj"])
        t = tuple(int(v) for v in g[f"win{i}_t"])
        # the library's knot update == traj_->incrementalUpdate
        after = T.incremental_update(knots, idx_opt, drotv)
        assert _qdist(after, g[f"win{i}_after"]) <= 1e-14
        # window origin as csrc/pgo.cu computes it (trajectory.cpp:60-63,255)
        t_beg_d = float(t_traj[0]) + 1e-9 * float(t_traj[1])
        t0_ns = int(1e9 * (t_beg_d + idx_traj * dtk))
        dt_ns = int(1e9 * dtk)
        q = T.evaluate(order, after[idx_traj:], t0_ns, dt_ns, t)
        assert _qdist(q, g[f"win{i}_q"]) <= 1e-13, (i, _qdist(q, g[f"win{i}_q"]))
        # the oracle's wrapper: same value, same first knot, Jacobian blocks equal after the f32 cast
        r = oracle.spline_eval(order, after[idx_traj:], t0_ns, dt_ns, t[0] * 1_000_000_000 + t[1], want_J=True)
        assert r is not None
        qo, _, idx, J = r
        assert _qdist(qo, g[f"win{i}_q"]) <= 1e-13 and idx == int(g[f"win{i}_idx"])
        Jf = np.concatenate([J[k] for k in range(order)], axis=1).astype(np.float32)     # [3, 3*order], as trajectory.cpp:99-106
        ref = g[f"win{i}_J"]
        assert np.abs(Jf - ref).max() <= 2e-7 * max(1.0, float(np.abs(ref).max())), (i, np.abs(Jf - ref).max())


def test_live_against_reference_source(oracle):
    if not oracle.have_ref_traj():
        pytest.skip("oracle/_ref/libref_traj.so not built (reference tree absent)")
    rng = np.random.default_rng(8)
    for trial in range(30):
        order = int(rng.choice([2, 4]))
        K = int(rng.integers(order + 1, 16))
        knots = np.zeros((K, 4)); knots[0] = T.incremental_update(np.array([[0, 0, 0, 1.0]]), 0, rng.normal(0, 1, 3))[0]
        for i in range(1, K):
            knots[i] = T.incremental_update(knots[i - 1][None, :], 0, rng.normal(0, 0.1, 3))[0]
        dtk = float(rng.choice([0.02, 0.05, 0.1]))
        t_beg = float(rng.uniform(1e9, 1.7e9))
        n_seg = K - order + 1
        ts = t_beg + rng.uniform(0.001, n_seg - 0.001) * dtk
        t = (int(np.floor(ts)), int((ts - np.floor(ts)) * 1e9))
        q_ref, idx_ref, J_ref = oracle.ref1p_evaluate(order, t_beg, dtk, knots, t)
        q = T.evaluate(order, knots, int(1e9 * t_beg), int(1e9 * dtk), t)
        assert _qdist(q, q_ref) <= 1e-13, trial
