"""Host-side mirror of the reference's trajectory initialisation over the C ABI.

Reference surface mirrored:
  PoseGraphOptimizer::integrateAngVel        src/backend/pose_graph_optimizer.cpp:191-222
  Trajectory::generateCtrlPoses / fitCtrlPoses   src/backend/trajectory.cpp:112-212, 357-489
  Trajectory::evaluate / incrementalUpdate   src/backend/trajectory.cpp:86-110, 221-238, 329-355, 491-499
All arithmetic is host C++ inside libcmax_b200.so (csrc/traj_init.cu); this module only marshals pointers.
Stamps are (sec, nsec) pairs like ros::Time."""
import ctypes as C

import numpy as np

from . import _capi


def _stamp(t):
    return _capi.Stamp(int(t[0]), int(t[1]))


def _stamps(a):
    a = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, 2)
    return a, a.ctypes.data_as(C.POINTER(_capi.Stamp))


def stamp_from_sec(t):
    """ros::Time::fromSec: floor + round-to-nearest nanosecond."""
    sec = int(np.floor(t))
    nsec = int(round((t - sec) * 1e9))
    sec += nsec // 1000000000
    nsec %= 1000000000
    return sec, nsec


def integrate_ang_vel(pose_latest, ang_vel_prev, first_time_window, stamps, ang_vels):
    """pose_latest = ((sec, nsec), xyzw); ang_vel_prev = ((sec, nsec), w[3]).
    Returns (pose_stamps (n,2) uint32, poses (n,4), new ang_vel_prev)."""
    L = _capi.lib()
    st, stp = _stamps(stamps)
    w = np.ascontiguousarray(ang_vels, dtype=np.float64).reshape(-1, 3)
    m = len(st)
    q0 = np.ascontiguousarray(pose_latest[1], dtype=np.float64)
    prev_t = _stamp(ang_vel_prev[0])
    prev_w = np.array(ang_vel_prev[1], dtype=np.float64)
    out_t = np.zeros((max(m, 1), 2), np.uint32)
    out_q = np.zeros((max(m, 1), 4))
    n = C.c_int(0)
    _capi.check(L.cmaxb_traj_integrate_ang_vel(_stamp(pose_latest[0]), _capi.dptr(q0), C.byref(prev_t), _capi.dptr(prev_w),
                                               int(bool(first_time_window)), stp, _capi.dptr(w), m,
                                               out_t.ctypes.data_as(C.POINTER(_capi.Stamp)), _capi.dptr(out_q), C.byref(n)))
    return out_t[: n.value].copy(), out_q[: n.value].copy(), ((prev_t.sec, prev_t.nsec), prev_w)


def num_ctrl_poses(spline_order, t_beg, t_end, dt_knots):
    n = _capi.lib().cmaxb_traj_num_ctrl_poses(int(spline_order), _stamp(t_beg), _stamp(t_end), float(dt_knots))
    if n < 0:
        _capi.check(n)
    return n


def fit_ctrl_poses(spline_order, dt_knots, t_beg_sec, num_cps, pose_stamps, poses_xyzw):
    st, stp = _stamps(pose_stamps)
    q = np.ascontiguousarray(poses_xyzw, dtype=np.float64).reshape(-1, 4)
    out = np.zeros((num_cps, 4))
    _capi.check(_capi.lib().cmaxb_traj_fit_ctrl_poses(int(spline_order), float(dt_knots), float(t_beg_sec), int(num_cps), stp,
                                                       _capi.dptr(q), len(q), _capi.dptr(out)))
    return out


def generate_ctrl_poses(spline_order, dt_knots, pose_stamps, poses_xyzw, t_beg, t_end):
    """Trajectory::generateCtrlPoses(poses, t_beg, t_end): t_beg / t_end are (sec, nsec)."""
    n = num_ctrl_poses(spline_order, t_beg, t_end, dt_knots)
    return fit_ctrl_poses(spline_order, dt_knots, float(t_beg[0]) + 1e-9 * float(t_beg[1]), n, pose_stamps, poses_xyzw)


def evaluate(spline_order, knots_xyzw, t0_ns, dt_ns, t):
    k = np.ascontiguousarray(knots_xyzw, dtype=np.float64).reshape(-1, 4)
    out = np.zeros(4)
    _capi.check(_capi.lib().cmaxb_traj_evaluate(int(spline_order), _capi.dptr(k), len(k), int(t0_ns), int(dt_ns), _stamp(t), _capi.dptr(out)))
    return out


def incremental_update(knots_xyzw, idx_beg, x):
    k = np.array(knots_xyzw, dtype=np.float64).reshape(-1, 4)
    xx = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    _capi.check(_capi.lib().cmaxb_traj_incremental_update(_capi.dptr(k), len(k), int(idx_beg), _capi.dptr(xx)))
    return k
