"""TEST INFRASTRUCTURE.  Independent Python restatement of the back-end window pipeline
(PoseGraphOptimizer: src/backend/pose_graph_optimizer.cpp:72-354) composed from the oracle's pieces -- the REAL
Sophus / Eigen code for the angular-velocity integration and the control-pose fit (oracle/_ref), the CPU oracle's
back-end cost, the restated GSL loop (oracle/gsl_fr.py) and the oracle's map upkeep -- used to check the
library's C++ pipeline (csrc/pgo.cu).  Times are (sec, nsec) tuples with ros::Time / ros::Duration arithmetic."""
import math

import numpy as np

from . import gsl_fr
from . import oracle_py as O


def dur(d):
    s = math.floor(d)
    ns = int(round((d - s) * 1e9))
    s = int(s) + ns // 1000000000
    return s, ns % 1000000000


def t_add(t, d):
    s, ns = t[0] + d[0], t[1] + d[1]
    while ns >= 1000000000:
        ns -= 1000000000; s += 1
    while ns < 0:
        ns += 1000000000; s -= 1
    return s, ns


def t_sec(t):
    return float(t[0]) + 1e-9 * float(t[1])


def t_nsec(t):
    return t[0] * 1000000000 + t[1]


def dur_sec(a, b):
    s, ns = a[0] - b[0], a[1] - b[1]
    if ns < 0:
        ns += 1000000000; s -= 1
    return float(s) + 1e-9 * float(ns)


def _qmul(a, b):
    w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]
    x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1]
    y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2]
    z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0]
    q = np.array([x, y, z, w])
    return q / np.linalg.norm(q)


class PipelineOracle:
    def __init__(self, lut, SW, SH, PW, PH, order, dt_knots, win_size, win_stride, y_angle_deg=0.0, max_update_times=255,
                 min_num_ev=0.0, batch_size=100, sample_rate=1, blur_sigma=1.0):
        self.lut, self.SW, self.SH, self.PW, self.PH = lut, SW, SH, PW, PH
        self.N, self.dtk = order, dt_knots
        self.win_size, self.win_stride = dur(win_size), dur(win_stride)
        self.y_angle, self.max_upd, self.min_ev = y_angle_deg, max_update_times, min_num_ev
        self.bs, self.sr, self.sigma = batch_size, sample_rate, blur_sigma
        self.cp_stride = int(round(win_stride / dt_knots))
        self.init = False
        self.first = True
        self.count = 0
        self.av = {}
        self.knots = np.zeros((0, 4))
        self.idx_opt = 0
        self.IG = np.zeros((PH, PW), np.float32)
        self.times = np.zeros((PH, PW), np.uint8)

    def push(self, ts, w):
        ts = (int(ts[0]), int(ts[1]))
        if not self.init:
            self.t_win_beg = ts
            self.t_win_end = t_add(ts, self.win_size)
            self.t_av_beg, self.t_av_end = self.t_win_beg, self.t_win_end
            self.traj_t_beg, self.traj_t_beg_ns = t_sec(ts), t_nsec(ts)
            self.traj_dt_ns = int(1e9 * self.dtk)
            self.prev = (ts, np.array(w, dtype=np.float64))
            th = self.y_angle * math.pi / 180
            self.latest = (ts, np.array([0.0, math.sin(th / 2), 0.0, math.cos(th / 2)]))
            self.init = True
        self.av.setdefault(ts, np.array(w, dtype=np.float64))

    def _traj_eval(self, t):
        r = O.spline_eval(self.N, self.knots, self.traj_t_beg_ns, self.traj_dt_ns, t_nsec(t), want_J=False)
        assert r is not None, "time outside the spline"
        return r[0]

    def process(self, events):
        rep = {}
        keys = sorted(self.av)
        sub = [k for k in keys if k > self.t_av_beg and k < self.t_av_end]        # upper_bound(beg) .. lower_bound(end)
        for k in [k for k in keys if k < self.t_av_end]:
            if k not in sub:
                del self.av[k]
        ws = np.array([self.av.pop(k) for k in sub]).reshape(-1, 3)
        st = np.array(sub, dtype=np.uint32).reshape(-1, 2)
        pst, pq, ps, pw = O.ref_integrate_ang_vel(self.latest[0], self.latest[1], self.prev[0], self.prev[1], self.first, st, ws)
        self.prev = ((int(ps[0]), int(ps[1])), pw)
        num = int(round(dur_sec(self.t_av_end, self.t_av_beg) / self.dtk)) + (3 if self.N == 4 else 1)
        ctrl = O.ref_fit_ctrl_poses(self.N, self.dtk, t_sec(self.t_av_beg), num, pst, pq)
        if self.first:
            self.idx_opt = 3 if self.N == 4 else 1
            self.first = False
        else:
            ctrl = ctrl[(3 if self.N == 4 else 1):]
        self.knots = np.concatenate([self.knots, ctrl])
        size = len(self.knots)
        idx_traj = self.count * self.cp_stride
        self.idx_opt = max(idx_traj, self.idx_opt)
        n_opt = size - self.idx_opt
        tnext = t_add(self.t_win_beg, self.win_stride)
        rep.update(n_ctrl_poses=size, idx_cp_traj_beg=idx_traj, idx_cp_opt_beg=self.idx_opt, num_cp_opt=n_opt, optimized=0)
        if len(events) > self.min_ev and n_opt > 0:
            t0_ns = int(1e9 * (self.traj_t_beg + idx_traj * self.dtk))
            dt_ns = int(1e9 * self.dtk)
            kn = self.knots[idx_traj:].copy()
            n_fixed = self.idx_opt - idx_traj
            mk = lambda alpha: O.be_args(events, self.lut, self.SW, self.SH, self.PW, self.PH, kn, t0_ns, dt_ns, self.N, n_fixed, tnext,
                                         self.IG.copy(), alpha, self.bs, self.sr, self.sigma)
            # first evaluation of the window: IGp <- IG, alpha from IL at x = 0 (updateAlpha)
            r0 = O.be_eval(mk(0.0), np.zeros(3 * n_opt), False, images=True)
            alpha = O.update_alpha(self.IG, r0["il_old"] + r0["il_new"])
            a = mk(alpha)
            last = {}

            def f(x):
                last["x"] = np.array(x)
                return -O.be_eval(a, x, False)["contrast"]

            def fdf(x):
                last["x"] = np.array(x)
                r = O.be_eval(a, x, True)
                return -r["contrast"], -r["grad"]

            x, stt = gsl_fr.minimize_fr(f, fdf, np.zeros(3 * n_opt), line_tol=0.1, epsabs_grad=1e-4)
            rep.update(opt=stt, alpha=alpha, optimized=1, x=x, x_last=last["x"])
            for i in range(n_opt):                                              # incrementalUpdate: exp(x_i) * K_i
                self.knots[self.idx_opt + i] = _qmul(_qexp(x[3 * i:3 * i + 3]), self.knots[self.idx_opt + i])
            il_old = O.be_eval(a, last["x"], False, images=True)["il_old"]
            O.update_ig(self.IG, il_old, self.times, self.max_upd)
            t, marks = self.t_win_beg, 0
            while t < tnext:
                O.set_update_times(self.lut, self.SW, self.SH, self.PW, self.PH, self._traj_eval(t), 3, self.times)
                t = t_add(t, dur(0.05))
                marks += 1
            rep["n_fov_marks"] = marks
        tl = t_add(self.t_win_end, (-dur(1e-6)[0], -dur(1e-6)[1]))
        self.latest = (tl, self._traj_eval(tl))
        rep["pose_latest"] = self.latest
        self.t_win_beg = t_add(self.t_win_beg, self.win_stride)
        self.t_av_beg = self.t_win_end
        self.t_win_end = t_add(self.t_win_end, self.win_stride)
        self.t_av_end = self.t_win_end
        self.count += 1
        return rep


def _qexp(w):
    w = np.asarray(w, dtype=np.float64)
    th2 = float(w @ w)
    if th2 < 1e-20:
        im, re = 0.5 - th2 / 48.0, 1.0 - th2 / 8.0
    else:
        th = math.sqrt(th2)
        im, re = math.sin(0.5 * th) / th, math.cos(0.5 * th)
    return np.array([im * w[0], im * w[1], im * w[2], re])
