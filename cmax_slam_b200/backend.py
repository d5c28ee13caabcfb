"""Host-side mirror of the reference's back-end cost interface over the C ABI.

Reference surface mirrored:
  EventWarper::computeImageOfWarpedEvents        src/backend/event_pano_warper.cpp:167-231
  cmax_slam::computeContrast (global)            src/backend/global_focus_funcs.cpp:52-80
  global_contrast_fdf / _f / _df (GSL callbacks) src/backend/global_optim_contrast_gsl_analytical.cpp:17-81
  PoseGraphOptimizer::copyAndUpdateTraj           src/backend/pose_graph_optimizer.cpp:239-242
All arithmetic happens in libcmax_b200.so on the GPU; this module only marshals pointers.
"""
import ctypes as C
import math

import numpy as np

from . import _capi
from ._capi import GRAD_ADJOINT, GRAD_DENSE, CmaxbError  # noqa: F401


class EventWarperCMax:
    """Device-resident back-end contrast functor for one sliding window at a time."""

    def __init__(self, sensor_width, sensor_height, lut_xyz, pano_width, pano_height, blur_sigma=1.0,
                 event_batch_size=100, event_sample_rate=1, spline_order=2, contrast_measure=0,
                 grad_mode=GRAD_ADJOINT, device=0, stream=None):
        self._L = _capi.lib()
        lut = np.ascontiguousarray(lut_xyz, dtype=np.float64).reshape(-1, 3)
        if lut.shape[0] != sensor_width * sensor_height:
            raise ValueError("lut_xyz must hold sensor_width*sensor_height bearing vectors")
        cfg = _capi.BeCfg(sensor_width, sensor_height, lut.ctypes.data, pano_width, pano_height, float(blur_sigma),
                          int(event_batch_size), int(event_sample_rate), int(spline_order), int(contrast_measure),
                          int(grad_mode), int(device), None if stream is None else C.c_void_p(int(stream)))
        h = C.c_void_p()
        _capi.check(self._L.cmaxb_be_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.pano_width, self.pano_height = pano_width, pano_height
        # radius of cv::GaussianBlur(Size(0,0), sigma) on CV_32F, as csrc/capi_common.cuh make_taps computes it
        self.blur_radius = ((int(round(float(blur_sigma) * 8 + 1)) | 1) // 2) if blur_sigma > 0 else 0
        self.spline_order = spline_order
        self.n_events = 0
        self.n_params = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.cmaxb_be_destroy(self._h)
            self._h = None

    __del__ = close

    def set_window(self, events, knots_xyzw, t0_ns, dt_ns, n_fixed, t_next_win_beg, IGp=None, alpha=float("nan")):
        """events: 16-byte dvs_msgs::Event records of the window; knots_xyzw: (K,4) control poses of the
        temporary trajectory; t_next_win_beg: (sec, nsec); alpha=NaN => updateAlpha on the first eval."""
        ev = np.ascontiguousarray(events)
        if ev.dtype.itemsize != 16:
            raise ValueError("events must be 16-byte dvs_msgs::Event records")
        kn = np.ascontiguousarray(knots_xyzw, dtype=np.float64).reshape(-1, 4)
        igp = None if IGp is None else np.ascontiguousarray(IGp, dtype=np.float32)
        if igp is not None and igp.size != self.pano_width * self.pano_height:
            raise ValueError("IGp must be a pano_height x pano_width float image")
        w = _capi.BeWindow(ev.ctypes.data, len(ev), kn.ctypes.data, kn.shape[0], int(t0_ns), int(dt_ns), int(n_fixed),
                           int(t_next_win_beg[0]), int(t_next_win_beg[1]), None if igp is None else igp.ctypes.data,
                           float(alpha))
        self._keep = (ev, kn, igp)
        _capi.check(self._L.cmaxb_be_set_window(self._h, C.byref(w)))
        self.n_events = len(ev)
        self.n_knots = kn.shape[0]
        self.n_params = 3 * (kn.shape[0] - n_fixed)

    def _x(self, x):
        if x is None:
            return None, 0
        xx = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        return xx, xx.size

    def eval(self, x=None, want_grad=True):
        """(+contrast, +gradient[3*K_opt] or None) at incremental rotation vectors x (None = zeros)."""
        xx, n = self._x(x)
        c = C.c_double()
        g = np.zeros(max(self.n_params, 1))
        _capi.check(self._L.cmaxb_be_eval(self._h, None if xx is None else _capi.dptr(xx), n, C.byref(c),
                                          _capi.dptr(g) if want_grad else None))
        return c.value, (g[: self.n_params] if want_grad else None)

    # -- event-sharded evaluation (one window split by time across GPUs) -------------------------------
    def eval_begin(self, x=None, want_grad=True):
        """Poses + scatter of this rank's events; IL assembled into the plane returned by il_plane()."""
        xx, n = self._x(x)
        _capi.check(self._L.cmaxb_be_eval_begin(self._h, None if xx is None else _capi.dptr(xx), n, int(want_grad)))
        self._split_grad = bool(want_grad)

    def il_plane(self):
        """(device pointer, element count) of the float32 IL plane to all-reduce across ranks."""
        ptr, cnt = C.c_void_p(), C.c_size_t()
        _capi.check(self._L.cmaxb_be_il_plane(self._h, C.byref(ptr), C.byref(cnt)))
        return ptr.value, cnt.value

    def il_plane_tensor(self):
        """The IL plane as a torch CUDA tensor VIEW of the library's buffer (for torch.distributed collectives)."""
        import torch
        ptr, cnt = self.il_plane()

        class _View:
            __cuda_array_interface__ = {"shape": (cnt,), "typestr": "<f4", "data": (ptr, False), "version": 2}

        return torch.as_tensor(_View(), device="cuda")

    def eval_end(self):
        """(contrast, this rank's PARTIAL gradient or None) after the caller summed the IL planes."""
        c = C.c_double()
        g = np.zeros(max(self.n_params, 1))
        _capi.check(self._L.cmaxb_be_eval_end(self._h, C.byref(c), _capi.dptr(g) if self._split_grad else None))
        return c.value, (g[: self.n_params] if self._split_grad else None)

    def eval_end_launch(self):
        _capi.check(self._L.cmaxb_be_eval_end_launch(self._h, int(self._split_grad)))

    def grad_tensor(self):
        """This rank's partial gradient as a torch CUDA tensor VIEW of the library's buffer (float64, 3*K_opt)."""
        import torch
        ptr, cnt = C.c_void_p(), C.c_size_t()
        _capi.check(self._L.cmaxb_be_grad_device(self._h, C.byref(ptr), C.byref(cnt)))
        if cnt.value == 0:
            return None

        class _View:
            __cuda_array_interface__ = {"shape": (cnt.value,), "typestr": "<f8", "data": (ptr.value, False), "version": 2}

        return torch.as_tensor(_View(), device="cuda")

    def eval_end_fetch(self):
        c = C.c_double()
        g = np.zeros(max(self.n_params, 1))
        _capi.check(self._L.cmaxb_be_eval_end_fetch(self._h, C.byref(c), _capi.dptr(g) if self._split_grad else None))
        return c.value, (g[: self.n_params] if self._split_grad else None)

    # -- the same with the image phases sharded by row band (cmaxb_be_shard_*) ------------------------------
    @staticmethod
    def _dev_tensor(ptr, count, typestr):
        import torch

        class _View:
            __cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2}

        return torch.as_tensor(_View(), device="cuda")

    def shard_begin(self, x, want_grad, world, rank):
        """Poses + scatter of this rank's events.  Returns (send, recv) float32 tensor views: `send` holds `world` extended
        bands, `recv` receives this rank's band summed over ranks (reduce_scatter_tensor(recv, send))."""
        xx, n = self._x(x)
        send, recv, chunk = C.c_void_p(), C.c_void_p(), C.c_size_t()
        _capi.check(self._L.cmaxb_be_shard_begin(self._h, None if xx is None else _capi.dptr(xx), n, int(want_grad), int(world),
                                                 int(rank), C.byref(send), C.byref(recv), C.byref(chunk)))
        self._split_grad = bool(want_grad)
        return (self._dev_tensor(send.value, chunk.value * world, "<f4"), self._dev_tensor(recv.value, chunk.value, "<f4"))

    def shard_image(self):
        """Blur of the band; returns the (S1, S2) float64 tensor view to all-reduce."""
        p = C.c_void_p()
        _capi.check(self._L.cmaxb_be_shard_image(self._h, C.byref(p)))
        return self._dev_tensor(p.value, 2, "<f8")

    def shard_adjoint(self, world):
        """Contrast + mean; gradient evaluations: adjoint blur of the band.  Returns (g_own, g_full) views or (None, None)."""
        own, full, cnt = C.c_void_p(), C.c_void_p(), C.c_size_t()
        _capi.check(self._L.cmaxb_be_shard_adjoint(self._h, C.byref(own), C.byref(full), C.byref(cnt)))
        if not own.value:
            return None, None
        return self._dev_tensor(own.value, cnt.value, "<f4"), self._dev_tensor(full.value, cnt.value * world, "<f4")

    def shard_gather(self):
        _capi.check(self._L.cmaxb_be_shard_gather(self._h))

    # -- exchange by the kernels over peer memory (cmaxb_be_exchange_* / cmaxb_be_xeval) ----------------------
    def exchange_connect(self, group=None):
        """Collective over `group` (torch.distributed, any backend: only the 64-byte IPC handles travel through it)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        h = C.create_string_buffer(64)
        _capi.check(self._L.cmaxb_be_exchange_init(self._h, world, rank, h))
        handles = [None] * world
        dist.all_gather_object(handles, h.raw, group=group)
        _capi.check(self._L.cmaxb_be_exchange_connect(self._h, b"".join(handles)))
        dist.barrier(group)
        self._xworld = world

    def exchange_close(self):
        _capi.check(self._L.cmaxb_be_exchange_close(self._h))
        self._xworld = 0

    def exchange_stats(self):
        """(panorama tiles this rank's events touched in the last xeval, tiles of the panorama)"""
        t = (C.c_int64 * 2)()
        _capi.check(self._L.cmaxb_be_exchange_stats(self._h, t))
        return int(t[0]), int(t[1])

    def xeval(self, x=None, want_grad=True):
        """Collective evaluation of the time-sharded window: (contrast, gradient) of the WHOLE window on every rank."""
        xx, n = self._x(x)
        c = C.c_double()
        g = np.zeros(max(self.n_params, 1))
        _capi.check(self._L.cmaxb_be_xeval(self._h, None if xx is None else _capi.dptr(xx), n, C.byref(c),
                                           _capi.dptr(g) if want_grad else None))
        return c.value, (g[: self.n_params] if want_grad else None)

    # -- device-resident global map (event_pano_warper.cpp:81-132) --------------------------------------------
    def resetIG(self):
        _capi.check(self._L.cmaxb_be_map_reset(self._h))

    def setMap(self, IG=None, visit_counts=None):
        ig = None if IG is None else np.ascontiguousarray(IG, dtype=np.float32)
        vc = None if visit_counts is None else np.ascontiguousarray(visit_counts, dtype=np.uint8)
        _capi.check(self._L.cmaxb_be_map_set(self._h, None if ig is None else C.c_void_p(ig.ctypes.data),
                                             None if vc is None else C.c_void_p(vc.ctypes.data)))

    def getIG(self):
        """(IG_ float32 pano, IG_update_times_map_ uint8 pano)"""
        ig = np.empty((self.pano_height, self.pano_width), np.float32)
        vc = np.empty((self.pano_height, self.pano_width), np.uint8)
        _capi.check(self._L.cmaxb_be_map_get(self._h, C.c_void_p(ig.ctypes.data), C.c_void_p(vc.ctypes.data)))
        return ig, vc

    def updateIGp(self, alpha=float("nan")):
        """IGp <- IG on the device (after set_window); alpha NaN => updateAlpha on the first evaluation."""
        _capi.check(self._L.cmaxb_be_map_use_as_igp(self._h, float(alpha)))

    def updateIG(self, x_opt=None, max_update_times=10):
        xx, n = self._x(x_opt)
        _capi.check(self._L.cmaxb_be_map_update(self._h, None if xx is None else _capi.dptr(xx), n, int(max_update_times)))

    def setUpdateTimesIG(self, rots_xyzw, radius=3):
        q = np.ascontiguousarray(rots_xyzw, dtype=np.float64).reshape(-1, 4)
        _capi.check(self._L.cmaxb_be_map_mark_fov(self._h, _capi.dptr(q), q.shape[0], int(radius)))

    def setupProblemAndOptimize(self, x0=None, params=None):
        """PoseGraphOptimizer::setupProblemAndOptimize_gsl (global_optim_contrast_gsl.cpp:15-145) without GSL.
        Returns (x_opt, stats dict); the caller applies x_opt with incrementalUpdate (trajectory.cpp:221-238)."""
        n = self.n_params
        xs = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).reshape(n)
        out = np.zeros(max(n, 1))
        res = _capi.OptResult()
        prm = None if params is None else C.byref(_capi.OptParams(*params))
        _capi.check(self._L.cmaxb_be_optimize(self._h, None if xs is None else _capi.dptr(xs), n, prm, _capi.dptr(out), C.byref(res)))
        return out[:n], {k: getattr(res, k) for k, _ in _capi.OptResult._fields_}

    @property
    def alpha(self):
        a = C.c_double()
        _capi.check(self._L.cmaxb_be_get_alpha(self._h, C.byref(a)))
        return a.value

    def computeImageOfWarpedEvents(self, x=None, blurred=True):
        xx, n = self._x(x)
        out = np.empty((self.pano_height, self.pano_width), np.float32)
        _capi.check(self._L.cmaxb_be_get_iwe(self._h, None if xx is None else _capi.dptr(xx), n, int(blurred),
                                             C.c_void_p(out.ctypes.data)))
        return out

    def local_iwe(self, x=None):
        """(IL_old_, IL_new_) -- what updateIG consumes (event_pano_warper.cpp:109-126)."""
        xx, n = self._x(x)
        a = np.empty((self.pano_height, self.pano_width), np.float32)
        b = np.empty_like(a)
        _capi.check(self._L.cmaxb_be_get_il(self._h, None if xx is None else _capi.dptr(xx), n,
                                            C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data)))
        return a, b

    def derivative_bands(self, x=None, blurred=True):
        xx, n = self._x(x)
        out = np.empty((self.n_params, self.pano_height, self.pano_width), np.float32)
        _capi.check(self._L.cmaxb_be_get_bands(self._h, None if xx is None else _capi.dptr(xx), n, int(blurred),
                                               C.c_void_p(out.ctypes.data)))
        return out

    def warped_cells(self, x=None):
        xx, n = self._x(x)
        out = np.empty(self.n_events, np.int32)
        _capi.check(self._L.cmaxb_be_get_cells(self._h, None if xx is None else _capi.dptr(xx), n, C.c_void_p(out.ctypes.data)))
        return out

    def batch_poses(self, x=None):
        """Per-batch (R [nb,3,3] f64, Jk [nb,3,3*order] f32, idx_cp_beg [nb]) of the device So3Spline."""
        xx, n = self._x(x)
        nb = C.c_int64()
        xp = None if xx is None else _capi.dptr(xx)
        _capi.check(self._L.cmaxb_be_get_poses(self._h, xp, n, C.byref(nb), None, None, None, 0))
        R = np.zeros((nb.value, 3, 3))
        Jk = np.zeros((nb.value, 3, 3 * self.spline_order), np.float32)
        idx = np.zeros(nb.value, np.int32)
        _capi.check(self._L.cmaxb_be_get_poses(self._h, xp, n, C.byref(nb), C.c_void_p(R.ctypes.data),
                                               C.c_void_p(Jk.ctypes.data), C.c_void_p(idx.ctypes.data), nb.value))
        return R, Jk, idx

    def profile(self, enable=True):
        _capi.check(self._L.cmaxb_be_profile(self._h, int(enable)))

    def kernel_times(self):
        ms = np.zeros(_capi.K_COUNT)
        n = np.zeros(_capi.K_COUNT, np.uint64)
        _capi.check(self._L.cmaxb_be_kernel_times(self._h, _capi.dptr(ms), n.ctypes.data_as(C.POINTER(C.c_uint64))))
        return {_capi.K_NAMES[i]: (float(ms[i]), int(n[i])) for i in range(_capi.K_COUNT) if n[i] > 0}


# GSL callback triple with the reference's semantics: -contrast / -gradient
# (global_optim_contrast_gsl_analytical.cpp:56-66); want_df=False is the `df == nullptr` path.
def global_contrast_fdf(v, warper, want_df=True):
    c, g = warper.eval(v, want_grad=want_df)
    return -c, (None if g is None else -g)


def global_contrast_f(v, warper):
    return global_contrast_fdf(v, warper, want_df=False)[0]


def global_contrast_df(v, warper):
    return global_contrast_fdf(v, warper, want_df=True)[1]


class PoseGraphOptimizerCMax:
    """Host-side mirror of PoseGraphOptimizer's window pipeline (src/backend/pose_graph_optimizer.cpp:72-354) over the
    C ABI (cmaxb_pgo_*): pushAngVel / isReadyFrontendPoses / processTimeWindow (+ getAngVelSubset, integrateAngVel,
    setUpdateTimesIG, slideWindow).  `warper` is an EventWarperCMax whose spline order matches."""

    def __init__(self, warper, spline_order, dt_knots, time_window_size, sliding_window_stride, y_angle_deg=0.0,
                 max_update_times=255, min_num_ev_per_win=0.0, opt_params=None):
        self._L = _capi.lib()
        self.warper = warper
        cfg = _capi.PgoCfg(int(spline_order), float(dt_knots), float(time_window_size), float(sliding_window_stride),
                           float(y_angle_deg), int(max_update_times), float(min_num_ev_per_win),
                           0 if opt_params is None else 1, _capi.OptParams(*(opt_params or (0.1, 0.1, 50, 1e-4, 1e-4))))
        h = C.c_void_p()
        _capi.check(self._L.cmaxb_pgo_create(C.byref(cfg), warper._h if warper is not None else None, C.byref(h)))
        self._p = h

    def close(self):
        if getattr(self, "_p", None):
            self._L.cmaxb_pgo_destroy(self._p)
            self._p = None

    __del__ = close

    def pushAngVel(self, ts, ang_vel):
        w = np.ascontiguousarray(ang_vel, dtype=np.float64).reshape(3)
        _capi.check(self._L.cmaxb_pgo_push_ang_vel(self._p, _capi.Stamp(int(ts[0]), int(ts[1])), _capi.dptr(w)))

    def window(self):
        """((sec, nsec) t_win_beg, (sec, nsec) t_win_end, ang_vel_ready)"""
        a, b, r = _capi.Stamp(), _capi.Stamp(), C.c_int(0)
        _capi.check(self._L.cmaxb_pgo_window(self._p, C.byref(a), C.byref(b), C.byref(r)))
        return (a.sec, a.nsec), (b.sec, b.nsec), bool(r.value)

    def processTimeWindow(self, events):
        ev = np.ascontiguousarray(events)
        rep = _capi.PgoReport()
        _capi.check(self._L.cmaxb_pgo_process_window(self._p, C.c_void_p(ev.ctypes.data), len(ev), C.byref(rep)))
        out = {k: getattr(rep, k) for k in ("window", "n_ang_vel", "n_frontend_poses", "n_ctrl_poses", "idx_cp_traj_beg",
                                            "idx_cp_opt_beg", "num_cp_opt", "optimized", "alpha", "n_fov_marks")}
        out["opt"] = {k: getattr(rep.opt, k) for k, _ in _capi.OptResult._fields_}
        out["pose_latest"] = ((rep.pose_latest_t.sec, rep.pose_latest_t.nsec), np.array(rep.pose_latest_xyzw[:]))
        out["t_win"] = ((rep.t_win_beg.sec, rep.t_win_beg.nsec), (rep.t_win_end.sec, rep.t_win_end.nsec))
        return out

    def ctrl_poses(self):
        n, t0, dt = C.c_int(0), C.c_int64(0), C.c_int64(0)
        _capi.check(self._L.cmaxb_pgo_get_ctrl_poses(self._p, None, 0, C.byref(n), C.byref(t0), C.byref(dt)))
        q = np.zeros((n.value, 4))
        if n.value:
            _capi.check(self._L.cmaxb_pgo_get_ctrl_poses(self._p, _capi.dptr(q), n.value, C.byref(n), C.byref(t0), C.byref(dt)))
        return q, t0.value, dt.value
