"""ctypes binding of libcmax_b200.so (include/cmax_b200.h).  No CPU fallback: if the library is
missing or there is no CUDA device the product fails loudly."""
import ctypes as C
import os

import numpy as np

from . import build as _build

K_NAMES = ["zero", "fe_scatter", "fe_gather", "blur_reduce", "adjoint_blur", "be_poses", "be_scatter",
           "be_gather", "be_grad_reduce", "misc", "fe_eval_fused", "be_eval_fused", "be_x_push", "be_x_sums", "be_x_pull",
           "be_x_grad"]
K_COUNT = len(K_NAMES)

GRAD_DENSE, GRAD_ADJOINT = 0, 1
CONTRAST_VARIANCE, CONTRAST_MEAN_SQUARE = 0, 1

ERR = {-1: "CMAXB_ERR_INVALID", -2: "CMAXB_ERR_CUDA", -3: "CMAXB_ERR_EVENT_RANGE", -4: "CMAXB_ERR_TIME_ORDER",
       -5: "CMAXB_ERR_SPLINE_RANGE", -6: "CMAXB_ERR_STATE"}


class CmaxbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR.get(code, code)}: {msg}")
        self.code = code


class FeCfg(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("lut_xyz", C.c_void_p), ("blur_sigma", C.c_double), ("batch_size", C.c_int32),
                ("contrast_measure", C.c_int32), ("grad_mode", C.c_int32), ("device", C.c_int32),
                ("stream", C.c_void_p), ("max_hypotheses", C.c_int32), ("lanes", C.c_int32), ("packet_slots", C.c_int32)]


class BeCfg(C.Structure):
    _fields_ = [("sensor_width", C.c_int32), ("sensor_height", C.c_int32), ("lut_xyz", C.c_void_p),
                ("pano_width", C.c_int32), ("pano_height", C.c_int32), ("blur_sigma", C.c_double),
                ("batch_size", C.c_int32), ("event_sample_rate", C.c_int32), ("spline_order", C.c_int32),
                ("contrast_measure", C.c_int32), ("grad_mode", C.c_int32), ("device", C.c_int32),
                ("stream", C.c_void_p)]


class OptParams(C.Structure):
    _fields_ = [("initial_step", C.c_double), ("line_tol", C.c_double), ("max_iterations", C.c_int32),
                ("epsabs_grad", C.c_double), ("tolfun", C.c_double), ("fused_trials", C.c_int32)]


class OptResult(C.Structure):
    _fields_ = [("cost_initial", C.c_double), ("cost_final", C.c_double), ("iterations", C.c_int32),
                ("f_evals", C.c_int32), ("g_evals", C.c_int32), ("stop_reason", C.c_int32), ("cost_launches", C.c_int32)]


COST_F = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int, C.c_void_p, C.POINTER(C.c_double))
COST_FDF = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double))


def optimize_callback(f, fdf, x0, params):
    """cmaxb_optimize_callback with Python callables: f(x)->cost, fdf(x)->(cost, grad); params = (step, line_tol, max_iter,
    epsabs_grad, tolfun).  Returns (x, stats dict)."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = len(x0)

    def _f(xp, nn, user, out):
        out[0] = float(f(np.array([xp[i] for i in range(nn)])))
        return 0

    def _fdf(xp, nn, user, out, g):
        v, gr = fdf(np.array([xp[i] for i in range(nn)]))
        out[0] = float(v)
        for i in range(nn):
            g[i] = float(gr[i])
        return 0

    cf, cfdf = COST_F(_f), COST_FDF(_fdf)
    xo = np.zeros(n)
    res = OptResult()
    prm = OptParams(*params)
    check(lib().cmaxb_optimize_callback(n, dptr(x0), cf, cfdf, None, C.byref(prm), dptr(xo), C.byref(res)))
    return xo, {k: getattr(res, k) for k, _ in OptResult._fields_}


class Stamp(C.Structure):
    _fields_ = [("sec", C.c_uint32), ("nsec", C.c_uint32)]


class CameraInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("K", C.c_double * 9), ("D", C.c_double * 12), ("n_D", C.c_int32),
                ("R", C.c_double * 9), ("P", C.c_double * 12)]


class StreamCfg(C.Structure):
    _fields_ = [("dt_ang_vel", C.c_double), ("num_events_per_packet", C.c_int32), ("event_sample_rate", C.c_int32)]


class PgoCfg(C.Structure):
    _fields_ = [("spline_order", C.c_int32), ("dt_knots", C.c_double), ("time_window_size", C.c_double),
                ("sliding_window_stride", C.c_double), ("y_angle_deg", C.c_double), ("max_update_times", C.c_int32),
                ("min_num_ev_per_win", C.c_double), ("use_opt_params", C.c_int32), ("opt_params", OptParams)]


class PgoReport(C.Structure):
    _fields_ = [("window", C.c_int32), ("t_win_beg", Stamp), ("t_win_end", Stamp), ("n_ang_vel", C.c_int32),
                ("n_frontend_poses", C.c_int32), ("n_ctrl_poses", C.c_int32), ("idx_cp_traj_beg", C.c_int32),
                ("idx_cp_opt_beg", C.c_int32), ("num_cp_opt", C.c_int32), ("optimized", C.c_int32), ("opt", OptResult),
                ("alpha", C.c_double), ("n_fov_marks", C.c_int32), ("pose_latest_t", Stamp), ("pose_latest_xyzw", C.c_double * 4)]


class BeWindow(C.Structure):
    _fields_ = [("events", C.c_void_p), ("n_events", C.c_size_t), ("knots_xyzw", C.c_void_p),
                ("n_knots", C.c_int32), ("t0_ns", C.c_int64), ("dt_ns", C.c_int64), ("n_fixed", C.c_int32),
                ("tnext_sec", C.c_uint32), ("tnext_nsec", C.c_uint32), ("IGp", C.c_void_p), ("alpha", C.c_double)]


# every symbol include/cmax_b200.h declares
EXPORTS = [
    "cmaxb_fe_create", "cmaxb_fe_destroy", "cmaxb_fe_set_packet", "cmaxb_fe_set_packet_async", "cmaxb_fe_set_packet_view",
    "cmaxb_fe_select_packet", "cmaxb_fe_lanes_fork", "cmaxb_fe_lanes_join", "cmaxb_fe_launch_info", "cmaxb_fe_eval", "cmaxb_fe_eval_batch",
    "cmaxb_fe_eval_launch", "cmaxb_fe_eval_fetch", "cmaxb_fe_exchange_init", "cmaxb_fe_exchange_connect", "cmaxb_fe_exchange_close",
    "cmaxb_fe_eval_fetch_all", "cmaxb_fe_set_result_mirror", "cmaxb_fe_get_iwe", "cmaxb_fe_get_deriv", "cmaxb_fe_get_cells",
    "cmaxb_be_create", "cmaxb_be_destroy", "cmaxb_be_set_window", "cmaxb_be_eval", "cmaxb_be_eval_begin", "cmaxb_be_il_plane", "cmaxb_be_eval_end", "cmaxb_be_eval_end_launch", "cmaxb_be_grad_device", "cmaxb_be_eval_end_fetch", "cmaxb_be_shard_begin", "cmaxb_be_shard_image", "cmaxb_be_shard_adjoint", "cmaxb_be_shard_gather", "cmaxb_be_exchange_init", "cmaxb_be_exchange_connect", "cmaxb_be_exchange_close", "cmaxb_be_xeval", "cmaxb_be_exchange_stats", "cmaxb_be_get_alpha",
    "cmaxb_be_get_il", "cmaxb_be_get_iwe", "cmaxb_be_get_bands", "cmaxb_be_get_cells", "cmaxb_be_get_poses",
    "cmaxb_be_map_reset", "cmaxb_be_map_set", "cmaxb_be_map_get", "cmaxb_be_map_use_as_igp", "cmaxb_be_map_update",
    "cmaxb_be_map_mark_fov",
    "cmaxb_fe_optimize", "cmaxb_be_optimize", "cmaxb_optimize_callback",
    "cmaxb_traj_integrate_ang_vel", "cmaxb_traj_num_ctrl_poses", "cmaxb_traj_fit_ctrl_poses", "cmaxb_traj_evaluate",
    "cmaxb_traj_incremental_update",
    "cmaxb_precompute_bearing_vectors",
    "cmaxb_stream_create", "cmaxb_stream_destroy", "cmaxb_stream_push", "cmaxb_stream_next_packet", "cmaxb_stream_window_events",
    "cmaxb_stream_state", "cmaxb_stream_attach_device", "cmaxb_stream_push_ex", "cmaxb_stream_next_packet_device", "cmaxb_stream_released", "cmaxb_stream_wait_copied",
    "cmaxb_pgo_create", "cmaxb_pgo_destroy", "cmaxb_pgo_push_ang_vel", "cmaxb_pgo_window", "cmaxb_pgo_process_window",
    "cmaxb_pgo_get_ctrl_poses", "cmaxb_be_last_eval_x",
    "cmaxb_last_error", "cmaxb_version", "cmaxb_device_count", "cmaxb_launch_count",
    "cmaxb_fe_profile", "cmaxb_fe_kernel_times", "cmaxb_fe_cta_times", "cmaxb_fe_phase_times", "cmaxb_be_profile", "cmaxb_be_kernel_times", "cmaxb_kernel_name",
]

_lib = None


def lib():
    """Load (building if needed) libcmax_b200.so.  Raises if it cannot be had -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    L = C.CDLL(path)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.cmaxb_fe_create.argtypes = [C.POINTER(FeCfg), C.POINTER(vp)]
    L.cmaxb_fe_destroy.argtypes = [vp]
    L.cmaxb_fe_destroy.restype = None
    L.cmaxb_fe_set_packet.argtypes = [vp, vp, C.c_size_t, C.c_double]
    L.cmaxb_fe_set_packet_async.argtypes = [vp, vp, C.c_size_t, C.c_double]
    L.cmaxb_fe_set_packet_view.argtypes = [vp, vp, C.c_size_t, C.c_double]
    L.cmaxb_fe_select_packet.argtypes = [vp, C.c_int]
    L.cmaxb_fe_lanes_fork.argtypes = [vp]
    L.cmaxb_fe_lanes_join.argtypes = [vp]
    L.cmaxb_fe_launch_info.argtypes = [vp, ip]
    L.cmaxb_fe_eval.argtypes = [vp, dp, dp, dp]
    L.cmaxb_fe_eval_batch.argtypes = [vp, dp, C.c_int, dp, dp]
    L.cmaxb_fe_eval_launch.argtypes = [vp, dp, C.c_int, C.c_int]
    L.cmaxb_fe_eval_fetch.argtypes = [vp, dp, dp]
    L.cmaxb_fe_set_result_mirror.argtypes = [vp, vp]
    L.cmaxb_fe_exchange_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.cmaxb_fe_exchange_connect.argtypes = [vp, vp, vp]
    L.cmaxb_fe_exchange_close.argtypes = [vp]
    L.cmaxb_fe_eval_fetch_all.argtypes = [vp, dp]
    L.cmaxb_fe_get_iwe.argtypes = [vp, dp, C.c_int, vp]
    L.cmaxb_fe_get_deriv.argtypes = [vp, dp, C.c_int, vp]
    L.cmaxb_fe_get_cells.argtypes = [vp, dp, vp]
    L.cmaxb_be_map_reset.argtypes = [vp]
    L.cmaxb_be_map_set.argtypes = [vp, vp, vp]
    L.cmaxb_be_map_get.argtypes = [vp, vp, vp]
    L.cmaxb_be_map_use_as_igp.argtypes = [vp, C.c_double]
    L.cmaxb_be_map_update.argtypes = [vp, dp, C.c_int, C.c_int]
    L.cmaxb_be_map_mark_fov.argtypes = [vp, dp, C.c_int, C.c_int]
    L.cmaxb_fe_optimize.argtypes = [vp, dp, C.POINTER(OptParams), dp, C.POINTER(OptResult)]
    L.cmaxb_be_optimize.argtypes = [vp, dp, C.c_int, C.POINTER(OptParams), dp, C.POINTER(OptResult)]
    L.cmaxb_optimize_callback.argtypes = [C.c_int, dp, COST_F, COST_FDF, vp, C.POINTER(OptParams), dp, C.POINTER(OptResult)]
    sp = C.POINTER(Stamp)
    L.cmaxb_traj_integrate_ang_vel.argtypes = [Stamp, dp, sp, dp, C.c_int, sp, dp, C.c_int, sp, dp, C.POINTER(C.c_int)]
    L.cmaxb_traj_num_ctrl_poses.argtypes = [C.c_int, Stamp, Stamp, C.c_double]
    L.cmaxb_traj_fit_ctrl_poses.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, sp, dp, C.c_int, dp]
    L.cmaxb_traj_evaluate.argtypes = [C.c_int, dp, C.c_int, C.c_int64, C.c_int64, Stamp, dp]
    L.cmaxb_traj_incremental_update.argtypes = [dp, C.c_int, C.c_int, dp]
    L.cmaxb_precompute_bearing_vectors.argtypes = [C.POINTER(CameraInfo), C.c_int, dp]
    L.cmaxb_stream_create.argtypes = [C.POINTER(StreamCfg), C.POINTER(vp)]
    L.cmaxb_stream_destroy.argtypes = [vp]
    L.cmaxb_stream_destroy.restype = None
    L.cmaxb_stream_push.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_int)]
    L.cmaxb_stream_next_packet.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), sp, C.POINTER(C.c_int)]
    L.cmaxb_stream_window_events.argtypes = [vp, Stamp, Stamp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.cmaxb_stream_state.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), sp]
    L.cmaxb_stream_attach_device.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.cmaxb_stream_push_ex.argtypes = [vp, vp, C.c_size_t, C.c_int, C.POINTER(C.c_int)]
    L.cmaxb_stream_next_packet_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), sp, C.POINTER(C.c_int)]
    L.cmaxb_stream_released.argtypes = [vp, C.POINTER(C.c_int64)]
    L.cmaxb_stream_wait_copied.argtypes = [vp, vp]
    L.cmaxb_pgo_create.argtypes = [C.POINTER(PgoCfg), vp, C.POINTER(vp)]
    L.cmaxb_pgo_destroy.argtypes = [vp]
    L.cmaxb_pgo_destroy.restype = None
    L.cmaxb_pgo_push_ang_vel.argtypes = [vp, Stamp, dp]
    L.cmaxb_pgo_window.argtypes = [vp, sp, sp, C.POINTER(C.c_int)]
    L.cmaxb_pgo_process_window.argtypes = [vp, vp, C.c_size_t, C.POINTER(PgoReport)]
    L.cmaxb_pgo_get_ctrl_poses.argtypes = [vp, dp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.cmaxb_be_last_eval_x.argtypes = [vp, dp, C.c_int]
    L.cmaxb_last_error.restype = C.c_char_p
    L.cmaxb_launch_count.restype = C.c_uint64
    L.cmaxb_kernel_name.restype = C.c_char_p
    L.cmaxb_kernel_name.argtypes = [C.c_int]
    L.cmaxb_fe_profile.argtypes = [vp, C.c_int]
    L.cmaxb_fe_kernel_times.argtypes = [vp, dp, C.POINTER(C.c_uint64)]
    L.cmaxb_fe_cta_times.argtypes = [vp, dp, C.c_int, C.POINTER(C.c_int)]
    L.cmaxb_fe_phase_times.argtypes = [vp, dp]
    if not hasattr(L, "cmaxb_be_create"):
        raise RuntimeError("libcmax_b200.so is stale (no back-end symbols): rebuild with cmax_slam_b200/build.py --force")
    L.cmaxb_be_create.argtypes = [C.POINTER(BeCfg), C.POINTER(vp)]
    L.cmaxb_be_destroy.argtypes = [vp]
    L.cmaxb_be_destroy.restype = None
    L.cmaxb_be_set_window.argtypes = [vp, C.POINTER(BeWindow)]
    L.cmaxb_be_eval.argtypes = [vp, dp, C.c_int, dp, dp]
    L.cmaxb_be_get_alpha.argtypes = [vp, dp]
    L.cmaxb_be_eval_begin.argtypes = [vp, dp, C.c_int, C.c_int]
    L.cmaxb_be_il_plane.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.cmaxb_be_eval_end.argtypes = [vp, dp, dp]
    L.cmaxb_be_eval_end_launch.argtypes = [vp, C.c_int]
    L.cmaxb_be_grad_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.cmaxb_be_shard_begin.argtypes = [vp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.cmaxb_be_shard_image.argtypes = [vp, C.POINTER(vp)]
    L.cmaxb_be_shard_adjoint.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.cmaxb_be_shard_gather.argtypes = [vp]
    L.cmaxb_be_exchange_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.cmaxb_be_exchange_connect.argtypes = [vp, C.c_char_p]
    L.cmaxb_be_exchange_close.argtypes = [vp]
    L.cmaxb_be_exchange_stats.argtypes = [vp, C.POINTER(C.c_int64)]
    L.cmaxb_be_xeval.argtypes = [vp, dp, C.c_int, C.POINTER(C.c_double), dp]
    L.cmaxb_be_eval_end_fetch.argtypes = [vp, dp, dp]
    L.cmaxb_be_get_il.argtypes = [vp, dp, C.c_int, vp, vp]
    L.cmaxb_be_get_iwe.argtypes = [vp, dp, C.c_int, C.c_int, vp]
    L.cmaxb_be_get_bands.argtypes = [vp, dp, C.c_int, C.c_int, vp]
    L.cmaxb_be_get_cells.argtypes = [vp, dp, C.c_int, vp]
    L.cmaxb_be_get_poses.argtypes = [vp, dp, C.c_int, C.POINTER(C.c_int64), vp, vp, vp, C.c_int64]
    L.cmaxb_be_profile.argtypes = [vp, C.c_int]
    L.cmaxb_be_kernel_times.argtypes = [vp, dp, C.POINTER(C.c_uint64)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise CmaxbError(rc, lib().cmaxb_last_error().decode())


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def launch_count():
    return int(lib().cmaxb_launch_count())


def device_count():
    return int(lib().cmaxb_device_count())
