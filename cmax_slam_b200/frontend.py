"""Host-side mirror of the reference's front-end cost interface over the C ABI.

Reference surface mirrored (same names, argument meaning and sign conventions):
  AngVelEstimator::computeImageOfWarpedEvents   src/frontend/local_image_warped_events.cpp:10-57
  cmax_slam::computeContrast                    src/frontend/local_focus_funcs.cpp:82-120
  local_contrast_fdf / _f / _df (GSL callbacks) src/frontend/local_optim_contrast_gsl.cpp:20-70
All arithmetic happens in libcmax_b200.so on the GPU; this module only marshals pointers.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import GRAD_ADJOINT, GRAD_DENSE, CmaxbError  # noqa: F401


class AngVelEstimatorCMax:
    """Device-resident front-end contrast functor (one handle = one CUDA stream)."""

    def __init__(self, width, height, K4, lut_xyz, blur_sigma=1.0, event_batch_size=100, contrast_measure=0,
                 grad_mode=GRAD_ADJOINT, device=0, stream=None, max_hypotheses=1, lanes=0, packet_slots=1):
        self._L = _capi.lib()
        lut = np.ascontiguousarray(lut_xyz, dtype=np.float64).reshape(-1, 3)
        if lut.shape[0] != width * height:
            raise ValueError("lut_xyz must hold width*height bearing vectors")
        cfg = _capi.FeCfg(width, height, K4[0], K4[1], K4[2], K4[3], lut.ctypes.data, float(blur_sigma),
                          int(event_batch_size), int(contrast_measure), int(grad_mode), int(device),
                          None if stream is None else C.c_void_p(int(stream)), int(max_hypotheses), int(lanes),
                          int(packet_slots))
        h = C.c_void_p()
        _capi.check(self._L.cmaxb_fe_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.width, self.height = width, height
        self.max_hypotheses = max(1, int(max_hypotheses))
        self.n_events = 0
        self._res_c = np.zeros(self.max_hypotheses)
        self._res_g = np.zeros((self.max_hypotheses, 3))
        self._om = np.zeros((self.max_hypotheses, 3))
        self._ks = []
        self._xworld = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.cmaxb_fe_destroy(self._h)
            self._h = None

    __del__ = close

    # -- packet ---------------------------------------------------------------------------------
    def set_packet(self, events, t_ref_sec, wait=True, view=False):
        """events: numpy structured array with the 16-byte dvs_msgs::Event layout (synth.EVENT_DTYPE)
        or a (ptr, n) tuple of pinned host / device memory; t_ref_sec = time_packet_.toSec().  wait=False queues the
        upload and returns (the validation verdict comes with the next evaluation).  view=True: (ptr, n) is DEVICE
        memory that is used in place (cmaxb_fe_set_packet_view)."""
        if isinstance(events, tuple):
            ptr, n = events
        else:
            ev = np.ascontiguousarray(events)
            if ev.dtype.itemsize != 16:
                raise ValueError("events must be 16-byte dvs_msgs::Event records")
            self._ev_keep = ev
            ptr, n = ev.ctypes.data, len(ev)
        fn = self._L.cmaxb_fe_set_packet_view if view else (self._L.cmaxb_fe_set_packet if wait else self._L.cmaxb_fe_set_packet_async)
        _capi.check(fn(self._h, C.c_void_p(ptr), n, float(t_ref_sec)))
        self.n_events = n

    def select_packet(self, slot):
        """Resident packet slot (0 .. packet_slots-1) the following set_packet / eval / getter calls act on."""
        _capi.check(self._L.cmaxb_fe_select_packet(self._h, int(slot)))

    def lanes_fork(self):
        _capi.check(self._L.cmaxb_fe_lanes_fork(self._h))

    def lanes_join(self):
        _capi.check(self._L.cmaxb_fe_lanes_join(self._h))

    def launch_info(self):
        info = np.zeros(7, np.int32)
        _capi.check(self._L.cmaxb_fe_launch_info(self._h, info.ctypes.data_as(C.POINTER(C.c_int32))))
        return {"grid_full": int(info[0]), "grid_lane": int(info[1]), "lanes": int(info[2]), "tma": bool(info[3]),
                "tile_h_full": int(info[4]), "tile_h_lane": int(info[5]), "gather_records": int(info[6])}

    # -- cost ------------------------------------------------------------------------------------
    def eval(self, ang_vel, want_grad=True):
        """(+contrast, +gradient[3] or None) at one angular velocity."""
        om = np.ascontiguousarray(ang_vel, dtype=np.float64).reshape(3)
        c = C.c_double()
        g = np.zeros(3)
        _capi.check(self._L.cmaxb_fe_eval(self._h, _capi.dptr(om), C.byref(c), _capi.dptr(g) if want_grad else None))
        return c.value, (g if want_grad else None)

    def eval_batch(self, ang_vels, want_grad=True):
        om = np.ascontiguousarray(ang_vels, dtype=np.float64).reshape(-1, 3)
        k = om.shape[0]
        c = np.zeros(k)
        g = np.zeros((k, 3))
        _capi.check(self._L.cmaxb_fe_eval_batch(self._h, _capi.dptr(om), k, _capi.dptr(c), _capi.dptr(g) if want_grad else None))
        return c, (g if want_grad else None)

    def eval_launch(self, ang_vels, want_grad=True):
        """Queue one evaluation (up to 8 may be outstanding); results come back in launch order from eval_fetch."""
        om = np.ascontiguousarray(ang_vels, dtype=np.float64).reshape(-1, 3)
        k = om.shape[0]
        self._om[:k] = om
        _capi.check(self._L.cmaxb_fe_eval_launch(self._h, _capi.dptr(self._om), k, int(want_grad)))
        self._ks.append(k)

    def eval_fetch(self):
        k = self._ks[0] if self._ks else 0
        _capi.check(self._L.cmaxb_fe_eval_fetch(self._h, _capi.dptr(self._res_c), _capi.dptr(self._res_g)))
        self._ks.pop(0)
        return self._res_c[:k].copy(), self._res_g[:k].copy()

    # -- fused multi-GPU result exchange (peer-to-peer stores from the evaluation kernel) ----------
    def exchange_connect(self, group=None, gathered_dev_ptr=None):
        """Collective over `group` (torch.distributed, any backend: only the 64-byte IPC handles travel
        through it).  Afterwards every rank must issue the same sequence of eval_launch calls."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        h = C.create_string_buffer(64)
        _capi.check(self._L.cmaxb_fe_exchange_init(self._h, world, rank, h))
        handles = [None] * world
        dist.all_gather_object(handles, h.raw, group=group)
        blob = b"".join(handles)
        _capi.check(self._L.cmaxb_fe_exchange_connect(self._h, blob, C.c_void_p(gathered_dev_ptr) if gathered_dev_ptr else None))
        self._xworld = world
        dist.barrier(group)

    def exchange_close(self):
        _capi.check(self._L.cmaxb_fe_exchange_close(self._h))
        self._xworld = 0

    def eval_fetch_all(self):
        """Rows (contrast, g0, g1, g2) of ALL ranks for the oldest outstanding launch: array [world, k, 4]."""
        k = self._ks[0] if self._ks else 0
        rows = np.zeros((self._xworld, max(k, 1), 4))
        _capi.check(self._L.cmaxb_fe_eval_fetch_all(self._h, _capi.dptr(rows)))
        self._ks.pop(0)
        return rows

    def set_result_mirror(self, device_ptr):
        """device_ptr: address of a device buffer of max_hypotheses*4 float64 (e.g. tensor.data_ptr()) that also
        receives every evaluation's (contrast, g0, g1, g2) rows -- input of the multi-GPU collective."""
        _capi.check(self._L.cmaxb_fe_set_result_mirror(self._h, C.c_void_p(device_ptr) if device_ptr else None))

    def setupProblemAndOptimize(self, ang_vel0, params=None):
        """AngVelEstimator::setupProblemAndOptimize_gsl (local_optim_contrast_gsl.cpp:74-233) without GSL:
        Fletcher-Reeves CG with the reference's constants.  Returns (ang_vel, stats dict)."""
        om = np.ascontiguousarray(ang_vel0, dtype=np.float64).reshape(3)
        out = np.zeros(3)
        res = _capi.OptResult()
        prm = None if params is None else C.byref(_capi.OptParams(*params))
        _capi.check(self._L.cmaxb_fe_optimize(self._h, _capi.dptr(om), prm, _capi.dptr(out), C.byref(res)))
        return out, {k: getattr(res, k) for k, _ in _capi.OptResult._fields_}

    # -- reference-named entry points ------------------------------------------------------------
    def computeImageOfWarpedEvents(self, ang_vel, with_deriv=False, blurred=True):
        """IWE (H,W) float32 [and derivative image (H,W,3)] as the reference function fills them."""
        om = np.ascontiguousarray(ang_vel, dtype=np.float64).reshape(3)
        iwe = np.empty((self.height, self.width), np.float32)
        _capi.check(self._L.cmaxb_fe_get_iwe(self._h, _capi.dptr(om), int(blurred), C.c_void_p(iwe.ctypes.data)))
        if not with_deriv:
            return iwe
        d = np.empty((self.height, self.width, 3), np.float32)
        _capi.check(self._L.cmaxb_fe_get_deriv(self._h, _capi.dptr(om), int(blurred), C.c_void_p(d.ctypes.data)))
        return iwe, d

    def warped_cells(self, ang_vel):
        om = np.ascontiguousarray(ang_vel, dtype=np.float64).reshape(3)
        out = np.empty(self.n_events, np.int32)
        _capi.check(self._L.cmaxb_fe_get_cells(self._h, _capi.dptr(om), C.c_void_p(out.ctypes.data)))
        return out

    # -- profiling -------------------------------------------------------------------------------
    def profile(self, enable=True):
        _capi.check(self._L.cmaxb_fe_profile(self._h, int(enable)))

    def phase_times(self):
        """Fused kernel (profiling on): us from kernel entry to each phase boundary of the last launch."""
        t = np.zeros(10)
        _capi.check(self._L.cmaxb_fe_phase_times(self._h, _capi.dptr(t)))
        return t

    def cta_times(self):
        """Fused kernel (profiling on): per-CTA us since kernel entry, array [ctas, 4] = scatter end, image end, gather start, gather end."""
        out = np.zeros((1024, 4))
        n = C.c_int(0)
        _capi.check(self._L.cmaxb_fe_cta_times(self._h, _capi.dptr(out), 1024, C.byref(n)))
        return out[: n.value].copy()

    def kernel_times(self):
        ms = np.zeros(_capi.K_COUNT)
        n = np.zeros(_capi.K_COUNT, np.uint64)
        _capi.check(self._L.cmaxb_fe_kernel_times(self._h, _capi.dptr(ms), n.ctypes.data_as(C.POINTER(C.c_uint64))))
        return {_capi.K_NAMES[i]: (float(ms[i]), int(n[i])) for i in range(_capi.K_COUNT) if n[i] > 0}


# GSL callback triple with the reference's exact semantics: returns -contrast / -gradient
# (local_optim_contrast_gsl.cpp:48-54); df=None means value only (:32,40).
def local_contrast_fdf(v, estimator, want_df=True):
    c, g = estimator.eval(v, want_grad=want_df)
    return -c, (None if g is None else -g)


def local_contrast_f(v, estimator):
    return local_contrast_fdf(v, estimator, want_df=False)[0]


def local_contrast_df(v, estimator):
    return local_contrast_fdf(v, estimator, want_df=True)[1]
