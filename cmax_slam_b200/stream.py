"""Host-side mirror of the reference's event ingestion over the C ABI.

Reference surface mirrored:
  CMaxSLAM::eventsCallback                       src/cmax_slam.cpp:147-161
  AngVelEstimator::pushEvent / getEventSubset / deleteOldEvents / slideWindow
                                                 src/frontend/ang_vel_estimator.cpp:68-183
  PoseGraphOptimizer::getEventSubset             src/backend/pose_graph_optimizer.cpp:133-166
The bookkeeping is host C++ inside libcmax_b200.so (csrc/stream.cu); packets / windows come back as numpy views of
the library's pinned staging buffers (valid until the call after next)."""
import ctypes as C

import numpy as np

from . import _capi
from .synth import EVENT_DTYPE


class EventStream:
    def __init__(self, dt_ang_vel, num_events_per_packet, event_sample_rate=1):
        self._L = _capi.lib()
        cfg = _capi.StreamCfg(float(dt_ang_vel), int(num_events_per_packet), int(event_sample_rate))
        h = C.c_void_p()
        _capi.check(self._L.cmaxb_stream_create(C.byref(cfg), C.byref(h)))
        self._s = h

    def close(self):
        if getattr(self, "_s", None):
            self._L.cmaxb_stream_destroy(self._s)
            self._s = None

    __del__ = close

    PUSH_BORROW, PUSH_SORTED = 1, 2

    def eventsCallback(self, msg_events, flags=0):
        """One dvs_msgs::EventArray (16-byte records, numpy array or a (ptr, n) tuple).  Returns the number of complete
        packets waiting.  flags: PUSH_BORROW (page-locked message referenced in place), PUSH_SORTED (time-sorted message)."""
        if isinstance(msg_events, tuple):
            ptr, n = msg_events
        else:
            ev = np.ascontiguousarray(msg_events)
            ptr, n = ev.ctypes.data, len(ev)
        k = C.c_int(0)
        _capi.check(self._L.cmaxb_stream_push_ex(self._s, C.c_void_p(ptr), n, int(flags), C.byref(k)))
        return k.value

    def attach_device(self, device=0, cuda_stream=None, ring_events=0):
        """Device-resident event store: every pushed event crosses PCIe once; next_packet_device hands out ring views."""
        _capi.check(self._L.cmaxb_stream_attach_device(self._s, int(device), C.c_void_p(int(cuda_stream)) if cuda_stream else None,
                                                       int(ring_events)))

    def next_packet_device(self):
        """None, or ((device pointer, n), (sec, nsec) time_packet, span_too_long) -- input of set_packet(..., view=True)."""
        p, n, t, f = C.c_void_p(), C.c_size_t(0), _capi.Stamp(), C.c_int(0)
        rc = self._L.cmaxb_stream_next_packet_device(self._s, C.byref(p), C.byref(n), C.byref(t), C.byref(f))
        if rc == 1:
            return None
        _capi.check(rc)
        return (p.value, n.value), (t.sec, t.nsec), bool(f.value)

    def wait_copied(self, consumer_stream):
        """consumer_stream (cudaStream_t) waits for the device copies of everything pushed so far."""
        _capi.check(self._L.cmaxb_stream_wait_copied(self._s, C.c_void_p(int(consumer_stream)) if consumer_stream else None))

    def released(self):
        k = C.c_int64(0)
        _capi.check(self._L.cmaxb_stream_released(self._s, C.byref(k)))
        return k.value

    def _view(self, ptr, n):
        if not n:
            return np.zeros(0, EVENT_DTYPE)
        buf = (C.c_uint8 * (16 * n)).from_address(ptr)
        return np.frombuffer(buf, dtype=EVENT_DTYPE)

    def next_packet(self):
        """None, or (events view, (sec, nsec) time_packet, span_too_long)."""
        p, n, t, f = C.c_void_p(), C.c_size_t(0), _capi.Stamp(), C.c_int(0)
        rc = self._L.cmaxb_stream_next_packet(self._s, C.byref(p), C.byref(n), C.byref(t), C.byref(f))
        if rc == 1:
            return None
        _capi.check(rc)
        return self._view(p.value, n.value), (t.sec, t.nsec), bool(f.value)

    def window_events(self, t_beg, t_end):
        """Events of the back-end window, or None when the store does not cover the window yet."""
        p, n = C.c_void_p(), C.c_size_t(0)
        rc = self._L.cmaxb_stream_window_events(self._s, _capi.Stamp(int(t_beg[0]), int(t_beg[1])),
                                                _capi.Stamp(int(t_end[0]), int(t_end[1])), C.byref(p), C.byref(n))
        if rc == 1:
            return None
        _capi.check(rc)
        return self._view(p.value, n.value)

    def state(self):
        a, b, c, t = C.c_int64(0), C.c_int64(0), C.c_int64(0), _capi.Stamp()
        _capi.check(self._L.cmaxb_stream_state(self._s, C.byref(a), C.byref(b), C.byref(c), C.byref(t)))
        return {"n_stored": a.value, "n_subsets_pending": b.value, "n_ts_map": c.value, "time_packet": (t.sec, t.nsec)}


def precompute_bearing_vectors(width, height, K, D, R=None, P=None, device=0):
    """CMaxSLAM::precomputeBearingVectors (src/cmax_slam.cpp:106-120) on the device: (H*W, 3) float64."""
    K = np.asarray(K, dtype=np.float64).reshape(3, 3)
    D = np.asarray(D, dtype=np.float64).reshape(-1)
    R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64).reshape(3, 3)
    P = np.hstack([K, np.zeros((3, 1))]) if P is None else np.asarray(P, dtype=np.float64).reshape(3, 4)
    info = _capi.CameraInfo()
    info.width, info.height, info.n_D = int(width), int(height), len(D)
    info.K[:] = K.ravel().tolist()
    info.D[:] = (D.tolist() + [0.0] * 12)[:12]
    info.R[:] = R.ravel().tolist()
    info.P[:] = P.ravel().tolist()
    out = np.zeros((width * height, 3))
    _capi.check(_capi.lib().cmaxb_precompute_bearing_vectors(C.byref(info), int(device), _capi.dptr(out)))
    return out
