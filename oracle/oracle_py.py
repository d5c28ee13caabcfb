"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (cmax_slam_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libcmax_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libref_basalt.so")

EVENT_DTYPE = np.dtype(
    [("x", "<u2"), ("y", "<u2"), ("sec", "<u4"), ("nsec", "<u4"), ("polarity", "u1"), ("pad", "u1", (3,))]
)
assert EVENT_DTYPE.itemsize == 16

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)


class FeArgs(C.Structure):
    _fields_ = [("events", C.c_void_p), ("n_events", C.c_int64), ("t_ref_sec", C.c_double),
                ("lut_xyz", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("batch_size", C.c_int32), ("blur_sigma", C.c_double), ("contrast_measure", C.c_int32)]


class FeOut(C.Structure):
    _fields_ = [("contrast", C.c_double), ("grad", C.c_double * 3), ("iwe", C.c_void_p),
                ("deriv", C.c_void_p), ("iwe_raw", C.c_void_p), ("deriv_raw", C.c_void_p),
                ("cells", C.c_void_p), ("n_inbounds", C.c_int64)]


class BeArgs(C.Structure):
    _fields_ = [("events", C.c_void_p), ("n_events", C.c_int64), ("lut_xyz", C.c_void_p),
                ("sensor_width", C.c_int32), ("sensor_height", C.c_int32),
                ("pano_width", C.c_int32), ("pano_height", C.c_int32),
                ("knots_xyzw", C.c_void_p), ("n_knots", C.c_int32),
                ("t0_ns", C.c_int64), ("dt_ns", C.c_int64), ("spline_order", C.c_int32),
                ("n_fixed", C.c_int32), ("tnext_sec", C.c_uint32), ("tnext_nsec", C.c_uint32),
                ("IGp", C.c_void_p), ("alpha", C.c_double),
                ("batch_size", C.c_int32), ("event_sample_rate", C.c_int32),
                ("blur_sigma", C.c_double), ("contrast_measure", C.c_int32)]


class BeOut(C.Structure):
    _fields_ = [("contrast", C.c_double), ("grad", C.c_void_p), ("iwe", C.c_void_p),
                ("bands", C.c_void_p), ("bands_raw", C.c_void_p), ("il_old", C.c_void_p),
                ("il_new", C.c_void_p), ("cells", C.c_void_p), ("n_inbounds", C.c_int64)]


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(
            os.path.join(_HERE, "cmax_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-j8"], env={**os.environ, "CXX": "g++"})
    return _LIB


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_fe_eval.restype = C.c_int
        L.orc_fe_eval.argtypes = [C.POINTER(FeArgs), _dp, C.c_int, C.POINTER(FeOut)]
        L.orc_fe_eval_batch.restype = C.c_int
        L.orc_fe_eval_batch.argtypes = [C.POINTER(FeArgs), _dp, C.c_int, C.c_int, _dp, _dp, C.c_int]
        L.orc_be_eval.restype = C.c_int
        L.orc_be_eval.argtypes = [C.POINTER(BeArgs), _dp, C.c_int, C.POINTER(BeOut)]
        L.orc_update_alpha.restype = C.c_double
        L.orc_update_alpha.argtypes = [_fp, _fp, C.c_int64]
        L.orc_gaussian_kernel.restype = C.c_int
        L.orc_gaussian_kernel.argtypes = [C.c_double, _fp]
        L.orc_gaussian_blur.restype = None
        L.orc_gaussian_blur.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_double]
        L.orc_mean_stddev.restype = None
        L.orc_mean_stddev.argtypes = [_fp, C.c_int64, _dp, _dp]
        L.orc_so3_spline_eval.restype = C.c_int
        L.orc_so3_spline_eval.argtypes = [C.c_int, _dp, C.c_int, C.c_int64, C.c_int64, C.c_int64, _dp, _dp, _ip, _dp]
        L.orc_so3_exp.argtypes = [_dp, _dp]
        L.orc_so3_log.argtypes = [_dp, _dp]
        L.orc_batch_mid_time.argtypes = [C.c_uint32] * 4 + [C.POINTER(C.c_uint32)] * 2
        L.orc_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def have_ref():
    return os.path.exists(_REF)


def ref():
    """The REAL basalt/Sophus code of the reference (oracle/_ref), or None."""
    global _ref
    if _ref is None and have_ref():
        L = C.CDLL(_REF)
        L.ref_so3_spline_eval.restype = C.c_int
        L.ref_so3_spline_eval.argtypes = [C.c_int, _dp, C.c_int, C.c_int64, C.c_int64, C.c_int64, _dp, _dp, _ip, _dp]
        L.ref_so3_exp.argtypes = [_dp, _dp]
        L.ref_so3_log.argtypes = [_dp, _dp]
        L.ref_knot_update.argtypes = [_dp, _dp, _dp]
        L.ref_left_jacobians.argtypes = [_dp, _dp, _dp]
        if hasattr(L, "ref_fit_ctrl_poses"):
            _up = C.POINTER(C.c_uint32)
            L.ref_fit_ctrl_poses.restype = C.c_int
            L.ref_fit_ctrl_poses.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, _up, _dp, C.c_int, _dp]
            L.ref_integrate_ang_vel.restype = C.c_int
            L.ref_integrate_ang_vel.argtypes = [_up, _dp, _up, _dp, C.c_int, _up, _dp, C.c_int, _up, _dp]
        _ref = L
    return _ref


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def _d(a):
    return a.ctypes.data_as(_dp)


def fe_args(events, t_ref_sec, lut, W, H, K4, batch_size=100, blur_sigma=1.0, measure=0):
    events = np.ascontiguousarray(events)
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    a = FeArgs(events.ctypes.data, len(events), float(t_ref_sec), lut.ctypes.data, W, H,
               K4[0], K4[1], K4[2], K4[3], batch_size, blur_sigma, measure)
    a._keep = (events, lut)
    return a


def fe_eval(args, omega, want_grad=True, images=False, cells=False):
    """Returns dict(contrast, grad, [iwe, deriv, iwe_raw, deriv_raw], [cells], n_inbounds)."""
    W, H = args.width, args.height
    o = FeOut()
    res = {}
    if images:
        res["iwe"] = np.zeros((H, W), np.float32)
        res["iwe_raw"] = np.zeros((H, W), np.float32)
        o.iwe, o.iwe_raw = res["iwe"].ctypes.data, res["iwe_raw"].ctypes.data
        if want_grad:
            res["deriv"] = np.zeros((H, W, 3), np.float32)
            res["deriv_raw"] = np.zeros((H, W, 3), np.float32)
            o.deriv, o.deriv_raw = res["deriv"].ctypes.data, res["deriv_raw"].ctypes.data
    if cells:
        res["cells"] = np.zeros(args.n_events, np.int32)
        o.cells = res["cells"].ctypes.data
    om = np.ascontiguousarray(omega, dtype=np.float64)
    rc = lib().orc_fe_eval(C.byref(args), _d(om), int(want_grad), C.byref(o))
    if rc != 0:
        raise RuntimeError("oracle: " + lib().orc_last_error().decode())
    res["contrast"] = o.contrast
    res["grad"] = np.array(list(o.grad)) if want_grad else None
    res["n_inbounds"] = o.n_inbounds
    return res


def fe_eval_batch(args, omegas, want_grad=True, n_threads=1):
    om = np.ascontiguousarray(omegas, dtype=np.float64).reshape(-1, 3)
    k = om.shape[0]
    c = np.zeros(k)
    g = np.zeros((k, 3))
    rc = lib().orc_fe_eval_batch(C.byref(args), _d(om), k, int(want_grad), _d(c), _d(g), n_threads)
    if rc != 0:
        raise RuntimeError("oracle: " + lib().orc_last_error().decode())
    return c, (g if want_grad else None)


def be_args(events, lut, SW, SH, PW, PH, knots_xyzw, t0_ns, dt_ns, order, n_fixed, tnext, IGp=None,
            alpha=0.0, batch_size=100, sample_rate=1, blur_sigma=1.0, measure=0):
    events = np.ascontiguousarray(events)
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    knots = np.ascontiguousarray(knots_xyzw, dtype=np.float64)
    igp = None if IGp is None else np.ascontiguousarray(IGp, dtype=np.float32)
    a = BeArgs(events.ctypes.data, len(events), lut.ctypes.data, SW, SH, PW, PH, knots.ctypes.data,
               knots.shape[0], int(t0_ns), int(dt_ns), order, n_fixed, int(tnext[0]), int(tnext[1]),
               None if igp is None else igp.ctypes.data, float(alpha), batch_size, sample_rate,
               blur_sigma, measure)
    a._keep = (events, lut, knots, igp)
    return a


def be_eval(args, x=None, want_grad=True, images=False, cells=False):
    W, H = args.pano_width, args.pano_height
    P = 3 * (args.n_knots - args.n_fixed)
    o = BeOut()
    res = {}
    g = np.zeros(max(P, 1))
    if want_grad:
        o.grad = g.ctypes.data
    if images:
        for k in ("iwe", "il_old", "il_new"):
            res[k] = np.zeros((H, W), np.float32)
            setattr(o, k, res[k].ctypes.data)
        if want_grad:
            res["bands"] = np.zeros((P, H, W), np.float32)
            res["bands_raw"] = np.zeros((P, H, W), np.float32)
            o.bands, o.bands_raw = res["bands"].ctypes.data, res["bands_raw"].ctypes.data
    if cells:
        res["cells"] = np.zeros(args.n_events, np.int32)
        o.cells = res["cells"].ctypes.data
    xx = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
    rc = lib().orc_be_eval(C.byref(args), None if xx is None else _d(xx), int(want_grad), C.byref(o))
    if rc != 0:
        raise RuntimeError("oracle: " + lib().orc_last_error().decode())
    res["contrast"] = o.contrast
    res["grad"] = g[:P].copy() if want_grad else None
    res["n_inbounds"] = o.n_inbounds
    return res


def gaussian_blur(img, sigma):
    img = np.ascontiguousarray(img, dtype=np.float32)
    H, W = img.shape[:2]
    Cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    lib().orc_gaussian_blur(img.ctypes.data_as(_fp), out.ctypes.data_as(_fp), W, H, Cn, float(sigma))
    return out


def gaussian_kernel(sigma):
    t = np.zeros(64, np.float32)
    k = lib().orc_gaussian_kernel(float(sigma), t.ctypes.data_as(_fp))
    return t[:k].copy()


def mean_stddev(img):
    img = np.ascontiguousarray(img, dtype=np.float32)
    m, s = C.c_double(), C.c_double()
    lib().orc_mean_stddev(img.ctypes.data_as(_fp), img.size, C.byref(m), C.byref(s))
    return m.value, s.value


def _spline_eval(fn, order, knots, t0_ns, dt_ns, t_ns, want_J=True):
    knots = np.ascontiguousarray(knots, dtype=np.float64)
    q = np.zeros(4)
    R = np.zeros(9)
    J = np.zeros(9 * order)
    idx = C.c_int32(0)
    rc = fn(order, _d(knots), knots.shape[0], int(t0_ns), int(dt_ns), int(t_ns), _d(q), _d(R),
            C.byref(idx), _d(J) if want_J else None)
    if rc != 0:
        return None
    return q, R.reshape(3, 3), idx.value, J.reshape(order, 3, 3)


def spline_eval(order, knots, t0_ns, dt_ns, t_ns, want_J=True):
    return _spline_eval(lib().orc_so3_spline_eval, order, knots, t0_ns, dt_ns, t_ns, want_J)


def ref_spline_eval(order, knots, t0_ns, dt_ns, t_ns, want_J=True):
    return _spline_eval(ref().ref_so3_spline_eval, order, knots, t0_ns, dt_ns, t_ns, want_J)


def set_update_times(lut, SW, SH, PW, PH, rot_xyzw, radius, times):
    """in place on `times` (uint8 [PH,PW])"""
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    q = np.ascontiguousarray(rot_xyzw, dtype=np.float64)
    assert times.dtype == np.uint8 and times.flags.c_contiguous
    L = lib()
    L.orc_set_update_times.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.c_void_p]
    L.orc_set_update_times.restype = None
    L.orc_set_update_times(_d(lut), SW, SH, PW, PH, _d(q), int(radius), times.ctypes.data)


def update_ig(IG, il_old, times, max_update_times):
    """in place on IG (float32)"""
    assert IG.dtype == np.float32 and IG.flags.c_contiguous
    il = np.ascontiguousarray(il_old, dtype=np.float32)
    t = np.ascontiguousarray(times, dtype=np.uint8)
    L = lib()
    L.orc_update_ig.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64]
    L.orc_update_ig.restype = None
    L.orc_update_ig(IG.ctypes.data, il.ctypes.data, t.ctypes.data, int(max_update_times), IG.size)


def update_alpha(IGp, IL):
    a = np.ascontiguousarray(IGp, dtype=np.float32)
    b = np.ascontiguousarray(IL, dtype=np.float32)
    return lib().orc_update_alpha(a.ctypes.data_as(_fp), b.ctypes.data_as(_fp), a.size)


# ---- trajectory initialisation against the real Eigen / Sophus code (oracle/_ref) -------------------------------
def ref_fit_ctrl_poses(order, dt_knots, t_beg, num_cps, stamps, poses_xyzw):
    """stamps: (n,2) uint32 (sec, nsec); poses: (n,4).  Returns (num_cps,4) or raises."""
    L = ref()
    st = np.ascontiguousarray(stamps, dtype=np.uint32)
    q = np.ascontiguousarray(poses_xyzw, dtype=np.float64)
    out = np.zeros((num_cps, 4))
    rc = L.ref_fit_ctrl_poses(order, float(dt_knots), float(t_beg), int(num_cps), st.ctypes.data_as(C.POINTER(C.c_uint32)), _d(q), len(q), _d(out))
    if rc != 0:
        raise ValueError(f"ref_fit_ctrl_poses rc={rc}")
    return out


def ref_integrate_ang_vel(latest_stamp, latest_xyzw, prev_stamp, prev_w, first, stamps, w):
    """Returns (pose_stamps (n,2), poses (n,4), new prev_stamp, new prev_w)."""
    L = ref()
    up = C.POINTER(C.c_uint32)
    ls = np.ascontiguousarray(latest_stamp, dtype=np.uint32)
    lq = np.ascontiguousarray(latest_xyzw, dtype=np.float64)
    ps = np.array(prev_stamp, dtype=np.uint32)
    pw = np.array(prev_w, dtype=np.float64)
    st = np.ascontiguousarray(stamps, dtype=np.uint32).reshape(-1, 2)
    ww = np.ascontiguousarray(w, dtype=np.float64).reshape(-1, 3)
    os_ = np.zeros((len(st), 2), np.uint32)
    oq = np.zeros((len(st), 4))
    n = L.ref_integrate_ang_vel(ls.ctypes.data_as(up), _d(lq), ps.ctypes.data_as(up), _d(pw), int(first), st.ctypes.data_as(up), _d(ww), len(st),
                                os_.ctypes.data_as(up), _d(oq))
    return os_[:n], oq[:n], ps, pw


# ---- first-party geometry against the reference's own sources (oracle/_ref/libref_geom.so) ----------------------
_REF_GEOM = os.path.join(_HERE, "_ref", "libref_geom.so")
_ref_geom = None


def have_ref_geom():
    return os.path.exists(_REF_GEOM)


def ref_geom():
    global _ref_geom
    if _ref_geom is None and have_ref_geom():
        _ref_geom = C.CDLL(_REF_GEOM)
    return _ref_geom


def _geom_pinhole(L, fn, p, K4, with_intr):
    p = np.ascontiguousarray(p, dtype=np.float64); K4 = np.ascontiguousarray(K4, dtype=np.float64)
    uv, px, Jp, Ji = np.zeros(2), np.zeros(2), np.zeros(6), np.zeros(4)
    if with_intr:
        getattr(L, fn)(_d(p), _d(K4), _d(uv), _d(px), _d(Jp), _d(Ji))
    else:
        getattr(L, fn)(_d(p), _d(K4), _d(uv), _d(px), _d(Jp))
    return uv, px, Jp.reshape(2, 3), Ji.reshape(2, 2)


def geom_pinhole(p, K4):
    """oracle restatement: (uv, pixel, Jproj 2x3)"""
    return _geom_pinhole(lib(), "orc_geom_pinhole", p, K4, False)[:3]


def ref_geom_pinhole(p, K4):
    """the reference's canonicalProjection + applyIntrinsics: (uv, pixel, Jproj 2x3, Jintr 2x2)"""
    return _geom_pinhole(ref_geom(), "ref_pinhole", p, K4, True)


def _geom_cross(L, fn, v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    M = np.zeros(9)
    getattr(L, fn)(_d(v), _d(M))
    return M.reshape(3, 3)


def geom_cross2matrix(v):
    return _geom_cross(lib(), "orc_geom_cross2matrix", v)


def ref_geom_cross2matrix(v):
    return _geom_cross(ref_geom(), "ref_cross2matrix", v)


def _geom_equirect(L, fn, w, PW, PH):
    w = np.ascontiguousarray(w, dtype=np.float64)
    px = np.zeros(2)
    J = np.zeros(6, np.float32)
    getattr(L, fn)(_d(w), int(PW), int(PH), _d(px), J.ctypes.data_as(_fp))
    return px, J.reshape(2, 3)


def geom_equirect(w, PW, PH):
    return _geom_equirect(lib(), "orc_geom_equirect", w, PW, PH)


def ref_geom_equirect(w, PW, PH):
    return _geom_equirect(ref_geom(), "ref_equirect", w, PW, PH)


# ---- first-party trajectory code of the reference (src/backend/trajectory.cpp compiled with stubs) -----------------
_REF_TRAJ = os.path.join(_HERE, "_ref", "libref_traj.so")
_ref_traj = None
_up = C.POINTER(C.c_uint32)


def have_ref_traj():
    return os.path.exists(_REF_TRAJ)


def ref_traj():
    global _ref_traj
    if _ref_traj is None and have_ref_traj():
        _ref_traj = C.CDLL(_REF_TRAJ)
    return _ref_traj


def _u2(t):
    return np.array([int(t[0]), int(t[1])], dtype=np.uint32)


def ref1p_generate_ctrl_poses(order, dt_knots, t_traj_beg, t_beg, t_end, stamps, poses_xyzw):
    L = ref_traj()
    st = np.ascontiguousarray(stamps, dtype=np.uint32).reshape(-1, 2)
    q = np.ascontiguousarray(poses_xyzw, dtype=np.float64).reshape(-1, 4)
    out = np.zeros((256, 4))
    a, b, c = _u2(t_traj_beg), _u2(t_beg), _u2(t_end)
    L.ref1p_generate_ctrl_poses.restype = C.c_int
    n = L.ref1p_generate_ctrl_poses(int(order), C.c_double(dt_knots), a.ctypes.data_as(_up), b.ctypes.data_as(_up), c.ctypes.data_as(_up),
                                    st.ctypes.data_as(_up), _d(q), len(q), _d(out), 256)
    if n < 0:
        raise ValueError("ref1p_generate_ctrl_poses failed")
    return out[:n].copy()


def ref1p_evaluate(order, t_beg, dt_knots, knots_xyzw, t, want_J=True):
    """(q, idx_beg, J float32 [3, 3*order]) from the real Linear/CubicTrajectory::evaluate"""
    L = ref_traj()
    k = np.ascontiguousarray(knots_xyzw, dtype=np.float64).reshape(-1, 4)
    q = np.zeros(4)
    idx = C.c_int32(0)
    J = np.zeros((3, 3 * order), np.float32)
    tt = _u2(t)
    L.ref1p_evaluate(int(order), C.c_double(t_beg), C.c_double(dt_knots), _d(k), len(k), tt.ctypes.data_as(_up), _d(q), C.byref(idx),
                     J.ctypes.data_as(_fp) if want_J else None)
    return q, idx.value, J


def ref1p_window_evaluate(order, t_traj_beg, dt_knots, knots_xyzw, idx_traj_beg, idx_opt_beg, drotv, t):
    """CopyAndIncrementalUpdate + evaluate on the temporary trajectory: (q, idx_beg, J f32, knots after incrementalUpdate)"""
    L = ref_traj()
    k = np.ascontiguousarray(knots_xyzw, dtype=np.float64).reshape(-1, 4)
    d = np.ascontiguousarray(drotv, dtype=np.float64).reshape(-1)
    q = np.zeros(4)
    idx = C.c_int32(0)
    J = np.zeros((3, 3 * order), np.float32)
    after = np.zeros_like(k)
    a, tt = _u2(t_traj_beg), _u2(t)
    L.ref1p_window_evaluate(int(order), a.ctypes.data_as(_up), C.c_double(dt_knots), _d(k), len(k), int(idx_traj_beg), int(idx_opt_beg), _d(d),
                            tt.ctypes.data_as(_up), _d(q), C.byref(idx), J.ctypes.data_as(_fp), _d(after))
    return q, idx.value, J, after


# ---- the reference's own back-end warp code (src/backend/event_pano_warper.cpp compiled with stubs) --------------------
_REF_WARPER = os.path.join(_HERE, "_ref", "libref_warper.so")
_ref_warper = None


def have_ref_warper():
    return os.path.exists(_REF_WARPER)


class RefEventWarper:
    """EventWarper of the reference (real translation unit) behind oracle/ref_warper_shim.cpp."""

    def __init__(self, lut, SW, SH, PW, PH, order=2, blur_sigma=1.0, batch_size=100, sample_rate=1, max_update_times=10):
        global _ref_warper
        if _ref_warper is None:
            _ref_warper = C.CDLL(_REF_WARPER)
            _ref_warper.ref1p_warper_create.restype = C.c_void_p
        self.L = _ref_warper
        self.lut = np.ascontiguousarray(lut, dtype=np.float64)
        self.PW, self.PH, self.order = PW, PH, order
        self.h = C.c_void_p(self.L.ref1p_warper_create(_d(self.lut), SW, SH, PW, PH, C.c_double(blur_sigma), batch_size, sample_rate,
                                                       max_update_times, order))

    def close(self):
        if self.h:
            self.L.ref1p_warper_destroy(self.h)
            self.h = None

    def set_ig(self, IG):
        a = np.ascontiguousarray(IG, dtype=np.float32)
        self.L.ref1p_warper_set_ig(self.h, a.ctypes.data_as(_fp))

    def get_map(self):
        IG = np.zeros((self.PH, self.PW), np.float32)
        t = np.zeros((self.PH, self.PW), np.uint8)
        self.L.ref1p_warper_get_map(self.h, IG.ctypes.data_as(_fp), t.ctypes.data_as(C.c_void_p))
        return IG, t

    def eval(self, events, t_beg, dt_knots, knots_xyzw, n_fixed, tnext, first_iter, want_grad):
        ev = np.ascontiguousarray(events)
        k = np.ascontiguousarray(knots_xyzw, dtype=np.float64).reshape(-1, 4)
        A = self.PW * self.PH
        P = 3 * (len(k) - n_fixed)
        iwe = np.zeros((self.PH, self.PW), np.float32)
        bands = np.zeros((max(P, 1), self.PH, self.PW), np.float32)
        ilo, iln = np.zeros_like(iwe), np.zeros_like(iwe)
        alpha = C.c_double(0)
        tn = np.array([int(tnext[0]), int(tnext[1])], dtype=np.uint32)
        self.L.ref1p_warper_eval.restype = C.c_int
        nb = self.L.ref1p_warper_eval(self.h, C.c_void_p(ev.ctypes.data), C.c_longlong(len(ev)), C.c_double(t_beg), C.c_double(dt_knots), _d(k), len(k),
                                      int(n_fixed), tn.ctypes.data_as(C.POINTER(C.c_uint32)), int(first_iter), int(want_grad), iwe.ctypes.data_as(_fp),
                                      bands.ctypes.data_as(_fp), ilo.ctypes.data_as(_fp), iln.ctypes.data_as(_fp), C.byref(alpha))
        return {"iwe": iwe, "bands": bands[:P] if want_grad else None, "il_old": ilo, "il_new": iln, "alpha": alpha.value, "n_bands": nb}

    def update_ig(self):
        self.L.ref1p_warper_update_ig(self.h)

    def mark_fov(self, q_xyzw, radius):
        q = np.ascontiguousarray(q_xyzw, dtype=np.float64)
        self.L.ref1p_warper_mark_fov(self.h, _d(q), int(radius))


# ---- the reference's own front-end image builder and focus functions (compiled with stubs) ----------------------------
_REF_FE = os.path.join(_HERE, "_ref", "libref_fe.so")
_REF_FOCUS = os.path.join(_HERE, "_ref", "libref_focus.so")
_ref_fe = _ref_focus = None


def have_ref_firstparty():
    return all(os.path.exists(p) for p in (_REF_FE, _REF_FOCUS, _REF_WARPER))


def ref1p_fe_images(events, t_ref, lut, W, H, K4, omega, want_grad=True, blur_sigma=1.0, batch_size=100):
    """AngVelEstimator::computeImageOfWarpedEvents of the reference (real translation unit): (iwe [H,W], deriv [H,W,3] or None).
    t_ref = (sec, nsec) of time_packet_."""
    global _ref_fe
    if _ref_fe is None:
        _ref_fe = C.CDLL(_REF_FE)
    ev = np.ascontiguousarray(events)
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    K = np.ascontiguousarray(K4, dtype=np.float64)
    om = np.ascontiguousarray(omega, dtype=np.float64)
    tr = np.array([int(t_ref[0]), int(t_ref[1])], dtype=np.uint32)
    iwe = np.zeros((H, W), np.float32)
    der = np.zeros((H, W, 3), np.float32)
    _ref_fe.ref1p_fe_images(C.c_void_p(ev.ctypes.data), C.c_longlong(len(ev)), tr.ctypes.data_as(C.POINTER(C.c_uint32)), _d(lut), int(W), int(H), _d(K),
                            C.c_double(blur_sigma), int(batch_size), _d(om), iwe.ctypes.data_as(_fp), der.ctypes.data_as(_fp) if want_grad else None)
    return iwe, (der if want_grad else None)


def _focus():
    global _ref_focus
    if _ref_focus is None:
        _ref_focus = C.CDLL(_REF_FOCUS)
        _ref_focus.ref1p_fe_contrast.restype = C.c_double
        _ref_focus.ref1p_be_contrast.restype = C.c_double
    return _ref_focus


def ref1p_fe_contrast(iwe, deriv, measure=0):
    """computeContrast of src/frontend/local_focus_funcs.cpp: (contrast, grad[3] or None)"""
    iwe = np.ascontiguousarray(iwe, dtype=np.float32)
    H, W = iwe.shape
    if deriv is None:
        return _focus().ref1p_fe_contrast(iwe.ctypes.data_as(_fp), None, W, H, int(measure), None), None
    d = np.ascontiguousarray(deriv, dtype=np.float32)
    g = np.zeros(3)
    c = _focus().ref1p_fe_contrast(iwe.ctypes.data_as(_fp), d.ctypes.data_as(_fp), W, H, int(measure), _d(g))
    return c, g


def ref1p_be_contrast(iwe, bands, measure=0):
    """computeContrast of src/backend/global_focus_funcs.cpp: (contrast, grad[P] or None)"""
    iwe = np.ascontiguousarray(iwe, dtype=np.float32)
    H, W = iwe.shape
    if bands is None:
        return _focus().ref1p_be_contrast(iwe.ctypes.data_as(_fp), None, 0, W, H, int(measure), None), None
    b = np.ascontiguousarray(bands, dtype=np.float32)
    g = np.zeros(b.shape[0])
    c = _focus().ref1p_be_contrast(iwe.ctypes.data_as(_fp), b.ctypes.data_as(_fp), b.shape[0], W, H, int(measure), _d(g))
    return c, g


# ---- the reference's own node classes driven without ROS (oracle/ref_node_shim.cpp) --------------------------------------
_REF_NODE = os.path.join(_HERE, "_ref", "libref_node.so")
_REF_FULL = os.path.join(_HERE, "_ref", "libref_full.so")


def have_ref_full():
    return os.path.exists(_REF_FULL)


def have_ref_node():
    return os.path.exists(_REF_NODE)


class RefNode:
    """AngVelEstimator + PoseGraphOptimizer of the reference (real translation units); the two GSL solves are stand-ins: the
    front-end returns the next row of `omegas`, the back-end leaves the control poses unchanged."""

    def __init__(self, W, H, K4, omegas, dt_ang_vel=0.01, num_events_per_packet=2000, fe_sample_rate=1, win_size=0.2, win_stride=0.1,
                 dt_knots=0.05, spline_degree=1, pano_height=64, y_angle=0.0, min_ev_rate=10, max_update_times=10, full=False):
        """full=True: oracle/_ref/libref_full.so -- the reference's own *_optim_contrast_gsl*.cpp run the solves (over the GSL
        stand-in of oracle/stubs/gsl); `omegas` is then unused."""
        self.L = C.CDLL(_REF_FULL if full else _REF_NODE)
        self.L.ref1p_node_create.restype = C.c_void_p
        K = np.ascontiguousarray(K4, dtype=np.float64)
        om = np.ascontiguousarray(omegas, dtype=np.float64).reshape(-1, 3)
        self.h = C.c_void_p(self.L.ref1p_node_create(int(W), int(H), _d(K), C.c_double(dt_ang_vel), int(num_events_per_packet), int(fe_sample_rate),
                                                     C.c_double(win_size), C.c_double(win_stride), C.c_double(dt_knots), int(spline_degree),
                                                     int(pano_height), C.c_double(y_angle), int(min_ev_rate), int(max_update_times), _d(om), len(om)))

    def close(self):
        if self.h:
            self.L.ref1p_node_destroy(self.h)
            self.h = None

    def events(self, ev):
        ev = np.ascontiguousarray(ev)
        self.L.ref1p_node_events(self.h, C.c_void_p(ev.ctypes.data), C.c_longlong(len(ev)))

    def counts(self):
        a, b, c = C.c_int(0), C.c_int(0), C.c_longlong(0)
        self.L.ref1p_node_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def packet(self, i):
        v = (C.c_longlong * 7)()
        h = C.c_ulonglong(0)
        self.L.ref1p_node_packet(self.h, int(i), v, C.byref(h))
        return list(v), h.value

    def packet_omega(self, i):
        w = np.zeros(3)
        self.L.ref1p_node_packet_omega(self.h, int(i), _d(w))
        return w

    def get_map(self, PW, PH):
        IG = np.zeros((PH, PW), np.float32)
        t = np.zeros((PH, PW), np.uint8)
        self.L.ref1p_node_get_map(self.h, IG.ctypes.data_as(_fp), t.ctypes.data_as(C.c_void_p))
        return IG, t

    def window(self, i):
        v = (C.c_longlong * 15)()
        h = C.c_ulonglong(0)
        q = np.zeros(4)
        knots = np.zeros((512, 4))
        self.L.ref1p_node_window(self.h, int(i), v, C.byref(h), _d(q), _d(knots), 2048)
        v = list(v)
        return v, h.value, q, knots[: v[9]].copy()


def hash_events(ev):
    """FNV-style hash of (x, y, sec, nsec) per event, as oracle/ref_node_shim.cpp computes it"""
    h = 1469598103934665603
    M = (1 << 64) - 1
    for e in ev:
        for x in (int(e["x"]), int(e["y"]), int(e["sec"]), int(e["nsec"])):
            h ^= x
            h = (h * 1099511628211) & M
    return h
