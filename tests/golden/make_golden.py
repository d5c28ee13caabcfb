"""Generates the committed golden fixtures under tests/golden/ (run HERE, in the build container).

  blur_cv2.npz     cv2 4.13 outputs (GaussianBlur / getGaussianKernel / meanStdDev / the
                   contrast_Variance op chain) -- pins the oracle's restatement of the un-vendored
                   OpenCV arithmetic the reference calls.
  spline_ref.npz   outputs of the REAL basalt::So3Spline / Sophus code of the reference, through
                   oracle/_ref/libref_basalt.so (built from /root/reference by oracle/Makefile).
  fe_oracle.npz, be_oracle.npz
                   regression pins of the oracle itself on small seeded inputs (the reference has
                   no golden vectors for the warp/contrast arithmetic: "parity unpinned").
Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
from cmax_slam_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def golden_blur():
    import cv2
    rng = np.random.default_rng(11)
    d = {"cv2_version": np.array(cv2.__version__)}
    # a, b: row length (W*channels) a multiple of the AVX2 width -> whole rows go through OpenCV's
    # vector body; c, d: ragged widths whose tail columns take OpenCV's scalar path
    for name, shape in (("a", (40, 64)), ("b", (64, 96, 3)), ("c", (37, 53)), ("d", (20, 19))):
        img = (rng.gamma(0.5, 2.0, shape) * (rng.random(shape) < 0.3)).astype(np.float32)
        d[f"img_{name}"] = img
        for sigma in (0.5, 1.0, 2.0):
            d[f"blur_{name}_{sigma}"] = cv2.GaussianBlur(img, (0, 0), sigma)
    for sigma in (0.5, 1.0, 1.5, 2.0, 3.0):
        ks = int(round(sigma * 8 + 1)) | 1
        d[f"kernel_{sigma}"] = cv2.getGaussianKernel(ks, sigma, cv2.CV_32F).ravel()
    img = d["blur_a_1.0"]
    m, s = cv2.meanStdDev(img)
    d["meanstd_a"] = np.array([m.ravel()[0], s.ravel()[0]])
    # contrast_Variance op chain (local_focus_funcs.cpp:26-44) with cv2 primitives
    deriv = cv2.GaussianBlur((rng.normal(0, 1, (40, 64, 3))).astype(np.float32), (0, 0), 1.0)
    d["var_deriv"] = deriv
    zm = (img * np.float32(2.0) + np.float32(-2.0 * m.ravel()[0])).astype(np.float32)  # convertTo(alpha=2, beta=-2 mean)
    g = []
    for c in range(3):
        ch = np.ascontiguousarray(deriv[..., c])
        mc = cv2.mean(ch)[0]
        g.append(cv2.mean(cv2.multiply(zm, cv2.subtract(ch, mc)))[0])
    d["var_grad"] = np.array(g)
    d["ms_contrast"] = np.array(cv2.norm(img, cv2.NORM_L2SQR) / img.size)
    d["ms_grad"] = np.array([2.0 * cv2.mean(cv2.multiply(img, np.ascontiguousarray(deriv[..., c])))[0] for c in range(3)])
    np.savez_compressed(os.path.join(OUT, "blur_cv2.npz"), **d)


def golden_spline():
    assert O.have_ref(), "oracle/_ref/libref_basalt.so missing: run `make -C oracle` with /root/reference present"
    rng = np.random.default_rng(5)
    d = {}
    for order in (2, 4):
        K = 10
        knots = np.zeros((K, 4))
        knots[0] = [0, 0, 0, 1]
        for k in range(1, K):
            knots[k] = synth._qmul(knots[k - 1], synth._qexp(rng.normal(0, 0.25, 3)))
            knots[k] /= np.linalg.norm(knots[k])
        t0, dt = 1_600_000_000_250_000_000, 50_000_000
        ts = rng.integers(t0, t0 + (K - order + 1) * dt - 1, 64)
        ts[0] = t0
        ts[1] = t0 + (K - order + 1) * dt - 1
        q, R, idx, J = [], [], [], []
        for t in ts:
            r = O.ref_spline_eval(order, knots, t0, dt, int(t))
            q.append(r[0]); R.append(r[1]); idx.append(r[2]); J.append(r[3])
        d[f"knots_{order}"] = knots
        d[f"t_{order}"] = ts
        d[f"q_{order}"] = np.array(q); d[f"R_{order}"] = np.array(R)
        d[f"idx_{order}"] = np.array(idx); d[f"J_{order}"] = np.array(J)
        d["t0"] = np.array(t0); d["dt"] = np.array(dt)
    # exp / log / left Jacobians, incl. the small-angle branches
    import ctypes as C
    ws = np.concatenate([rng.normal(0, 1.0, (40, 3)), rng.normal(0, 1e-6, (8, 3)), np.zeros((1, 3)),
                         rng.normal(0, 1e-12, (4, 3))])
    qs, logs, Jl, Jli = [], [], [], []
    L = O.ref()
    for w in ws:
        w = np.ascontiguousarray(w)
        q = np.zeros(4); L.ref_so3_exp(O._d(w), O._d(q)); qs.append(q)
        lg = np.zeros(3); L.ref_so3_log(O._d(q), O._d(lg)); logs.append(lg)
        a = np.zeros(9); b = np.zeros(9); L.ref_left_jacobians(O._d(w), O._d(a), O._d(b)); Jl.append(a); Jli.append(b)
    d["w"] = ws; d["exp_w"] = np.array(qs); d["log_exp_w"] = np.array(logs)
    d["Jl"] = np.array(Jl); d["Jlinv"] = np.array(Jli)
    np.savez_compressed(os.path.join(OUT, "spline_ref.npz"), **d)


def golden_fe():
    pk = synth.make_fe_packet(5000, 64, 48, (60.0, 61.0, 31.5, 23.5), 21, 150)
    a = O.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K)
    d = {"events": pk.events, "t_ref_sec": np.array(pk.t_ref_sec), "K": np.array(pk.K), "W": np.array(64), "H": np.array(48)}
    oms = np.array([[0.6, -1.1, 2.3], [0, 0, 0], [1.0, 0.5, -3.0]])
    d["omegas"] = oms
    for i, om in enumerate(oms):
        for measure in (0, 1):
            a.contrast_measure = measure
            r = O.fe_eval(a, om, True, images=(measure == 0), cells=True)
            d[f"contrast_{i}_{measure}"] = np.array(r["contrast"]); d[f"grad_{i}_{measure}"] = r["grad"]
            if measure == 0:
                d[f"cells_{i}"] = r["cells"]; d[f"iwe_{i}"] = r["iwe"]; d[f"deriv_{i}"] = r["deriv"]
    np.savez_compressed(os.path.join(OUT, "fe_oracle.npz"), **d)


def golden_be():
    d = {}
    for order in (2, 4):
        w = synth.make_be_window(4001, 8, 128, 64, 31 + order, order=order, sensor=(64, 48), K4=(60.0, 61.0, 31.5, 23.5),
                                 n_landmarks=300, n_fixed=1 if order == 2 else 3)
        rng = np.random.default_rng(order)
        IGp = np.abs(rng.normal(0, 0.3, (64, 128))).astype(np.float32)
        P = 3 * (8 - w.n_fixed)
        x = rng.normal(0, 0.02, P)
        a = O.be_args(w.events, w.lut, 64, 48, 128, 64, w.knots_xyzw, w.t0_ns, w.dt_ns, order, w.n_fixed, w.tnext, IGp, 0.5,
                      sample_rate=2 if order == 4 else 1)
        r = O.be_eval(a, x, True, images=True, cells=True)
        d[f"events_{order}"] = w.events; d[f"knots_{order}"] = w.knots_xyzw
        d[f"t0_{order}"] = np.array(w.t0_ns); d[f"dt_{order}"] = np.array(w.dt_ns); d[f"tnext_{order}"] = np.array(w.tnext)
        d[f"IGp_{order}"] = IGp; d[f"x_{order}"] = x
        d[f"contrast_{order}"] = np.array(r["contrast"]); d[f"grad_{order}"] = r["grad"]
        d[f"cells_{order}"] = r["cells"]; d[f"iwe_{order}"] = r["iwe"]
        d[f"il_old_{order}"] = r["il_old"]; d[f"il_new_{order}"] = r["il_new"]
    np.savez_compressed(os.path.join(OUT, "be_oracle.npz"), **d)


def _rand_walk_poses(rng, n, t0, span, sigma):
    """n poses on a random SO(3) walk at sorted stamps in [t0, t0 + span] (ns-integral)."""
    ts = np.sort(rng.uniform(t0, t0 + span, n))
    stamps = np.array([[int(np.floor(t)), int(round((t - np.floor(t)) * 1e9)) % 1000000000] for t in ts], dtype=np.uint32)
    L = O.ref()
    q = np.zeros((n, 4)); q[0] = [0, 0, 0, 1]
    cur = np.array([0.0, 0.0, 0.0, 1.0])
    # start from a non-trivial rotation
    e = np.zeros(4); L.ref_so3_exp(O._d(np.array([0.3, -0.2, 0.5])), O._d(e)); cur = e.copy(); q[0] = cur
    for i in range(1, n):
        out = np.zeros(4)
        L.ref_knot_update(O._d(rng.normal(0, sigma, 3)), O._d(cur), O._d(out))
        cur = out; q[i] = cur
    return stamps, q


def golden_traj():
    """Trajectory initialisation through the REAL Eigen FullPivHouseholderQR / Sophus code
    of the reference (oracle/_ref): control-pose fits (linear, cubic; well- and ill-posed) and the angular-velocity
    integration."""
    assert O.have_ref()
    rng = np.random.default_rng(17)
    d = {}
    cases = []
    # (order, dt_knots, span, n_poses, sigma)
    for ci, (order, dtk, span, n, sig) in enumerate([(2, 0.1, 0.5, 40, 0.02), (4, 0.1, 0.5, 40, 0.02), (2, 0.05, 0.2, 9, 0.05),
                                                     (4, 0.05, 0.2, 9, 0.05), (4, 0.1, 1.0, 200, 0.01), (2, 0.02, 0.2, 14, 0.03)]):
        t0 = 1.6e9 + 37.123456789 + ci
        stamps, q = _rand_walk_poses(rng, n, t0 + 1e-4, span - 2e-4, sig)
        num = int(round(span / dtk)) + (3 if order == 4 else 1)
        ctrl = O.ref_fit_ctrl_poses(order, dtk, t0, num, stamps, q)
        d[f"fit{ci}_in"] = np.array([order, dtk, t0, num])
        d[f"fit{ci}_stamps"] = stamps; d[f"fit{ci}_poses"] = q; d[f"fit{ci}_ctrl"] = ctrl
        cases.append(ci)
    # rank-deficient system: all poses inside the first knot interval of a 6-control-pose linear fit
    stamps, q = _rand_walk_poses(rng, 12, 100.0 + 1e-3, 0.09, 0.02)
    d["fitdef_in"] = np.array([2, 0.1, 100.0, 6]); d["fitdef_stamps"] = stamps; d["fitdef_poses"] = q
    d["fitdef_ctrl"] = O.ref_fit_ctrl_poses(2, 0.1, 100.0, 6, stamps, q)
    d["n_fit"] = np.array(len(cases))
    # integration: first window (nothing skipped) and a later window with stale entries
    for ci, first in enumerate((1, 0)):
        m = 30
        ts = np.sort(rng.uniform(50.0, 50.6, m))
        stamps = np.array([[int(np.floor(t)), int((t - np.floor(t)) * 1e9)] for t in ts], dtype=np.uint32)
        w = rng.normal(0, 1.5, (m, 3))
        latest_stamp = np.array([49, 999000000], np.uint32)
        latest_q = np.array([0.1, -0.2, 0.05, 0.0]); latest_q[3] = np.sqrt(1 - (latest_q[:3] ** 2).sum())
        prev_stamp = stamps[4].copy() if not first else np.array([49, 990000000], np.uint32)
        prev_w = rng.normal(0, 1.0, 3)
        os_, oq, ps, pw = O.ref_integrate_ang_vel(latest_stamp, latest_q, prev_stamp, prev_w, first, stamps, w)
        d[f"int{ci}_stamps"] = stamps; d[f"int{ci}_w"] = w; d[f"int{ci}_latest_stamp"] = latest_stamp; d[f"int{ci}_latest_q"] = latest_q
        d[f"int{ci}_prev_stamp"] = prev_stamp; d[f"int{ci}_prev_w"] = prev_w; d[f"int{ci}_first"] = np.array(first)
        d[f"int{ci}_out_stamps"] = os_; d[f"int{ci}_out_q"] = oq; d[f"int{ci}_new_prev_stamp"] = ps; d[f"int{ci}_new_prev_w"] = pw
    np.savez_compressed(os.path.join(OUT, "traj_ref.npz"), **d)


def golden_lut():
    """cv2.undistortPoints (4.13) on every pixel of small sensors with DAVIS-like calibrations: pins the bearing-vector
    LUT restatement (oracle/lut_py.py) and the device kernel (csrc/camera.cu)."""
    import cv2
    d = {"cv2_version": np.array(cv2.__version__)}
    cams = {
        # DAVIS240C-like plumb_bob calibration, identity rectification, P = [K | 0]
        "a": dict(W=240, H=180, K=[199.09, 0, 132.19, 0, 198.83, 110.71, 0, 0, 1], D=[-0.368, 0.150, -0.0003, -0.0002, 0.0],
                  R=np.eye(3).ravel(), P=[199.09, 0, 132.19, 0, 0, 198.83, 110.71, 0, 0, 0, 1, 0]),
        # rectified stereo-style: rotated R, new projection with a baseline term, 8-coefficient rational model
        "b": dict(W=96, H=64, K=[80.5, 0, 47.2, 0, 81.1, 31.7, 0, 0, 1], D=[0.12, -0.25, 0.001, -0.002, 0.05, 0.01, -0.02, 0.003],
                  R=cv2.Rodrigues(np.array([0.01, -0.02, 0.005]))[0].ravel(), P=[78.0, 0, 48.0, -3.9, 0, 78.0, 32.0, 0, 0, 0, 1, 0]),
    }
    for name, c in cams.items():
        W, H = c["W"], c["H"]
        K = np.array(c["K"], float).reshape(3, 3); D = np.array(c["D"], float)
        R = np.array(c["R"], float).reshape(3, 3); P = np.array(c["P"], float).reshape(3, 4)
        ys, xs = np.mgrid[0:H, 0:W]
        pts = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float32).reshape(-1, 1, 2)
        # one point per call, as rectifyPoint does (1x1 CV_32FC2 Mat), would be 43200 calls; a single N x 1 call runs the same scalar code
        rect = cv2.undistortPoints(pts, K, D, R=R, P=P).reshape(-1, 2)
        for k in ("K", "D", "R", "P"):
            d[f"{name}_{k}"] = np.array(c[k], float)
        d[f"{name}_WH"] = np.array([W, H])
        d[f"{name}_rect"] = rect.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "lut_cv2.npz"), **d)


def golden_geom():
    """Outputs of the reference's OWN geometry sources (src/utils/image_geom_util.cpp, include/utils/image_geom_util.h,
    include/backend/equirectangular_camera.h) compiled with container stubs into oracle/_ref/libref_geom.so."""
    assert O.have_ref_geom(), "oracle/_ref/libref_geom.so missing: run `make -C oracle` with /root/reference present"
    rng = np.random.default_rng(23)
    n = 400
    P = np.concatenate([rng.normal(0, 0.4, (n, 2)), rng.uniform(0.3, 2.0, (n, 1))], 1)      # camera-frame points, z > 0
    K4 = np.array([588.0999, 593.9887, 339.8259, 242.4252])
    pin = [O.ref_geom_pinhole(p, K4) for p in P]
    V = rng.normal(0, 1, (50, 3))
    Wd = rng.normal(0, 1, (n, 3)); Wd /= np.linalg.norm(Wd, axis=1, keepdims=True)
    Wd = np.concatenate([Wd, Wd[:50] * rng.uniform(0.5, 3.0, (50, 1))])                      # non-unit rays too
    eq = [O.ref_geom_equirect(w, 1280, 720) for w in Wd] + [O.ref_geom_equirect(w, 4096, 2048) for w in Wd[:100]]
    np.savez_compressed(os.path.join(OUT, "geom_ref.npz"), P=P, K4=K4, uv=np.array([a[0] for a in pin]), px=np.array([a[1] for a in pin]),
                        Jproj=np.array([a[2] for a in pin]), Jintr=np.array([a[3] for a in pin]), V=V,
                        cross=np.array([O.ref_geom_cross2matrix(v) for v in V]), W=Wd, eq_px=np.array([a[0] for a in eq]),
                        eq_J=np.array([a[1] for a in eq]))


def golden_traj_firstparty():
    """Outputs of the reference's OWN src/backend/trajectory.cpp (Linear/CubicTrajectory) compiled with ROS / OpenCV / glog
    stand-ins into oracle/_ref/libref_traj.so: generateCtrlPoses, evaluate (+ f32 knot Jacobians), CopyAndIncrementalUpdate."""
    assert O.have_ref_traj(), "oracle/_ref/libref_traj.so missing: run `make -C oracle` with /root/reference present"
    rng = np.random.default_rng(41)
    d = {}
    n_gen = 0
    for order in (2, 4):
        for dtk, span, n in ((0.1, 0.5, 40), (0.05, 0.2, 12), (0.02, 0.1, 30)):
            t0 = 1.6e9 + 12.25 + n_gen * 0.37
            tb = (int(np.floor(t0)), int(round((t0 - np.floor(t0)) * 1e9)))
            te_s = t0 + span
            te = (int(np.floor(te_s)), int(round((te_s - np.floor(te_s)) * 1e9)) % 1000000000)
            stamps, q = _rand_walk_poses(rng, n, t0 + 1e-4, span - 2e-4, 0.03)
            ctrl = O.ref1p_generate_ctrl_poses(order, dtk, tb, tb, te, stamps, q)
            d[f"gen{n_gen}_in"] = np.array([order, dtk]); d[f"gen{n_gen}_tb"] = np.array(tb, np.uint32); d[f"gen{n_gen}_te"] = np.array(te, np.uint32)
            d[f"gen{n_gen}_stamps"] = stamps; d[f"gen{n_gen}_poses"] = q; d[f"gen{n_gen}_ctrl"] = ctrl
            n_gen += 1
    d["n_gen"] = np.array(n_gen)
    # evaluate / window evaluation
    n_ev = 0
    for order in (2, 4):
        K = 12
        _, knots = _rand_walk_poses(rng, K, 0.0, 1.0, 0.08)
        dtk = 0.05
        t_traj = (1600000123, 456789012)
        t_beg_d = t_traj[0] + 1e-9 * t_traj[1]
        for idx_traj, idx_opt in ((0, 1 if order == 2 else 3), (2, 3), (4, 6)):
            drotv = rng.normal(0, 0.02, 3 * (K - idx_opt))
            n_seg = (K - idx_traj) - order + 1
            for u in rng.uniform(0.01, n_seg - 0.01, 6):
                t_s = (t_beg_d + idx_traj * dtk) + u * dtk
                t = (int(np.floor(t_s)), int((t_s - np.floor(t_s)) * 1e9))
                q, idx, J, after = O.ref1p_window_evaluate(order, t_traj, dtk, knots, idx_traj, idx_opt, drotv, t)
                d[f"win{n_ev}_in"] = np.array([order, dtk, idx_traj, idx_opt]); d[f"win{n_ev}_ttraj"] = np.array(t_traj, np.uint32)
                d[f"win{n_ev}_knots"] = knots; d[f"win{n_ev}_drotv"] = drotv; d[f"win{n_ev}_t"] = np.array(t, np.uint32)
                d[f"win{n_ev}_q"] = q; d[f"win{n_ev}_idx"] = np.array(idx); d[f"win{n_ev}_J"] = J; d[f"win{n_ev}_after"] = after
                n_ev += 1
    d["n_win"] = np.array(n_ev)
    np.savez_compressed(os.path.join(OUT, "traj_firstparty.npz"), **d)


def golden_hotpath_firstparty():
    """Outputs of the reference's OWN hot-path translation units -- src/frontend/local_image_warped_events.cpp,
    src/frontend/local_focus_funcs.cpp, src/backend/event_pano_warper.cpp, src/backend/global_focus_funcs.cpp (+ trajectory.cpp,
    image_geom_util.cpp, equirectangular_camera.h, real basalt/Sophus/Eigen) -- compiled unmodified with the stand-in headers of
    oracle/stubs/ into oracle/_ref/libref_{fe,focus,warper}.so, on small seeded inputs (the inputs are regenerated from the seeds by
    the test; only the outputs are stored)."""
    assert O.have_ref_firstparty(), "oracle/_ref/libref_{fe,focus,warper}.so missing: run `make -C oracle` with /root/reference present"
    d = {}
    # front-end: 4000 events on a 48x36 sensor
    pk = synth.make_fe_packet(4000, 48, 36, (40.0, 41.0, 23.5, 17.5), 71, 120)
    sec = int(np.floor(pk.t_ref_sec)); nsec = int(round((pk.t_ref_sec - sec) * 1e9))
    for k, om in enumerate((pk.omega_true, np.zeros(3), pk.omega_true + np.array([0.3, -0.2, 0.4]))):
        iwe, der = O.ref1p_fe_images(pk.events, (sec, nsec), pk.lut, 48, 36, pk.K, om, True)
        d[f"fe{k}_omega"] = np.asarray(om, float); d[f"fe{k}_iwe"] = iwe; d[f"fe{k}_deriv"] = der
        for m in (0, 1):
            c, g = O.ref1p_fe_contrast(iwe, der, m)
            d[f"fe{k}_contrast{m}"] = np.array(c); d[f"fe{k}_grad{m}"] = g
    d["fe_tref"] = np.array([sec, nsec], np.uint32)
    # back-end: 5000 events, 6 knots, 64x32 panorama, 32x24 sensor; linear and cubic; first iteration with a non-empty map
    for order in (2, 4):
        w = synth.make_be_window(5000, 6, 64, 32, 72 + order, order=order, sensor=(32, 24), K4=(30.0, 30.5, 15.5, 11.5), n_landmarks=150,
                                 n_fixed=1 if order == 2 else 3)
        dtk = w.dt_ns / 1e9
        t_beg = w.t0_ns / 1e9
        rng = np.random.default_rng(5)
        IG = np.abs(rng.normal(0, 0.3, (32, 64))).astype(np.float32)
        rw = O.RefEventWarper(w.lut, 32, 24, 64, 32, order=order)
        rw.set_ig(IG)
        r = rw.eval(w.events, t_beg, dtk, w.knots_xyzw, w.n_fixed, w.tnext, True, True)
        rw.close()
        c, g = O.ref1p_be_contrast(r["iwe"], r["bands"], 0)
        d[f"be{order}_tbeg"] = np.array([t_beg, dtk]); d[f"be{order}_IG"] = IG; d[f"be{order}_alpha"] = np.array(r["alpha"])
        for k in ("iwe", "bands", "il_old", "il_new"):
            d[f"be{order}_{k}"] = r[k]
        d[f"be{order}_contrast"] = np.array(c); d[f"be{order}_grad"] = g
    # the same at the sizes the GPU parity tests use (tests/test_gpu_firstparty.py): no band images stored
    pk = synth.make_fe_packet(6000, 64, 48, (60.0, 61.0, 31.5, 23.5), 81, 200)
    sec = int(np.floor(pk.t_ref_sec)); nsec = int(round((pk.t_ref_sec - sec) * 1e9))
    for k, om in enumerate((pk.omega_true, pk.omega_true + np.array([0.3, -0.2, 0.4]))):
        iwe, der = O.ref1p_fe_images(pk.events, (sec, nsec), pk.lut, 64, 48, pk.K, om, True)
        d[f"gfe{k}_omega"] = np.asarray(om, float); d[f"gfe{k}_iwe"] = iwe; d[f"gfe{k}_deriv"] = der
        for m in (0, 1):
            c, g = O.ref1p_fe_contrast(iwe, der, m)
            d[f"gfe{k}_contrast{m}"] = np.array(c); d[f"gfe{k}_grad{m}"] = g
    for order in (2, 4):
        w = synth.make_be_window(20000, 8, 128, 64, 90 + order, order=order, sensor=(64, 48), K4=(60.0, 61.0, 31.5, 23.5), n_landmarks=300,
                                 n_fixed=1 if order == 2 else 3)
        dtk, t_beg = w.dt_ns / 1e9, w.t0_ns / 1e9
        IG = np.abs(np.random.default_rng(6).normal(0, 0.3, (64, 128))).astype(np.float32)
        rw = O.RefEventWarper(w.lut, 64, 48, 128, 64, order=order)
        rw.set_ig(IG)
        r = rw.eval(w.events, t_beg, dtk, w.knots_xyzw, w.n_fixed, w.tnext, True, True)
        rw.close()
        c, g = O.ref1p_be_contrast(r["iwe"], r["bands"], 0)
        d[f"gbe{order}_tbeg"] = np.array([t_beg, dtk]); d[f"gbe{order}_IG"] = IG; d[f"gbe{order}_alpha"] = np.array(r["alpha"])
        for k in ("iwe", "il_old", "il_new"):
            d[f"gbe{order}_{k}"] = r[k]
        d[f"gbe{order}_contrast"] = np.array(c); d[f"gbe{order}_grad"] = g
    np.savez_compressed(os.path.join(OUT, "hotpath_firstparty.npz"), **d)


def golden_node_firstparty():
    """Packets, windows and control poses produced by the reference's OWN node classes (ang_vel_estimator.cpp,
    pose_graph_optimizer.cpp, ... compiled with stand-in headers, oracle/_ref/libref_node.so) for two seeded event streams."""
    assert O.have_ref_node()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_node_firstparty import _stream
    d = {}
    for tag, degree in (("lin", 1), ("cub", 3)):
        ev = _stream(130000, 21 + degree)
        omegas = np.random.default_rng(4).normal(0, 1.5, (37, 3))
        ref = O.RefNode(64, 48, (60.0, 61.0, 31.5, 23.5), omegas, dt_ang_vel=0.01, num_events_per_packet=2000, fe_sample_rate=1, dt_knots=0.05,
                        spline_degree=degree, y_angle=15.0)
        ref.events(ev)
        n_pk, n_win, n_stored = ref.counts()
        pk = [ref.packet(i) for i in range(n_pk)]
        d[f"{tag}_packets"] = np.array([v for v, _ in pk], dtype=np.int64)
        d[f"{tag}_packet_hash"] = np.array([h for _, h in pk], dtype=np.uint64)
        wins = [ref.window(i) for i in range(n_win)]
        d[f"{tag}_windows"] = np.array([v for v, _, _, _ in wins], dtype=np.int64)
        d[f"{tag}_window_hash"] = np.array([h for _, h, _, _ in wins], dtype=np.uint64)
        d[f"{tag}_latest_q"] = np.array([q for _, _, q, _ in wins])
        for i, (_, _, _, k) in enumerate(wins):
            d[f"{tag}_knots{i}"] = k
        d[f"{tag}_n_stored"] = np.array(n_stored)
        ref.close()
    np.savez_compressed(os.path.join(OUT, "node_firstparty.npz"), **d)


def golden_pipeline_firstparty():
    """The reference run AS A WHOLE on the CPU (node classes + warp + focus + its own optimiser glue over the GSL stand-in,
    oracle/_ref/libref_full.so) on a seeded synthetic sequence: angular velocity per packet, control poses per window, final map."""
    assert O.have_ref_full()
    K_T = (120.0, 122.0, 63.0, 47.0)
    d = {}
    for tag, degree, seed in (("lin", 1, 31), ("cub", 3, 32)):
        w = synth.make_be_window(120000, 9, 256, 128, seed, order=2, sensor=(128, 96), K4=K_T, n_landmarks=800, knot_sigma=0.1)
        ref = O.RefNode(128, 96, K_T, np.zeros((1, 3)), dt_ang_vel=0.01, num_events_per_packet=6000, dt_knots=0.05, spline_degree=degree,
                        pano_height=128, min_ev_rate=10, max_update_times=30, full=True)
        ref.events(w.events)
        n_pk, n_win, _ = ref.counts()
        d[f"{tag}_stamps"] = np.array([ref.packet(i)[0][:2] for i in range(n_pk)], dtype=np.int64)
        d[f"{tag}_omegas"] = np.array([ref.packet_omega(i) for i in range(n_pk)])
        for i in range(n_win):
            v, _, lq, kn = ref.window(i)
            d[f"{tag}_win{i}"] = np.array(v, dtype=np.int64); d[f"{tag}_knots{i}"] = kn; d[f"{tag}_latest{i}"] = lq
        d[f"{tag}_n_win"] = np.array(n_win)
        IG, times = ref.get_map(256, 128)
        d[f"{tag}_IG"] = IG; d[f"{tag}_times"] = times
        ref.close()
    np.savez_compressed(os.path.join(OUT, "pipeline_firstparty.npz"), **d)


if __name__ == "__main__":
    if "--only-traj" in sys.argv:
        golden_traj()
        sys.exit(0)
    if "--only-pipeline1p" in sys.argv:
        golden_pipeline_firstparty()
        sys.exit(0)
    if "--only-node1p" in sys.argv:
        golden_node_firstparty()
        sys.exit(0)
    if "--only-hotpath1p" in sys.argv:
        golden_hotpath_firstparty()
        sys.exit(0)
    if "--only-traj1p" in sys.argv:
        golden_traj_firstparty()
        sys.exit(0)
    if "--only-geom" in sys.argv:
        golden_geom()
        sys.exit(0)
    if "--only-lut" in sys.argv:
        golden_lut()
        sys.exit(0)
    golden_lut()
    golden_geom()
    golden_traj_firstparty()
    golden_hotpath_firstparty()
    golden_node_firstparty()
    golden_pipeline_firstparty()
    golden_traj()
    golden_blur()
    golden_spline()
    golden_fe()
    golden_be()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
