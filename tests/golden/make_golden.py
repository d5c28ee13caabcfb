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


if __name__ == "__main__":
    golden_blur()
    golden_spline()
    golden_fe()
    golden_be()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
