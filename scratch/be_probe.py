import sys, time
sys.path.insert(0, '.')
import numpy as np
from cmax_slam_b200 import synth, _capi
from cmax_slam_b200.backend import EventWarperCMax, GRAD_ADJOINT, GRAD_DENSE
from oracle import oracle_py as O

def run(n_ev, K, PW, PH, order, seed, dense_check=True, sample_rate=1):
    w = synth.make_be_window(n_ev, K, PW, PH, seed, order=order, n_landmarks=max(2000, n_ev // 200), n_fixed=1 if order == 2 else 3)
    rng = np.random.default_rng(seed)
    IGp = np.abs(rng.normal(0, 0.3, (PH, PW))).astype(np.float32)
    P = 3 * (K - w.n_fixed)
    x = rng.normal(0, 0.01, P)
    a = O.be_args(w.events, w.lut, w.sensor_width, w.sensor_height, PW, PH, w.knots_xyzw, w.t0_ns, w.dt_ns, order, w.n_fixed, w.tnext, IGp, 0.5, sample_rate=sample_rate)
    t = time.time(); ro = O.be_eval(a, x, True, images=dense_check, cells=True); t_cpu = time.time() - t
    for mode in ((GRAD_DENSE, GRAD_ADJOINT) if dense_check else (GRAD_ADJOINT,)):
        be = EventWarperCMax(w.sensor_width, w.sensor_height, w.lut, PW, PH, spline_order=order, grad_mode=mode, event_sample_rate=sample_rate)
        be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
        cells = be.warped_cells(x)
        c, g = be.eval(x, True)
        c0, _ = be.eval(x, False)
        print(f"order {order} n {len(w.events)} K {K} pano {PW}x{PH} mode {mode}: cells mismatch {(cells != ro['cells']).sum()} | C {c} C0 {c0} oracle {ro['contrast']} rel {abs(c-ro['contrast'])/ro['contrast']:.2e}")
        print("   grad relerr(max/maxabs)", np.abs(g - ro["grad"]).max() / np.abs(ro["grad"]).max(), "max |g|", np.abs(ro["grad"]).max())
        if dense_check:
            ilo, iln = be.local_iwe(x)
            print("   il_old maxdiff", np.abs(ilo - ro["il_old"]).max(), "il_new", np.abs(iln - ro["il_new"]).max(), "iwe", np.abs(be.computeImageOfWarpedEvents(x) - ro["iwe"]).max())
            bands = be.derivative_bands(x, True)
            print("   bands maxdiff", np.abs(bands - ro["bands"]).max(), "max", np.abs(ro["bands"]).max())
        for want in (False, True):
            for _ in range(3): be.eval(x, want)
            t = time.time(); N = 10
            for _ in range(N): be.eval(x, want)
            dt = (time.time() - t) / N
            print(f"   eval want_grad={want}: {dt*1e6:.1f} us -> {len(w.events)/dt:.3e} ev/s (cpu oracle f+g {t_cpu*1e3:.1f} ms)")
        be.profile(True)
        for _ in range(5): be.eval(x, True)
        print("   kernels:", {k: (round(v[0] / v[1] * 1e3, 2), v[1]) for k, v in be.kernel_times().items()}, "us avg")
        be.close()
    # alpha = NaN path
    be = EventWarperCMax(w.sensor_width, w.sensor_height, w.lut, PW, PH, spline_order=order, event_sample_rate=sample_rate)
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp)
    c, _ = be.eval(x, False)
    ilo, iln = be.local_iwe(x)
    print("   alpha gpu", be.alpha, "oracle", O.update_alpha(IGp, ro["il_old"] + ro["il_new"]) if dense_check else None)
    be.close()

run(60000, 8, 512, 256, 2, 7)
run(60001, 8, 512, 256, 4, 8, sample_rate=3)
run(1_000_000, 64, 1280, 720, 2, 4, dense_check=False)
run(1_000_000, 64, 1280, 720, 4, 4, dense_check=False)
print("launches", _capi.launch_count())
