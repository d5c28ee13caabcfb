# back-end evaluation latency on small windows (the sizes the reference's launch files produce) and on C4, graph replay on / off
import os, sys, time; sys.path.insert(0, '.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.backend import EventWarperCMax
tag = os.environ.get("PROBE_TAG", "")
def case(name, n, knots, pw, ph, dev=None):
    w = synth.make_be_window(n, knots, pw, ph, 7, order=2, n_landmarks=2000) if dev is None else synth.make_be_window_torch(n, knots, pw, ph, 7, order=2, n_landmarks=50000, device=dev)
    rng = np.random.default_rng(1)
    IGp = np.abs(rng.normal(0, 0.3, (ph, pw))).astype(np.float32)
    be = EventWarperCMax(w.sensor_width, w.sensor_height, w.lut, pw, ph, spline_order=2)
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
    x = rng.normal(0, 0.01, 3 * (knots - w.n_fixed))
    for grad in (True, False):
        for _ in range(5): be.eval(x, grad)
        t = time.perf_counter(); k = 200 if n < 1e6 else 20
        for _ in range(k): c, g = be.eval(x, grad)
        us = (time.perf_counter() - t) / k * 1e6
        print(f"{tag} {name}: n={n} {'f+g' if grad else 'value'} {us:.1f} us contrast {c:.9f}" + (f" |g| {np.abs(g).max():.6e}" if grad else ""), flush=True)
    be.close()
case("small window", 9000, 8, 1024, 512)
case("medium window", 200000, 12, 1024, 512)
if len(sys.argv) > 1: case("C4", 10_000_000, 64, 1280, 720, dev="cuda")
