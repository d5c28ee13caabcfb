# per-kernel-class times of small / medium back-end windows (profiler on: launch by launch, CUDA events around each class)
import sys; sys.path.insert(0, '.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.backend import EventWarperCMax
for name, n, knots in (("small", 9000, 8), ("medium", 200000, 12), ("1M", 1000000, 24)):
    w = synth.make_be_window(n, knots, 1024, 512, 7, order=2, n_landmarks=2000)
    rng = np.random.default_rng(1)
    IGp = np.abs(rng.normal(0, 0.3, (512, 1024))).astype(np.float32)
    be = EventWarperCMax(w.sensor_width, w.sensor_height, w.lut, 1024, 512, spline_order=2)
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
    x = rng.normal(0, 0.01, 3 * (knots - w.n_fixed))
    for _ in range(5): be.eval(x, True)
    be.profile(True)
    for _ in range(20): be.eval(x, True)
    kt = be.kernel_times(); be.profile(False)
    print(name, n, {k: round(v[0] / v[1] * 1e3, 1) for k, v in kt.items()}, "sum %.1f us" % sum(v[0] / 20 * 1e3 for v in kt.values()))
    be.close()
