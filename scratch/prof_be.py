import sys, time; sys.path.insert(0,'.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.backend import EventWarperCMax
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
w = synth.be_config("C4", scale=scale)
rng = np.random.default_rng(4)
IGp = np.abs(rng.normal(0, 0.3, (720, 1280))).astype(np.float32)
be = EventWarperCMax(w.sensor_width, w.sensor_height, w.lut, 1280, 720, spline_order=2)
be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
x = rng.normal(0, 0.01, 3*63)
for i in range(3): be.eval(x, True)
t=time.time(); N=5
for i in range(N): be.eval(x, True)
dt=(time.time()-t)/N
t=time.time()
for i in range(N): be.eval(x, False)
dt0=(time.time()-t)/N
print(f"BE C4 x{scale}: n={len(w.events)} f+g {dt*1e6:.1f} us ({len(w.events)/dt:.3e} ev/s)  value {dt0*1e6:.1f} us ({len(w.events)/dt0:.3e} ev/s)")
be.profile(True)
for i in range(3): be.eval(x, True)
print({k: (round(v[0]/v[1]*1e3,1), v[1]) for k,v in be.kernel_times().items()})
be.close()
